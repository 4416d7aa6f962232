"""CPU-side checks of the C-ABI library and the Python mirror module (no GPU needed).

Covers: libgsfield.so loads and exports every symbol include/gsfield.h declares; argument
validation mirrors the reference's asserts/panics (src/field.rs:44-46,104-106,177-181,163);
and the product path fails loudly -- no CPU fallback -- when no CUDA device is present.
"""
import ctypes
import os
import re

import numpy as np
import pytest

import gstools_core as gc
from conftest import ROOT

HAVE_GPU = gc.device_count() > 0


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gsfield.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(gsf_[a-z_0-9]+)\s*\(", hdr))
    assert {"gsf_summate", "gsf_summate_incompr", "gsf_summate_fourier", "gsf_summate_on_stream",
            "gsf_dfma_peak", "gsf_get_last_stats"} <= names
    lib = ctypes.CDLL(os.path.join(ROOT, "gstools-core_b200", "gstools_core", "libgsfield.so"))
    for n in sorted(names):
        assert hasattr(lib, n), "libgsfield.so does not export %s" % n
    assert lib.gsf_abi_version() == 2


def test_module_surface_matches_reference():
    # src/lib.rs:29-84: __version__, summate, summate_incompr, summate_fourier
    for n in ("__version__", "summate", "summate_incompr", "summate_fourier"):
        assert hasattr(gc, n)
    import inspect
    assert list(inspect.signature(gc.summate).parameters) == ["cov_samples", "z1", "z2", "pos", "num_threads"]
    assert list(inspect.signature(gc.summate_incompr).parameters) == ["cov_samples", "z1", "z2", "pos", "num_threads"]
    assert list(inspect.signature(gc.summate_fourier).parameters) == [
        "spectrum_factor", "modes", "z1", "z2", "pos", "num_threads"]


def test_dtype_and_rank_are_type_errors():
    k = np.ones((2, 3)); z = np.ones(3); pos = np.ones((2, 4))
    with pytest.raises(TypeError):
        gc.summate(k.astype(np.float32), z, z, pos)       # PyReadonlyArray<f64> never casts
    with pytest.raises(TypeError):
        gc.summate(k, z, z, [[1.0, 2.0], [3.0, 4.0]])     # not an ndarray
    with pytest.raises(TypeError):
        gc.summate(k, z.reshape(3, 1), z, pos)            # rank mismatch
    with pytest.raises(OverflowError):
        gc.summate(k, z, z, pos, num_threads=-1)          # Option<usize>


def test_shape_mismatch_is_value_error():
    k = np.ones((2, 3)); z = np.ones(3)
    with pytest.raises(ValueError):
        gc.summate(k, z, z, np.ones((3, 4)))              # field.rs:44
    with pytest.raises(ValueError):
        gc.summate(k, np.ones(4), z, np.ones((2, 4)))     # field.rs:45
    with pytest.raises(ValueError):
        gc.summate_incompr(k, z, np.ones(2), np.ones((2, 4)))  # field.rs:106
    with pytest.raises(ValueError):
        gc.summate_fourier(np.ones(2), k, z, z, np.ones((2, 4)))


def test_c_abi_validation_codes():
    L = gc._load()
    z = np.ones(3); k = np.ones((9, 3)); pos = np.ones((9, 4)); out = np.zeros(4)
    rc = L.gsf_summate(0, 3, 4, k.ctypes.data, 3, 1, z.ctypes.data, 1, z.ctypes.data, 1,
                       pos.ctypes.data, 4, 1, out.ctypes.data, 0)
    assert rc == 1 and b"dim" in L.gsf_last_error()                       # GSF_ERR_DIM
    if gc.device_count() == 0:   # dim = 9 passes validation (any dim works, as in the reference)
        rc = L.gsf_summate(9, 3, 4, k.ctypes.data, 3, 1, z.ctypes.data, 1, z.ctypes.data, 1,
                           pos.ctypes.data, 4, 1, out.ctypes.data, 0)
        assert rc == 4                                                     # GSF_ERR_NO_DEVICE
    rc = L.gsf_summate_incompr(1, 3, 4, k.ctypes.data, 3, 1, z.ctypes.data, 1, z.ctypes.data, 1,
                               pos.ctypes.data, 4, 1, out.ctypes.data, 1, 1, 0)
    assert rc == 1 and b"two- and three-dimensional" in L.gsf_last_error()  # field.rs:180
    rc = L.gsf_summate_incompr(2, 0, 4, k.ctypes.data, 3, 1, z.ctypes.data, 1, z.ctypes.data, 1,
                               pos.ctypes.data, 4, 1, out.ctypes.data, 1, 2, 0)
    assert rc == 3                                                         # GSF_ERR_EMPTY_MODES, field.rs:163
    rc = L.gsf_summate(2, -1, 4, k.ctypes.data, 3, 1, z.ctypes.data, 1, z.ctypes.data, 1,
                       pos.ctypes.data, 4, 1, out.ctypes.data, 0)
    assert rc == 2                                                         # GSF_ERR_SHAPE
    rc = L.gsf_summate(2, 3, 4, None, 3, 1, z.ctypes.data, 1, z.ctypes.data, 1,
                       pos.ctypes.data, 4, 1, out.ctypes.data, 0)
    assert rc == 6                                                         # GSF_ERR_ARG


def test_incompr_python_errors():
    with pytest.raises(ValueError):
        gc.summate_incompr(np.ones((1, 3)), np.ones(3), np.ones(3), np.ones((1, 4)))
    with pytest.raises(ValueError):
        gc.summate_incompr(np.ones((2, 0)), np.ones(0), np.ones(0), np.ones((2, 4)))


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-device behaviour")
def test_no_device_fails_loudly():
    k = np.ones((2, 3)); z = np.ones(3); pos = np.ones((2, 4))
    for call in (lambda: gc.summate(k, z, z, pos),
                 lambda: gc.summate_incompr(k, z, z, pos),
                 lambda: gc.summate_fourier(z, k, z, z, pos),
                 lambda: gc.dfma_peak(0, 1.0)):
        with pytest.raises(RuntimeError, match="no CUDA device"):
            call()


def test_krige_surface_and_validation():
    # src/lib.rs:86-118: calc_field_krige_and_variance / calc_field_krige (krige_mat, krig_vecs, cond, num_threads)
    import inspect
    for fn in (gc.calc_field_krige, gc.calc_field_krige_and_variance):
        assert list(inspect.signature(fn).parameters) == ["krige_mat", "krig_vecs", "cond", "num_threads"]
    mat = np.eye(3); vecs = np.ones((3, 5)); cond = np.ones(3)
    with pytest.raises(ValueError):
        gc.calc_field_krige(mat, vecs[:2], cond)            # src/krige.rs:31
    with pytest.raises(ValueError):
        gc.calc_field_krige_and_variance(mat, vecs, np.ones(4))   # :32
    with pytest.raises(TypeError):
        gc.calc_field_krige(mat.astype(np.float32), vecs, cond)
    L = gc._load()
    f = np.zeros(5)
    assert L.gsf_krige(-1, 5, mat.ctypes.data, 3, 1, vecs.ctypes.data, 5, 1, cond.ctypes.data, 1,
                       f.ctypes.data, None, 0) == 2          # GSF_ERR_SHAPE
    assert L.gsf_krige(3, 5, None, 3, 1, vecs.ctypes.data, 5, 1, cond.ctypes.data, 1,
                       f.ctypes.data, None, 0) == 6          # GSF_ERR_ARG
    if not HAVE_GPU:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            gc.calc_field_krige(mat, vecs, cond)


def test_variogram_surface_and_validation():
    # src/lib.rs:119-216: names, positional order and Option<> defaults of the four bindings
    import inspect
    assert list(inspect.signature(gc.variogram_structured).parameters) == ["f", "estimator_type", "num_threads"]
    assert list(inspect.signature(gc.variogram_ma_structured).parameters) == [
        "f", "mask", "estimator_type", "num_threads"]
    assert list(inspect.signature(gc.variogram_directional).parameters) == [
        "f", "bin_edges", "pos", "direction", "angles_tol", "bandwidth", "separate_dirs", "estimator_type",
        "num_threads"]
    assert list(inspect.signature(gc.variogram_unstructured).parameters) == [
        "f", "bin_edges", "pos", "estimator_type", "distance_type", "num_threads"]
    f = np.ones((1, 5)); pos = np.ones((2, 5)); edges = np.linspace(0.0, 1.0, 4)
    with pytest.raises(TypeError):
        gc.variogram_structured(np.ones((3, 3), dtype=np.float32))
    with pytest.raises(TypeError):
        gc.variogram_ma_structured(np.ones((3, 3)), np.zeros((3, 3)))          # mask must be bool
    with pytest.raises(ValueError):
        gc.variogram_ma_structured(np.ones((3, 3)), np.zeros((3, 2), dtype=bool))
    with pytest.raises(ValueError):
        gc.variogram_structured(np.ones((3, 3)), "mm")                         # Option<char>
    with pytest.raises(ValueError):
        gc.variogram_unstructured(f, edges, np.ones((2, 4)))                   # src/variogram.rs:473
    with pytest.raises(ValueError):
        gc.variogram_unstructured(f, np.array([1.0]), pos)                     # :480
    with pytest.raises(ValueError):
        gc.variogram_directional(f, edges, pos, np.ones((1, 3)))               # :326
    with pytest.raises(ValueError):
        gc.variogram_directional(f, edges, pos, np.ones((1, 2)), angles_tol=0.0)   # :345
    # an empty field never reaches the device: [0.0] (src/variogram.rs:144-146)
    assert gc.variogram_structured(np.ones((0, 3))).tolist() == [0.0]


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-device failure mode")
def test_variograms_fail_loudly_without_device():
    f = np.ones((1, 5)); pos = np.ones((2, 5)); edges = np.linspace(0.0, 1.0, 4)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        gc.variogram_structured(np.ones((4, 3)))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        gc.variogram_unstructured(f, edges, pos)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        gc.variogram_directional(f, edges, pos, np.array([[1.0, 0.0]]))


def test_native_binding_fast_lane_and_its_fallbacks():
    """csrc/gsf_pybind.c: the native CPython binding handles plain float64 ndarrays with small results
    and declines (None) everything else, so that the ctypes path raises the reference-shaped errors."""
    gc._load()
    nat = gc._native
    assert nat is not None, "native binding not built (make -C gstools-core_b200/csrc)"
    k, z1, z2, pos = np.random.rand(2, 10), np.random.rand(10), np.random.rand(10), np.random.rand(2, 50)
    sf = np.random.rand(10)
    # declined: wrong dtype / not an ndarray / shape mismatch / rank mismatch / result beyond the fast lane
    assert nat.summate(k.astype(np.float32), z1, z2, pos, None) is None
    assert nat.summate([[1.0]], z1, z2, pos, None) is None
    assert nat.summate(k, z1[:5], z2, pos, None) is None
    assert nat.summate(k, z1, z2, pos[0], None) is None
    assert nat.summate(k, z1, z2, np.random.rand(2, 40000), None) is None
    assert nat.summate_fourier(sf[:3], k, z1, z2, pos, None) is None
    assert nat.summate(k, z1, z2, np.zeros((2, 0)), None) is None
    with pytest.raises(OverflowError):
        nat.summate(k, z1, z2, pos, -1)
    with pytest.raises(TypeError):
        nat.summate(k, z1, z2)
    # handled: strided views are fine; without a device the C ABI's status comes back as an int
    r = nat.summate(k[:, ::2], z1[::2], z2[::2], np.asfortranarray(pos), 4)
    if gc.device_count() == 0:
        assert r == 4                                     # GSF_ERR_NO_DEVICE
        for fn, a in ((gc.summate, (k, z1, z2, pos)), (gc.summate_incompr, (k, z1, z2, pos)),
                      (gc.summate_fourier, (sf, k, z1, z2, pos))):
            with pytest.raises(RuntimeError, match="no CUDA device"):
                fn(*a)
    else:
        assert isinstance(r, np.ndarray) and r.shape == (50,)
    # the reference-shaped errors still come from the general path
    with pytest.raises(TypeError):
        gc.summate(k.astype(np.float32), z1, z2, pos)
    with pytest.raises(ValueError):
        gc.summate(k, z1[:5], z2, pos)


def test_native_binding_marshalling_against_recording_callbacks():
    """The native binding with the three C entry points replaced by ctypes callbacks that record what they
    were given: dimensions, ELEMENT strides of arbitrary views (incl. negative), the (d, M) Fortran-ordered
    result of summate_incompr, num_threads, and status propagation -- no GPU involved."""
    from gstools_core import _gsf_native as nat
    i64, vp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
    D = ctypes.POINTER(ctypes.c_double)
    seen = {}

    def read(ptr, n, stride):
        return [ptr[i * stride] for i in range(n)]

    @ctypes.CFUNCTYPE(ci, ci, i64, i64, D, i64, i64, D, i64, D, i64, D, i64, i64, D, ci)
    def fake_summate(d, n, m, k, ks0, ks1, z1, z1s, z2, z2s, pos, ps0, ps1, out, nt):
        seen.update(kind=0, d=d, n=n, m=m, ks=(ks0, ks1), zs=(z1s, z2s), ps=(ps0, ps1), nt=nt,
                    k00=k[0], k_row1=read(k, n, ks1)[:3] if d > 0 else None, z1=read(z1, n, z1s), pos_last=pos[(d - 1) * ps0 + (m - 1) * ps1])
        for j in range(m):
            out[j] = 100.0 + j
        return seen.get("rc", 0)

    @ctypes.CFUNCTYPE(ci, ci, i64, i64, D, i64, i64, D, i64, D, i64, D, i64, i64, D, i64, i64, ci)
    def fake_incompr(d, n, m, k, ks0, ks1, z1, z1s, z2, z2s, pos, ps0, ps1, out, os0, os1, nt):
        seen.update(kind=1, d=d, n=n, m=m, os=(os0, os1), nt=nt)
        for a in range(d):
            for j in range(m):
                out[a * os0 + j * os1] = 10.0 * a + j
        return 0

    @ctypes.CFUNCTYPE(ci, ci, i64, i64, D, i64, D, i64, i64, D, i64, D, i64, D, i64, i64, D, ci)
    def fake_fourier(d, n, m, sf, sfs, k, ks0, ks1, z1, z1s, z2, z2s, pos, ps0, ps1, out, nt):
        seen.update(kind=2, d=d, n=n, m=m, sf=read(sf, n, sfs), sfs=sfs)
        for j in range(m):
            out[j] = -1.0 * j
        return 0

    addr = lambda f: ctypes.cast(f, vp).value                      # noqa: E731
    try:
        nat.bind(addr(fake_summate), addr(fake_incompr), addr(fake_fourier), np.ndarray, np.empty, np.float64, 256 * 1024)
        rng = np.random.default_rng(3)
        kbig, zbig, pbig = rng.normal(size=(3, 14)), rng.normal(size=30), rng.normal(size=(3, 40))
        k, z1, z2 = kbig[:, ::2], zbig[0:21:3], zbig[20::-3]        # views: strides 2, 3, -3 (7 modes)
        pos = pbig[::-1, ::4]                                        # rows reversed, every 4th point (10 points)
        out = nat.summate(k, z1, z2, pos, 5)
        assert isinstance(out, np.ndarray) and out.shape == (10,) and np.array_equal(out, 100.0 + np.arange(10))
        assert (seen["d"], seen["n"], seen["m"], seen["nt"]) == (3, 7, 10, 5)
        assert seen["ks"] == (14, 2) and seen["zs"] == (3, -3) and seen["ps"] == (-40, 4)
        assert seen["k00"] == k[0, 0] and seen["z1"] == list(z1) and seen["pos_last"] == pos[2, 9]
        outi = nat.summate_incompr(k, z1, z2, pos, None)
        assert outi.shape == (3, 10) and outi.flags.f_contiguous and seen["os"] == (1, 3) and seen["nt"] == 0
        assert np.array_equal(outi, 10.0 * np.arange(3)[:, None] + np.arange(10)[None, :])     # src/field.rs:166-174 layout
        sf = np.linspace(1, 2, 14)[::2]
        outf = nat.summate_fourier(sf, k, z1, z2, pos, None)
        assert seen["kind"] == 2 and seen["sf"] == list(sf) and seen["sfs"] == 2 and np.array_equal(outf, -1.0 * np.arange(10))
        seen["rc"] = 3                                               # a non-zero status comes back as an int for the wrapper to raise
        assert nat.summate(k, z1, z2, pos, None) == 3
    finally:
        seen.pop("rc", None)
        gc._native = None
        gc._bind_native(gc._load())                                  # restore the real entry points
    assert gc._native is not None
