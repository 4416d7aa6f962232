"""gstools_core -- B200-native drop-in for the field-summation part of GSTools-Core's pyo3 module.

Mirrors the three Python functions GSTools imports from the reference's `gstools_core`
(/root/reference/src/lib.rs:33-84): same names, same positional order, same return shapes

    summate(cov_samples, z1, z2, pos, num_threads=None)                 -> ndarray (M,)
    summate_incompr(cov_samples, z1, z2, pos, num_threads=None)         -> ndarray (d, M), F-ordered
    summate_fourier(spectrum_factor, modes, z1, z2, pos, num_threads=None) -> ndarray (M,)

but the work is done by hand-written sm_100a CUDA kernels behind the C ABI in include/gsfield.h
(libgsfield.so, built in-tree by csrc/Makefile).  There is no CPU fallback: if the library is
missing or no CUDA device is usable the call raises.

Differences from the reference, all deliberate:
  * a shape / dim mismatch raises ValueError instead of aborting the process (the reference
    panics with panic="abort", Cargo.toml:22);
  * `pos` (and the mode arrays) may also be objects exposing ``__cuda_array_interface__``
    (e.g. torch CUDA tensors); the result is then returned in device memory as the same kind of
    object via the ``out=`` argument of the ``*_device`` helpers.
"""
from __future__ import annotations

import collections
import ctypes
import os
import weakref

import numpy as np

__version__ = "1.1.0+b200.0"   # reference crate version (Cargo.toml:3) + local build tag

_HERE = os.path.dirname(os.path.abspath(__file__))
# GSF_LIB: load another build of the same ABI (A/B measurements, e.g. csrc `make variants`)
_LIB_PATH = os.environ.get("GSF_LIB") or os.path.join(_HERE, "libgsfield.so")

_i64 = ctypes.c_int64
_vp = ctypes.c_void_p
_int = ctypes.c_int


class GsfStats(ctypes.Structure):
    _fields_ = [
        ("total_ms", ctypes.c_double), ("kernel_ms", ctypes.c_double), ("prep_ms", ctypes.c_double),
        ("point_modes", _i64), ("h2d_bytes", _i64), ("d2h_bytes", _i64),
        ("kernel_launches", ctypes.c_int32), ("n_devices", ctypes.c_int32), ("n_chunks", ctypes.c_int32),
        ("points_per_thread", ctypes.c_int32), ("lanes_per_point", ctypes.c_int32),
        ("pos_memory", ctypes.c_int32), ("out_memory", ctypes.c_int32), ("grid_path", ctypes.c_int32),
        ("poly_degree", ctypes.c_int32), ("fp64_slots", ctypes.c_int32), ("staging_threads", ctypes.c_int32),
        ("mode_group", ctypes.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class GsfRequest(ctypes.Structure):
    """mirror of gsf_request (include/gsfield.h)"""
    _fields_ = [
        ("struct_size", ctypes.c_int32), ("kind", ctypes.c_int32), ("dim", ctypes.c_int32),
        ("num_threads", ctypes.c_int32),
        ("n_modes", _i64), ("n_points", _i64),
        ("spectrum_factor", _vp), ("sf_s", _i64),
        ("modes", _vp), ("modes_s0", _i64), ("modes_s1", _i64),
        ("z1", _vp), ("z1_s", _i64),
        ("z2", _vp), ("z2_s", _i64),
        ("pos", _vp), ("pos_s0", _i64), ("pos_s1", _i64),
        ("out", _vp), ("out_s0", _i64), ("out_s1", _i64),
        ("scale", ctypes.c_double), ("offset", ctypes.c_double * 3),
        ("n_axes", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("axis", _vp * 3), ("axis_n", _i64 * 3), ("axis_s", _i64 * 3),
    ]


_PINNED_MIN = 256 * 1024          # below this a staging copy is cheaper than a pool round trip
_PINNED_MAX = 1 << 30

_lib = None
_native = None   # csrc/gsf_pybind.c: native fast lane for small calls (bound in _load)


def _bind_native(L):
    """Native CPython binding (the counterpart of the reference's pyo3 layer): argument conversion
    in C for small calls, where ctypes marshalling (~20 us) would dominate.  Optional -- every call
    it declines (returns None) goes through the ctypes path below.  GSF_NATIVE_BINDING=0 disables it."""
    global _native
    if os.environ.get("GSF_NATIVE_BINDING", "1") == "0":
        return
    try:
        from . import _gsf_native as nat
    except ImportError:
        return
    addr = [ctypes.cast(getattr(L, n), _vp).value for n in ("gsf_summate", "gsf_summate_incompr", "gsf_summate_fourier")]
    nat.bind(addr[0], addr[1], addr[2], np.ndarray, np.empty, np.float64, _PINNED_MIN)
    _native = nat


def _load():
    """Load libgsfield.so; fail loudly (no fallback) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            "gstools_core (B200): %s not found -- build it with `make -C gstools-core_b200/csrc` "
            "or `python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % _LIB_PATH)
    L = ctypes.CDLL(_LIB_PATH)
    a2 = [_vp, _i64, _i64]
    a1 = [_vp, _i64]
    head = [_int, _i64, _i64]
    L.gsf_summate.argtypes = head + a2 + a1 + a1 + a2 + [_vp, _int]
    L.gsf_summate_incompr.argtypes = head + a2 + a1 + a1 + a2 + [_vp, _i64, _i64, _int]
    L.gsf_summate_fourier.argtypes = head + a1 + a2 + a1 + a1 + a2 + [_vp, _int]
    L.gsf_krige.argtypes = [_i64, _i64] + a2 + a2 + a1 + [_vp, _vp, _int]
    L.gsf_krige.restype = _int
    L.gsf_summate_ex.argtypes = [ctypes.POINTER(GsfRequest)]
    L.gsf_set_grid_detection.argtypes = [_int]
    L.gsf_summate_on_stream.argtypes = [_int] + head + a1 + a2 + a1 + a1 + a2 + [_vp, _i64, _i64, _vp]
    L.gsf_set_devices.argtypes = [ctypes.POINTER(_int), _int]
    L.gsf_shard_bounds.argtypes = [_i64, _int, _int, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]
    L.gsf_host_alloc.argtypes = [_i64, ctypes.POINTER(_vp)]
    L.gsf_host_free.argtypes = [_vp]
    L.gsf_set_chunk_points.argtypes = [_i64]
    L.gsf_set_variant.argtypes = [_int, _int]
    L.gsf_set_profiling.argtypes = [_int]
    L.gsf_set_poly_degree.argtypes = [_int]
    L.gsf_host_register.argtypes = [_vp, _i64]
    L.gsf_host_unregister.argtypes = [_vp]
    L.gsf_get_last_stats.argtypes = [ctypes.POINTER(GsfStats)]
    L.gsf_dfma_peak.argtypes = [_int, ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                                ctypes.POINTER(ctypes.c_double)]
    L.gsf_dmma_peak.argtypes = L.gsf_dfma_peak.argtypes
    L.gsf_dmma_peak.restype = _int
    L.gsf_last_error.restype = ctypes.c_char_p
    _a2, _a1 = [_vp, _i64, _i64], [_vp, _i64]
    L.gsf_variogram_structured.argtypes = [_i64, _i64] + _a2 + _a2 + [ctypes.c_char, _vp, _int]
    L.gsf_variogram_unstructured.argtypes = ([_int, _i64, _i64, _i64] + _a2 + _a1 + _a2
                                             + [ctypes.c_char, ctypes.c_char, _vp, _vp, _int])
    L.gsf_variogram_directional.argtypes = ([_int, _i64, _i64, _i64, _i64] + _a2 + _a1 + _a2 + _a2
                                            + [ctypes.c_double, ctypes.c_double, _int, ctypes.c_char, _vp, _vp, _int])
    L.gsf_debug_variogram_thresholds.argtypes = [ctypes.c_double, ctypes.c_double,
                                                 ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    for name in ("gsf_variogram_structured", "gsf_variogram_unstructured", "gsf_variogram_directional",
                 "gsf_debug_variogram_thresholds"):
        getattr(L, name).restype = _int
    for name in ("gsf_summate", "gsf_summate_incompr", "gsf_summate_fourier", "gsf_summate_on_stream",
                 "gsf_summate_ex", "gsf_set_grid_detection",
                 "gsf_set_devices", "gsf_shard_bounds", "gsf_host_alloc", "gsf_host_free", "gsf_set_chunk_points", "gsf_set_variant", "gsf_set_profiling",
                 "gsf_set_poly_degree", "gsf_host_register", "gsf_host_unregister",
                 "gsf_get_last_stats", "gsf_dfma_peak", "gsf_abi_version", "gsf_device_count",
                 "gsf_shutdown"):
        getattr(L, name).restype = _int
    _lib = L
    _bind_native(L)
    return L


_ERR = {1: ValueError, 2: ValueError, 3: ValueError, 4: RuntimeError, 5: RuntimeError, 6: ValueError,
        7: MemoryError}


def _raise(rc):
    msg = _load().gsf_last_error().decode("utf-8", "replace")
    raise _ERR.get(rc, RuntimeError)("gstools_core (B200): %s [status %d]" % (msg, rc))


# ---------------------------------------------------------------------------------------------
# argument marshalling

class _Arr:
    """pointer + shape + element strides of a float64 host (numpy) or device (CUDA array interface) array"""
    __slots__ = ("ptr", "shape", "strides", "keep", "device")

    def __init__(self, obj, ndim, name):
        cai = getattr(obj, "__cuda_array_interface__", None)
        if cai is not None and not isinstance(obj, np.ndarray):
            if cai["typestr"] not in ("<f8", "=f8", "f8"):
                raise TypeError("argument '%s': device array must be float64, got %s" % (name, cai["typestr"]))
            shape = tuple(cai["shape"])
            st = cai.get("strides")
            if st is None:
                st, acc = [], 8
                for s in reversed(shape):
                    st.append(acc)
                    acc *= s
                st = tuple(reversed(st))
            self.ptr = cai["data"][0]
            self.device = True
            self.keep = obj
        else:
            if not isinstance(obj, np.ndarray):
                raise TypeError("argument '%s': 'ndarray' expected, got %s" % (name, type(obj).__name__))
            if obj.dtype != np.float64:
                # PyReadonlyArray<f64> does not cast (src/lib.rs:38-41): wrong dtype is a TypeError
                raise TypeError("argument '%s': type mismatch: expected float64, got %s" % (name, obj.dtype))
            shape, st = obj.shape, obj.strides
            self.ptr = obj.ctypes.data
            self.device = False
            self.keep = obj
        if len(shape) != ndim:
            raise TypeError("argument '%s': dimensionality mismatch: expected %d, got %d" % (name, ndim, len(shape)))
        if any(s % 8 for s in st):
            raise ValueError("argument '%s': strides must be multiples of 8 bytes" % name)
        self.shape = shape
        self.strides = tuple(s // 8 for s in st)

    def args(self):
        return (self.ptr,) + self.strides




def _result_array(shape, order="C"):
    """Freshly allocated float64 result, like the reference's owned Array1/Array2 handed to numpy
    (src/lib.rs:47).  Mid-sized results are backed by pinned memory from the library's caching
    pool so the device->host copy lands in them directly; the block returns to the pool when the
    array (and every view of it) is garbage collected.  GSF_PINNED_OUTPUT=0 disables this."""
    n = 1
    for s in shape:
        n *= int(s)
    nbytes = n * 8
    if not (_PINNED_MIN <= nbytes <= _PINNED_MAX) or os.environ.get("GSF_PINNED_OUTPUT", "1") == "0":
        return np.empty(shape, dtype=np.float64, order=order)
    L = _load()
    ptr = _vp()
    if L.gsf_host_alloc(nbytes, ctypes.byref(ptr)) != 0 or not ptr.value:
        return np.empty(shape, dtype=np.float64, order=order)
    buf = (ctypes.c_char * nbytes).from_address(ptr.value)
    weakref.finalize(buf, L.gsf_host_free, ptr.value)
    return np.frombuffer(buf, dtype=np.float64).reshape(shape, order=order)


def _threads(num_threads):
    if num_threads is None:
        return 0
    n = int(num_threads)
    if n < 0:
        raise OverflowError("can't convert negative int to unsigned")   # Option<usize>, src/lib.rs:41
    return n


def _check_shapes(cov, z1, z2, pos):
    # the reference's assert_eq!s, src/field.rs:44-46 / 104-106 / 227-229
    if cov.shape[0] != pos.shape[0]:
        raise ValueError("dim mismatch: cov_samples has %d rows, pos has %d" % (cov.shape[0], pos.shape[0]))
    if cov.shape[1] != z1.shape[0] or cov.shape[1] != z2.shape[0]:
        raise ValueError("mode count mismatch: cov_samples has %d modes, z1 %d, z2 %d"
                         % (cov.shape[1], z1.shape[0], z2.shape[0]))


def summate(cov_samples, z1, z2, pos, num_threads=None):
    """Scalar randomization method (reference: summate_py, src/lib.rs:33-48 -> field::summator)."""
    L = _load()
    if _auto_pin_cap:
        _auto_pin(pos)
    if _native is not None:
        r = _native.summate(cov_samples, z1, z2, pos, num_threads)
        if r.__class__ is np.ndarray:
            return r
        if r is not None:
            _raise(r)
    cov, a1, a2, p = _Arr(cov_samples, 2, "cov_samples"), _Arr(z1, 1, "z1"), _Arr(z2, 1, "z2"), _Arr(pos, 2, "pos")
    _check_shapes(cov, a1, a2, p)
    if p.device:
        raise TypeError("summate: pos is a device array; use summate_device(..., out=...)")
    d, n = cov.shape
    m = p.shape[1]
    out = _result_array((m,))
    rc = L.gsf_summate(d, n, m, *cov.args(), *a1.args(), *a2.args(), *p.args(), out.ctypes.data,
                       _threads(num_threads))
    if rc:
        _raise(rc)
    return out


def summate_incompr(cov_samples, z1, z2, pos, num_threads=None):
    """Incompressible vector field (reference: summate_incompr_py, src/lib.rs:50-65).

    Returns shape (d, M) in Fortran order, like the reference (src/field.rs:166-174)."""
    L = _load()
    if _auto_pin_cap:
        _auto_pin(pos)
    if _native is not None:
        r = _native.summate_incompr(cov_samples, z1, z2, pos, num_threads)
        if r.__class__ is np.ndarray:
            return r
        if r is not None:
            _raise(r)
    cov, a1, a2, p = _Arr(cov_samples, 2, "cov_samples"), _Arr(z1, 1, "z1"), _Arr(z2, 1, "z2"), _Arr(pos, 2, "pos")
    _check_shapes(cov, a1, a2, p)
    if p.device:
        raise TypeError("summate_incompr: pos is a device array; use summate_incompr_device(..., out=...)")
    d, n = cov.shape
    m = p.shape[1]
    out = _result_array((d, m), order="F")
    rc = L.gsf_summate_incompr(d, n, m, *cov.args(), *a1.args(), *a2.args(), *p.args(), out.ctypes.data,
                               out.strides[0] // 8, out.strides[1] // 8, _threads(num_threads))
    if rc:
        _raise(rc)
    return out


def summate_fourier(spectrum_factor, modes, z1, z2, pos, num_threads=None):
    """Periodic Fourier method (reference: summate_fourier_py, src/lib.rs:67-84)."""
    L = _load()
    if _auto_pin_cap:
        _auto_pin(pos)
    if _native is not None:
        r = _native.summate_fourier(spectrum_factor, modes, z1, z2, pos, num_threads)
        if r.__class__ is np.ndarray:
            return r
        if r is not None:
            _raise(r)
    sf = _Arr(spectrum_factor, 1, "spectrum_factor")
    cov, a1, a2, p = _Arr(modes, 2, "modes"), _Arr(z1, 1, "z1"), _Arr(z2, 1, "z2"), _Arr(pos, 2, "pos")
    _check_shapes(cov, a1, a2, p)
    if sf.shape[0] != cov.shape[1]:
        # the reference panics inside ndarray's Zip::and (part dimension mismatch)
        raise ValueError("spectrum_factor has %d entries, modes has %d" % (sf.shape[0], cov.shape[1]))
    if p.device:
        raise TypeError("summate_fourier: pos is a device array; use summate_fourier_device(..., out=...)")
    d, n = cov.shape
    m = p.shape[1]
    out = _result_array((m,))
    rc = L.gsf_summate_fourier(d, n, m, *sf.args(), *cov.args(), *a1.args(), *a2.args(), *p.args(),
                               out.ctypes.data, _threads(num_threads))
    if rc:
        _raise(rc)
    return out


# ---------------------------------------------------------------------------------------------
# extended calls: fused post-scale/offset and structured grids (8 f3)

_KINDS = {"summate": 0, "summate_incompr": 1, "summate_fourier": 2}


def _extended(kind, sf, cov_samples, z1, z2, pos, axes, scale, offset, num_threads, out=None):
    L = _load()
    cov, a1, a2 = _Arr(cov_samples, 2, "cov_samples"), _Arr(z1, 1, "z1"), _Arr(z2, 1, "z2")
    d, n = cov.shape
    if cov.shape[1] != a1.shape[0] or cov.shape[1] != a2.shape[0]:
        raise ValueError("mode count mismatch")
    r = GsfRequest()
    r.struct_size = ctypes.sizeof(GsfRequest)
    r.kind, r.dim, r.num_threads, r.n_modes = kind, d, _threads(num_threads), n
    keep = [cov, a1, a2]
    r.modes, r.modes_s0, r.modes_s1 = cov.ptr, cov.strides[0], cov.strides[1]
    r.z1, r.z1_s, r.z2, r.z2_s = a1.ptr, a1.strides[0], a2.ptr, a2.strides[0]
    if kind == 2:
        s = _Arr(sf, 1, "spectrum_factor")
        if s.shape[0] != n:
            raise ValueError("spectrum_factor has %d entries, modes has %d" % (s.shape[0], n))
        r.spectrum_factor, r.sf_s = s.ptr, s.strides[0]
        keep.append(s)
    if axes is not None:
        axes = [np.ascontiguousarray(a, dtype=np.float64) for a in axes]
        if len(axes) != d:
            raise ValueError("need %d axis vectors, got %d" % (d, len(axes)))
        if any(a.ndim != 1 for a in axes):
            raise ValueError("axis vectors must be one-dimensional")
        m = 1
        for i, a in enumerate(axes):
            r.axis[i], r.axis_n[i], r.axis_s[i] = a.ctypes.data, a.shape[0], 1
            m *= a.shape[0]
        r.n_axes = d
        keep.append(axes)
    else:
        p = _Arr(pos, 2, "pos")
        if p.shape[0] != d:
            raise ValueError("dim mismatch: cov_samples has %d rows, pos has %d" % (d, p.shape[0]))
        if p.device:
            raise TypeError("pos is a device array; use the *_device functions")
        m = p.shape[1]
        r.pos, r.pos_s0, r.pos_s1 = p.ptr, p.strides[0], p.strides[1]
        keep.append(p)
    r.n_points = m
    if out is not None:
        # caller-provided result (host ndarray or device array), e.g. to keep the field on the GPU
        o = _Arr(out, 2 if kind == 1 else 1, "out")
        if o.shape != ((d, m) if kind == 1 else (m,)):
            raise ValueError("out has shape %s, expected %s" % (o.shape, (d, m) if kind == 1 else (m,)))
        if kind == 1:
            r.out_s0, r.out_s1 = o.strides
        else:
            r.out_s0, r.out_s1 = 0, o.strides[0] if m > 1 else 1
        r.out = o.ptr
        keep.append(o)
    elif kind == 1:
        out = _result_array((d, m), order="F")
        r.out_s0, r.out_s1 = out.strides[0] // 8, out.strides[1] // 8
        r.out = out.ctypes.data
    else:
        out = _result_array((m,))
        r.out_s0, r.out_s1 = 0, 1
        r.out = out.ctypes.data
    r.scale = float(scale)
    off = np.zeros(3)
    off[:np.size(offset)] = np.ravel(offset)[:3]
    for i in range(3):
        r.offset[i] = float(off[i])
    rc = L.gsf_summate_ex(ctypes.byref(r))
    if rc:
        _raise(rc)
    return out


def summate_scaled(cov_samples, z1, z2, pos, scale=1.0, offset=0.0, num_threads=None):
    """scale * summate(...) + offset in one pass (GSTools: sqrt(var/N) * summed_modes + mean)."""
    return _extended(0, None, cov_samples, z1, z2, pos, None, scale, offset, num_threads)


def summate_incompr_scaled(cov_samples, z1, z2, pos, scale=1.0, offset=(0.0, 0.0, 0.0), num_threads=None):
    """scale * summate_incompr(...) + offset[component] in one pass."""
    return _extended(1, None, cov_samples, z1, z2, pos, None, scale, offset, num_threads)


def summate_fourier_scaled(spectrum_factor, modes, z1, z2, pos, scale=1.0, offset=0.0, num_threads=None):
    return _extended(2, spectrum_factor, modes, z1, z2, pos, None, scale, offset, num_threads)


def summate_grid(cov_samples, z1, z2, axes, scale=1.0, offset=0.0, num_threads=None, out=None):
    """summate on the rectilinear grid axes[0] x axes[1] (x axes[2]) (GSTools mesh_type="structured"):
    same values as summate(..., pos=expanded grid flattened in C order), shape (prod(n_a),)."""
    return _extended(0, None, cov_samples, z1, z2, None, axes, scale, offset, num_threads, out)


def summate_incompr_grid(cov_samples, z1, z2, axes, scale=1.0, offset=(0.0, 0.0, 0.0), num_threads=None, out=None):
    return _extended(1, None, cov_samples, z1, z2, None, axes, scale, offset, num_threads, out)


def summate_fourier_grid(spectrum_factor, modes, z1, z2, axes, scale=1.0, offset=0.0, num_threads=None, out=None):
    return _extended(2, spectrum_factor, modes, z1, z2, None, axes, scale, offset, num_threads, out)


def set_grid_detection(enabled=True):
    """Automatic structured-grid detection in summate*/(host pos): True/False, None = GSF_GRID_DETECT."""
    _load().gsf_set_grid_detection(-1 if enabled is None else (1 if enabled else 0))


# ---------------------------------------------------------------------------------------------
# kriging: mirrors calc_field_krige / calc_field_krige_and_variance, src/lib.rs:86-118

def _krige(krige_mat, krig_vecs, cond, want_error, num_threads):
    L = _load()
    mat, vecs, cnd = _Arr(krige_mat, 2, "krige_mat"), _Arr(krig_vecs, 2, "krig_vecs"), _Arr(cond, 1, "cond")
    c = mat.shape[0]
    if mat.shape[1] != c or vecs.shape[0] != c or cnd.shape[0] != c:
        raise ValueError("shape mismatch: krige_mat %s, krig_vecs %s, cond %s (src/krige.rs:30-32)"
                         % (mat.shape, vecs.shape, cnd.shape))
    m = vecs.shape[1]
    field = np.empty(m, dtype=np.float64)
    error = np.empty(m, dtype=np.float64) if want_error else None
    rc = L.gsf_krige(c, m, *mat.args(), *vecs.args(), *cnd.args(), field.ctypes.data,
                     error.ctypes.data if want_error else None, _threads(num_threads))
    if rc:
        _raise(rc)
    return (field, error) if want_error else field


def calc_field_krige(krige_mat, krig_vecs, cond, num_threads=None):
    """Kriging field (reference: calc_field_krige_py, src/lib.rs:104-118)."""
    return _krige(krige_mat, krig_vecs, cond, False, num_threads)


def calc_field_krige_and_variance(krige_mat, krig_vecs, cond, num_threads=None):
    """Kriging field and error variance (reference: calc_field_krige_and_variance_py, src/lib.rs:86-102)."""
    return _krige(krige_mat, krig_vecs, cond, True, num_threads)


# ---------------------------------------------------------------------------------------------
# empirical variograms (reference: src/variogram.rs, bindings src/lib.rs:119-216)

def _char(value, default, name):
    # Option<char>: None -> default; pyo3 wants a str of length 1
    if value is None:
        return default.encode()
    if not isinstance(value, str):
        raise TypeError("argument '%s': 'str' expected, got %s" % (name, type(value).__name__))
    if len(value) != 1:
        raise ValueError("argument '%s': expected a string of length 1" % name)
    return value.encode("utf-8")[:1]


def _host(arr, name):
    if arr.device:
        raise TypeError("argument '%s': the variogram estimators take host (numpy) arrays" % name)
    return arr


def _variogram_structured(f, mask, estimator_type, num_threads):
    L = _load()
    fa = _host(_Arr(f, 2, "f"), "f")
    est = _char(estimator_type, "m", "estimator_type")
    n0, n1 = fa.shape
    margs = (None, 0, 0)
    if mask is not None:
        if not isinstance(mask, np.ndarray):
            raise TypeError("argument 'mask': 'ndarray' expected, got %s" % type(mask).__name__)
        if mask.dtype != np.bool_:
            raise TypeError("argument 'mask': type mismatch: expected bool, got %s" % mask.dtype)
        if mask.ndim != 2:
            raise TypeError("argument 'mask': dimensionality mismatch: expected 2, got %d" % mask.ndim)
        if mask.shape != (n0, n1):
            # the reference panics inside ndarray's Zip::and (src/variogram.rs:218-221)
            raise ValueError("mask has shape %s, f has %s" % (mask.shape, (n0, n1)))
        margs = (mask.ctypes.data, mask.strides[0], mask.strides[1])
    out = np.empty(max(n0, 1), dtype=np.float64)
    out[0] = 0.0   # size 0 still yields [0.0] (src/variogram.rs:144-146)
    if n0 > 0:
        rc = L.gsf_variogram_structured(n0, n1, *fa.args(), *margs, est, out.ctypes.data, _threads(num_threads))
        if rc:
            _raise(rc)
    return out


def variogram_structured(f, estimator_type=None, num_threads=None):
    """Variogram along axis 0 of a structured field (reference: variogram_structured_py,
    src/lib.rs:119-131 -> src/variogram.rs:136-178).  estimator_type 'm' (default) or 'c'."""
    return _variogram_structured(f, None, estimator_type, num_threads)


def variogram_ma_structured(f, mask, estimator_type=None, num_threads=None):
    """Masked structured variogram (reference: variogram_ma_structured_py, src/lib.rs:133-147 ->
    src/variogram.rs:190-240).  mask: bool array of f's shape, True = excluded."""
    if mask is None:
        raise TypeError("argument 'mask': 'ndarray' expected, got NoneType")
    return _variogram_structured(f, mask, estimator_type, num_threads)


def variogram_unstructured(f, bin_edges, pos, estimator_type=None, distance_type=None, num_threads=None):
    """Isotropic variogram of scattered data (reference: variogram_unstructured_py,
    src/lib.rs:188-216 -> src/variogram.rs:465-545).  Returns (variogram (n_bins,) float64,
    counts (n_bins,) uint64).  distance_type 'e' (default) Euclid, else Haversine (pos in degrees)."""
    L = _load()
    fa, e, p = _host(_Arr(f, 2, "f"), "f"), _host(_Arr(bin_edges, 1, "bin_edges"), "bin_edges"), _host(_Arr(pos, 2, "pos"), "pos")
    est, dist = _char(estimator_type, "m", "estimator_type"), _char(distance_type, "e", "distance_type")
    if p.shape[1] != fa.shape[1]:
        raise ValueError("len(pos) = %d != len(f) = %d" % (p.shape[1], fa.shape[1]))       # src/variogram.rs:473-479
    if e.shape[0] < 2:
        raise ValueError("len(bin_edges) = %d < 2 too small" % e.shape[0])                 # :480-484
    nb = e.shape[0] - 1
    v, c = np.empty(nb, dtype=np.float64), np.empty(nb, dtype=np.uint64)
    rc = L.gsf_variogram_unstructured(p.shape[0], fa.shape[0], fa.shape[1], nb, *fa.args(), *e.args(), *p.args(),
                                      est, dist, v.ctypes.data, c.ctypes.data, _threads(num_threads))
    if rc:
        _raise(rc)
    return v, c


def variogram_directional(f, bin_edges, pos, direction, angles_tol=None, bandwidth=None, separate_dirs=None,
                          estimator_type=None, num_threads=None):
    """Directional variogram of scattered data (reference: variogram_directional_py,
    src/lib.rs:149-186 -> src/variogram.rs:315-447).  direction: (n_dirs, dim), normed.  Defaults
    as in the binding: angles_tol pi/8, bandwidth -1 (off), separate_dirs False.  Returns
    (variogram, counts), both (n_dirs, n_bins)."""
    L = _load()
    fa, e, p = _host(_Arr(f, 2, "f"), "f"), _host(_Arr(bin_edges, 1, "bin_edges"), "bin_edges"), _host(_Arr(pos, 2, "pos"), "pos")
    dr = _host(_Arr(direction, 2, "direction"), "direction")
    est = _char(estimator_type, "m", "estimator_type")
    tol = float(np.pi / 8.0 if angles_tol is None else angles_tol)
    bw = float(-1.0 if bandwidth is None else bandwidth)
    if p.shape[0] != dr.shape[1]:
        raise ValueError("dim(pos) = %d != dim(direction) = %d" % (p.shape[0], dr.shape[1]))   # src/variogram.rs:326-332
    if p.shape[1] != fa.shape[1]:
        raise ValueError("len(pos) = %d != len(f) = %d" % (p.shape[1], fa.shape[1]))           # :333-339
    if e.shape[0] < 2:
        raise ValueError("len(bin_edges) = %d < 2 too small" % e.shape[0])                     # :340-344
    if not tol > 0.0:
        raise ValueError("tolerance for angle search masks must be > 0")                       # :345-348
    nb, nd = e.shape[0] - 1, dr.shape[0]
    v, c = np.empty((nd, nb), dtype=np.float64), np.empty((nd, nb), dtype=np.uint64)
    rc = L.gsf_variogram_directional(p.shape[0], fa.shape[0], fa.shape[1], nb, nd, *fa.args(), *e.args(), *p.args(),
                                     *dr.args(), tol, bw, int(bool(separate_dirs)), est, v.ctypes.data,
                                     c.ctypes.data, _threads(num_threads))
    if rc:
        _raise(rc)
    return v, c


# ---------------------------------------------------------------------------------------------
# device-resident, stream-ordered entry points

def _on_stream(kind, sf, cov_samples, z1, z2, pos, out, stream, sync=False):
    L = _load()
    cov, a1, a2, p = _Arr(cov_samples, 2, "cov_samples"), _Arr(z1, 1, "z1"), _Arr(z2, 1, "z2"), _Arr(pos, 2, "pos")
    _check_shapes(cov, a1, a2, p)
    d, n = cov.shape
    m = p.shape[1]
    o = _Arr(out, 2 if kind == 1 else 1, "out")
    if not (p.device and o.device):
        raise TypeError("*_device functions need pos and out in device memory")
    if kind == 1:
        if o.shape != (d, m):
            raise ValueError("out must have shape (%d, %d)" % (d, m))
        ostr = o.strides
    else:
        if o.shape != (m,) or (m > 1 and o.strides[0] != 1):
            raise ValueError("out must be a contiguous array of %d float64" % m)
        ostr = (0, 1)
    if kind == 2:
        s = _Arr(sf, 1, "spectrum_factor")
        if s.shape[0] != n:
            raise ValueError("spectrum_factor has %d entries, modes has %d" % (s.shape[0], n))
        sargs = s.args()
    else:
        sargs = (None, 0)
    if sync:
        # synchronous host-style entry points with device pointers: returns when the result is
        # complete, and shards the points over set_devices([...]) through NVLink peer mappings
        if kind == 0:
            rc = L.gsf_summate(d, n, m, *cov.args(), *a1.args(), *a2.args(), *p.args(), o.ptr, 0)
        elif kind == 1:
            rc = L.gsf_summate_incompr(d, n, m, *cov.args(), *a1.args(), *a2.args(), *p.args(), o.ptr,
                                       ostr[0], ostr[1], 0)
        else:
            rc = L.gsf_summate_fourier(d, n, m, *sargs, *cov.args(), *a1.args(), *a2.args(), *p.args(), o.ptr, 0)
    else:
        rc = L.gsf_summate_on_stream(kind, d, n, m, *sargs, *cov.args(), *a1.args(), *a2.args(), *p.args(),
                                     o.ptr, ostr[0], ostr[1], _vp(int(stream) if stream else 0))
    if rc:
        _raise(rc)
    return out


def summate_device(cov_samples, z1, z2, pos, out, stream=0, sync=False):
    """summate on device-resident pos/out.

    sync=False: enqueued on `stream` (a cudaStream_t handle) of the arrays' device, no host sync.
    sync=True : blocking call; with set_devices([...]) naming several GPUs every device takes a
                contiguous point shard and reads/writes the owner's memory over NVLink (peer
                mapping), no staging copies and no collective."""
    return _on_stream(0, None, cov_samples, z1, z2, pos, out, stream, sync)


def summate_incompr_device(cov_samples, z1, z2, pos, out, stream=0, sync=False):
    return _on_stream(1, None, cov_samples, z1, z2, pos, out, stream, sync)


def summate_fourier_device(spectrum_factor, modes, z1, z2, pos, out, stream=0, sync=False):
    return _on_stream(2, spectrum_factor, modes, z1, z2, pos, out, stream, sync)


# ---------------------------------------------------------------------------------------------
# context / diagnostics

def device_count():
    return _load().gsf_device_count()


def set_devices(device_ids=None):
    """Devices that host-memory calls shard their points over (None = default: GSF_DEVICES or [0])."""
    L = _load()
    ids = list(device_ids or [])
    arr = (_int * max(len(ids), 1))(*ids)
    rc = L.gsf_set_devices(arr, len(ids))
    if rc:
        _raise(rc)


def shard_bounds(n_points, n_shards, shard):
    """Contiguous point range [begin, end) of one shard: the partition used across devices
    inside one call and across ranks (one process per GPU) by bench.py. No collective needed."""
    b, e = _i64(), _i64()
    rc = _load().gsf_shard_bounds(int(n_points), int(n_shards), int(shard), ctypes.byref(b), ctypes.byref(e))
    if rc:
        _raise(rc)
    return b.value, e.value


def chunk_schedule(n_points, forced_chunk=0):
    """Pipeline chunk sizes the library uses for n_points host-resident points (host logic only)."""
    L = _load()
    L.gsf_debug_chunk_schedule.argtypes = [_i64, _i64, ctypes.POINTER(_i64), _int]
    L.gsf_debug_chunk_schedule.restype = _int
    cap = 4096
    buf = (_i64 * cap)()
    n = L.gsf_debug_chunk_schedule(int(n_points), int(forced_chunk), buf, cap)
    if n < 0:
        cap = -n
        buf = (_i64 * cap)()
        n = L.gsf_debug_chunk_schedule(int(n_points), int(forced_chunk), buf, cap)
    return list(buf[:n])


def set_chunk_points(n):
    rc = _load().gsf_set_chunk_points(int(n))
    if rc:
        _raise(rc)


def set_variant(points_per_thread=0, lanes_per_point=0):
    rc = _load().gsf_set_variant(int(points_per_thread), int(lanes_per_point))
    if rc:
        _raise(rc)


def set_poly_degree(degree=0):
    """Degree of the cosine polynomial of the point x mode kernels: 0 automatic (6 below 2^27
    point*modes, 5 above), or force 5 / 6.  See include/gsfield.h."""
    rc = _load().gsf_set_poly_degree(int(degree))
    if rc:
        _raise(rc)


class pinned:
    """Page-lock a caller-owned numpy array for as long as the object lives (context manager):

        with gstools_core.pinned(pos):
            for z1, z2 in ensemble: field = gstools_core.summate(k, z1, z2, pos)

    The GPU then reads `pos` in place instead of staging it through the library's pinned ring."""

    def __init__(self, array):
        if not isinstance(array, np.ndarray) or not (array.flags.c_contiguous or array.flags.f_contiguous):
            raise TypeError("pinned(): a contiguous numpy array is required")
        self._arr = array
        rc = _load().gsf_host_register(array.ctypes.data, array.nbytes)
        if rc:
            _raise(rc)
        self._live = True

    def release(self):
        if self._live:
            self._live = False
            _load().gsf_host_unregister(self._arr.ctypes.data)

    def __enter__(self):
        return self._arr

    def __exit__(self, *exc):
        self.release()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# Opt-in: page-lock position arrays that come back.  GSTools ensembles evaluate many fields at the
# SAME positions; the second time an array (same object memory, same size) is passed it is
# registered with cudaHostRegister (a few ms, once) and from then on the GPU reads it in place --
# no staging copy, a third of the host memory traffic.  The cache holds a reference to every
# registered array, so its memory cannot be freed or reused while it is page-locked; least recently
# used entries are unregistered beyond the byte budget.  Off by default (pinned memory is a scarce,
# unswappable resource): set_auto_pin(max_mb) or GSF_AUTO_PIN_MB=<mb>.
_AUTO_PIN_MIN = 1 << 20
_auto_pin_cap = int(float(os.environ.get("GSF_AUTO_PIN_MB", "0")) * (1 << 20))
_auto_pin_seen = {}                 # (ptr, nbytes) -> sightings of not-yet-registered arrays
_auto_pin_live = collections.OrderedDict()   # (ptr, nbytes) -> pinned handle (LRU order)
_auto_pin_bytes = 0


def set_auto_pin(max_mb=0):
    """Byte budget (MiB) of automatically page-locked position arrays; 0 switches the feature off and
    releases everything it holds."""
    global _auto_pin_cap
    _auto_pin_cap = int(float(max_mb) * (1 << 20))
    _auto_pin_trim()
    if _auto_pin_cap == 0:
        _auto_pin_seen.clear()


def _auto_pin_trim():
    global _auto_pin_bytes
    while _auto_pin_live and _auto_pin_bytes > _auto_pin_cap:
        (_, nbytes), handle = _auto_pin_live.popitem(last=False)
        handle.release()
        _auto_pin_bytes -= nbytes


def _auto_pin(pos):
    """Called with the `pos` argument of a host call when the feature is on."""
    global _auto_pin_bytes
    if not isinstance(pos, np.ndarray) or pos.dtype != np.float64 or not pos.flags.c_contiguous:
        return
    nbytes = pos.nbytes
    if nbytes < _AUTO_PIN_MIN or nbytes > _auto_pin_cap:
        return
    key = (pos.ctypes.data, nbytes)
    if key in _auto_pin_live:
        _auto_pin_live.move_to_end(key)
        return
    n = _auto_pin_seen.get(key, 0) + 1
    if n < 2:
        if len(_auto_pin_seen) > 64:
            _auto_pin_seen.clear()
        _auto_pin_seen[key] = n
        return
    _auto_pin_seen.pop(key, None)
    try:
        handle = pinned(pos)
    except Exception:
        return                      # registration refused (limits): keep staging
    _auto_pin_live[key] = handle
    _auto_pin_bytes += nbytes
    _auto_pin_trim()


def set_profiling(enabled=True):
    """False/0 off, True/1 per call, 2 accumulate kernel_ms over all calls until the next set_profiling."""
    _load().gsf_set_profiling(int(enabled))


def last_stats():
    s = GsfStats()
    rc = _load().gsf_get_last_stats(ctypes.byref(s))
    if rc:
        _raise(rc)
    return s.as_dict()


def dfma_peak(device=0, min_ms=200.0):
    """Measured FP64 DFMA rate of `device` in thread-level DFMA/s (the roofline denominator)."""
    rate, ms = ctypes.c_double(), ctypes.c_double()
    rc = _load().gsf_dfma_peak(int(device), float(min_ms), ctypes.byref(rate), ctypes.byref(ms))
    if rc:
        _raise(rc)
    return rate.value, ms.value


def dmma_peak(device=0, min_ms=200.0):
    """Measured FP64 tensor-path (DMMA) rate in thread-level FMA/s: roofline of the grid kernel."""
    rate, ms = ctypes.c_double(), ctypes.c_double()
    rc = _load().gsf_dmma_peak(int(device), float(min_ms), ctypes.byref(rate), ctypes.byref(ms))
    if rc:
        _raise(rc)
    return rate.value, ms.value


def shutdown():
    cap = _auto_pin_cap
    set_auto_pin(0)                   # unregister automatically page-locked arrays
    set_auto_pin(cap / (1 << 20))
    _load().gsf_shutdown()
