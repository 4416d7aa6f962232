"""Host-side logic of the CUDA library that needs no device: the pipeline chunk schedule and the
exact structured-grid detector (both exported for introspection through the C ABI)."""
import ctypes

import numpy as np
import pytest

import gstools_core as gc

L = gc._load()
L.gsf_debug_chunk_schedule.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
L.gsf_debug_chunk_schedule.restype = ctypes.c_int
L.gsf_debug_detect_grid.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.POINTER(ctypes.c_int64)]
L.gsf_debug_detect_grid.restype = ctypes.c_int


def schedule(m, forced=0):
    buf = (ctypes.c_int64 * 4096)()
    n = L.gsf_debug_chunk_schedule(m, forced, buf, 4096)
    assert n >= 0
    return list(buf[:n])


@pytest.mark.parametrize("m", [1, 1000, 32768, 40000, 70001, 10**5, 3 * 10**5, 10**6, 12_500_000, 10**8])
def test_chunk_schedule_covers_all_points(m):
    s = schedule(m)
    assert sum(s) == m and all(c > 0 for c in s)
    assert max(s) <= (1 << 20) or len(s) == 1
    if m >= 10**6:
        assert s[0] == 32768 and s[-1] == 32768          # small first / last chunk: short exposed copies
        assert s[:3] == [32768, 65536, 131072]             # ramp up by doubling
    if m <= 32768:
        assert s == [m]
    if 32768 < m <= 131072:
        assert len(s) == 3 and s[0] == s[-1] == max(8192, m // 4) // 1024 * 1024   # quarter-sized first / last chunk


def test_forced_chunk_size():
    assert schedule(300_000, 33 * 1024) == [33792] * 8 + [300_000 - 8 * 33792]
    assert schedule(5000, 1024) == [1024] * 4 + [904]
    assert schedule(0) == []


def detect(pos):
    n = (ctypes.c_int64 * 3)()
    ok = L.gsf_debug_detect_grid(pos.shape[0], pos.shape[1], pos.ctypes.data, pos.strides[0] // 8,
                                 pos.strides[1] // 8, n)
    return tuple(n[: pos.shape[0]]) if ok else None


def expand(axes, indexing="ij"):
    g = np.meshgrid(*axes, indexing=indexing)
    return np.ascontiguousarray(np.stack([x.ravel() for x in g]))


def test_detect_grid_shapes():
    rng = np.random.default_rng(0)
    for shape in [(64, 100), (17, 300), (40, 50, 60), (1, 90, 80), (70, 1, 90), (300, 20, 8)]:
        axes = [np.sort(rng.uniform(-5, 5, n)) for n in shape]
        assert detect(expand(axes)) == shape
    # repeated axis values are still a grid as long as the C-order structure holds
    axes = [np.array([0.0, 1.0, 1.0, 2.0] * 10), np.linspace(0, 1, 128)]
    got = detect(expand(axes))
    assert got is None or got == (40, 128)


def test_detect_grid_rejects():
    rng = np.random.default_rng(1)
    axes = [np.linspace(0, 1, 40), np.linspace(0, 2, 50), np.linspace(0, 3, 60)]
    pos = expand(axes)
    assert detect(pos) == (40, 50, 60)
    q = pos.copy(); q[2, 77777] = np.nextafter(q[2, 77777], 10.0)
    assert detect(q) is None                                  # one ulp off anywhere => not a grid
    q = pos.copy(); q[0, -1] += 1.0
    assert detect(q) is None
    assert detect(rng.uniform(size=(3, 50000))) is None       # scattered points
    assert detect(np.stack([np.linspace(0, 1, 50000)] * 2)) is None      # a line (C1-like)
    assert detect(expand(axes, "xy")) is None                 # Fortran-style expansion
    assert detect(pos[:, :4000]) is None                      # too few points to bother
    assert detect(np.asfortranarray(pos)) is None             # needs unit stride along points
    assert detect(expand([np.linspace(0, 1, 5000), np.linspace(0, 1, 4)])) is None   # rows of 4: no gain
    # -0.0 vs +0.0 differ bitwise: treated as a mismatch (falls back to the general kernel)
    q = pos.copy(); q[1, 5] = -0.0
    assert pos[1, 5] == 0.0 and q[1, 5] == pos[1, 5] and detect(q) is None


def _thresholds(edge, tol=1.0):
    a, b = ctypes.c_double(), ctypes.c_double()
    assert L.gsf_debug_variogram_thresholds(ctypes.c_double(edge), ctypes.c_double(tol), ctypes.byref(a),
                                              ctypes.byref(b)) == 0
    return a.value, b.value


def test_variogram_sqrt_threshold_is_exact():
    # t(e) = min{x : sqrt(x) >= e}: what makes `d2 >= t(e)` the same test as `sqrt(d2) >= e`
    # (gsf_variogram_kernels.cuh; reference comparison src/variogram.rs:397,518)
    rng = np.random.default_rng(7)
    edges = np.concatenate([10.0 ** rng.uniform(-300, 150, 400), rng.uniform(0, 10, 400),
                            [1.0, 2.0, 3.0, 5.0 / 3.0, 1e-320, 5e-324, 1.3e154, 1.5e154, 1.7976931348623157e308]])
    for e in edges:
        t, _ = _thresholds(float(e))
        assert np.sqrt(t) >= e
        assert t == 0.0 or np.sqrt(np.nextafter(t, 0.0)) < e
    assert _thresholds(0.0)[0] == 0.0 and _thresholds(-3.0)[0] == 0.0
    assert _thresholds(np.inf)[0] == np.inf and np.isnan(_thresholds(np.nan)[0])


def test_variogram_acos_threshold_is_exact():
    # a*(tol) = max{a in [0,1) : acos(a) >= tol}: `angle <= a*` replaces `acos(angle) >= tol`
    # (reference: src/variogram.rs:281-285)
    import math
    for tol in [math.pi / 8, math.pi / 4, 1e-3, 1e-9, 1.0, math.pi / 2, 0.3]:
        _, a = _thresholds(1.0, tol)
        assert 0.0 <= a < 1.0 and math.acos(a) >= tol
        nxt = float(np.nextafter(a, 2.0))
        assert nxt >= 1.0 or math.acos(nxt) < tol
    assert _thresholds(1.0, 2.0)[1] == -1.0            # acos(a) <= pi/2 < tol for every a >= 0: never rejected
    assert _thresholds(1.0, 1e-300)[1] == float(np.nextafter(1.0, 0.0))


def test_cpulist_parser_of_the_numa_binding():
    """bind_rank_to_gpu_node reads /sys/bus/pci/devices/<gpu>/local_cpulist; its parser on the formats sysfs prints."""
    L.gsf_debug_parse_cpulist.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    L.gsf_debug_parse_cpulist.restype = ctypes.c_int

    def parse(text):
        buf = (ctypes.c_int * 1024)()
        n = L.gsf_debug_parse_cpulist(text.encode(), buf, 1024)
        return list(buf[:n])

    assert parse("0-3") == [0, 1, 2, 3]
    assert parse("0-15,64-79\n") == list(range(16)) + list(range(64, 80))
    assert parse("5") == [5]
    assert parse("1,3,5-6") == [1, 3, 5, 6]
    assert parse("2-2,2") == [2]
    assert parse("") == [] and parse("\n") == []
    assert parse("0-1,x") == [0, 1]


def test_mode_group_detector():
    """detect_mode_group (structured-grid path): run length of consecutive modes sharing all but the last
    wave-vector component -- exact, and 1 whenever the structure is not perfect."""
    L.gsf_debug_mode_group.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
    L.gsf_debug_mode_group.restype = ctypes.c_int64

    def group(k):
        return L.gsf_debug_mode_group(k.shape[0], k.shape[1], k.ctypes.data, k.strides[0] // 8, k.strides[1] // 8)

    def lattice(shape, order="C"):
        axes = [np.linspace(0.1, 3.0, s) * (a + 1) for a, s in enumerate(shape)]
        return np.ascontiguousarray(np.stack([g.ravel(order=order) for g in np.meshgrid(*axes, indexing="ij")]))

    assert group(lattice((12, 16))) == 16                      # Fourier lattice, last component fastest
    assert group(lattice((5, 4, 8))) == 8
    assert group(lattice((100, 100))) == 100                   # C4
    assert group(lattice((12, 16), order="F")) == 1            # first component fastest: no runs
    assert group(np.random.default_rng(0).normal(size=(3, 1000))) == 1      # randomization method
    k = lattice((12, 16)); k[0, 17] += 1e-12
    assert group(k) == 1                                       # one perturbed component
    k = lattice((12, 16))[:, :-16]; k = np.ascontiguousarray(np.concatenate([k, k[:, :8]], axis=1))
    assert group(k) == 1                                       # last run shorter than the others
    assert group(lattice((4, 8))) == 1                         # fewer than 64 modes: not worth it
    assert group(lattice((40, 2))) == 1                        # runs shorter than 4
    k = lattice((12, 16))
    assert group(k[:, ::-1]) == 16                             # negative mode stride (a view)
    two = np.ascontiguousarray(np.concatenate([lattice((6, 16)), lattice((6, 16))], axis=1))
    assert group(two) == 16                                    # repeated outer values in separate runs are fine
    merged = np.ascontiguousarray(np.concatenate([lattice((1, 16)), lattice((1, 16)), lattice((10, 16))[:, 16:]], axis=1))
    assert group(merged) == 1                                  # one run twice as long as the rest
