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
    if m <= 65536:
        assert s == [m]


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
