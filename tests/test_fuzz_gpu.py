"""Seeded fuzz over shapes, strides, kinds and paths: CUDA result vs CPU oracle (1e-9 sigma)."""
import numpy as np
import pytest

import gstools_core as gc
import oracle

import os

pytestmark = pytest.mark.gpu
TOL = 1e-9
N_GENERAL = int(os.environ.get("GSF_FUZZ_GENERAL", "40"))     # raise for a soak run
N_GRID = int(os.environ.get("GSF_FUZZ_GRID", "20"))


def rel_err(got, ref):
    s = float(np.std(ref))
    if not np.isfinite(s) or s == 0.0:
        s = max(1.0, float(np.max(np.abs(ref)))) if ref.size else 1.0
    return float(np.max(np.abs(got - ref))) / s if ref.size else 0.0


def strided(rng, a):
    """Return an array equal to `a` but living in a randomly laid-out buffer."""
    mode = rng.integers(0, 4)
    if mode == 0:
        return a
    if mode == 1:
        return np.asfortranarray(a)
    if mode == 2:                                   # every other element of a larger buffer
        big = np.zeros(tuple(2 * s for s in a.shape))
        sl = tuple(slice(None, None, 2) for _ in a.shape)
        big[sl] = a
        return big[sl]
    big = np.zeros(tuple(s + 3 for s in a.shape))   # interior window
    sl = tuple(slice(1, 1 + s) for s in a.shape)
    big[sl] = a
    return big[sl]


@pytest.mark.parametrize("seed", range(N_GENERAL))
def test_fuzz_general(seed):
    if gc.device_count() < 1:
        pytest.fail("no CUDA device")
    rng = np.random.default_rng(1000 + seed)
    kind = ["summate", "summate_incompr", "summate_fourier"][seed % 3]
    d = int(rng.integers(2, 4)) if kind == "summate_incompr" else int(rng.integers(1, 9))
    n = int(rng.choice([1, 2, 7, 33, 255, 256, 257, 600]))
    m = int(rng.choice([1, 2, 31, 127, 128, 129, 383, 384, 385, 1000, 5003, 40001]))
    k = rng.normal(size=(d, n)) * rng.choice([0.1, 1.0, 30.0])
    z1, z2, sf = rng.normal(size=n), rng.normal(size=n), rng.normal(size=n)
    pos = rng.uniform(-20, 20, size=(d, m)) * rng.choice([1.0, 100.0])
    gc.set_grid_detection(False)
    gc.set_chunk_points(int(rng.choice([0, 0, 1024, 7168])))
    if d <= 3:
        P, L = [(0, 0), (1, 1), (2, 1), (3, 1), (4, 1), (1, 8), (2, 4), (1, 32)][int(rng.integers(0, 8))]
        gc.set_variant(P, L)
    args = [strided(rng, k), strided(rng, z1), strided(rng, z2), strided(rng, pos)]
    try:
        if kind == "summate_fourier":
            got = gc.summate_fourier(strided(rng, sf), *args)
            ref = oracle.summate_fourier(sf, k, z1, z2, pos, oracle.max_threads())
        else:
            got = getattr(gc, kind)(*args)
            ref = getattr(oracle, kind)(k, z1, z2, pos, oracle.max_threads() if kind == "summate" else 1)
    finally:
        gc.set_variant(0, 0)
        gc.set_chunk_points(0)
        gc.set_grid_detection(None)
    assert got.shape == ref.shape
    assert rel_err(got, ref) <= TOL, (kind, d, n, m)


@pytest.mark.parametrize("seed", range(N_GRID))
def test_fuzz_grid(seed):
    rng = np.random.default_rng(2000 + seed)
    kind = ["summate", "summate_incompr", "summate_fourier"][seed % 3]
    d = int(rng.integers(2, 4))
    shape = [int(rng.choice([1, 2, 5, 8, 17, 32, 33, 64, 100])) for _ in range(d)]
    shape[-1] = int(rng.choice([1, 7, 8, 9, 40, 100, 128, 129, 300]))
    n = int(rng.choice([1, 15, 16, 17, 100, 333]))
    axes = [np.sort(rng.uniform(-10, 10, s)) for s in shape]
    g = np.meshgrid(*axes, indexing="ij")
    pos = np.ascontiguousarray(np.stack([x.ravel() for x in g]))
    k = rng.normal(size=(d, n)); z1, z2, sf = rng.normal(size=n), rng.normal(size=n), rng.normal(size=n)
    scale, off = float(rng.normal()), float(rng.normal())
    if kind == "summate":
        got = gc.summate_grid(k, z1, z2, axes, scale=scale, offset=off)
        ref = scale * oracle.summate(k, z1, z2, pos) + off
    elif kind == "summate_fourier":
        got = gc.summate_fourier_grid(sf, k, z1, z2, axes, scale=scale, offset=off)
        ref = scale * oracle.summate_fourier(sf, k, z1, z2, pos) + off
    else:
        offs = rng.normal(size=d)
        got = gc.summate_incompr_grid(k, z1, z2, axes, scale=scale, offset=offs)
        ref = scale * oracle.summate_incompr(k, z1, z2, pos) + offs[:, None]
    assert got.shape == ref.shape
    s = float(np.std(ref)) or 1.0
    assert float(np.max(np.abs(got - ref))) <= TOL * max(s, abs(scale)), (kind, shape, n)
