"""GPU tests of the structured-grid path and the fused post-scale (8 f1).

The grid path must return the same field as the general kernel / the CPU oracle on the expanded
grid (tolerance 1e-9 sigma), both when requested explicitly with axis vectors and when the plain
reference-shaped call detects the grid in `pos`.
"""
import numpy as np
import pytest

import gstools_core as gc
import oracle
from gstools_core import workloads

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(autouse=True)
def _reset():
    if gc.device_count() < 1:
        pytest.fail("no CUDA device")
    gc.set_variant(0, 0)
    gc.set_chunk_points(0)
    gc.set_devices(None)
    gc.set_grid_detection(None)
    yield
    gc.set_grid_detection(None)


def rel_err(got, ref):
    s = float(np.std(ref)) or 1.0
    return float(np.max(np.abs(got - ref))) / s


def expand(axes):
    g = np.meshgrid(*axes, indexing="ij")
    return np.ascontiguousarray(np.stack([x.ravel() for x in g]))


def modes(seed, d, n, heavy=False):
    rng = np.random.default_rng(seed)
    k = rng.normal(size=(d, n))
    if heavy:
        k = k / np.abs(rng.normal(size=n))
    return k, rng.normal(size=n), rng.normal(size=n), rng.normal(size=n)


@pytest.mark.parametrize("shape", [(33, 20), (64, 100), (5, 7), (100, 8), (40, 300), (17, 1031)])
@pytest.mark.parametrize("n", [1, 40, 257])
def test_grid_2d_explicit(shape, n):
    rng = np.random.default_rng(1)
    axes = [np.sort(rng.uniform(-30, 30, s)) for s in shape]      # non-uniform rectilinear grid
    k, z1, z2, sf = modes(2, 2, n)
    pos = expand(axes)
    ref = oracle.summate(k, z1, z2, pos, oracle.max_threads())
    got = gc.summate_grid(k, z1, z2, axes)
    assert gc.last_stats()["grid_path"] == 2
    assert got.shape == ref.shape and rel_err(got, ref) <= TOL
    reff = oracle.summate_fourier(sf, k, z1, z2, pos, oracle.max_threads())
    assert rel_err(gc.summate_fourier_grid(sf, k, z1, z2, axes), reff) <= TOL
    refi = oracle.summate_incompr(k, z1, z2, pos)
    goti = gc.summate_incompr_grid(k, z1, z2, axes)
    assert goti.shape == (2, pos.shape[1]) and goti.flags.f_contiguous
    assert rel_err(goti, refi) <= TOL


@pytest.mark.parametrize("shape", [(10, 11, 12), (3, 50, 100), (32, 1, 64), (1, 40, 40), (21, 22, 130), (9, 9, 700)])
@pytest.mark.parametrize("n", [33, 1000])
def test_grid_3d_explicit(shape, n):
    rng = np.random.default_rng(3)
    axes = [np.sort(rng.uniform(0, 100, s)) for s in shape]
    k, z1, z2, sf = modes(4, 3, n, heavy=True)
    pos = expand(axes)
    ref = oracle.summate(k, z1, z2, pos, oracle.max_threads())
    got = gc.summate_grid(k, z1, z2, axes)
    assert rel_err(got, ref) <= TOL
    refi = oracle.summate_incompr(k, z1, z2, pos, 1)
    assert rel_err(gc.summate_incompr_grid(k, z1, z2, axes), refi) <= TOL


def test_auto_detection_matches_general_kernel():
    axes = [np.linspace(0, 50, 41), np.linspace(-5, 5, 37), np.linspace(2, 9, 64)]
    pos = expand(axes)
    k, z1, z2, sf = modes(5, 3, 300, heavy=True)
    gc.set_grid_detection(False)
    gen = gc.summate(k, z1, z2, pos)
    assert gc.last_stats()["grid_path"] == 0
    gc.set_grid_detection(True)
    auto = gc.summate(k, z1, z2, pos)
    assert gc.last_stats()["grid_path"] == 1
    assert rel_err(auto, gen) <= TOL
    assert rel_err(auto, oracle.summate(k, z1, z2, pos, oracle.max_threads())) <= TOL
    # incompr + fourier through the reference-shaped calls
    assert rel_err(gc.summate_incompr(k, z1, z2, pos), oracle.summate_incompr(k, z1, z2, pos)) <= TOL
    assert gc.last_stats()["grid_path"] == 1
    assert rel_err(gc.summate_fourier(sf, k, z1, z2, pos), oracle.summate_fourier(sf, k, z1, z2, pos)) <= TOL


def test_detection_rejects_non_grids():
    axes = [np.linspace(0, 50, 41), np.linspace(-5, 5, 37), np.linspace(2, 9, 64)]
    pos = expand(axes)
    k, z1, z2, _ = modes(6, 3, 256)
    gc.set_grid_detection(True)
    gc.summate(k, z1, z2, pos)
    assert gc.last_stats()["grid_path"] == 1                       # the unmodified grid is detected
    for mutate in (lambda p: p.__setitem__((2, 12345), p[2, 12345] + 1e-13),     # one perturbed point
                   lambda p: p.__setitem__((0, -1), np.nextafter(p[0, -1], 1e9)),
                   lambda p: p.__setitem__((1, 64 * 5 + 3), -p[1, 64 * 5 + 3] - 1.0)):
        q = pos.copy()
        mutate(q)
        got = gc.summate(k, z1, z2, q)
        assert gc.last_stats()["grid_path"] == 0                   # exact check failed -> general kernel
        assert rel_err(got, oracle.summate(k, z1, z2, q, oracle.max_threads())) <= TOL
    # random points, a line (C1-like), F-ordered meshgrid (xy indexing) are not C-order grids
    rng = np.random.default_rng(0)
    rnd = rng.uniform(0, 10, size=(3, 70000))
    gc.summate(k, z1, z2, rnd)
    assert gc.last_stats()["grid_path"] == 0
    line = np.stack([np.linspace(0, 10, 50000)] * 3)
    gc.summate(k, z1, z2, line)
    assert gc.last_stats()["grid_path"] == 0
    gx = np.meshgrid(*axes, indexing="xy")
    posxy = np.ascontiguousarray(np.stack([g.ravel() for g in gx]))
    got = gc.summate(k, z1, z2, posxy)
    assert rel_err(got, oracle.summate(k, z1, z2, posxy, oracle.max_threads())) <= TOL


def test_grid_edge_cases():
    k, z1, z2, sf = modes(9, 3, 50)
    # empty grid: one axis of length 0
    out = gc.summate_grid(k, z1, z2, [np.linspace(0, 1, 5), np.zeros(0), np.linspace(0, 1, 7)])
    assert out.shape == (0,)
    # single-point axes, N = 0 (fold identity), NaN mode, zero mode in incompr
    axes = [np.array([1.5]), np.linspace(0, 1, 40), np.linspace(0, 2, 33)]
    pos = expand(axes)
    assert rel_err(gc.summate_grid(k, z1, z2, axes), oracle.summate(k, z1, z2, pos)) <= TOL
    z0 = np.zeros(0)
    assert np.array_equal(gc.summate_grid(np.zeros((3, 0)), z0, z0, axes), np.zeros(pos.shape[1]))
    kn = k.copy(); kn[1, 7] = np.nan
    assert np.isnan(gc.summate_grid(kn, z1, z2, axes)).all()
    kz = k.copy(); kz[:, 3] = 0.0
    assert np.isnan(gc.summate_incompr_grid(kz, z1, z2, axes)).all()       # 0/0 as in the reference
    with pytest.raises(ValueError):
        gc.summate_incompr_grid(np.zeros((3, 0)), z0, z0, axes)             # field.rs:163
    with pytest.raises(ValueError):
        gc.summate_grid(k, z1, z2, axes[:2])                               # axis count != dim
    # an infinite coordinate poisons exactly the points that use it
    ax = [np.linspace(0, 1, 20), np.array([0.0, np.inf, 1.0]), np.linspace(0, 1, 16)]
    out = gc.summate_grid(k, z1, z2, ax).reshape(20, 3, 16)
    assert np.isnan(out[:, 1, :]).all() and np.isfinite(out[:, 0, :]).all() and np.isfinite(out[:, 2, :]).all()


def test_scale_and_offset_fusion():
    k, z1, z2, sf = modes(7, 3, 200)
    rng = np.random.default_rng(1)
    pos = rng.uniform(0, 20, size=(3, 5000))
    base = gc.summate(k, z1, z2, pos)
    got = gc.summate_scaled(k, z1, z2, pos, scale=0.37, offset=2.5)
    assert np.max(np.abs(got - (0.37 * base + 2.5))) <= 1e-12 * max(1.0, np.std(base))
    basei = gc.summate_incompr(k, z1, z2, pos)
    goti = gc.summate_incompr_scaled(k, z1, z2, pos, scale=1.5, offset=(3.0, 0.0, -1.0))
    want = 1.5 * basei + np.array([[3.0], [0.0], [-1.0]])
    assert np.max(np.abs(goti - want)) <= 1e-12 * max(1.0, np.std(basei))
    axes = [np.linspace(0, 9, 30), np.linspace(0, 5, 20), np.linspace(0, 1, 16)]
    g0 = gc.summate_grid(k, z1, z2, axes)
    g1 = gc.summate_grid(k, z1, z2, axes, scale=2.0, offset=-0.5)
    assert np.max(np.abs(g1 - (2.0 * g0 - 0.5))) <= 1e-12 * max(1.0, np.std(g0))
    # every lanes-per-point variant adds the offset exactly once
    for P, L in [(1, 4), (2, 8), (1, 32)]:
        gc.set_variant(P, L)
        v = gc.summate_scaled(k, z1, z2, pos, scale=0.37, offset=2.5)
        assert np.max(np.abs(v - (0.37 * base + 2.5))) <= 1e-11 * max(1.0, np.std(base))


@pytest.mark.parametrize("cfg", ["c2", "c3", "c4", "c5"])
def test_baseline_configs_grid_vs_general(cfg):
    """Full-size BASELINE grids: auto-detected grid path vs the oracle on a strided subset (and vs
    the general kernel for the configs where that takes well under a second)."""
    w = workloads.make(cfg)
    fn = getattr(gc, w["kind"])
    gc.set_grid_detection(True)
    got = fn(*w["args"])
    st = gc.last_stats()
    assert st["grid_path"] == 1
    m = w["m"]
    idx = np.unique(np.concatenate([np.arange(0, m, max(1, m // 6000)), np.arange(2048), np.arange(m - 2048, m)]))
    ref = getattr(oracle, w["kind"])(*workloads.subset_points(w, idx)["args"], oracle.max_threads())
    sub = got[:, idx] if got.ndim == 2 else got[idx]
    e = rel_err(sub, ref)
    print("%s grid path: max|d|/sigma=%.3g total_ms=%.2f chunks=%d" % (cfg, e, st["total_ms"], st["n_chunks"]))
    assert e <= TOL
    if cfg in ("c2", "c3"):
        gc.set_grid_detection(False)
        gen = fn(*w["args"])
        assert gc.last_stats()["grid_path"] == 0
        assert rel_err(got, gen) <= TOL


def test_device_resident_grid_detection():
    """Synchronous API with device-resident pos/out: the grid is detected ON the device (two small
    kernels) and routed to the GEMM path; the stream-ordered API never detects (it must not sync)."""
    torch = pytest.importorskip("torch")
    axes = [np.linspace(0, 50, 41), np.linspace(-5, 5, 37), np.linspace(2, 9, 64)]
    pos = expand(axes)
    k, z1, z2, sf = modes(12, 3, 300, heavy=True)
    ref = oracle.summate(k, z1, z2, pos, oracle.max_threads())
    dpos = torch.from_numpy(pos).cuda()
    out = torch.empty(pos.shape[1], dtype=torch.float64, device="cuda")
    gc.set_grid_detection(True)
    gc.summate_device(k, z1, z2, dpos, out, sync=True)
    assert gc.last_stats()["grid_path"] == 1
    assert rel_err(out.cpu().numpy(), ref) <= TOL
    # incompressible, F-ordered device output
    outi = torch.empty((pos.shape[1], 3), dtype=torch.float64, device="cuda").t()
    gc.summate_incompr_device(k, z1, z2, dpos, outi, sync=True)
    assert gc.last_stats()["grid_path"] == 1
    assert rel_err(outi.cpu().numpy(), oracle.summate_incompr(k, z1, z2, pos)) <= TOL
    # a perturbed point => general kernel
    q = pos.copy(); q[1, 4321] += 1e-9
    dq = torch.from_numpy(q).cuda()
    gc.summate_device(k, z1, z2, dq, out, sync=True)
    assert gc.last_stats()["grid_path"] == 0
    assert rel_err(out.cpu().numpy(), oracle.summate(k, z1, z2, q, oracle.max_threads())) <= TOL
    # 2-D grid with AoS (Fortran-ordered) device positions: strides handled on the device
    ax2 = [np.linspace(0, 3, 300), np.linspace(0, 2, 200)]
    p2 = expand(ax2)
    k2, a, b, _ = modes(13, 2, 400)
    d2 = torch.from_numpy(np.ascontiguousarray(p2.T)).cuda().t()
    o2 = torch.empty(p2.shape[1], dtype=torch.float64, device="cuda")
    gc.summate_device(k2, a, b, d2, o2, sync=True)
    assert gc.last_stats()["grid_path"] == 1
    assert rel_err(o2.cpu().numpy(), oracle.summate(k2, a, b, p2, oracle.max_threads())) <= TOL
    # stream-ordered call: no detection
    gc.summate_device(k, z1, z2, dpos, out, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert gc.last_stats()["grid_path"] == 0
    assert rel_err(out.cpu().numpy(), ref) <= TOL


@pytest.mark.parametrize("shape,n", [((50, 40, 100), 1000),      # 63 row tiles x 63 K blocks
                                     ((3, 21, 300), 1000),       # 2 row tiles x 3 column blocks: 16-way split
                                     ((90, 130), 250),           # 2-D, 3 row tiles, two column blocks
                                     ((7, 9, 40), 40)])          # 3 K blocks only
def test_mode_split_launches_are_exact_and_deterministic(shape, n):
    """Small grids split the modes over grid.z; the last-arriving CTA of a tile adds the partial sums in
    split order.  The result must not depend on arrival order (bit-identical repeats, self-resetting
    tickets) and must match the oracle."""
    d = len(shape)
    axes = [np.sort(np.random.default_rng(3 + i).uniform(0, 60, s)) for i, s in enumerate(shape)]
    k, z1, z2, _ = modes(31, d, n, heavy=True)
    pos = expand(axes)
    ref = oracle.summate(k, z1, z2, pos, oracle.max_threads())
    first = gc.summate_grid(k, z1, z2, axes)
    assert gc.last_stats()["grid_path"] == 2
    assert rel_err(first, ref) <= TOL
    for _ in range(4):
        assert np.array_equal(first, gc.summate_grid(k, z1, z2, axes))
    refi = oracle.summate_incompr(k, z1, z2, pos, oracle.max_threads())
    goti = gc.summate_incompr_grid(k, z1, z2, axes)
    assert rel_err(goti, refi) <= TOL
    assert np.array_equal(goti, gc.summate_incompr_grid(k, z1, z2, axes))
    # device-resident result: one launch over all rows
    import torch
    out = torch.empty(pos.shape[1], dtype=torch.float64, device="cuda")
    gc.summate_grid(k, z1, z2, axes, out=out)
    torch.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), ref) <= TOL


def _lattice_modes(shape, seed):
    """wave vectors on a tensor lattice, flattened in C order (last component fastest) -- what the Fourier
    generator passes: modes = meshgrid(kx, ky[, kz], indexing='ij').reshape(dim, -1)"""
    rng = np.random.default_rng(seed)
    axes_k = [np.sort(rng.uniform(-2.0, 2.0, s)) for s in shape]
    k = np.ascontiguousarray(np.stack([g.ravel() for g in np.meshgrid(*axes_k, indexing="ij")]))
    n = k.shape[1]
    return k, rng.normal(size=n), rng.normal(size=n), rng.uniform(0.1, 1.0, size=n)


@pytest.mark.parametrize("kshape,pshape", [((12, 16), (70, 90)), ((5, 4, 8), (9, 10, 40)), ((3, 100), (33, 64))])
def test_tensor_structured_modes_are_summed_per_group(kshape, pshape):
    """Modes on a lattice (Fourier method): runs of modes sharing all but the last wave-vector component
    are summed in the F table and the GEMM contracts over the runs only.  Same field as the general
    kernel / the oracle; anything that breaks the structure falls back to one row per mode."""
    d = len(kshape)
    k, z1, z2, sf = _lattice_modes(kshape, 41)
    axes = [np.linspace(0.0, 25.0, s, endpoint=False) for s in pshape]
    pos = expand(axes)
    group = kshape[-1]
    ref = oracle.summate_fourier(sf, k, z1, z2, pos, oracle.max_threads())
    got = gc.summate_fourier_grid(sf, k, z1, z2, axes)
    st = gc.last_stats()
    assert st["grid_path"] == 2 and st["mode_group"] == group, st
    assert rel_err(got, ref) <= TOL
    assert np.array_equal(got, gc.summate_fourier_grid(sf, k, z1, z2, axes))
    gc.set_grid_detection(False)
    gen = gc.summate_fourier(sf, k, z1, z2, pos)                      # general point x mode kernel
    gc.set_grid_detection(None)
    assert rel_err(got, gen) <= TOL
    # scalar and incompressible calls see the same structure
    assert rel_err(gc.summate_grid(k, z1, z2, axes), oracle.summate(k, z1, z2, pos, oracle.max_threads())) <= TOL
    assert gc.last_stats()["mode_group"] == group
    goti = gc.summate_incompr_grid(k, z1, z2, axes)
    assert gc.last_stats()["mode_group"] == group
    refi = oracle.summate_incompr(k, z1, z2, pos, oracle.max_threads())
    ok = np.isfinite(refi)                                             # (a zero wave vector gives NaN, as in the reference)
    assert np.array_equal(ok, np.isfinite(goti)) and rel_err(goti[ok], refi[ok]) <= TOL
    # one perturbed component breaks the lattice: every mode on its own again, same answer
    k2 = k.copy(); k2[0, group + 1] += 1e-3
    got2 = gc.summate_fourier_grid(sf, k2, z1, z2, axes)
    assert gc.last_stats()["mode_group"] == 1
    assert rel_err(got2, oracle.summate_fourier(sf, k2, z1, z2, pos, oracle.max_threads())) <= TOL
    # lattice flattened with the FIRST component fastest: no runs, plain path
    k3 = np.ascontiguousarray(np.stack([g.ravel(order="F") for g in np.meshgrid(*[np.unique(k[a]) for a in range(d)], indexing="ij")]))
    got3 = gc.summate_fourier_grid(sf, k3, z1, z2, axes)
    assert gc.last_stats()["mode_group"] == 1
    assert rel_err(got3, oracle.summate_fourier(sf, k3, z1, z2, pos, oracle.max_threads())) <= TOL


def test_c4_fourier_lattice_through_the_default_api():
    """BASELINE configs[3] scaled down: the plain reference-shaped call detects the grid in `pos` AND the
    lattice in `modes` (100 x 100 wave vectors -> 100 groups of 100)."""
    w = workloads.make("c4", 1.0 / 256)
    got = gc.summate_fourier(*w["args"])
    st = gc.last_stats()
    assert st["grid_path"] == 1 and st["mode_group"] == 100, st
    ref = oracle.summate_fourier(*w["args"], oracle.max_threads())
    assert rel_err(got, ref) <= TOL
