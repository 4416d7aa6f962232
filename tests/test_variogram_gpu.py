"""GPU parity of the variogram estimators against the CPU oracle (oracle/variogram_oracle.c) and the
reference's own known-answer tests (src/variogram.rs:577-842, tests/golden/variogram_rs_kat.json).

Bar: counts bit-exact (the kernels bin on squared distances against host-computed exact thresholds,
gsf_variogram_kernels.cuh); variogram values within REL_TOL of the oracle -- the GPU adds the same
terms in a different (fixed) order, so the difference is a few ulp times sqrt(#pairs).
"""
import json
import os

import numpy as np
import pytest

import gstools_core as gc
import oracle

pytestmark = pytest.mark.gpu

REL_TOL = 1e-12
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def vkat():
    with open(os.path.join(HERE, "golden", "variogram_rs_kat.json")) as fh:
        return json.load(fh)


def close(a, b, tol=REL_TOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    scale = max(float(np.nanmax(np.abs(b), initial=0.0)), 1e-300)
    assert np.all(np.isnan(a) == np.isnan(b))
    assert np.nanmax(np.abs(a - b), initial=0.0) <= tol * scale, (a, b)


def same_counts(a, b):
    assert a.dtype == np.uint64 and a.shape == b.shape
    assert np.array_equal(a, b), (a, b)


def scattered(rng, d, m, nf=1, nan_frac=0.0):
    pos = rng.uniform(0.0, 100.0, (d, m))
    f = rng.normal(size=(nf, m))
    if nan_frac:
        f[rng.uniform(size=f.shape) < nan_frac] = np.nan
    return pos, f


# ---- the reference's known answers ---------------------------------------------------------------

def test_kat_structured(vkat):
    f = np.array(vkat["struct_field"]).reshape(-1, 1)
    close(gc.variogram_structured(f, "m"), vkat["struct_gamma"], 2e-15)
    close(gc.variogram_ma_structured(f, np.zeros((10, 1), dtype=bool), "m"), vkat["struct_gamma"], 2e-15)
    mask2 = np.array(vkat["ma_struct_mask2"]).reshape(-1, 1)
    close(gc.variogram_ma_structured(f, mask2, "m"), vkat["ma_struct_gamma2"], 2e-15)


def test_kat_unstructured_and_directional(vkat):
    pos = np.stack([np.arange(0.0, 10.0, 1.0), np.arange(0.0, 10.0, 1.0)])
    f = np.array([vkat["unstruct_field"]])
    edges = np.linspace(0.0, 5.0, 4)
    g, c = gc.variogram_unstructured(f, edges, pos, "m", "e")
    close(g, vkat["unstruct_gamma"], 2e-15)
    assert c.tolist() == vkat["unstruct_counts"]
    direction = np.array([[0.0, np.pi], [0.0, 0.0]])
    g, c = gc.variogram_directional(f, edges, pos, direction, np.pi / 8.0, -1.0, False, "m")
    close(g, vkat["directional_gamma"], 2e-15)
    assert c.tolist() == vkat["directional_counts"]
    # src/variogram.rs:711-816: multi-field = mean of the single-field estimates
    f2 = np.array([vkat["unstruct_field2"]])
    g1, _ = gc.variogram_unstructured(f, edges, pos)
    g2, _ = gc.variogram_unstructured(f2, edges, pos)
    gm, _ = gc.variogram_unstructured(np.concatenate([f, f2]), edges, pos)
    close(gm, 0.5 * (g1 + g2), 4e-15)


# ---- unstructured --------------------------------------------------------------------------------

@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 8])      # d >= 4: generic pair kernel, points in input order
@pytest.mark.parametrize("est", ["m", "c"])
@pytest.mark.parametrize("m", [1, 2, 129, 1500])
def test_unstructured_random(d, est, m):
    rng = np.random.default_rng(100 * d + m)
    pos, f = scattered(rng, d, m)
    edges = np.linspace(0.0, 60.0 * max(1.0, np.sqrt(d / 3.0)), 17)   # typical distances grow like sqrt(d)
    g, c = gc.variogram_unstructured(f, edges, pos, est, "e")
    go, co = oracle.variogram_unstructured(f, edges, pos, est, "e", oracle.max_threads())
    same_counts(c, co)
    close(g, go)


def test_space_time_lattice_ties_and_dim_limit():
    """4-D integer lattice (exact ties at the bin edges, several j chunks); dim 9 is refused loudly"""
    rng = np.random.default_rng(44)
    pos = rng.integers(0, 6, size=(4, 2500)).astype(np.float64)
    f = rng.normal(size=(1, 2500))
    edges = np.arange(0.0, 9.0)
    g, c = gc.variogram_unstructured(f, edges, pos, "m", "e")
    go, co = oracle.variogram_unstructured(f, edges, pos, "m", "e", oracle.max_threads())
    same_counts(c, co)
    close(g, go)
    with pytest.raises(ValueError):
        gc.variogram_unstructured(f[:, :10], edges, rng.normal(size=(9, 10)), "m", "e")


def test_unstructured_integer_lattice_ties():
    # distances that hit bin edges exactly (3-4-5 triangles etc.): membership must follow
    # `dist < lo || dist >= hi` on the rounded sqrt, src/variogram.rs:518
    rng = np.random.default_rng(5)
    xs, ys = np.meshgrid(np.arange(40.0), np.arange(30.0), indexing="ij")
    pos = np.stack([xs.ravel(), ys.ravel()])
    f = rng.normal(size=(1, pos.shape[1]))
    edges = np.array([0.0, 1.0, np.sqrt(2.0), 2.0, np.sqrt(5.0), 3.0, 5.0, np.sqrt(50.0), 13.0, 25.0])
    g, c = gc.variogram_unstructured(f, edges, pos)
    go, co = oracle.variogram_unstructured(f, edges, pos, "m", "e", oracle.max_threads())
    same_counts(c, co)
    close(g, go)


def test_unstructured_multi_field_with_nodata():
    rng = np.random.default_rng(11)
    pos, f = scattered(rng, 2, 700, nf=5, nan_frac=0.1)
    edges = np.linspace(0.0, 50.0, 9)
    for est in ("m", "c"):
        g, c = gc.variogram_unstructured(f, edges, pos, est)
        go, co = oracle.variogram_unstructured(f, edges, pos, est, "e", oracle.max_threads())
        same_counts(c, co)
        close(g, go)


def test_unstructured_strided_inputs_and_many_bins():
    # F-ordered / transposed / sliced views, and more bins than one pass of the kernel holds
    rng = np.random.default_rng(12)
    pos, f = scattered(rng, 3, 900, nf=2)
    posv = np.asfortranarray(pos)
    fv = np.repeat(f, 2, axis=1)[:, ::2]
    edges_full = np.linspace(0.0, 90.0, 2 * 131)
    edges = edges_full[::2]                               # 131 edges, stride 2
    g, c = gc.variogram_unstructured(fv, edges, posv)
    go, co = oracle.variogram_unstructured(f, np.ascontiguousarray(edges), pos, "m", "e", oracle.max_threads())
    same_counts(c, co)
    close(g, go)


def test_unstructured_irregular_bin_edges():
    rng = np.random.default_rng(13)
    pos, f = scattered(rng, 2, 600)
    for edges in (np.array([10.0, 5.0, 20.0, 15.0, 40.0]),      # overlapping / inverted bins
                  np.array([0.0, 10.0, 10.0, 30.0]),            # empty bin
                  np.array([-5.0, 0.0, 1e-300, np.inf]),        # non-positive and infinite edges
                  np.array([0.0, np.nan, 30.0]),                # NaN edge: that side never excludes
                  np.array([7.0, 7.0])):                        # single empty bin
        g, c = gc.variogram_unstructured(f, edges, pos)
        go, co = oracle.variogram_unstructured(f, edges, pos, "m", "e", oracle.max_threads())
        same_counts(c, co)
        close(g, go)


def test_unstructured_repeated_points_and_nan_positions():
    rng = np.random.default_rng(14)
    pos, f = scattered(rng, 2, 300)
    pos[:, 100:150] = pos[:, 0:50]            # repeated points: dist == 0 lands in a bin starting at 0
    pos[0, 7] = np.nan                        # NaN distance: neither `<` nor `>=` excludes it (reference quirk)
    edges = np.linspace(0.0, 40.0, 6)
    g, c = gc.variogram_unstructured(f, edges, pos)
    go, co = oracle.variogram_unstructured(f, edges, pos, "m", "e", oracle.max_threads())
    same_counts(c, co)
    close(g, go)


def test_unstructured_larger_than_one_tile_row():
    rng = np.random.default_rng(15)
    pos, f = scattered(rng, 2, 5000)          # 256-point j chunks, 40 i blocks, diagonal + off-diagonal tiles
    edges = np.linspace(0.0, 30.0, 13)
    g, c = gc.variogram_unstructured(f, edges, pos)
    go, co = oracle.variogram_unstructured(f, edges, pos, "m", "e", oracle.max_threads())
    same_counts(c, co)
    close(g, go)
    g2, c2 = gc.variogram_unstructured(f, edges, pos)
    assert np.array_equal(g, g2) and np.array_equal(c, c2)      # run-to-run deterministic
    assert int(c.sum()) <= 5000 * 4999 // 2


def test_haversine():
    rng = np.random.default_rng(16)
    m = 800
    pos = np.stack([rng.uniform(-80.0, 80.0, m), rng.uniform(-180.0, 180.0, m)])
    f = rng.normal(size=(2, m))
    edges = np.linspace(0.0, 2.0, 11)         # radians on the unit sphere
    for est in ("m", "c"):
        g, c = gc.variogram_unstructured(f, edges, pos, est, "h")
        go, co = oracle.variogram_unstructured(f, edges, pos, est, "h", oracle.max_threads())
        # CUDA's sin/cos/atan2 differ from glibc's in the last ulp, so a pair sitting within an ulp of
        # an edge could move to the neighbouring bin; for these seeded inputs none does (the chance
        # is ~1e-9), and with equal counts the sums contain the same terms
        same_counts(c, co)
        close(g, go)
    with pytest.raises(ValueError, match="Haversine"):
        gc.variogram_unstructured(np.ones((1, 4)), edges, np.ones((3, 4)), "m", "h")


# ---- directional ---------------------------------------------------------------------------------

def unit(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


@pytest.mark.parametrize("d", [1, 2, 3, 4, 6, 7])
@pytest.mark.parametrize("bandwidth,separate", [(-1.0, False), (8.0, False), (8.0, True), (-1.0, True)])
def test_directional_random(d, bandwidth, separate):
    rng = np.random.default_rng(200 + d)
    pos, f = scattered(rng, d, 900, nf=2, nan_frac=0.05)
    direction = unit(rng.normal(size=(3, d)))
    edges = np.linspace(0.0, 50.0 * max(1.0, np.sqrt(d / 3.0)), 11)
    for est in ("m", "c"):
        g, c = gc.variogram_directional(f, edges, pos, direction, np.pi / 6.0, bandwidth, separate, est)
        go, co = oracle.variogram_directional(f, edges, pos, direction, np.pi / 6.0, bandwidth, separate, est,
                                              oracle.max_threads())
        same_counts(c, co)
        close(g, go)


def test_directional_axis_aligned_lattice():
    # angles and band distances that sit exactly on the tolerance: the threshold form of the acos
    # and sqrt tests must agree with libm's on every pair
    xs, ys = np.meshgrid(np.arange(25.0), np.arange(25.0), indexing="ij")
    pos = np.stack([xs.ravel(), ys.ravel()])
    rng = np.random.default_rng(21)
    f = rng.normal(size=(1, pos.shape[1]))
    direction = np.array([[1.0, 0.0], [0.0, 1.0], [np.sqrt(0.5), np.sqrt(0.5)]])
    edges = np.array([0.0, 1.5, 3.0, 6.0, 12.0])
    for tol in (np.pi / 4.0, np.pi / 8.0, np.arctan(0.5), 1e-6):
        for bw in (-1.0, 1.0, 2.0):
            g, c = gc.variogram_directional(f, edges, pos, direction, tol, bw)
            go, co = oracle.variogram_directional(f, edges, pos, direction, tol, bw, False, "m",
                                                  oracle.max_threads())
            same_counts(c, co)
            close(g, go)


def test_directional_defaults_and_many_directions():
    rng = np.random.default_rng(22)
    pos, f = scattered(rng, 2, 400)
    ang = np.linspace(0.0, np.pi, 30, endpoint=False)
    direction = np.stack([np.cos(ang), np.sin(ang)], axis=1)            # 30 directions x 40 bins: several passes
    edges = np.linspace(0.0, 80.0, 41)
    g, c = gc.variogram_directional(f, edges, pos, direction)
    go, co = oracle.variogram_directional(f, edges, pos, direction, None, None, None, None, oracle.max_threads())
    same_counts(c, co)
    close(g, go)
    assert g.shape == (30, 40)


# ---- structured ----------------------------------------------------------------------------------

@pytest.mark.parametrize("shape", [(1, 1), (2, 7), (10, 1), (64, 33), (300, 257)])
@pytest.mark.parametrize("est", ["m", "c"])
def test_structured_random(shape, est):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    f = rng.normal(size=shape)
    close(gc.variogram_structured(f, est), oracle.variogram_structured(f, est, oracle.max_threads()))
    mask = rng.uniform(size=shape) < 0.3
    close(gc.variogram_ma_structured(f, mask, est), oracle.variogram_ma_structured(f, mask, est, oracle.max_threads()))
    full = np.ones(shape, dtype=bool)
    close(gc.variogram_ma_structured(f, full, est), oracle.variogram_ma_structured(f, full, est))   # count 0 everywhere


def test_structured_views_and_empty_columns():
    rng = np.random.default_rng(31)
    big = rng.normal(size=(80, 90))
    f = big[::2, 5:65:3]                                   # strided view
    mask = (rng.uniform(size=big.shape) < 0.2)[::2, 5:65:3]
    close(gc.variogram_structured(f), oracle.variogram_structured(np.ascontiguousarray(f)))
    close(gc.variogram_ma_structured(f.T, mask.T), oracle.variogram_ma_structured(np.ascontiguousarray(f.T),
                                                                                   np.ascontiguousarray(mask.T)))
    assert gc.variogram_structured(np.ones((5, 0))).tolist() == [0.0] * 5
    f_nan = rng.normal(size=(20, 4))
    f_nan[3, 1] = np.nan                                   # structured estimators do not skip NaN (src/variogram.rs:159-164)
    close(gc.variogram_structured(f_nan), oracle.variogram_structured(f_nan))


def test_structured_large_field_splits():
    rng = np.random.default_rng(32)
    f = rng.normal(size=(700, 1200))                       # several splits per lag
    g = gc.variogram_structured(f)
    close(g, oracle.variogram_structured(f, "m", oracle.max_threads()))
    assert np.array_equal(g, gc.variogram_structured(f))   # deterministic


def test_field_summation_still_works_after_variograms():
    # the variogram paths borrow the field paths' device workspaces
    rng = np.random.default_rng(33)
    k = rng.normal(size=(2, 50)); z1 = rng.normal(size=50); z2 = rng.normal(size=50)
    pos = rng.uniform(0, 10, (2, 3000))
    before = gc.summate(k, z1, z2, pos)
    p2, f2 = scattered(rng, 2, 500)
    gc.variogram_unstructured(f2, np.linspace(0, 50, 8), p2)
    gc.variogram_structured(rng.normal(size=(50, 60)))
    assert np.array_equal(before, gc.summate(k, z1, z2, pos))


# ---- seeded fuzz over shapes, estimators, edges and directions -----------------------------------

@pytest.mark.parametrize("seed", range(24))
def test_fuzz_pairs(seed):
    rng = np.random.default_rng(9000 + seed)
    d = int(rng.integers(1, 4))
    m = int(rng.choice([3, 50, 127, 128, 129, 300, 777, 2100]))
    nf = int(rng.choice([1, 1, 2, 4]))
    pos, f = scattered(rng, d, m, nf=nf, nan_frac=float(rng.choice([0.0, 0.0, 0.2])))
    if rng.uniform() < 0.3:
        pos = np.round(pos / 5.0) * 5.0                     # lattice: repeated points and exact ties
    nb = int(rng.integers(1, 40))
    edges = np.sort(rng.uniform(0.0, 120.0, nb + 1))
    if rng.uniform() < 0.3:
        edges = np.linspace(0.0, float(rng.uniform(10.0, 150.0)), nb + 1)
    if rng.uniform() < 0.15:
        rng.shuffle(edges)                                  # non-monotone: generic kernel, bin-by-bin test
    est = str(rng.choice(["m", "c"]))
    if rng.uniform() < 0.5:
        g, c = gc.variogram_unstructured(f, edges, pos, est, "e")
        go, co = oracle.variogram_unstructured(f, edges, pos, est, "e", oracle.max_threads())
    else:
        nd = int(rng.choice([1, 2, 3, 6]))
        direction = unit(rng.normal(size=(nd, d)))
        tol = float(rng.choice([np.pi / 8.0, np.pi / 4.0, 0.2, 1.7]))
        bw = float(rng.choice([-1.0, 3.0, 20.0]))
        sep = bool(rng.integers(0, 2))
        g, c = gc.variogram_directional(f, edges, pos, direction, tol, bw, sep, est)
        go, co = oracle.variogram_directional(f, edges, pos, direction, tol, bw, sep, est, oracle.max_threads())
    same_counts(c, co)
    close(g, go)


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_structured(seed):
    rng = np.random.default_rng(9500 + seed)
    shape = (int(rng.integers(1, 200)), int(rng.integers(1, 150)))
    f = rng.normal(size=shape)
    est = str(rng.choice(["m", "c"]))
    close(gc.variogram_structured(f, est), oracle.variogram_structured(f, est, oracle.max_threads()))
    mask = rng.uniform(size=shape) < float(rng.uniform(0.0, 0.9))
    close(gc.variogram_ma_structured(f, mask, est), oracle.variogram_ma_structured(f, mask, est, oracle.max_threads()))
