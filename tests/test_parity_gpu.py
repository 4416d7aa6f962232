"""GPU parity tests: the CUDA path (through the C ABI / the gstools_core mirror module) against
the CPU oracle on identical inputs.

Acceptance (BASELINE.json north_star):
    max |gpu - oracle| <= 1e-9 * sigma,  sigma = population std of the oracle output
and the reference's own known-answer vectors (src/field.rs:346-355,371-380,396-427) to 1e-13 abs.
Bitwise equality is NOT expected: the kernels use FMA and a single-cos formulation.
"""
import numpy as np
import pytest

import gstools_core as gc
import oracle
from gstools_core import workloads

pytestmark = pytest.mark.gpu

TOL = 1e-9          # x sigma, the north_star tolerance
KAT_ABS = 1e-13


def _need_gpu():
    if gc.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")


@pytest.fixture(autouse=True)
def _reset():
    _need_gpu()
    gc.set_variant(0, 0)
    gc.set_chunk_points(0)
    gc.set_devices(None)
    gc.set_grid_detection(None)
    gc.set_poly_degree(0)
    yield
    gc.set_grid_detection(None)
    gc.set_poly_degree(0)


def rel_err(got, ref):
    sigma = float(np.std(ref))
    if not np.isfinite(sigma) or sigma == 0.0:
        sigma = 1.0
    return float(np.max(np.abs(got - ref))) / sigma


# ------------------------------------------------------------------------------- goldens
def test_kat_summate(kat):
    out = gc.summate(kat["cov_samples"], kat["z_1"], kat["z_2"], kat["pos"])
    assert out.shape == (8,) and out.dtype == np.float64
    assert np.max(np.abs(out - kat["summate"])) <= KAT_ABS


def test_kat_fourier(kat):
    out = gc.summate_fourier(kat["spectrum_factor"], kat["cov_samples"], kat["z_1"], kat["z_2"], kat["pos"])
    assert np.max(np.abs(out - kat["summate_fourier"])) <= KAT_ABS


def test_kat_incompr(kat):
    out = gc.summate_incompr(kat["cov_samples"], kat["z_1"], kat["z_2"], kat["pos"])
    assert out.shape == (3, 8) and out.flags.f_contiguous     # src/field.rs:166-174
    assert np.max(np.abs(out - kat["summate_incompr"])) <= KAT_ABS


# ------------------------------------------------------------------------------- random parity
def _rand(seed, d, n, m, heavy=False, span=50.0):
    rng = np.random.default_rng(seed)
    k = rng.normal(size=(d, n))
    if heavy:
        k = k / np.abs(rng.normal(size=n))                    # multivariate-t tails (C2/C5-like)
    z1, z2 = rng.normal(size=n), rng.normal(size=n)
    pos = rng.uniform(-span, span, size=(d, m))
    return k, z1, z2, pos


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 8, 9, 17])     # d > 8: gsf_sum_kernel_anyd (any dim, like the reference)
@pytest.mark.parametrize("n,m", [(1, 1), (10, 8), (257, 1031), (1000, 4099)])
def test_summate_random(d, n, m):
    k, z1, z2, pos = _rand(100 + d, d, n, m, heavy=(d == 3))
    ref = oracle.summate(k, z1, z2, pos)
    got = gc.summate(k, z1, z2, pos)
    assert got.shape == ref.shape
    assert rel_err(got, ref) <= TOL


def test_any_dim_strided_scaled_and_chunked():
    """dim > 8 through every host route: strided views, forced chunks, fused scale/offset"""
    d, n, m = 12, 130, 70001
    k, z1, z2, pos = _rand(77, d, n, m)
    ref = oracle.summate(k, z1, z2, pos)
    big = np.zeros((d, 2 * m))
    big[:, ::2] = pos
    assert rel_err(gc.summate(k, z1, z2, big[:, ::2]), ref) <= TOL          # strided pos
    gc.set_chunk_points(8192)
    try:
        got = gc.summate(k, z1, z2, pos)
    finally:
        gc.set_chunk_points(0)
    assert rel_err(got, ref) <= TOL
    got = gc.summate_scaled(k, z1, z2, pos, scale=0.25, offset=3.0)
    assert rel_err(got, 0.25 * ref + 3.0) <= TOL


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("n,m", [(1, 1), (10, 8), (257, 1031), (1000, 4099)])
def test_incompr_random(d, n, m):
    k, z1, z2, pos = _rand(200 + d, d, n, m)
    ref = oracle.summate_incompr(k, z1, z2, pos)
    got = gc.summate_incompr(k, z1, z2, pos)
    assert got.shape == (d, m) and got.flags.f_contiguous
    assert rel_err(got, ref) <= TOL


@pytest.mark.parametrize("d", [1, 2, 3, 11])
def test_fourier_random(d):
    k, z1, z2, pos = _rand(300 + d, d, 513, 2050)
    sf = np.random.default_rng(9).normal(size=513)            # the reference's KAT uses signed factors
    ref = oracle.summate_fourier(sf, k, z1, z2, pos)
    got = gc.summate_fourier(sf, k, z1, z2, pos)
    assert rel_err(got, ref) <= TOL


# ------------------------------------------------------------------------------- kernel variants
@pytest.mark.parametrize("P,L", [(4, 1), (3, 1), (2, 1), (1, 1), (2, 2), (2, 4), (2, 8), (2, 16), (2, 32),
                                 (1, 2), (1, 4), (1, 8), (1, 16), (1, 32)])
def test_all_variants_agree(P, L):
    k, z1, z2, pos = _rand(7, 3, 700, 3001, heavy=True)
    ref = oracle.summate(k, z1, z2, pos)
    refi = oracle.summate_incompr(k, z1, z2, pos)
    gc.set_variant(P, L)
    got = gc.summate(k, z1, z2, pos)
    assert gc.last_stats()["points_per_thread"] == P and gc.last_stats()["lanes_per_point"] == L
    assert rel_err(got, ref) <= TOL
    goti = gc.summate_incompr(k, z1, z2, pos)
    assert rel_err(goti, refi) <= TOL
    k2, z1, z2, pos2 = _rand(8, 2, 300, 1500)
    gc.set_variant(P, L)
    assert rel_err(gc.summate(k2, z1, z2, pos2), oracle.summate(k2, z1, z2, pos2)) <= TOL


def test_deterministic_repeat():
    k, z1, z2, pos = _rand(11, 3, 999, 70001, heavy=True)
    a = gc.summate(k, z1, z2, pos)
    b = gc.summate(k, z1, z2, pos)
    assert np.array_equal(a, b)                               # fixed-order sums: bit-reproducible


# ------------------------------------------------------------------------------- layouts / memory kinds
def test_strided_views():
    k, z1, z2, pos = _rand(21, 3, 64, 5000)
    ref = oracle.summate(k, z1, z2, pos)
    kf = np.asfortranarray(k)
    posf = np.asfortranarray(pos)                             # AoS positions
    zz = np.zeros(2 * 64); zz[::2] = z1
    big = np.zeros((3, 10000)); big[:, ::2] = pos
    assert rel_err(gc.summate(kf, zz[::2], z2, posf), ref) <= TOL
    assert rel_err(gc.summate(k, z1, z2, big[:, ::2]), ref) <= TOL
    assert rel_err(gc.summate(k[:, ::-1], z1[::-1], z2[::-1], pos), ref) <= 1e-12 * 64  # other mode order


def test_chunk_boundaries_and_pipeline():
    k, z1, z2, pos = _rand(31, 3, 50, 300_000)
    ref = oracle.summate(k, z1, z2, pos, oracle.max_threads())
    refi = oracle.summate_incompr(k, z1, z2, pos)
    for chunk in (1024, 33 * 1024, 1 << 20):
        gc.set_chunk_points(chunk)
        got = gc.summate(k, z1, z2, pos)
        st = gc.last_stats()
        assert st["n_chunks"] == -(-300_000 // min(chunk, 300_000))   # chunk is already a multiple of 1024
        assert rel_err(got, ref) <= TOL
        assert rel_err(gc.summate_incompr(k, z1, z2, pos), refi) <= TOL


def test_pinned_and_device_memory():
    torch = pytest.importorskip("torch")
    k, z1, z2, pos = _rand(41, 3, 128, 200_000)
    ref = oracle.summate(k, z1, z2, pos, oracle.max_threads())
    refi = oracle.summate_incompr(k, z1, z2, pos)
    # pinned host input
    tp = torch.from_numpy(pos).pin_memory()
    got = gc.summate(k, z1, z2, tp.numpy())
    assert gc.last_stats()["pos_memory"] == 1
    assert rel_err(got, ref) <= TOL
    # device-resident, stream ordered
    dpos = torch.from_numpy(pos).cuda()
    dout = torch.empty(pos.shape[1], dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        gc.summate_device(k, z1, z2, dpos, dout, stream=s.cuda_stream)
    s.synchronize()
    assert rel_err(dout.cpu().numpy(), ref) <= TOL
    # device modes too, incompr with both output layouts
    dk, dz1, dz2 = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (k, z1, z2))
    douti = torch.empty((3, pos.shape[1]), dtype=torch.float64, device="cuda")
    gc.summate_incompr_device(dk, dz1, dz2, dpos, douti, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rel_err(douti.cpu().numpy(), refi) <= TOL
    doutf = torch.empty((pos.shape[1], 3), dtype=torch.float64, device="cuda").t()   # F-ordered (3, M)
    gc.summate_incompr_device(dk, dz1, dz2, dpos, doutf, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rel_err(doutf.cpu().numpy(), refi) <= TOL


# ------------------------------------------------------------------------------- edge cases
def test_edge_cases():
    pos = np.ones((2, 5))
    z0 = np.zeros(0)
    assert np.array_equal(gc.summate(np.zeros((2, 0)), z0, z0, pos), np.zeros(5))       # fold identity
    assert np.array_equal(gc.summate_fourier(z0, np.zeros((2, 0)), z0, z0, pos), np.zeros(5))
    assert gc.summate(np.ones((2, 3)), np.ones(3), np.ones(3), np.ones((2, 0))).shape == (0,)
    assert gc.summate_incompr(np.ones((2, 3)), np.ones(3), np.ones(3), np.ones((2, 0))).shape == (2, 0)
    out = gc.summate_incompr(np.zeros((2, 1)), np.ones(1), np.ones(1), pos)
    assert np.isnan(out).all()                                                          # k = 0 -> 0/0
    # non-finite inputs propagate like sin/cos(inf) = NaN
    k = np.array([[1.0, np.inf], [0.5, 1.0]])
    got = gc.summate(k, np.ones(2), np.ones(2), pos)
    assert np.isnan(got).all()
    got = gc.summate(np.ones((2, 2)), np.array([1.0, np.nan]), np.ones(2), pos)
    assert np.isnan(got).all()
    # zero amplitude mode contributes nothing
    k = np.array([[1.0, 2.0], [0.5, 1.0]])
    a = gc.summate(k, np.array([1.0, 0.0]), np.array([2.0, 0.0]), pos)
    b = gc.summate(k[:, :1], np.array([1.0]), np.array([2.0]), pos)
    assert np.allclose(a, b, rtol=0, atol=1e-15)


def test_large_phases():
    # |phase| up to ~1e6 rad as in C5 (SURVEY.md section 7 "heavy-tailed spectra")
    rng = np.random.default_rng(5)
    k = rng.normal(size=(3, 400)) * 1e4
    z1, z2 = rng.normal(size=400), rng.normal(size=400)
    pos = rng.uniform(0, 100, size=(3, 2000))
    ref = oracle.summate(k, z1, z2, pos)
    assert rel_err(gc.summate(k, z1, z2, pos), ref) <= TOL


# ------------------------------------------------------------------------------- BASELINE configs
@pytest.mark.parametrize("detect", [False, True])
@pytest.mark.parametrize("cfg,scale", [("c1", 1.0), ("c2", 0.02), ("c3", 0.02), ("c4", 0.0005), ("c5", 0.00005)])
def test_baseline_configs_scaled(cfg, scale, detect):
    w = workloads.make(cfg, scale)
    ref = getattr(oracle, w["kind"])(*w["args"], oracle.max_threads())
    gc.set_grid_detection(detect)
    got = getattr(gc, w["kind"])(*w["args"])
    st = gc.last_stats()
    e = rel_err(got, ref)
    print(cfg, "m=%d n=%d grid_path=%d deg=%d max|d|/sigma=%.3g" % (w["m"], w["n"], st["grid_path"], st["poly_degree"], e))
    assert e <= TOL
    if not detect:
        assert st["grid_path"] == 0


@pytest.mark.parametrize("deg", [5, 6])
def test_both_polynomial_degrees_meet_the_contract(deg):
    """The two shipped degrees of the cosine polynomial on every kind / dim, and the 1e-11 sigma
    bar the throughput degree was chosen by (two orders inside the 1e-9 contract)."""
    gc.set_poly_degree(deg)
    for d, kind in ((2, "summate"), (3, "summate"), (2, "summate_incompr"), (3, "summate_incompr"), (3, "summate_fourier")):
        k, z1, z2, pos = _rand(500 + d, d, 1000, 40_000, heavy=(d == 3))
        args = (k, z1, z2, pos) if kind != "summate_fourier" else (np.abs(z1) + 0.1, k, z1, z2, pos)
        ref = getattr(oracle, kind)(*args, oracle.max_threads())
        got = getattr(gc, kind)(*args)
        st = gc.last_stats()
        assert st["poly_degree"] == deg and st["fp64_slots"] == d + 5 + deg + (d if kind == "summate_incompr" else 1)
        e = rel_err(got, ref)
        print("deg %d %s d=%d: max|d|/sigma=%.3g" % (deg, kind, d, e))
        assert e <= 1e-11
    # the degree is ONE decision per call: chunking must not change a bit
    k, z1, z2, pos = _rand(9, 3, 300, 200_000, heavy=True)
    one = gc.summate(k, z1, z2, pos)
    gc.set_chunk_points(16 * 1024)
    assert np.array_equal(one, gc.summate(k, z1, z2, pos))


def test_degree_rule():
    """Automatic rule: degree 6 below 2^27 point*modes (KATs to 1e-13), degree 5 from there on."""
    k, z1, z2, pos = _rand(10, 3, 1000, 100_000)
    gc.summate(k, z1, z2, pos)
    assert gc.last_stats()["poly_degree"] == 6
    k, z1, z2, pos = _rand(10, 3, 1000, 140_000)
    gc.summate(k, z1, z2, pos)
    assert gc.last_stats()["poly_degree"] == 5
    gc.set_variant(1, 4)                                   # lanes split the modes: only built at degree 6
    got = gc.summate(k, z1, z2, pos)
    assert gc.last_stats()["poly_degree"] == 6 and gc.last_stats()["lanes_per_point"] == 4
    assert rel_err(got, oracle.summate(k, z1, z2, pos, oracle.max_threads())) <= TOL


# ------------------------------------------------------------------------------- full-size configs
# the oracle runs on >= 2^17 evenly strided points + the first / last 1024
# points of every GPU shard and of every pipeline chunk.  Every test states (and asserts) which
# device path produced the result: `grid_path` 0 = the general point x mode kernel (what bench.py
# headlines), 1 = the auto-detected structured-grid GEMM.
def _acceptance_idx(m, n_strided=1 << 17, edge=1024):
    idx = [np.arange(0, m, max(1, m // n_strided)), np.arange(min(m, edge)), np.arange(max(0, m - edge), m)]
    bounds = set()
    run = 0
    for c in gc.chunk_schedule(m):                     # the pipeline's own chunk boundaries
        run += c
        bounds.add(run)
    for g in (2, 4, 8):                                # shard boundaries of 2 / 4 / 8 devices
        for r in range(1, g):
            bounds.add(gc.shard_bounds(m, g, r)[0])
    for e in sorted(bounds):
        if 0 < e < m:
            idx.append(np.arange(max(0, e - edge), min(m, e + edge)))
    return np.unique(np.concatenate(idx))


def _full_size_check(cfg, detect, expect_grid_path):
    w = workloads.make(cfg)
    fn, ofn = getattr(gc, w["kind"]), getattr(oracle, w["kind"])
    gc.set_grid_detection(detect)
    try:
        got = fn(*w["args"])
        st = gc.last_stats()
    finally:
        gc.set_grid_detection(None)
    assert st["grid_path"] == expect_grid_path, st
    if expect_grid_path == 0:
        assert st["poly_degree"] == 5                  # throughput-bound problems run the degree-5 kernels
    idx = _acceptance_idx(w["m"])
    assert idx.size >= (1 << 17)
    ref = ofn(*workloads.subset_points(w, idx)["args"], oracle.max_threads())
    sub = got[:, idx] if got.ndim == 2 else got[idx]
    e = rel_err(sub, ref)
    print("%s full size, grid_path=%d: %d oracle points, max|d|/sigma=%.3g, total_ms=%.2f chunks=%d deg=%d"
          % (cfg, st["grid_path"], idx.size, e, st["total_ms"], st["n_chunks"], st["poly_degree"]))
    assert e <= TOL
    assert np.isfinite(got).all()
    return w, got, idx, ref


@pytest.mark.parametrize("detect,path", [(False, 0), (True, 1)])
def test_c2_full_size(detect, path):
    """C2 at full size on both device paths + linearity in (z1, z2)."""
    w, got, idx, ref = _full_size_check("c2", detect, path)
    k, z1, z2, pos = w["args"]
    # linearity: f(2*z1 + a, 2*z2 + b) = 2 f(z1, z2) + f(a, b)
    gc.set_grid_detection(detect)
    try:
        rng = np.random.default_rng(0)
        a, b = rng.normal(size=z1.size), rng.normal(size=z1.size)
        lhs = gc.summate(k, 2 * z1 + a, 2 * z2 + b, pos)
        rhs = 2 * got + gc.summate(k, a, b, pos)
    finally:
        gc.set_grid_detection(None)
    assert np.max(np.abs(lhs - rhs)) <= TOL * np.std(got)


@pytest.mark.parametrize("detect,path", [(False, 0), (True, 1)])
def test_c3_full_size(detect, path):
    w, got, idx, ref = _full_size_check("c3", detect, path)
    assert got.shape == (3, w["m"]) and got.flags.f_contiguous
    # size-independent property: a repeated call is bit-identical
    gc.set_grid_detection(detect)
    try:
        assert np.array_equal(got, gc.summate_incompr(*w["args"]))
    finally:
        gc.set_grid_detection(None)


@pytest.mark.parametrize("detect,path", [(False, 0), (True, 1)])
def test_c4_full_size(detect, path):
    w, got, idx, ref = _full_size_check("c4", detect, path)
    # periodicity of the Fourier method: the field repeats with period L = 100 along each axis
    sf, modes, z1, z2, pos = w["args"]
    sub = idx[:: max(1, idx.size // 8192)]
    shifted = np.ascontiguousarray(pos[:, sub] + np.array([[100.0], [200.0]]))
    again = gc.summate_fourier(sf, modes, z1, z2, shifted)
    assert np.max(np.abs(again - got[sub])) <= 1e-9 * np.std(got)


@pytest.mark.parametrize("detect,path", [(False, 0), (True, 1)])
def test_c5_full_size(detect, path):
    """C5 (1e4 modes x 1e8 points): the general kernel at 40 ring wraps per tile and ~100 pipeline
    chunks -- the configuration bench.py headlines -- and the grid path the default API takes."""
    w, got, idx, ref = _full_size_check("c5", detect, path)
    assert w["m"] == 10 ** 8 and w["n"] == 10 ** 4


def test_dfma_peak_runs():
    rate, ms = gc.dfma_peak(0, 20.0)
    assert rate > 1e12 and ms > 0


def test_multi_device_sharding_matches_single():
    if gc.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    k, z1, z2, pos = _rand(51, 3, 200, 400_000, heavy=True)
    gc.set_devices([0])
    one = gc.summate(k, z1, z2, pos)
    onei = gc.summate_incompr(k, z1, z2, pos)
    gc.set_devices(list(range(gc.device_count())))
    many = gc.summate(k, z1, z2, pos)
    assert gc.last_stats()["n_devices"] == min(gc.device_count(), 400_000 // 65536)
    assert np.array_equal(one, many)              # per-point mode order does not depend on the shard
    assert np.array_equal(onei, gc.summate_incompr(k, z1, z2, pos))


def test_peer_sharding_of_device_resident_data():
    """pos/out resident on GPU 0, points sharded over all GPUs through NVLink peer
    mappings (kernels on GPU g read positions from / write results to GPU 0 directly)."""
    torch = pytest.importorskip("torch")
    if gc.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    k, z1, z2, pos = _rand(61, 3, 300, 600_000, heavy=True)
    dpos = torch.from_numpy(pos).cuda(0)
    dk, dz1, dz2 = (torch.from_numpy(x).cuda(0) for x in (k, z1, z2))
    one = torch.empty(pos.shape[1], dtype=torch.float64, device="cuda:0")
    many = torch.empty_like(one)
    torch.cuda.synchronize()
    gc.set_devices([0])
    gc.summate_device(k, z1, z2, dpos, one, sync=True)
    assert gc.last_stats()["n_devices"] == 1
    gc.set_devices(list(range(gc.device_count())))
    gc.summate_device(k, z1, z2, dpos, many, sync=True)
    assert gc.last_stats()["n_devices"] == min(gc.device_count(), 600_000 // 65536)
    assert torch.equal(one, many)
    ref = oracle.summate(k, z1, z2, pos, oracle.max_threads())
    assert rel_err(many.cpu().numpy(), ref) <= TOL
    # modes resident on the owner too, incompressible field with F-ordered output
    outi = torch.empty((pos.shape[1], 3), dtype=torch.float64, device="cuda:0").t()
    gc.summate_incompr_device(dk, dz1, dz2, dpos, outi, sync=True)
    assert gc.last_stats()["n_devices"] > 1
    assert rel_err(outi.cpu().numpy(), oracle.summate_incompr(k, z1, z2, pos, 1)) <= TOL


def test_mode_record_cache_invalidation():
    """Identical modes skip the upload/pre-pass (records cached on the device); any change of the
    mode arrays -- even in place, same pointers -- must be noticed."""
    gc.shutdown()                                             # drop any cached records
    k, z1, z2, pos = _rand(71, 3, 200, 60000)                 # beyond the one-launch small path (tests/test_small_gpu.py)
    a1 = gc.summate(k, z1, z2, pos)
    n1 = gc.last_stats()["kernel_launches"]
    a2 = gc.summate(k, z1, z2, pos)
    n2 = gc.last_stats()["kernel_launches"]
    assert np.array_equal(a1, a2) and n2 == n1 - 1            # second call: no gsf_prep_modes launch
    z1[17] += 0.5                                             # in-place edit, same buffer
    b = gc.summate(k, z1, z2, pos)
    assert gc.last_stats()["kernel_launches"] == n1
    assert rel_err(b, oracle.summate(k, z1, z2, pos)) <= TOL and not np.array_equal(a1, b)
    # same raw modes, other kind / scale => other records
    c = gc.summate_incompr(k, z1, z2, pos)
    assert rel_err(c, oracle.summate_incompr(k, z1, z2, pos)) <= TOL
    d = gc.summate_scaled(k, z1, z2, pos, scale=3.0)
    assert np.max(np.abs(d - 3.0 * b)) <= 1e-12 * np.std(b)
    e = gc.summate(k, z1, z2, pos)
    assert np.array_equal(e, b)


def test_auto_pin_of_repeated_position_arrays():
    """Opt-in (set_auto_pin / GSF_AUTO_PIN_MB): a position array passed for the second time is page-locked
    and read in place from then on; same results, in-place edits are seen, switching off releases it."""
    k, z1, z2, pos = _rand(81, 3, 300, 200_000)
    ref = gc.summate(k, z1, z2, pos)
    assert gc.last_stats()["pos_memory"] == 0
    gc.set_auto_pin(64)
    try:
        a = gc.summate(k, z1, z2, pos)                       # first sighting: still staged
        assert gc.last_stats()["pos_memory"] == 0
        b = gc.summate(k, z1, z2, pos)                       # second: registered, read in place
        assert gc.last_stats()["pos_memory"] == 1
        assert np.array_equal(a, ref) and np.array_equal(b, ref)
        pos[:, 1000:2000] += 0.25                            # in-place edit of a pinned array
        c = gc.summate(k, z1, z2, pos)
        assert gc.last_stats()["pos_memory"] == 1
        assert rel_err(c, oracle.summate(k, z1, z2, pos, oracle.max_threads())) <= TOL
        other = pos.copy()                                   # budget: 64 MB holds both (4.8 MB each)
        gc.summate(k, z1, z2, other); gc.summate(k, z1, z2, other)
        assert gc.last_stats()["pos_memory"] == 1
        gc.set_auto_pin(6)                                   # shrink: the least recently used one goes
        gc.summate(k, z1, z2, pos)
        assert gc.last_stats()["pos_memory"] == 0
    finally:
        gc.set_auto_pin(0)
    d = gc.summate(k, z1, z2, pos)
    assert gc.last_stats()["pos_memory"] == 0 and np.array_equal(c, d)


@pytest.mark.parametrize("m", [3000, 250_000])            # one-launch small path / chunk pipeline
def test_reversed_position_views(m):
    """ArrayView strides may be negative (numpy [::-1] views, src/lib.rs:43-46 binds whatever numpy passes)."""
    k, z1, z2, pos = _rand(91, 3, 120, m)
    ref = oracle.summate(k, z1, z2, pos, oracle.max_threads())
    rev_pts = pos[:, ::-1]                                  # points in reverse order: pos_s1 < 0
    assert rel_err(gc.summate(k, z1, z2, rev_pts), ref[::-1]) <= TOL
    rev_dim = pos[::-1]                                     # coordinates swapped: pos_s0 < 0
    ref_d = oracle.summate(k, z1, z2, np.ascontiguousarray(rev_dim), oracle.max_threads())
    assert rel_err(gc.summate(k, z1, z2, rev_dim), ref_d) <= TOL
    refi = oracle.summate_incompr(k, z1, z2, pos, oracle.max_threads())
    assert rel_err(gc.summate_incompr(k, z1, z2, rev_pts), refi[:, ::-1]) <= TOL
    every_other = pos[:, ::-2]
    ref_e = oracle.summate(k, z1, z2, np.ascontiguousarray(every_other), oracle.max_threads())
    assert rel_err(gc.summate(k, z1, z2, every_other), ref_e) <= TOL
