"""One-launch path for small host-resident problems (gsf_small_kernel; C1-sized calls): the raw
modes travel in the kernel parameters and every CTA builds the mode records itself.  Checked against
the CPU oracle (1e-9 sigma) AND bit-for-bit against the general two-launch
path (gsf_prep_modes + gsf_sum_kernel<D,NC,1,L,6>) with the same lane count."""
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
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    gc.set_variant(0, 0)
    gc.set_chunk_points(0)
    gc.set_devices(None)
    gc.set_grid_detection(None)
    gc.set_poly_degree(0)
    yield
    gc.set_variant(0, 0)


def _modes(seed, d, n, heavy=True):
    rng = np.random.default_rng(seed)
    k = rng.normal(size=(d, n))
    if heavy:
        k = k / np.abs(rng.normal(size=n))
    return k, rng.normal(size=n), rng.normal(size=n), rng.uniform(0.1, 1.0, size=n)


def _check(kind, d, n, m, seed=0, pos=None):
    k, z1, z2, sf = _modes(seed, d, n)
    if pos is None:
        pos = np.random.default_rng(seed + 1).uniform(-40.0, 40.0, size=(d, m))
    args = (sf, k, z1, z2, pos) if kind == "summate_fourier" else (k, z1, z2, pos)
    got = getattr(gc, kind)(*args)
    st = gc.last_stats()
    assert st["kernel_launches"] == 1 and st["n_chunks"] == 1 and st["points_per_thread"] == 1, st
    assert st["poly_degree"] == 6 and st["grid_path"] == 0
    ref = getattr(oracle, kind)(*args, oracle.max_threads())
    sigma = float(np.std(ref)) or 1.0
    assert got.shape == ref.shape
    assert float(np.max(np.abs(got - ref))) <= TOL * sigma
    # the general path with the same (P, L): prep kernel + summation kernel, bit-identical
    gc.set_variant(1, st["lanes_per_point"])
    two = getattr(gc, kind)(*args)
    st2 = gc.last_stats()
    gc.set_variant(0, 0)
    assert st2["kernel_launches"] == 2 and st2["lanes_per_point"] == st["lanes_per_point"], st2
    assert np.array_equal(got, two)
    return st


@pytest.mark.parametrize("kind,d", [("summate", 1), ("summate", 2), ("summate", 3), ("summate_incompr", 2),
                                    ("summate_incompr", 3), ("summate_fourier", 2), ("summate_fourier", 3)])
@pytest.mark.parametrize("n,m", [(1, 1), (7, 130), (100, 10000), (256, 3001), (33, 16384), (150, 33000)])
def test_small_path_matches_oracle_and_general_kernel(kind, d, n, m):
    if d * m * 8 > 800 * 1024:
        pytest.skip("beyond the small-path size")
    _check(kind, d, n, m, seed=n + m)


def test_c1_takes_the_one_launch_path():
    gc.shutdown()                                             # forget every cached / recently seen mode set
    w = workloads.make("c1")
    got = gc.summate(*w["args"])
    st = gc.last_stats()
    assert st["kernel_launches"] == 1 and st["lanes_per_point"] == 4, st
    ref = oracle.summate(*w["args"], oracle.max_threads())
    assert float(np.max(np.abs(got - ref))) <= TOL * float(np.std(ref))


def test_lane_counts_cover_1_to_32():
    seen = set()
    for m in (50000, 25000, 12000, 5000, 2000):
        st = _check("summate", 1, 256, m, seed=m)
        seen.add(st["lanes_per_point"])
    assert seen == {1, 2, 4, 8, 16}, seen


def test_strided_inputs_scale_offset_and_f_order_output():
    d, n, m = 3, 90, 2500
    k, z1, z2, _ = _modes(5, d, n)
    kb = np.zeros((d, 2 * n)); kb[:, ::2] = k
    zb = np.zeros((2, 3 * n)); zb[0, ::3] = z1; zb[1, 1::3] = z2
    pos = np.asfortranarray(np.random.default_rng(6).uniform(0, 30, size=(d, m)))     # pos_s1 = d
    a = (kb[:, ::2], zb[0, ::3], zb[1, 1::3], pos)
    got = gc.summate_incompr(*a)
    assert gc.last_stats()["kernel_launches"] == 1
    assert got.shape == (d, m) and got.flags.f_contiguous
    ref = oracle.summate_incompr(k, z1, z2, np.ascontiguousarray(pos), oracle.max_threads())
    assert float(np.max(np.abs(got - ref))) <= TOL * float(np.std(ref))
    s = gc.summate_scaled(a[0], a[1], a[2], pos, scale=0.25, offset=3.0)
    assert gc.last_stats()["kernel_launches"] == 1
    ref_s = 0.25 * oracle.summate(k, z1, z2, np.ascontiguousarray(pos), oracle.max_threads()) + 3.0
    assert float(np.max(np.abs(s - ref_s))) <= TOL * float(np.std(ref_s))


def test_edge_values_propagate_like_the_general_path():
    d, n, m = 2, 20, 500
    k, z1, z2, _ = _modes(9, d, n)
    k[:, 3] = 0.0                      # zero mode: NaN projector for incompr (src/field.rs:138)
    pos = np.random.default_rng(10).uniform(0, 10, size=(d, m))
    pos[0, 7] = np.inf
    pos[1, 11] = np.nan
    for kind in ("summate", "summate_incompr"):
        got = getattr(gc, kind)(k, z1, z2, pos)
        assert gc.last_stats()["kernel_launches"] == 1
        gc.set_variant(1, gc.last_stats()["lanes_per_point"])
        two = getattr(gc, kind)(k, z1, z2, pos)
        gc.set_variant(0, 0)
        assert np.array_equal(got, two, equal_nan=True)
    out = gc.summate(k, z1, z2, pos)
    assert np.isnan(out[7]) and np.isnan(out[11]) and np.isfinite(np.delete(out, [7, 11])).all()


def test_fresh_modes_every_call_and_repeat_identity():
    w = workloads.make("c1")
    k, z1, z2, pos = w["args"]
    first = gc.summate(k, z1, z2, pos)
    for seed in range(5):
        zz = np.random.default_rng(seed).normal(size=(2, z1.size))
        got = gc.summate(k, zz[0], zz[1], pos)
        ref = oracle.summate(k, zz[0], zz[1], pos, oracle.max_threads())
        assert float(np.max(np.abs(got - ref))) <= TOL * float(np.std(ref))
    assert np.array_equal(first, gc.summate(k, z1, z2, pos))


def test_repeated_modes_are_promoted_to_cached_records():
    """Fresh modes (ensembles) stay on the one-launch kernel; the same modes for the second time in a
    row go through gsf_prep_modes once and are served from the cached records afterwards.  All three
    paths give bit-identical fields."""
    gc.shutdown()
    d, n, m = 2, 100, 10000
    k, z1, z2, _ = _modes(21, d, n)
    rng = np.random.default_rng(22)
    pos = [rng.uniform(0, 50, size=(d, m)) for _ in range(3)]
    raw_bytes = (d + 2) * n * 8
    a = gc.summate(k, z1, z2, pos[0]); st_a = gc.last_stats()
    b = gc.summate(k, z1, z2, pos[1]); st_b = gc.last_stats()
    c = gc.summate(k, z1, z2, pos[2]); st_c = gc.last_stats()
    assert st_a["kernel_launches"] == 1 and st_a["h2d_bytes"] == d * m * 8 + raw_bytes      # one-launch kernel
    assert st_b["kernel_launches"] == 2                                                     # upload + pre-pass + kernel
    assert st_c["kernel_launches"] == 1 and st_c["h2d_bytes"] == d * m * 8                  # cached records
    for got, p in zip((a, b, c), pos):
        ref = oracle.summate(k, z1, z2, p, oracle.max_threads())
        assert float(np.max(np.abs(got - ref))) <= TOL * float(np.std(ref))
    assert np.array_equal(a, gc.summate(k, z1, z2, pos[0]))                                 # cached path == one-launch path
    z1b = z1.copy(); z1b[3] += 1.0
    e = gc.summate(k, z1b, z2, pos[0])                                                      # new modes: one-launch kernel again
    assert gc.last_stats()["kernel_launches"] == 1 and gc.last_stats()["h2d_bytes"] == d * m * 8 + raw_bytes
    assert not np.array_equal(a, e)


def test_concurrent_python_threads():
    """The native binding releases the GIL around the C call; the library serialises calls on its
    context mutex.  Four threads mixing C1-sized, mid-sized and incompressible calls must all get the
    single-threaded answers."""
    import threading
    cases = []
    for i, (kind, d, n, m) in enumerate([("summate", 2, 100, 10000), ("summate", 3, 300, 60000),
                                         ("summate_incompr", 3, 64, 4000), ("summate_fourier", 2, 200, 9000)]):
        k, z1, z2, sf = _modes(100 + i, d, n)
        pos = np.random.default_rng(200 + i).uniform(0, 30, size=(d, m))
        args = (sf, k, z1, z2, pos) if kind == "summate_fourier" else (k, z1, z2, pos)
        cases.append((kind, args, getattr(gc, kind)(*args)))
    errors = []

    def worker(idx):
        try:
            for it in range(25):
                kind, args, want = cases[(idx + it) % len(cases)]
                got = getattr(gc, kind)(*args)
                if not np.array_equal(got, want):
                    errors.append("thread %d iteration %d: %s differs" % (idx, it, kind))
        except Exception as exc:                      # noqa: BLE001
            errors.append("thread %d: %r" % (idx, exc))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
