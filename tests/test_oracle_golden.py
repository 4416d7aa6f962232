"""Pin the CPU oracle against the reference's own known-answer tests (src/field.rs:334-430).

The reference asserts bitwise equality for summator and summator_fourier and <= 6 ulp (or
<= f64::EPSILON absolute, approx's ulps_eq) for summator_incompr; the oracle must meet the same.
"""
import os

import numpy as np
import pytest

import oracle
from conftest import ulp_diff


def test_summate_bitwise(kat):
    out = oracle.summate(kat["cov_samples"], kat["z_1"], kat["z_2"], kat["pos"])
    assert out.shape == (8,)
    assert np.array_equal(out, kat["summate"]), ulp_diff(out, kat["summate"])  # :363-381 assert_eq!


def test_summate_fourier_bitwise(kat):
    out = oracle.summate_fourier(kat["spectrum_factor"], kat["cov_samples"], kat["z_1"], kat["z_2"],
                                 kat["pos"])
    assert np.array_equal(out, kat["summate_fourier"]), ulp_diff(out, kat["summate_fourier"])  # :337-356


def _ulps_eq(a, b, max_ulps=6):
    # approx::ulps_eq for f64: |a-b| <= f64::EPSILON  OR  same sign and ulp distance <= max_ulps
    return (np.abs(a - b) <= np.finfo(np.float64).eps) | (
        (np.signbit(a) == np.signbit(b)) & (ulp_diff(a, b) <= max_ulps))


def test_summate_incompr_6ulp(kat):
    out = oracle.summate_incompr(kat["cov_samples"], kat["z_1"], kat["z_2"], kat["pos"])
    assert out.shape == (3, 8) and out.flags.f_contiguous  # src/field.rs:166-174
    assert _ulps_eq(out, kat["summate_incompr"]).all(), ulp_diff(out, kat["summate_incompr"])  # :388-429


@pytest.mark.parametrize("threads", [1, 2, 8])
def test_thread_count_invariance(kat, threads):
    # summator / fourier fold modes per point in index order: independent of the thread count
    rng = np.random.default_rng(7)
    k = rng.normal(size=(3, 257))
    z1, z2, sf = rng.normal(size=257), rng.normal(size=257), rng.uniform(0.1, 2, size=257)
    pos = rng.uniform(-50, 50, size=(3, 1031))
    a = oracle.summate(k, z1, z2, pos, 1)
    assert np.array_equal(a, oracle.summate(k, z1, z2, pos, threads))
    f = oracle.summate_fourier(sf, k, z1, z2, pos, 1)
    assert np.array_equal(f, oracle.summate_fourier(sf, k, z1, z2, pos, threads))
    # incompr: split accumulators change the order (reference :130-163) => only close
    i1 = oracle.summate_incompr(k, z1, z2, pos, 1)
    it = oracle.summate_incompr(k, z1, z2, pos, threads)
    assert np.allclose(i1, it, rtol=0, atol=1e-12)


def test_strided_views(kat):
    # the reference accepts arbitrary-stride views (src/lib.rs:43-46)
    k = np.asfortranarray(kat["cov_samples"])
    pos = np.ascontiguousarray(kat["pos"].T).T
    big = np.zeros(20)
    big[::2] = kat["z_1"]
    out = oracle.summate(k, big[::2], kat["z_2"], pos)
    assert np.array_equal(out, kat["summate"])


def test_edge_cases():
    k = np.zeros((2, 0)); z = np.zeros(0); pos = np.ones((2, 5))
    assert np.array_equal(oracle.summate(k, z, z, pos), np.zeros(5))       # fold identity :55
    with pytest.raises(ValueError):
        oracle.summate_incompr(k, z, z, pos)                               # unwrap on None :163
    with pytest.raises(ValueError):
        oracle.summate_incompr(np.ones((1, 3)), np.ones(3), np.ones(3), np.ones((1, 4)))  # :180
    out = oracle.summate_incompr(np.zeros((2, 1)), np.ones(1), np.ones(1), pos)
    assert np.isnan(out).all()                                             # k = 0 => 0/0, :138
    assert oracle.summate(np.ones((2, 3)), np.ones(3), np.ones(3), np.ones((2, 0))).shape == (0,)


# ------------------------------------------------------------------------------- krige
@pytest.fixture(scope="module")
def krige_kat():
    import json, os
    from conftest import ROOT
    raw = json.load(open(os.path.join(ROOT, "tests", "golden", "krige_rs_kat.json")))
    return {k: np.array(v, dtype=np.float64) for k, v in raw.items() if not k.startswith("_")}


def test_krige_golden_6ulp(krige_kat):
    f, e = oracle.calc_field_krige_and_variance(krige_kat["krig_mat"], krige_kat["krig_vecs"], krige_kat["cond"])
    assert _ulps_eq(f, krige_kat["field"]).all(), ulp_diff(f, krige_kat["field"])      # src/krige.rs:198-208
    assert _ulps_eq(e, krige_kat["error"]).all(), ulp_diff(e, krige_kat["error"])      # :210-220
    f2 = oracle.calc_field_krige(krige_kat["krig_mat"], krige_kat["krig_vecs"], krige_kat["cond"])
    assert np.array_equal(f, f2)                                                      # :228-244


# ---- variogram estimators: the reference's own tests, src/variogram.rs:577-842 -----------------

@pytest.fixture(scope="module")
def vkat():
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "variogram_rs_kat.json")) as fh:
        return json.load(fh)


def ulps(a, b):
    return int(ulp_diff(a, b).max())


def _vsetup(vkat):
    pos = np.stack([np.arange(0.0, 10.0, 1.0), np.arange(0.0, 10.0, 1.0)])       # :672-676
    return pos, np.array([vkat["unstruct_field"]]), np.linspace(0.0, 5.0, 4)     # :677-689


def test_variogram_structured_golden(vkat):
    f = np.array(vkat["struct_field"]).reshape(-1, 1)
    assert ulps(oracle.variogram_structured(f, "m"), np.array(vkat["struct_gamma"])) <= vkat["max_ulps"]
    no_mask = np.zeros((10, 1), dtype=bool)
    assert ulps(oracle.variogram_ma_structured(f, no_mask, "m"), np.array(vkat["struct_gamma"])) <= vkat["max_ulps"]
    mask2 = np.array(vkat["ma_struct_mask2"]).reshape(-1, 1)
    assert ulps(oracle.variogram_ma_structured(f, mask2, "m"), np.array(vkat["ma_struct_gamma2"])) <= vkat["max_ulps"]


def test_variogram_unstructured_golden(vkat):
    pos, f, edges = _vsetup(vkat)
    gamma, cnts = oracle.variogram_unstructured(f, edges, pos, "m", "e")
    assert ulps(gamma, np.array(vkat["unstruct_gamma"])) <= vkat["max_ulps"]
    assert cnts.tolist() == vkat["unstruct_counts"]


def test_variogram_directional_golden(vkat):
    pos, f, edges = _vsetup(vkat)
    direction = np.array([[0.0, np.pi], [0.0, 0.0]])                             # :821
    gamma, cnts = oracle.variogram_directional(f, edges, pos, direction, np.pi / 8.0, -1.0, False, "m")
    assert ulps(gamma, np.array(vkat["directional_gamma"])) <= vkat["max_ulps"]
    assert cnts.tolist() == vkat["directional_counts"]


def test_variogram_multi_field_property(vkat):
    # src/variogram.rs:711-816: the multi-field estimate is the mean of the single-field ones
    pos, f, edges = _vsetup(vkat)
    f2 = np.array([vkat["unstruct_field2"]])
    both = np.concatenate([f, f2])
    g1, _ = oracle.variogram_unstructured(f, edges, pos, "m", "e")
    g2, _ = oracle.variogram_unstructured(f2, edges, pos, "m", "e")
    gm, _ = oracle.variogram_unstructured(both, edges, pos, "m", "e")
    assert ulps(gm, 0.5 * (g1 + g2)) <= vkat["max_ulps"]
    direction = np.array([[0.0, np.pi], [0.0, 0.0]])
    d1, _ = oracle.variogram_directional(f, edges, pos, direction)
    d2, _ = oracle.variogram_directional(f2, edges, pos, direction)
    dm, _ = oracle.variogram_directional(both, edges, pos, direction)
    assert ulps(dm, 0.5 * (d1 + d2)) <= vkat["max_ulps"]
