"""GPU parity tests of the kriging kernels against the CPU oracle and the
reference's own known-answer vectors (src/krige.rs:127-245, 6 ulp there).

Tolerance: the GPU sums the C conditions in a different order (DMMA tiles, fixed butterflies), so
results agree to rounding of a length-C^2 dot product: max|d| <= 1e-12 * max|reference|."""
import json
import os

import numpy as np
import pytest

import gstools_core as gc
import oracle
from conftest import ROOT

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(autouse=True)
def _reset():
    if gc.device_count() < 1:
        pytest.fail("no CUDA device")
    gc.set_devices(None)
    yield


def close(got, ref):
    scale = max(1e-300, float(np.max(np.abs(ref))))
    return float(np.max(np.abs(got - ref))) / scale


def test_krige_kat():
    raw = json.load(open(os.path.join(ROOT, "tests", "golden", "krige_rs_kat.json")))
    k = {n: np.array(v, dtype=np.float64) for n, v in raw.items() if not n.startswith("_")}
    f, e = gc.calc_field_krige_and_variance(k["krig_mat"], k["krig_vecs"], k["cond"])
    assert np.max(np.abs(f - k["field"])) <= 1e-15 and np.max(np.abs(e - k["error"])) <= 1e-15
    f2 = gc.calc_field_krige(k["krig_mat"], k["krig_vecs"], k["cond"])
    assert np.max(np.abs(f2 - k["field"])) <= 1e-15


def _problem(seed, c, m):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(c, c))
    mat = a @ a.T / c + np.eye(c)                  # symmetric positive definite like a kriging matrix
    mat = np.linalg.inv(mat)
    vecs = rng.uniform(0, 1, size=(c, m)) * rng.uniform(0, 1, size=(c, 1))
    cond = rng.normal(size=c)
    return mat, vecs, cond


@pytest.mark.parametrize("c,m", [(1, 1), (3, 6), (31, 65), (32, 64), (33, 1000), (100, 4097), (500, 10000)])
def test_krige_random(c, m):
    mat, vecs, cond = _problem(c * 7 + m, c, m)
    rf, re = oracle.calc_field_krige_and_variance(mat, vecs, cond, oracle.max_threads())
    f, e = gc.calc_field_krige_and_variance(mat, vecs, cond)
    assert f.shape == (m,) and e.shape == (m,)
    assert close(f, rf) <= RTOL and close(e, re) <= RTOL
    f2 = gc.calc_field_krige(mat, vecs, cond)
    assert close(f2, rf) <= RTOL


def test_krige_strided_and_nonsymmetric():
    rng = np.random.default_rng(5)
    c, m = 70, 900
    mat = rng.normal(size=(c, c))                  # NOT symmetric: K[j,i] vs K[i,j] must not be mixed up
    vecs = rng.normal(size=(c, m))
    cond = rng.normal(size=c)
    rf, re = oracle.calc_field_krige_and_variance(mat, vecs, cond)
    f, e = gc.calc_field_krige_and_variance(np.asfortranarray(mat), np.asfortranarray(vecs), cond)
    assert close(f, rf) <= RTOL and close(e, re) <= RTOL
    big = np.zeros(2 * c); big[::2] = cond
    f, e = gc.calc_field_krige_and_variance(mat, vecs, big[::2])
    assert close(f, rf) <= RTOL and close(e, re) <= RTOL


def test_krige_device_resident_and_errors():
    torch = pytest.importorskip("torch")
    mat, vecs, cond = _problem(3, 200, 5000)
    rf, re = oracle.calc_field_krige_and_variance(mat, vecs, cond, oracle.max_threads())
    L = gc._load()
    dv = torch.from_numpy(vecs).cuda(); df = torch.empty(5000, dtype=torch.float64, device="cuda"); de = torch.empty_like(df)
    rc = L.gsf_krige(200, 5000, mat.ctypes.data, 200, 1, dv.data_ptr(), 5000, 1, cond.ctypes.data, 1,
                     df.data_ptr(), de.data_ptr(), 0)
    assert rc == 0
    assert close(df.cpu().numpy(), rf) <= RTOL and close(de.cpu().numpy(), re) <= RTOL
    with pytest.raises(ValueError):
        gc.calc_field_krige(mat, vecs[:100], cond)                       # src/krige.rs:31
    with pytest.raises(ValueError):
        gc.calc_field_krige_and_variance(mat[:, :50], vecs, cond)        # :30
    assert gc.calc_field_krige(mat, vecs[:, :0], cond).shape == (0,)
