"""Accuracy against EXACT arithmetic (mpmath, 50 digits) on small heavy-tailed cases: the GPU
result must be as close to the true sum as the reference's own f64 algorithm (the oracle) is,
up to a small factor -- i.e. the 1e-9 sigma parity budget is not silently spent on GPU error."""
import numpy as np
import pytest

import gstools_core as gc
import oracle

mp = pytest.importorskip("mpmath")
pytestmark = pytest.mark.gpu


def exact_summate(k, z1, z2, pos):
    mp.mp.dps = 50
    d, n = k.shape
    out = []
    for j in range(pos.shape[1]):
        s = mp.mpf(0)
        for i in range(n):
            ph = sum(mp.mpf(float(k[a, i])) * mp.mpf(float(pos[a, j])) for a in range(d))
            s += mp.mpf(float(z1[i])) * mp.cos(ph) + mp.mpf(float(z2[i])) * mp.sin(ph)
        out.append(s)
    return out


@pytest.mark.parametrize("scale,label", [(1.0, "moderate phases"), (1e3, "phases ~1e5"), (3e4, "phases ~1e6+")])
def test_error_vs_exact_is_comparable_to_the_references(scale, label):
    if gc.device_count() < 1:
        pytest.fail("no CUDA device")
    rng = np.random.default_rng(42)
    n, m = 40, 48
    k = rng.normal(size=(3, n)) / np.abs(rng.normal(size=n)) * scale / 10.0      # heavy-tailed like C2/C5
    z1, z2 = rng.normal(size=n), rng.normal(size=n)
    pos = rng.uniform(0, 100, size=(3, m))
    exact = exact_summate(k, z1, z2, pos)
    got = gc.summate(k, z1, z2, pos)
    ref = oracle.summate(k, z1, z2, pos)
    sigma = float(np.std(ref))
    e_gpu = max(abs(mp.mpf(float(g)) - e) for g, e in zip(got, exact)) / sigma
    e_ref = max(abs(mp.mpf(float(r)) - e) for r, e in zip(ref, exact)) / sigma
    print("%s: max|phase| %.3g  gpu-vs-exact %.3g sigma, oracle-vs-exact %.3g sigma" %
          (label, float(np.max(np.abs(k.T @ pos))), float(e_gpu), float(e_ref)))
    assert e_gpu <= 1e-9
    # the GPU formulation is not (much) less accurate than the reference's own f64 algorithm:
    # both are dominated by the rounding of the phase, |phase| * 2^-53 per term
    assert e_gpu <= 8 * e_ref + 1e-14
