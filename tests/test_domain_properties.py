"""Size-independent properties the domain offers, checked on the CPU oracle (not gpu) and, with the
SAME code, on the CUDA path (gpu) -- they hold at any size, so they also run at BASELINE's full
configuration sizes where the oracle cannot follow:

* the incompressible field is divergence free: sum_a k_a p_a(k) = 0 for the projector of
  src/field.rs:138-152, so every mode's contribution has zero divergence;
* the scalar field is an even/odd superposition:  z2 = 0 gives a field even in x, z1 = 0 an odd one;
* the Fourier field with modes on the lattice 2 pi m / L is L-periodic;
* a rotation of (z1, z2) by an angle a equals a phase shift: the field of (z1 cos a + z2 sin a,
  z2 cos a - z1 sin a) at phase phi equals the original at phi - a  =>  summing both at a = pi
  cancels exactly the field (z -> -z).
"""
import numpy as np
import pytest

import gstools_core as gc
import oracle


def _modes(seed, d, n, ell=10.0):
    rng = np.random.default_rng(seed)
    k = rng.normal(size=(d, n)) * np.sqrt(np.pi / 2.0) / ell
    return k, rng.normal(size=n), rng.normal(size=n), rng


def _divergence(fn, d, n, centers, h):
    """central-difference divergence of the vector field fn at the given centers, and its scale"""
    k, z1, z2, _ = _modes(31 + d, d, n)
    m = centers.shape[1]
    pts = np.empty((d, 2 * d * m))
    for a in range(d):
        for s, sign in enumerate((+1.0, -1.0)):
            blk = centers.copy()
            blk[a] += sign * h
            pts[:, (2 * a + s) * m:(2 * a + s + 1) * m] = blk
    u = np.asarray(fn(k, z1, z2, pts))
    div = np.zeros(m)
    grad_scale = 0.0
    for a in range(d):
        du = (u[a, (2 * a) * m:(2 * a + 1) * m] - u[a, (2 * a + 1) * m:(2 * a + 2) * m]) / (2.0 * h)
        div += du
        grad_scale = max(grad_scale, float(np.max(np.abs(du))))
    return float(np.max(np.abs(div))), grad_scale


def _check_divergence_free(fn, d, n, m):
    rng = np.random.default_rng(5)
    centers = rng.uniform(0.0, 100.0, size=(d, m))
    div, scale = _divergence(fn, d, n, centers, 1e-3)
    assert scale > 0.0
    assert div <= 1e-6 * scale, (div, scale)


def _check_parity_in_x(fn):
    k, z1, z2, rng = _modes(7, 3, 200)
    pos = rng.uniform(-50.0, 50.0, size=(3, 3000))
    zero = np.zeros_like(z1)
    even_p, even_m = fn(k, z1, zero, pos), fn(k, z1, zero, -pos)
    odd_p, odd_m = fn(k, zero, z2, pos), fn(k, zero, z2, -pos)
    s = float(np.std(even_p)) + float(np.std(odd_p))
    assert np.max(np.abs(even_p - even_m)) <= 1e-11 * s       # cos part: even
    assert np.max(np.abs(odd_p + odd_m)) <= 1e-11 * s         # sin part: odd
    both = fn(k, z1, z2, pos)
    assert np.max(np.abs(both - (even_p + odd_p))) <= 1e-11 * s   # linearity in (z1, z2)
    assert np.max(np.abs(fn(k, -z1, -z2, pos) + both)) <= 1e-11 * s   # rotation of (z1, z2) by pi


def _check_fourier_periodic(fn, n_side, m):
    period = 100.0
    mx, my = np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij")
    modes = 2.0 * np.pi / period * np.stack([mx.ravel(), my.ravel()]).astype(np.float64)
    rng = np.random.default_rng(11)
    n = modes.shape[1]
    sf, z1, z2 = rng.uniform(0.1, 1.0, n), rng.normal(size=n), rng.normal(size=n)
    pos = rng.uniform(0.0, period, size=(2, m))
    base = fn(sf, modes, z1, z2, pos)
    shifted = fn(sf, modes, z1, z2, pos + np.array([[period], [-2.0 * period]]))
    # phases grow by 2 pi m: the rounding of (x + L) k is what is left, ~1e-16 * |phase| per term
    assert np.max(np.abs(base - shifted)) <= 1e-9 * float(np.std(base))


# ---------------------------------------------------------------------------- oracle (CPU)
@pytest.mark.parametrize("d", [2, 3])
def test_oracle_incompressible_field_is_divergence_free(d):
    _check_divergence_free(oracle.summate_incompr, d, 150, 400)


def test_oracle_even_odd_and_linearity():
    _check_parity_in_x(oracle.summate)


def test_oracle_fourier_periodicity():
    _check_fourier_periodic(oracle.summate_fourier, 12, 2000)


# ---------------------------------------------------------------------------- CUDA path (GPU)
@pytest.fixture
def gpu():
    if gc.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    gc.set_variant(0, 0)
    gc.set_chunk_points(0)
    gc.set_devices(None)


@pytest.mark.gpu
@pytest.mark.parametrize("d,n,m", [(2, 150, 400), (3, 150, 400), (3, 1000, 170_000)])   # last: C3-sized call (1e6 points)
def test_gpu_incompressible_field_is_divergence_free(gpu, d, n, m):
    _check_divergence_free(gc.summate_incompr, d, n, m)


@pytest.mark.gpu
def test_gpu_even_odd_and_linearity(gpu):
    _check_parity_in_x(gc.summate)


@pytest.mark.gpu
def test_gpu_fourier_periodicity(gpu):
    _check_fourier_periodic(gc.summate_fourier, 12, 2000)
    _check_fourier_periodic(gc.summate_fourier, 100, 300_000)      # C4's 1e4 modes
