"""ctypes loader for the CPU oracle (oracle/field_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs.  The product package (gstools-core_b200/) never imports this.

Functions mirror the reference's Python names (src/lib.rs:34,51,68) so parity tests read like the
reference's own: summate / summate_incompr / summate_fourier, numpy in, numpy out.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgsf_oracle.so")
_lib = None

_i64 = ctypes.c_int64
_dp = ctypes.c_void_p


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, n) for n in ("field_oracle.c", "variogram_oracle.c", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(map(os.path.getmtime, srcs)):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        common = [ctypes.c_int, _i64, _i64]
        arr2 = [_dp, _i64, _i64]
        arr1 = [_dp, _i64]
        L.gso_summator.argtypes = common + arr2 + arr1 + arr1 + arr2 + [_dp, ctypes.c_int]
        L.gso_summator_incompr.argtypes = common + arr2 + arr1 + arr1 + arr2 + [_dp, ctypes.c_int]
        L.gso_summator_fourier.argtypes = common + arr1 + arr2 + arr1 + arr1 + arr2 + [_dp, ctypes.c_int]
        L.gso_krige.argtypes = [_i64, _i64] + arr2 + arr2 + arr1 + [_dp, _dp, ctypes.c_int]
        L.gso_krige.restype = ctypes.c_int
        L.gso_variogram_structured.argtypes = [_i64, _i64] + arr2 + arr2 + [ctypes.c_char, _dp, ctypes.c_int]
        L.gso_variogram_unstructured.argtypes = ([ctypes.c_int, _i64, _i64, _i64] + arr2 + arr1 + arr2
                                                 + [ctypes.c_char, ctypes.c_char, _dp, _dp, ctypes.c_int])
        L.gso_variogram_directional.argtypes = ([ctypes.c_int, _i64, _i64, _i64, _i64] + arr2 + arr1 + arr2 + arr2
                                                + [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_char,
                                                   _dp, _dp, ctypes.c_int])
        for f in (L.gso_variogram_structured, L.gso_variogram_unstructured, L.gso_variogram_directional):
            f.restype = ctypes.c_int
        for f in (L.gso_summator, L.gso_summator_incompr, L.gso_summator_fourier, L.gso_max_threads):
            f.restype = ctypes.c_int
        _lib = L
    return _lib


def _a(x, ndim):
    x = np.asarray(x)
    if x.dtype != np.float64 or x.ndim != ndim:
        raise TypeError("oracle expects float64 arrays of rank %d" % ndim)
    return x


def _s(x):
    return [x.ctypes.data] + [s // 8 for s in x.strides]


def _check(cov, z1, z2, pos):
    if cov.shape[0] != pos.shape[0] or cov.shape[1] != z1.shape[0] or cov.shape[1] != z2.shape[0]:
        raise ValueError("shape mismatch (reference: assert_eq!, src/field.rs:44-46)")


def max_threads() -> int:
    return lib().gso_max_threads()


def summate(cov_samples, z1, z2, pos, num_threads=None):
    cov, z1, z2, pos = _a(cov_samples, 2), _a(z1, 1), _a(z2, 1), _a(pos, 2)
    _check(cov, z1, z2, pos)
    d, n = cov.shape
    m = pos.shape[1]
    out = np.empty(m, dtype=np.float64)
    rc = lib().gso_summator(d, n, m, *_s(cov), *_s(z1), *_s(z2), *_s(pos), out.ctypes.data,
                            int(num_threads or 1))
    if rc:
        raise ValueError("oracle summator failed rc=%d" % rc)
    return out


def summate_incompr(cov_samples, z1, z2, pos, num_threads=None):
    cov, z1, z2, pos = _a(cov_samples, 2), _a(z1, 1), _a(z2, 1), _a(pos, 2)
    _check(cov, z1, z2, pos)
    d, n = cov.shape
    m = pos.shape[1]
    out = np.empty((d, m), dtype=np.float64, order="F")  # src/field.rs:166-174
    rc = lib().gso_summator_incompr(d, n, m, *_s(cov), *_s(z1), *_s(z2), *_s(pos),
                                    out.ctypes.data, int(num_threads or 1))
    if rc:
        raise ValueError("oracle summator_incompr failed rc=%d" % rc)
    return out


def summate_fourier(spectrum_factor, modes, z1, z2, pos, num_threads=None):
    sf = _a(spectrum_factor, 1)
    cov, z1, z2, pos = _a(modes, 2), _a(z1, 1), _a(z2, 1), _a(pos, 2)
    _check(cov, z1, z2, pos)
    if sf.shape[0] != cov.shape[1]:
        raise ValueError("spectrum_factor length mismatch")
    d, n = cov.shape
    m = pos.shape[1]
    out = np.empty(m, dtype=np.float64)
    rc = lib().gso_summator_fourier(d, n, m, *_s(sf), *_s(cov), *_s(z1), *_s(z2), *_s(pos),
                                    out.ctypes.data, int(num_threads or 1))
    if rc:
        raise ValueError("oracle summator_fourier failed rc=%d" % rc)
    return out


def _krige(krige_mat, krig_vecs, cond, want_error, num_threads):
    mat, vecs, cond = _a(krige_mat, 2), _a(krig_vecs, 2), _a(cond, 1)
    c = mat.shape[0]
    if mat.shape[1] != c or vecs.shape[0] != c or cond.shape[0] != c:
        raise ValueError("shape mismatch (reference: assert_eq!, src/krige.rs:30-32)")
    m = vecs.shape[1]
    field = np.empty(m, dtype=np.float64)
    error = np.empty(m, dtype=np.float64) if want_error else None
    rc = lib().gso_krige(c, m, *_s(mat), *_s(vecs), *_s(cond), field.ctypes.data,
                         error.ctypes.data if want_error else None, int(num_threads or 1))
    if rc:
        raise ValueError("oracle krige failed rc=%d" % rc)
    return (field, error) if want_error else field


def calc_field_krige(krige_mat, krig_vecs, cond, num_threads=None):
    """reference: calc_field_krige_py, src/lib.rs:104-118 -> krige::calculator_field_krige"""
    return _krige(krige_mat, krig_vecs, cond, False, num_threads)


def calc_field_krige_and_variance(krige_mat, krig_vecs, cond, num_threads=None):
    """reference: calc_field_krige_and_variance_py, src/lib.rs:86-102"""
    return _krige(krige_mat, krig_vecs, cond, True, num_threads)


# ---- variogram estimators (reference: src/variogram.rs, bindings src/lib.rs:119-216) -----------

def _ch(c, default):
    return (default if c is None else c).encode()[:1]


def variogram_structured(f, estimator_type=None, num_threads=None):
    f = _a(f, 2)
    out = np.empty(max(f.shape[0], 1), dtype=np.float64)
    lib().gso_variogram_structured(f.shape[0], f.shape[1], *_s(f), None, 0, 0, _ch(estimator_type, "m"),
                                   out.ctypes.data, int(num_threads or 1))
    return out


def variogram_ma_structured(f, mask, estimator_type=None, num_threads=None):
    f = _a(f, 2)
    mask = np.asarray(mask)
    if mask.dtype != np.bool_ or mask.shape != f.shape:
        raise TypeError("mask must be a bool array of f's shape")
    out = np.empty(max(f.shape[0], 1), dtype=np.float64)
    lib().gso_variogram_structured(f.shape[0], f.shape[1], *_s(f), mask.ctypes.data, mask.strides[0],
                                   mask.strides[1], _ch(estimator_type, "m"), out.ctypes.data,
                                   int(num_threads or 1))
    return out


def variogram_unstructured(f, bin_edges, pos, estimator_type=None, distance_type=None, num_threads=None):
    f, e, pos = _a(f, 2), _a(bin_edges, 1), _a(pos, 2)
    if pos.shape[1] != f.shape[1] or e.shape[0] < 2:
        raise ValueError("shape mismatch (reference: assert!, src/variogram.rs:473-484)")
    nb = e.shape[0] - 1
    v, c = np.empty(nb, dtype=np.float64), np.empty(nb, dtype=np.uint64)
    rc = lib().gso_variogram_unstructured(pos.shape[0], f.shape[0], f.shape[1], nb, *_s(f), *_s(e), *_s(pos),
                                          _ch(estimator_type, "m"), _ch(distance_type, "e"), v.ctypes.data,
                                          c.ctypes.data, int(num_threads or 1))
    if rc:
        raise ValueError("oracle variogram_unstructured failed rc=%d" % rc)
    return v, c


def variogram_directional(f, bin_edges, pos, direction, angles_tol=None, bandwidth=None, separate_dirs=None,
                          estimator_type=None, num_threads=None):
    f, e, pos, direction = _a(f, 2), _a(bin_edges, 1), _a(pos, 2), _a(direction, 2)
    if pos.shape[0] != direction.shape[1] or pos.shape[1] != f.shape[1] or e.shape[0] < 2:
        raise ValueError("shape mismatch (reference: assert!, src/variogram.rs:326-346)")
    nb, nd = e.shape[0] - 1, direction.shape[0]
    v, c = np.empty((nd, nb), dtype=np.float64), np.empty((nd, nb), dtype=np.uint64)
    rc = lib().gso_variogram_directional(pos.shape[0], f.shape[0], f.shape[1], nb, nd, *_s(f), *_s(e), *_s(pos),
                                         *_s(direction), float(np.pi / 8 if angles_tol is None else angles_tol),
                                         float(-1.0 if bandwidth is None else bandwidth),
                                         int(bool(separate_dirs)), _ch(estimator_type, "m"), v.ctypes.data,
                                         c.ctypes.data, int(num_threads or 1))
    if rc:
        raise ValueError("oracle variogram_directional failed rc=%d" % rc)
    return v, c
