"""Synthetic inputs for the five BASELINE.json configs.

numpy `default_rng(seed)` (PCG64), all float64, z1/z2 ~ N(0,1).  The generators take a `scale`
so tests can run the same distributions at oracle-friendly sizes.

  C1  summate          d=2  N=100    M=1e4        Gaussian modes, points on a line
  C2  summate          d=3  N=1000   M=1e6        3-D Exponential (heavy-tailed) modes, 100^3 grid
  C3  summate_incompr  d=3  N=1000   M=1e6        3-D Gaussian modes, 100^3 grid
  C4  summate_fourier  d=2  N=1e4    M=4096^2     periodic grid modes + spectrum factor
  C5  summate          d=3  N=1e4    M=1e8        as C2, 1000x1000x100 grid with spacing 0.1
"""
from __future__ import annotations

import numpy as np

CONFIGS = ("c1", "c2", "c3", "c4", "c5")
KIND = {"c1": "summate", "c2": "summate", "c3": "summate_incompr", "c4": "summate_fourier", "c5": "summate"}
# FP64-pipe slots per point*mode: SURVEY.md 8(d4) "algorithmic" figure (libdevice-like sincos pair),
# and what this implementation's single-cos formulation actually issues (gsf_kernels.cuh header).
W_SURVEY = {"c1": 22, "c2": 23, "c3": 26, "c4": 22, "c5": 23}
# W_exec = dim + 5 + degree + nc; degree 6 (high) / 5 (throughput, chosen from 2^27 point*modes)
W_EXEC = {"c1": 14, "c2": 14, "c3": 16, "c4": 13, "c5": 14}


def _grid(shape, spacing, out=None):
    """C-order flattened meshgrid(indexing='ij') as a (d, M) array, built without temporaries."""
    d = len(shape)
    m = int(np.prod(shape))
    pos = np.empty((d, m), dtype=np.float64) if out is None else out
    for a in range(d):
        ax = np.arange(shape[a], dtype=np.float64) * spacing[a]
        view = pos[a].reshape(shape)
        idx = [None] * d
        idx[a] = slice(None)
        view[...] = ax[tuple(idx)]
    return pos


def _grid_range(shape, spacing, j0, j1):
    """Columns [j0, j1) of `_grid(shape, spacing)` without building the rest (bit-identical values:
    both evaluate float64(index) * spacing)."""
    d = len(shape)
    idx = np.arange(j0, j1, dtype=np.int64)
    pos = np.empty((d, j1 - j0), dtype=np.float64)
    for a in range(d - 1, -1, -1):
        np.multiply(idx % shape[a], spacing[a], out=pos[a])
        idx //= shape[a]
    return pos


def _axes(shape, spacing):
    """The axis vectors whose C-order expansion `_grid` returns (bit-identical values)."""
    return [np.arange(n, dtype=np.float64) * h for n, h in zip(shape, spacing)]


def gaussian_modes(rng, d, n, len_scale):
    # GSTools Gaussian model with rescale sqrt(pi)/2: k ~ N(0, (pi/2)/l^2 I)
    return rng.normal(size=(d, n)) * np.sqrt(np.pi / 2.0) / len_scale


def exponential_modes_3d(rng, n, len_scale):
    # 3-D Exponential spectral density ~ (1+(kl)^2)^-2: multivariate t with nu=1: k = g/|w|/l
    g = rng.normal(size=(3, n))
    w = rng.normal(size=n)
    return g / np.abs(w) / len_scale


def make(config, scale=1.0, pos_out=None, point_range=None):
    """Return dict(kind=..., args=(...)) for gstools_core.<kind>(*args).

    `scale` < 1 shrinks the number of points (grids keep their spacing, fewer cells per axis).
    `point_range=(j0, j1)` builds only that contiguous shard of the positions (`m` stays the total;
    `m_local` is the shard's size); pass a callable (m_total -> (j0, j1)) when the total is not
    known to the caller."""
    w = _make(config, scale, pos_out, point_range)
    w.setdefault("m_local", w["args"][-1].shape[1])
    w.setdefault("j0", 0)
    return w


def _shard(pr, m):
    j0, j1 = pr(m) if callable(pr) else pr
    if not (0 <= j0 <= j1 <= m):
        raise ValueError("point_range %r outside [0, %d]" % ((j0, j1), m))
    return int(j0), int(j1)


def _make(config, scale, pos_out, point_range):
    c = config.lower()
    if c == "c1":
        rng = np.random.default_rng(1)
        n, m = 100, max(8, int(round(10_000 * scale)))
        k = gaussian_modes(rng, 2, n, 1.0)
        z1, z2 = rng.normal(size=n), rng.normal(size=n)
        pos = np.stack([np.linspace(0.0, 10.0, m), np.linspace(-5.0, 5.0, m)])
        j0 = 0
        if point_range is not None:
            j0, j1 = _shard(point_range, m)
            pos = np.ascontiguousarray(pos[:, j0:j1])
        return dict(kind="summate", args=(k, z1, z2, pos), d=2, n=n, m=m, axes=None, j0=j0)
    if c in ("c2", "c5"):
        rng = np.random.default_rng(2 if c == "c2" else 5)
        n = 1000 if c == "c2" else 10_000
        k = exponential_modes_3d(rng, n, 10.0)
        z1, z2 = rng.normal(size=n), rng.normal(size=n)
        if c == "c2":
            side = max(2, int(round(100 * scale ** (1 / 3))))
            shape, spacing = (side, side, side), (1.0, 1.0, 1.0)
        else:
            f = scale ** (1 / 3)
            shape = (max(2, int(round(1000 * f))), max(2, int(round(1000 * f))), max(2, int(round(100 * f))))
            spacing = (0.1, 0.1, 0.1)
        m = int(np.prod(shape))
        if point_range is not None:
            j0, j1 = _shard(point_range, m)
            return dict(kind="summate", args=(k, z1, z2, _grid_range(shape, spacing, j0, j1)), d=3, n=n, m=m,
                        axes=None, j0=j0)
        pos = _grid(shape, spacing, pos_out)
        return dict(kind="summate", args=(k, z1, z2, pos), d=3, n=n, m=pos.shape[1], axes=_axes(shape, spacing))
    if c == "c3":
        rng = np.random.default_rng(3)
        n = 1000
        k = gaussian_modes(rng, 3, n, 10.0)
        z1, z2 = rng.normal(size=n), rng.normal(size=n)
        side = max(2, int(round(100 * scale ** (1 / 3))))
        if point_range is not None:
            j0, j1 = _shard(point_range, side ** 3)
            return dict(kind="summate_incompr", args=(k, z1, z2, _grid_range((side,) * 3, (1.0,) * 3, j0, j1)), d=3,
                        n=n, m=side ** 3, axes=None, j0=j0)
        pos = _grid((side, side, side), (1.0, 1.0, 1.0), pos_out)
        return dict(kind="summate_incompr", args=(k, z1, z2, pos), d=3, n=n, m=pos.shape[1],
                    axes=_axes((side, side, side), (1.0, 1.0, 1.0)))
    if c == "c4":
        rng = np.random.default_rng(4)
        period, ell, nm = 100.0, 5.0, 100
        dk = 2.0 * np.pi / period
        mx, my = np.meshgrid(np.arange(nm), np.arange(nm), indexing="ij")
        modes = np.ascontiguousarray(np.stack([mx.ravel(), my.ravel()]).astype(np.float64) * dk)
        kabs = np.sqrt((modes ** 2).sum(axis=0))
        spec = (ell / np.pi) ** 2 * np.exp(-(kabs * ell) ** 2 / np.pi)
        sf = np.sqrt(2.0 * spec * dk ** 2)
        n = modes.shape[1]
        z1, z2 = rng.normal(size=n), rng.normal(size=n)
        side = max(2, int(round(4096 * scale ** 0.5)))
        if point_range is not None:
            j0, j1 = _shard(point_range, side * side)
            return dict(kind="summate_fourier",
                        args=(sf, modes, z1, z2, _grid_range((side, side), (period / side,) * 2, j0, j1)), d=2, n=n,
                        m=side * side, axes=None, j0=j0)
        pos = _grid((side, side), (period / side, period / side), pos_out)
        return dict(kind="summate_fourier", args=(sf, modes, z1, z2, pos), d=2, n=n, m=pos.shape[1],
                    axes=_axes((side, side), (period / side, period / side)))
    raise ValueError("unknown config %r" % config)


def subset_points(w, idx):
    """Same workload restricted to the point columns `idx` (for oracle checks of big configs)."""
    args = list(w["args"])
    args[-1] = np.ascontiguousarray(args[-1][:, idx])
    return dict(w, args=tuple(args), m=len(idx), m_local=len(idx), axes=None)
