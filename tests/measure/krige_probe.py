"""Kriging kernels: time vs the CPU oracle (reference bench shape: 500 conditions x 1e4 points)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, gstools_core as gc, oracle
rng = np.random.default_rng(0)
dmma, _ = gc.dmma_peak(0, 200.0)
for c, m in ((500, 10_000), (500, 1_000_000), (2000, 100_000)):
    mat = rng.normal(size=(c, c)); vecs = rng.normal(size=(c, m)); cond = rng.normal(size=c)
    gc.set_profiling(True)
    for _ in range(2): gc.calc_field_krige_and_variance(mat, vecs, cond)
    ts, ks = [], []
    for _ in range(3):
        t0 = time.perf_counter(); gc.calc_field_krige_and_variance(mat, vecs, cond); ts.append(time.perf_counter() - t0)
        ks.append(gc.last_stats()["kernel_ms"])
    for _ in range(2): gc.calc_field_krige(mat, vecs, cond)
    tf = []
    for _ in range(3):
        t0 = time.perf_counter(); gc.calc_field_krige(mat, vecs, cond); tf.append(time.perf_counter() - t0)
    kf = gc.last_stats()["kernel_ms"]
    gc.set_profiling(False)
    sub = min(m, 2000)
    t0 = time.perf_counter(); oracle.calc_field_krige_and_variance(mat, np.ascontiguousarray(vecs[:, :sub]), cond, oracle.max_threads()); tc = (time.perf_counter() - t0) * m / sub
    fma = float(c) * c * m
    print("C=%d M=%d: variance e2e %.2f ms, kernels %.3f ms (%.2f T FMA/s, %.0f%% of DMMA peak) | field-only e2e %.2f ms kernels %.3f ms | CPU oracle (%d threads, extrapolated) %.0f ms"
          % (c, m, min(ts) * 1e3, min(ks), fma / min(ks) / 1e9, 100 * fma / (min(ks) * 1e-3) / dmma, min(tf) * 1e3, kf, oracle.max_threads(), tc * 1e3), flush=True)
