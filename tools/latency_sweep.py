"""End-to-end call latency (pageable numpy in/out) for small and medium problems, general vs grid path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, gstools_core as gc
rng = np.random.default_rng(0)
def bench(fn, *a, reps=30):
    for _ in range(3): fn(*a)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(*a); ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2] * 1e6
for shape in [(100, 100), (200, 200), (500, 500), (1000, 1000), (20, 20, 25), (50, 50, 40), (100, 100, 100)]:
    d = len(shape)
    axes = [np.linspace(0, 10, n) for n in shape]
    g = np.meshgrid(*axes, indexing="ij"); pos = np.ascontiguousarray(np.stack([x.ravel() for x in g]))
    scat = rng.uniform(0, 10, size=pos.shape)
    for n in (100, 1000):
        k = rng.normal(size=(d, n)); z1 = rng.normal(size=n); z2 = rng.normal(size=n)
        gc.set_grid_detection(False); tg = bench(gc.summate, k, z1, z2, pos)
        ts_ = bench(gc.summate, k, z1, z2, scat)
        gc.set_grid_detection(True); ta = bench(gc.summate, k, z1, z2, pos); used = gc.last_stats()["grid_path"]
        tx = bench(gc.summate_grid, k, z1, z2, axes)
        print("shape %-16s N=%4d pm=%.0e | general %7.0f us (scattered %7.0f) | auto-detect %7.0f us (grid_path=%d) | explicit axes %7.0f us"
              % (shape, n, pos.shape[1] * n, tg, ts_, ta, used, tx), flush=True)
