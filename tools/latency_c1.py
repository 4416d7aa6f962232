"""C1 (100 modes x 1e4 points) per-call latency from pageable memory: same modes every call (records
cached on the device) and fresh z1/z2 every call (the ensemble use: upload + pre-pass each call)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, gstools_core as gc
from gstools_core import workloads
w = workloads.make("c1"); k, z1, z2, pos = w["args"]
for _ in range(20): gc.summate(k, z1, z2, pos)
def med(f, n=400):
    ts = []
    for i in range(n):
        t0 = time.perf_counter(); f(i); ts.append(time.perf_counter() - t0)
    ts.sort(); return ts[n // 2] * 1e6, ts[n // 10] * 1e6
print("C1 same modes      : median %.1f us  p10 %.1f us" % med(lambda i: gc.summate(k, z1, z2, pos)))
zs = [np.random.default_rng(i).normal(size=(2, z1.size)) for i in range(400)]
print("C1 fresh z1/z2     : median %.1f us  p10 %.1f us" % med(lambda i: gc.summate(k, zs[i][0], zs[i][1], pos)))
print("stats", gc.last_stats())
