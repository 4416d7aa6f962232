"""Device-resident C5-like problem on GPU 0, sharded over G GPUs through NVLink peer mappings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch, gstools_core as gc
from gstools_core import workloads
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
gc.set_grid_detection(False)   # measure the general kernel (a detected grid would run the GEMM path on the owner GPU)
w = workloads.make("c5", scale)
k, z1, z2, pos = w["args"]; pm = w["m"] * w["n"]
dpos = torch.from_numpy(pos).cuda(0); out = torch.empty(w["m"], dtype=torch.float64, device="cuda:0")
torch.cuda.synchronize()
base = None
for g in (1, 2, 4, 8):
    if g > gc.device_count(): break
    gc.set_devices(list(range(g)))
    gc.summate_device(k, z1, z2, dpos, out, sync=True)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); gc.summate_device(k, z1, z2, dpos, out, sync=True); ts.append(time.perf_counter() - t0)
    t = min(ts); base = base or t
    print("peer-sharded device-resident c5 x%.2f: G=%d %.1f ms %.0f Gpm/s speedup %.2fx devices=%d"
          % (scale, g, t * 1e3, pm / t / 1e9, base / t, gc.last_stats()["n_devices"]), flush=True)
