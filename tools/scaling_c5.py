"""Strong-scaling sweep of ONE call sharded over the GPUs of one box (single process, one host
thread per device, no collective): gstools_core.summate on C5 (1e4 modes x 1e8 points)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch
import gstools_core as gc
from gstools_core import workloads

cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
w = workloads.make(cfg, scale)
fn = getattr(gc, w["kind"])
pm = w["m"] * w["n"]
args = list(w["args"])
ndev = gc.device_count()
print("devices", ndev, "workload", cfg, "points", w["m"], "modes", w["n"], flush=True)
t0 = time.perf_counter(); pinned = torch.from_numpy(args[-1]).pin_memory().numpy(); print("pin_memory %.2f s" % (time.perf_counter() - t0), flush=True)
res = {}
mode = sys.argv[3] if len(sys.argv) > 3 else "general"
gc.set_grid_detection(mode == "grid")
print("path:", mode, flush=True)
for label, pos in (("pinned", pinned), ("pageable", args[-1])):
    base = None
    for g in [1, 2, 4, 8]:
        if g > ndev: break
        gc.set_devices(list(range(g)))
        a = args[:-1] + [pos]
        out = fn(*a)                      # warm-up: contexts, buffers
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); out = fn(*a); ts.append(time.perf_counter() - t0)
        t = min(ts)
        if base is None: base = (t, out.copy())
        same = bool(np.array_equal(out, base[1]))
        st = gc.last_stats()
        print("%s G=%d: %.1f ms  %.0f Gpm/s  speedup %.2fx  devices_used=%d chunks=%d grid_path=%d identical_to_G1=%s"
              % (label, g, t * 1e3, pm / t / 1e9, base[0] / t, st["n_devices"], st["n_chunks"], st["grid_path"], same), flush=True)
        res["%s_G%d" % (label, g)] = {"ms": t * 1e3, "gpm_s": pm / t / 1e9, "speedup": base[0] / t}
        del out
print(json.dumps(res))
