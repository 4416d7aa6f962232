import sys, time, os, gc as pygc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch, gstools_core as gc
from gstools_core import workloads
gc.set_grid_detection(False)
w = workloads.make("c2"); fn = gc.summate
args = list(w["args"]); pins = [torch.from_numpy(args[-1]).pin_memory().numpy() for _ in range(2)]
for label in ("gc on", "gc off"):
    if label == "gc off": pygc.disable()
    for i in range(5): fn(*args[:-1], pins[i % 2])
    ts = []
    for i in range(400):
        t0 = time.perf_counter(); res = fn(*args[:-1], pins[i % 2]); gc.last_stats(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    print("%s: mean %.4f median %.4f p99 %.4f max %.4f  n>1.2ms=%d  worst idx %s" % (label, ts.mean(), np.median(ts), np.percentile(ts, 99), ts.max(), (ts > 1.2).sum(), np.argsort(ts)[-5:]))
