import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch, gstools_core as gc
from gstools_core import workloads
gc.set_grid_detection(False)
w = workloads.make("c2"); fn = gc.summate; pm = w["m"] * w["n"]
args = list(w["args"]); pin = torch.from_numpy(args[-1]).pin_memory().numpy()
a = args[:-1] + [pin]
def run(label):
    for _ in range(3): fn(*a)
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); o = fn(*a); ts.append(time.perf_counter() - t0)
    st = gc.last_stats()
    print("%-34s min %.3f ms  median %.3f ms  chunks=%d P=%d launches=%d" % (label, min(ts) * 1e3, sorted(ts)[15] * 1e3, st["n_chunks"], st["points_per_thread"], st["kernel_launches"]), flush=True)
for chunk in (32768, 63488, 126976, 253952, 507904):
    gc.set_chunk_points(chunk); gc.set_variant(0, 0); run("fixed %d auto" % chunk)
    gc.set_variant(1, 1); run("fixed %d P=1" % chunk)
gc.set_chunk_points(0)
gc.set_variant(0, 0); run("ramp auto")
gc.set_variant(1, 1); run("ramp P=1")
gc.set_variant(3, 1); run("ramp P=3")
