"""Per-phase host timeline (GSF_TRACE=1) of small / default-path calls: C1 and the C2 grid path."""
import os, sys, time
os.environ["GSF_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, gstools_core as gc
from gstools_core import workloads
def run(cfg, detect, label, n=6):
    w = workloads.make(cfg); fn = getattr(gc, w["kind"]); a = w["args"]
    gc.set_grid_detection(detect)
    for i in range(n):
        sys.stderr.write("== %s call %d\n" % (label, i)); sys.stderr.flush()
        t0 = time.perf_counter(); fn(*a); dt = time.perf_counter() - t0
        sys.stderr.write("== %s call %d python wall %.1f us\n" % (label, i, dt * 1e6)); sys.stderr.flush()
run("c1", False, "C1")
run("c2", True, "C2 default(grid)")
run("c2", False, "C2 general pageable")
os.environ["GSF_TRACE"] = "0"
