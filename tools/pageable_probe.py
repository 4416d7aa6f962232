import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, gstools_core as gc
from gstools_core import workloads
gc.set_grid_detection(False)
for cfg, scale in (("c2", 1.0), ("c2", 0.25), ("c3", 1.0), ("c4", 0.25)):
    w = workloads.make(cfg, scale); fn = getattr(gc, w["kind"]); a = w["args"]; pm = w["m"] * w["n"]
    for _ in range(3): fn(*a)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); fn(*a); ts.append(time.perf_counter() - t0)
    st = gc.last_stats()
    print("%s x%.2f pageable direct=%s: median %.3f ms min %.3f -> %.0f Gpm/s chunks=%d staging_threads=%d env=%s" % (cfg, scale, os.environ.get("GSF_PAGEABLE_DIRECT", "0"), sorted(ts)[10] * 1e3, min(ts) * 1e3, pm / sorted(ts)[10] / 1e9, st["n_chunks"], st["staging_threads"], os.environ.get("GSF_STAGING_THREADS", "-")), flush=True)
