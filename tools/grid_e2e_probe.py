"""Default API on gridded input (exact detection on, pageable positions, pinned-pool result): end-to-end
call time at C2 / C3 vs the row-chunk size (GSF_GRID_CHUNK_MB) and the mode split (GSF_GRID_KSPLIT).
Re-runs itself per setting (the overrides are read once per process)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
if os.environ.get("GSF_PROBE_CHILD") != "1":
    for mb in ("0", "1", "2", "4", "8"):
        for ks in ("0", "1", "2", "4", "8"):
            subprocess.run([sys.executable, __file__], env=dict(os.environ, GSF_PROBE_CHILD="1", GSF_GRID_KSPLIT=ks, GSF_GRID_CHUNK_MB=mb))
    sys.exit(0)
import numpy as np, gstools_core as gc
from gstools_core import workloads
res = []
for cfg in ("c2", "c3"):
    w = workloads.make(cfg); fn = getattr(gc, w["kind"]); a = w["args"]
    r = None
    for _ in range(5): r = fn(*a)
    assert gc.last_stats()["grid_path"] == 1
    ts = []
    for _ in range(60):
        t0 = time.perf_counter(); r = fn(*a); ts.append(time.perf_counter() - t0)
    ts.sort()
    res.append("%s median %.3f ms p10 %.3f chunks %d" % (cfg, ts[30] * 1e3, ts[6] * 1e3, gc.last_stats()["n_chunks"]))
print("CHUNK_MB=%s KSPLIT=%s: %s" % (os.environ.get("GSF_GRID_CHUNK_MB"), os.environ.get("GSF_GRID_KSPLIT"), " | ".join(res)), flush=True)
