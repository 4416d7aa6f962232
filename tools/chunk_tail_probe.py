"""C5-sized pageable calls: pipeline chunk size x short-tile tail.  GSF_TAIL_WAVES is read once per
process, so the script re-runs itself per tail setting; chunk sizes are swept in-process."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
if os.environ.get("GSF_PROBE_CHILD") != "1":
    for tw in ("0.5", "0.25", "0.1", "0"):
        subprocess.run([sys.executable, __file__], env=dict(os.environ, GSF_PROBE_CHILD="1", GSF_TAIL_WAVES=tw))
    sys.exit(0)
import numpy as np, gstools_core as gc
from gstools_core import workloads
gc.set_grid_detection(False)
w = workloads.make("c5", point_range=(0, 25_000_000))
k, z1, z2, pos = w["args"]
pm = w["n"] * w["m_local"]
wave = 148 * 9 * 384
for chunk in (0, 1 << 20, 2 * wave, 3 * wave, 4 * wave, 1 << 21):
    gc.set_chunk_points(chunk)
    gc.summate(k, z1, z2, pos)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); gc.summate(k, z1, z2, pos); ts.append(time.perf_counter() - t0)
    st = gc.last_stats()
    print("tail_waves=%s chunk=%8d: %.2f ms (%.0f Gpm/s) chunks=%d" % (os.environ.get("GSF_TAIL_WAVES"), chunk, min(ts) * 1e3, pm / min(ts) / 1e9, st["n_chunks"]), flush=True)
