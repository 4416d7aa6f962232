"""Pageable-input pipeline: chunk schedule knobs (GSF_CHUNK_FIRST / GROWTH / CAP_DIV / DOWN) at C2 and C3.
The knobs are read once per process, so the script re-runs itself per setting."""
import itertools, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
if os.environ.get("GSF_PROBE_CHILD") != "1":
    combos = [(32768, 2, 6, 64)]   # shipped default first
    for first, growth, div, down in itertools.product((16384, 32768, 65536), (2, 3, 4), (2, 3, 4, 6), (0, 1, 2, 64)):
        if (first, growth, div, down) != combos[0]:
            combos.append((first, growth, div, down))
    for first, growth, div, down in combos:
        subprocess.run([sys.executable, __file__], env=dict(os.environ, GSF_PROBE_CHILD="1", GSF_CHUNK_FIRST=str(first),
                       GSF_CHUNK_GROWTH=str(growth), GSF_CHUNK_CAP_DIV=str(div), GSF_CHUNK_DOWN=str(down)))
    sys.exit(0)
import numpy as np, gstools_core as gc
from gstools_core import workloads
gc.set_grid_detection(False)
out = []
for cfg in ("c2", "c3"):
    w = workloads.make(cfg); fn = getattr(gc, w["kind"]); a = w["args"]
    for _ in range(5): fn(*a)
    ts = []
    for _ in range(40):
        t0 = time.perf_counter(); fn(*a); ts.append(time.perf_counter() - t0)
    ts.sort()
    out.append("%s median %.3f p10 %.3f ms chunks=%d" % (cfg, ts[20] * 1e3, ts[4] * 1e3, gc.last_stats()["n_chunks"]))
e = os.environ
print("first=%s growth=%s cap_div=%s down=%s: %s" % (e["GSF_CHUNK_FIRST"], e["GSF_CHUNK_GROWTH"], e["GSF_CHUNK_CAP_DIV"], e["GSF_CHUNK_DOWN"], " | ".join(out)), flush=True)
