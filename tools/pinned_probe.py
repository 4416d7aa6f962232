"""Pinned-input policy: ONE zero-copy launch over PCIe vs the chunked copy pipeline, for positions
pinned with cudaHostAlloc (torch pin_memory) and with cudaHostRegister (gstools_core.pinned).
Usage: python tools/pinned_probe.py  (re-runs itself with GSF_ZERO_COPY=0 for the pipeline rows)"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
if os.environ.get("GSF_PROBE_CHILD") != "1":
    for zc in ("1", "0"):
        subprocess.run([sys.executable, __file__], env=dict(os.environ, GSF_PROBE_CHILD="1", GSF_ZERO_COPY=zc))
    sys.exit(0)
import numpy as np, torch, gstools_core as gc
from gstools_core import workloads
gc.set_grid_detection(False)
zc = os.environ.get("GSF_ZERO_COPY")
for cfg, rng, reps in (("c2", None, 30), ("c5", (0, 1_000_000), 20), ("c5", (0, 12_500_000), 5), ("c5", (0, 50_000_000), 2)):
    w = workloads.make(cfg, point_range=rng)
    k, z1, z2, pos = w["args"]
    pm = w["n"] * w["m_local"]
    tp = torch.from_numpy(pos).pin_memory().numpy()
    rows = []
    for label, arr, ctx in (("pageable", pos, None), ("hostalloc", tp, None), ("registered", pos, gc.pinned)):
        h = ctx(arr) if ctx else None
        for _ in range(2): gc.summate(k, z1, z2, arr)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); gc.summate(k, z1, z2, arr); ts.append(time.perf_counter() - t0)
        st = gc.last_stats()
        rows.append("%s %.3f ms (chunks %d)" % (label, sorted(ts)[len(ts) // 2] * 1e3, st["n_chunks"]))
        if h: h.release()
    print("GSF_ZERO_COPY=%s %s m=%d n=%d: %s" % (zc, cfg, w["m_local"], w["n"], " | ".join(rows)), flush=True)
