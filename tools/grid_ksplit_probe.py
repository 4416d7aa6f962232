"""Structured-grid GEMM on small grids: kernel time vs the number of mode splits (GSF_GRID_KSPLIT).
Re-runs itself per setting (the override is read once per process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
if os.environ.get("GSF_PROBE_CHILD") != "1":
    for ks in ("0", "1", "2", "3", "4", "5", "6", "7", "8", "10", "12"):
        subprocess.run([sys.executable, __file__], env=dict(os.environ, GSF_PROBE_CHILD="1", GSF_GRID_KSPLIT=ks))
    sys.exit(0)
import numpy as np, torch, gstools_core as gc
from gstools_core import workloads
res = []
for cfg in ("c2", "c3"):
    w = workloads.make(cfg, 0.001)
    axes = [np.arange(100.0)] * 3
    m = 10 ** 6
    kind = w["kind"]; nc = 3 if kind == "summate_incompr" else 1
    fn = getattr(gc, kind + "_grid")
    out = torch.empty((m, nc) if nc > 1 else (m,), dtype=torch.float64, device="cuda")
    out = out.t() if nc > 1 else out
    gc.set_profiling(True)
    ks = []
    for i in range(12):
        fn(*w["args"][:-1], axes, out=out); torch.cuda.synchronize()
        if i >= 2: ks.append(gc.last_stats()["kernel_ms"])
    res.append("%s kernel %.4f ms" % (cfg, sorted(ks)[len(ks) // 2]))
print("GSF_GRID_KSPLIT=%s: %s" % (os.environ.get("GSF_GRID_KSPLIT"), " | ".join(res)), flush=True)
