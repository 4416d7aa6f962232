"""ncu target: a few launches of the structured-grid GEMM on one config (explicit axes, device output)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch
import gstools_core as gc
from gstools_core import workloads
cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = workloads.make(cfg, 0.001)
full = {"c2": ((100, 100, 100), (1.0,) * 3), "c3": ((100, 100, 100), (1.0,) * 3),
        "c4": ((4096, 4096), (100 / 4096,) * 2), "c5": ((1000, 1000, 100), (0.1,) * 3)}[cfg]
axes = workloads._axes(*full)
m = int(np.prod(full[0]))
nc = 3 if w["kind"] == "summate_incompr" else 1
out = torch.empty((m, nc) if nc > 1 else (m,), dtype=torch.float64, device="cuda")
fn = getattr(gc, w["kind"] + "_grid")
for _ in range(reps):
    fn(*w["args"][:-1], axes, out=out.t() if nc > 1 else out)
torch.cuda.synchronize()
print("done", gc.last_stats())
