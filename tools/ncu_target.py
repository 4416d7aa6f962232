"""Small single-GPU target for ncu: a few device-resident launches of one config."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch
import gstools_core as gc
from gstools_core import workloads

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
scale = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
w = workloads.make(cfg, scale)
dargs = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in w["args"]]
out = torch.empty((3, w["m"]) if w["kind"] == "summate_incompr" else (w["m"],), dtype=torch.float64, device="cuda")
gc.set_variant(P, L)
fn = getattr(gc, w["kind"] + "_device")
for _ in range(reps):
    fn(*dargs, out, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("done", cfg, gc.last_stats())
