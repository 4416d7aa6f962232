"""GPU sweep: throughput of each P variant per (kind, dim, M) -> data for choose_variant's table."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch
import gstools_core as gc

def time_device(kind, args, out_shape, reps=4):
    dargs = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in args]
    out = torch.empty(out_shape, dtype=torch.float64, device="cuda")
    fn = getattr(gc, kind + "_device")
    st = torch.cuda.current_stream().cuda_stream
    fn(*dargs, out, stream=st); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(*dargs, out, stream=st); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

rng = np.random.default_rng(0)
N = 1000
for kind, d in [("summate", 1), ("summate", 2), ("summate", 3), ("summate_incompr", 2), ("summate_incompr", 3)]:
    for m in (50_000, 200_000, 500_000, 1_000_000, 4_000_000):
        k = rng.normal(size=(d, N)); z1 = rng.normal(size=N); z2 = rng.normal(size=N)
        pos = rng.uniform(0, 100, size=(d, m))
        oshape = (d, m) if kind == "summate_incompr" else (m,)
        res = []
        for P in (0, 1, 2, 3, 4, 6):
            if d == 1 and P == 6: continue
            gc.set_variant(P, 1 if P else 0)
            ms = time_device(kind, (k, z1, z2, pos), oshape)
            s = gc.last_stats()
            res.append("%s%d:%4.0f" % ("*" if P == 0 else "P", s["points_per_thread"], m * N / ms / 1e6))
        print("%-16s d=%d m=%8d  %s" % (kind, d, m, "  ".join(res)), flush=True)
