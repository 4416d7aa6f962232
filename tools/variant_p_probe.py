"""Points-per-thread variants of the degree-5 kernels at throughput sizes (4e6 points x 1000 modes,
device-resident): is P = 4 worth it now that the polynomial is one DFMA shorter?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch
import gstools_core as gc
rng = np.random.default_rng(0)
N, m = 1000, 4_000_000
for kind, d in [("summate", 3), ("summate", 2), ("summate_incompr", 3), ("summate_incompr", 2)]:
    k = rng.normal(size=(d, N)); z1 = rng.normal(size=N); z2 = rng.normal(size=N)
    pos = rng.uniform(0, 100, size=(d, m))
    dargs = [torch.from_numpy(a).cuda() for a in (k, z1, z2, pos)]
    nc = d if kind == "summate_incompr" else 1
    out = torch.empty((m, nc) if nc > 1 else (m,), dtype=torch.float64, device="cuda")
    out = out.t() if nc > 1 else out
    fn = getattr(gc, kind + "_device")
    st = torch.cuda.current_stream().cuda_stream
    res = []
    for P in (2, 3, 4):
        gc.set_variant(P, 1)
        fn(*dargs, out, stream=st); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record(); fn(*dargs, out, stream=st); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        s = gc.last_stats()
        res.append("P%d deg%d: %6.1f G pm/s" % (s["points_per_thread"], s["poly_degree"], m * N / best / 1e6))
    gc.set_variant(0, 0)
    print("%-16s d=%d  %s" % (kind, d, "   ".join(res)), flush=True)
