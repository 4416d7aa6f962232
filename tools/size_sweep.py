"""Kernel-only time (library CUDA events around the launch) vs number of points: fixed overhead?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch
import gstools_core as gc
rng = np.random.default_rng(0)
N = 1000
k = rng.normal(size=(3, N)); z1 = rng.normal(size=N); z2 = rng.normal(size=N)
dk, dz1, dz2 = (torch.from_numpy(x).cuda() for x in (k, z1, z2))
gc.set_profiling(True)
st = torch.cuda.current_stream().cuda_stream
for P in (3, 1):
    gc.set_variant(P, 1)
    prev = None
    for ctas_per_sm in (1, 2, 4, 9, 17, 18, 19, 27, 36, 72):
        m = 148 * ctas_per_sm * 128 * P
        pos = torch.from_numpy(rng.uniform(0, 100, size=(3, m))).cuda()
        out = torch.empty(m, dtype=torch.float64, device="cuda")
        ks = []
        for i in range(6):
            gc.summate_device(dk, dz1, dz2, pos, out, stream=st); torch.cuda.synchronize()
            if i >= 2: ks.append(gc.last_stats()["kernel_ms"])
        t = min(ks)
        print("P=%d ctas/SM=%3d m=%8d kernel %.4f ms  %.0f Gpm/s  ms per (CTA/SM) %.4f" % (P, ctas_per_sm, m, t, m * N / t / 1e6, t / ctas_per_sm), flush=True)
# C2 exact size
for P in (3, 1):
    gc.set_variant(P, 1)
    m = 1000000
    pos = torch.from_numpy(rng.uniform(0, 100, size=(3, m))).cuda(); out = torch.empty(m, dtype=torch.float64, device="cuda")
    ks = []
    for i in range(6):
        gc.summate_device(dk, dz1, dz2, pos, out, stream=st); torch.cuda.synchronize()
        if i >= 2: ks.append(gc.last_stats()["kernel_ms"])
    print("P=%d m=1000000 kernel %.4f ms %.0f Gpm/s" % (P, min(ks), m * N / min(ks) / 1e6))
