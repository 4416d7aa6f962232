import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch
import gstools_core as gc
rng = np.random.default_rng(0)
N = 1000
st = torch.cuda.current_stream().cuda_stream
gc.set_profiling(True)
res = []
for d, kind in ((3, "summate"), (3, "summate_incompr"), (2, "summate")):
    k = rng.normal(size=(d, N)); z1 = rng.normal(size=N); z2 = rng.normal(size=N)
    dk, dz1, dz2 = (torch.from_numpy(x).cuda() for x in (k, z1, z2))
    for m in (1000000, 2000000, 16777216 if d == 2 else 4000000):
        pos = torch.from_numpy(rng.uniform(0, 100, size=(d, m))).cuda()
        nc = d if kind == "summate_incompr" else 1
        out = torch.empty((m, nc) if nc > 1 else (m,), dtype=torch.float64, device="cuda")
        fn = getattr(gc, kind + "_device")
        ks = []
        for i in range(7):
            fn(dk, dz1, dz2, pos, out.t() if nc > 1 else out, stream=st); torch.cuda.synchronize()
            if i >= 2: ks.append(gc.last_stats()["kernel_ms"])
        res.append("%s d%d m=%d: %.0f" % (kind[8:] or "scalar", d, m, m * N / min(ks) / 1e6))
print("TAIL=%s | " % os.environ.get("GSF_TAIL_WAVES", "default") + " | ".join(res))
