"""Scratch GPU probe: DFMA peak, kernel-only throughput of each (P, L) variant on C2/C3, e2e."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch
import gstools_core as gc
from gstools_core import workloads

def time_device(kind, args, out_shape, reps=5):
    dargs = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in args]
    out = torch.empty(out_shape, dtype=torch.float64, device="cuda")
    fn = getattr(gc, kind + "_device")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        fn(*dargs, out, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(*dargs, out, stream=st); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

print("devices", gc.device_count(), torch.cuda.get_device_name(0))
for ms in (50, 500, 2000):
    r, t = gc.dfma_peak(0, ms)
    print("dfma_peak min_ms=%d: %.3f T DFMA/s (%.1f ms) -> %.1f DFMA/clk/SM @1965MHz" % (ms, r / 1e12, t, r / 148 / 1.965e9))
for cfg in ("c2", "c3", "c4s"):
    w = workloads.make("c4", 0.0625) if cfg == "c4s" else workloads.make(cfg)
    pm = w["m"] * w["n"]
    oshape = (3, w["m"]) if cfg == "c3" else (w["m"],)
    for P, L in [(0, 0), (8, 1), (6, 1), (4, 1), (3, 1), (2, 1), (1, 1)]:
        gc.set_variant(P, L)
        ms = time_device(w["kind"], w["args"], oshape)
        s = gc.last_stats()
        print("%s P=%d L=%d (used %d,%d): %.4f ms  %.1f Gpm/s" % (cfg, P, L, s["points_per_thread"], s["lanes_per_point"], ms, pm / ms / 1e6))
    gc.set_variant(0, 0)
# size sweep (3-D scalar, N=1000): heuristic vs forced variants
rng = np.random.default_rng(0)
for m in (10000, 100000, 300000, 1000000, 2000000, 3000000, 10000000):
    k = rng.normal(size=(3, 1000)); z1 = rng.normal(size=1000); z2 = rng.normal(size=1000)
    pos = rng.uniform(0, 100, size=(3, m))
    res = []
    for P, L in [(0, 0), (8, 1), (6, 1), (4, 1), (3, 1), (2, 1), (1, 1), (1, 4), (1, 16)]:
        gc.set_variant(P, L)
        ms = time_device("summate", (k, z1, z2, pos), (m,))
        s = gc.last_stats()
        res.append("(%d,%d)%s %.0f" % (s["points_per_thread"], s["lanes_per_point"], "*" if P == 0 else "", m * 1000 / ms / 1e6))
    print("sweep m=%d Gpm/s:" % m, "  ".join(res))
gc.set_variant(0, 0)
# e2e host paths on c2
w = workloads.make("c2")
k, z1, z2, pos = w["args"]
pm = w["m"] * w["n"]
for label, p in (("pageable", pos), ("pinned", torch.from_numpy(pos).pin_memory().numpy())):
    for chunk in (0, 1 << 16, 1 << 17, 1 << 18):
        gc.set_chunk_points(chunk)
        gc.summate(k, z1, z2, p)
        t = []
        for _ in range(5):
            t0 = time.perf_counter(); out = gc.summate(k, z1, z2, p); t.append(time.perf_counter() - t0)
        print("e2e c2 %s chunk=%d: %.3f ms  %.1f Gpm/s  chunks=%d" % (label, chunk, min(t) * 1e3, pm / min(t) / 1e9, gc.last_stats()["n_chunks"]))
