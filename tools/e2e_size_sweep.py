"""End to end (plain pageable numpy in, host array out) vs device-resident kernel time over problem sizes:
where between the one-launch small path and the chunk pipeline does a call lose the most?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch, gstools_core as gc
rng = np.random.default_rng(0)
gc.set_grid_detection(False)
st = torch.cuda.current_stream().cuda_stream
print("GSF_SMALL_KB=%s GSF_STAGING_THREADS=%s" % (os.environ.get("GSF_SMALL_KB", "-"), os.environ.get("GSF_STAGING_THREADS", "-")))
print("d=3 scattered points, pageable input; e2e = median of 40 calls; kernel = device-resident call incl. prep (CUDA events, best of 5)")
for n in (100, 1000):
    k = rng.normal(size=(3, n)); z1 = rng.normal(size=n); z2 = rng.normal(size=n)
    dk, dz1, dz2 = (torch.from_numpy(a).cuda() for a in (k, z1, z2))
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else \
        [1000, 3000, 10000, 16000, 20000, 30000, 60000, 100000, 200000, 300000, 600000, 1000000, 3000000, 10000000]
    for m in sizes:
        pos = rng.uniform(0, 100, size=(3, m))
        r = None
        for _ in range(4): r = gc.summate(k, z1, z2, pos)
        ts = []
        for _ in range(40):
            t0 = time.perf_counter(); r = gc.summate(k, z1, z2, pos); ts.append(time.perf_counter() - t0)
        s = gc.last_stats()
        e2e = sorted(ts)[20] * 1e6
        dpos = torch.from_numpy(pos).cuda(); out = torch.empty(m, dtype=torch.float64, device="cuda")
        gc.summate_device(dk, dz1, dz2, dpos, out, stream=st); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record(); gc.summate_device(dk, dz1, dz2, dpos, out, stream=st); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3)
        print("N=%4d M=%8d  e2e %9.1f us  kernel %9.1f us  overhead %8.1f us (x%.2f)  chunks=%d launches=%d P=%d L=%d deg=%d staging=%d"
              % (n, m, e2e, best, e2e - best, e2e / best, s["n_chunks"], s["kernel_launches"], s["points_per_thread"], s["lanes_per_point"],
                 s["poly_degree"], s["staging_threads"]), flush=True)
