"""C2 (1000 modes x 1e6 points) on every rank at once, one process per GPU (run under torchrun):
per-rank end-to-end call time from pageable and from caller-pinned memory, max over ranks.
    GSF_ZERO_COPY=0|1  GSF_STAGING_THREADS=n  python -m torch.distributed.run --nproc-per-node 8 tools/c2_weak_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch, torch.distributed as dist
import gstools_core as gc
from gstools_core import workloads
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist.init_process_group("gloo")
torch.cuda.set_device(local); gc.set_devices([local]); gc.set_grid_detection(False)
w = workloads.make("c2"); k, z1, z2, pos = w["args"]
pages = [pos, pos.copy()]

def run(label, k_steps=50):
    r = None
    for i in range(6): r = gc.summate(k, z1, z2, pages[i % 2])
    reps = []
    for _ in range(3):
        dist.barrier()
        t0 = time.perf_counter()
        for i in range(k_steps): r = gc.summate(k, z1, z2, pages[i % 2])
        t = torch.tensor([(time.perf_counter() - t0) * 1e3 / k_steps], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        reps.append(float(t))
    st = gc.last_stats()
    if rank == 0:
        print("%-28s max-over-ranks ms/call %s  (chunks %d, staging threads %d, pos_memory %d)"
              % (label, " ".join("%.3f" % x for x in reps), st["n_chunks"], st["staging_threads"], st["pos_memory"]), flush=True)

if rank == 0:
    print("world=%d GSF_ZERO_COPY=%s GSF_STAGING_THREADS=%s GSF_PAGEABLE_DIRECT=%s" % (world, os.environ.get("GSF_ZERO_COPY", "-"),
          os.environ.get("GSF_STAGING_THREADS", "-"), os.environ.get("GSF_PAGEABLE_DIRECT", "-")), flush=True)
run("pageable")
if len(sys.argv) > 1 and sys.argv[1] == "pageable":
    dist.destroy_process_group(); sys.exit(0)
h = [gc.pinned(p) for p in pages]
run("caller-pinned")
for x in h: x.release()
gc.set_grid_detection(True)
run("default API (grid path)")
dist.destroy_process_group()
