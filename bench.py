#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 field summation (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c5|c1..c4]

A "step" is one pass of the hot path (field::summator) over the whole synthetic workload.  The
default workload is the north star's scaling case, BASELINE.json configs[4] (C5): 3-D Exponential
covariance, 10^4 modes x 10^8 points (1000 x 1000 x 100 grid), 10^12 point*modes per step.  The
points are sharded contiguously over the N GPUs, one process per GPU, rank r taking
gsf_shard_bounds(M, N, r) -- the split that replaces `Zip::from(pos.columns()).par_map_collect`
(/root/reference/src/field.rs:53).  The path has no exchange step, so there is no data-path
collective: STRONG scaling, value = total point*modes / max-over-ranks time.

Numbers per run:
  value : kernel path, the rank's shard already resident in HBM, CUDA events on the launching stream.
  e2e   : the reference-facing call gstools_core.summate(...) on plain PAGEABLE numpy positions
          (what GSTools passes, /root/reference/src/lib.rs:43-46), host result; H2D and D2H inside
          the timed call.  `e2e.pinned_input` is the same call on page-locked positions.
  in_process : (N > 1, rank 0) ONE gstools_core.summate call sharding over all N devices inside one
          process (gsf_set_devices) -- the design the library ships -- with a parity flag against
          the single-device result at every shard boundary.
  c2    : configs[1] (1000 modes x 10^6 points), per rank, for the per-call overheads a small problem exposes.

`--impl reference` times the reference algorithm's CPU restatement (oracle/, OpenMP over all host
cores; the Rust crate itself cannot be built in this image -- no cargo/rustc) on a bounded sample
of the same workload, scaled linearly in the number of points.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "gstools-core_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "field::summator Gpoint*modes/s (f64)"
UNIT = "Gpoint*modes/s"
L2_BYTES = 126 * 1024 * 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the number of points (debug)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the c2 / structured_grid / in_process blocks")
    ap.add_argument("--in-process-child", type=int, default=0, help=argparse.SUPPRESS)
    return ap.parse_args()


def peak_file():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


WORKLOADS = {
    "c1": ("C1 field::summator 2D Gaussian, 100 modes x 1e4 points", "summate", 2, 100),
    "c2": ("C2 field::summator 3D Exponential, 1000 modes x 1e6 points (100^3 grid)", "summate", 3, 1000),
    "c3": ("C3 field::summator_incompr 3D, 1000 modes x 1e6 points (100^3 grid)", "summate_incompr", 3, 1000),
    "c4": ("C4 field::summator_fourier 2D, 1e4 modes x 4096^2 points", "summate_fourier", 2, 10000),
    "c5": ("C5 field::summator 3D Exponential, 1e4 modes x 1e8 points (1000x1000x100 grid)", "summate", 3, 10000),
}


def config_for(args, m_total):
    """The workload description -- identical in the `ours` and `reference` arms."""
    name, kind, d, n = WORKLOADS[args.workload]
    nc = d if kind == "summate_incompr" else 1
    per_rank_bytes = (d + nc) * 8 * m_total // max(1, args.gpus)
    return {
        "workload": name + (" [scaled x%g]" % args.scale if args.scale != 1.0 else ""),
        "kind": kind, "dim": d, "modes": n, "points_total": m_total, "point_modes_total": n * m_total,
        "sharding": "contiguous point shards, one per GPU, no collective (strong scaling: total work fixed)",
        "l2": ("inputs larger than L2: every step streams %.0f MB of positions/results per GPU through a %d MB L2"
               % (per_rank_bytes / 1e6, L2_BYTES >> 20)) if per_rank_bytes > 2 * L2_BYTES else
              "input/output sets rotated so that no step re-reads L2-resident data",
    }


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="gsf_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # "under load": samples whose power is above the midpoint between idle and max
            thr = (min(pw) + max(pw)) / 2 if max(pw) - min(pw) > 50 else -1
            load = [s for s, p in zip(sm, pw) if p >= thr] or sm
            out.update(sm_mhz=statistics.median(load), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=max(pw))
        return out


# ------------------------------------------------------------------------------------------------
def host_threads():
    """All host cores this process may use.  (torchrun exports OMP_NUM_THREADS=1 to its workers, so
    omp_get_max_threads() would under-report; the oracle takes the thread count explicitly.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(cfg, scale, seconds, threads):
    """A bounded, evenly strided point sample of the workload sized to ~`seconds` of CPU time.
    Returns (oracle function, argument tuple, sample size, total points, modes)."""
    import oracle
    from gstools_core import workloads

    # probe on a small shard to learn the rate, then build the sample from strided single points
    probe = workloads.make(cfg, scale, point_range=lambda m: (0, min(m, max(threads * 64, 4096))))
    fn = getattr(oracle, probe["kind"])
    n, m = probe["n"], probe["m"]
    t0 = time.perf_counter(); fn(*probe["args"], threads); dt = time.perf_counter() - t0
    rate = probe["m_local"] * n / max(dt, 1e-9)
    ms = int(min(m, max(probe["m_local"], rate * seconds / n)))
    # evenly strided over the whole domain: a few hundred contiguous runs, so that the generator
    # never has to materialise the full position array (2.4 GB for C5)
    runs = int(min(ms, 512))
    run_len = max(1, ms // runs)
    starts = np.linspace(0, m - run_len, runs).astype(np.int64)
    parts = [workloads.make(cfg, scale, point_range=(int(s), int(s) + run_len))["args"][-1] for s in starts]
    pos = np.ascontiguousarray(np.concatenate(parts, axis=1))
    return fn, probe["args"][:-1] + (pos,), pos.shape[1], m, n


def cpu_reference_rate(cfg, scale, seconds, threads=None):
    """Oracle (C restatement of the Rayon path) on a bounded point sample; returns (Gpm/s, info, sample)."""
    threads = threads or host_threads()
    fn, a, ms, m, n = cpu_sample(cfg, scale, seconds, threads)
    t0 = time.perf_counter(); fn(*a, threads); dt = time.perf_counter() - t0
    gpm = ms * n / dt / 1e9
    info = {"value": gpm, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d of %d points x %d modes (512 evenly spaced runs of consecutive points), %.2f s; cost is "
                      "linear in points" % (ms, m, n, dt)}
    return gpm, info, (fn, a, ms, m, n)


def run_reference(args):
    """--impl reference: the CPU restatement with all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    steps = max(1, args.steps)
    warm = max(0, args.warmup)
    # bounded sample per step: the whole run (warm-up + K steps) is sized to ~2 minutes of CPU time
    per_step = max(0.05, min(20.0, 120.0 / (steps + warm)))
    _, info, (fn, a, ms, m, n) = cpu_reference_rate(args.workload, args.scale, per_step, threads)
    for _ in range(warm):
        fn(*a, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn(*a, threads)
    dt = (time.perf_counter() - t0) / steps
    gpm = ms * n / dt / 1e9
    info.update(value=gpm, sample="%d of %d points x %d modes per step (512 evenly spaced runs of consecutive points); "
                "cost is linear in points; ms_per_step is extrapolated to the full point count" % (ms, m, n))
    line = {
        "impl": "reference", "metric": METRIC, "value": gpm, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3 * (m / ms), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_for(args, m),
        "cpu_baseline": info,
        "e2e": {"value": gpm, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import gstools_core as gc
    from gstools_core import workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or gc.device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    gc.set_devices([local])
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")   # host-side barrier for the phases that must leave the GPUs idle

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W, K = max(args.warmup, 3), max(args.steps, 1)
    shard = lambda m: gc.shard_bounds(m, world, rank)                       # noqa: E731
    w = workloads.make(args.workload, args.scale, point_range=shard if world > 1 else None)
    kind, n, m_total, m, d = w["kind"], w["n"], w["m"], w["m_local"], w["d"]
    pm_total, pm_local = n * m_total, n * m
    nc = d if kind == "summate_incompr" else 1
    margs = w["args"][:-1]
    pos_host = w["args"][-1]                                                  # plain numpy: pageable

    # The headline numbers measure the GENERAL point x mode kernel: the positions happen to be a
    # grid, which the default API would detect and route to the structured-grid GEMM path -- that
    # path is measured separately below ("structured_grid").
    gc.set_grid_detection(False)
    gc.set_profiling(False)

    # ---- value: device-resident shard.  Sets are rotated unless one set already exceeds 2x L2.
    pos_bytes, out_bytes = d * m * 8, nc * m * 8
    n_sets = 1 if pos_bytes + out_bytes > 2 * L2_BYTES else min(16, max(2, -(-2 * L2_BYTES // (pos_bytes + out_bytes))))
    dev_modes = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in margs]
    dev_pos = [torch.from_numpy(pos_host).cuda() for _ in range(n_sets)]
    oshape = (m, nc) if nc > 1 else (m,)
    dev_out = [torch.empty(oshape, dtype=torch.float64, device="cuda") for _ in range(n_sets)]
    dev_fn = getattr(gc, kind + "_device")
    stream = torch.cuda.current_stream()

    def dev_step(i):
        o = dev_out[i % n_sets]
        dev_fn(*dev_modes, dev_pos[i % n_sets], o.t() if nc > 1 else o, stream=stream.cuda_stream)

    sampler = ClockSampler(local)
    if rank == 0 and os.environ.get("GSF_BENCH_NO_SAMPLER") != "1":
        sampler.start()
        time.sleep(0.4)          # let nvidia-smi come up so that it samples the timed region
    for i in range(W):
        dev_step(i)
    barrier()
    # kernel time and step time come from ONE loop: the library records its own CUDA events around
    # every summation kernel on the launching stream (profiling mode 2 = accumulate over calls)
    gc.set_profiling(2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(K):
        dev_step(i)
    e1.record(stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1)) / K
    variant = gc.last_stats()                      # kernel_ms: summed over the K calls (accumulate mode)
    launches = K * variant["kernel_launches"]      # per call: gsf_prep_modes + gsf_sum_kernel
    kernel_ms = max_over_ranks(variant["kernel_ms"]) / K
    gc.set_profiling(False)
    # The sampler covers the device-timed region only: nvidia-smi polling takes driver locks and
    # visibly disturbs the host-synchronous end-to-end calls measured below.
    clocks = sampler.stop() if rank == 0 else {}
    del dev_pos
    torch.cuda.empty_cache()

    # ---- e2e: the reference-facing API on PAGEABLE host positions, host result
    host_fn = getattr(gc, kind)
    step_s = dev_ms * 1e-3
    Ke = K if step_s < 0.2 else max(3, min(K, 10))
    pages = [pos_host] if pos_bytes > (512 << 20) else [pos_host, pos_host.copy()]

    e2e_launches = 0

    def timed_calls(fn, arrays, k, warm):
        nonlocal e2e_launches
        # warm up the way the timed loop runs: the previous result stays alive while the next call
        # allocates its own, so two result blocks alternate in the library's pinned pool (a first-time
        # cudaMallocHost of an 800 MB result costs ~0.4 s and is not part of the steady state)
        res = None
        for i in range(warm):
            res = fn(*margs, arrays[i % len(arrays)])
        barrier()
        t0 = time.perf_counter()
        for i in range(k):
            res = fn(*margs, arrays[i % len(arrays)])
        torch.cuda.synchronize()
        local_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        e2e_launches += k * gc.last_stats()["kernel_launches"]
        return max_over_ranks(local_ms) / k, res

    # short steps run on the host's clock and a noisy neighbour moves one repetition by +-10 %:
    # three repetitions, the median is reported (all three are listed); long steps need one
    reps = []
    res = None
    for rep in range(3 if step_s < 0.05 else 1):
        ms_rep, res = timed_calls(host_fn, pages, Ke, min(W, 3) if rep == 0 else 0)
        reps.append(ms_rep)
    e2e_ms = sorted(reps)[len(reps) // 2]
    st = gc.last_stats()
    checksum = float(res.sum())
    del res

    # same call with the positions page-locked by the caller (gstools_core.pinned): the GPU reads them in place
    Kp = Ke if step_s < 0.2 else max(2, min(Ke, 5))
    pin_h = gc.pinned(pages[0])
    e2e_pinned_ms, _ = timed_calls(host_fn, pages[:1], Kp, 2)
    st_pin = gc.last_stats()
    pin_h.release()

    # ---- c2 block: configs[1] per rank (1000 modes x 1e6 points): per-call overheads
    c2 = None
    if not args.no_extras and args.scale == 1.0:
        c2 = c2_block(gc, workloads, torch, timed_calls_factory=(barrier, max_over_ranks), world=world)

    # ---- structured-grid path of the default API (different algorithmic work: reported separately)
    grid = None
    if world == 1 and w.get("axes") is not None and not args.no_extras:
        grid = grid_block(gc, torch, kind, margs, pos_host, w["axes"], nc, m, pm_local, W, 5 if step_s > 0.2 else K, local)
        gc.set_grid_detection(False)

    # ---- roofline denominator: measured DFMA issue rate (same box, same run)
    dfma_rate, dfma_ms = gc.dfma_peak(local, 300.0)

    # ---- in-process multi-device call (rank 0 drives all N GPUs; the other ranks wait on the host)
    inproc = None
    if world > 1 and not args.no_extras:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        if rank == 0:
            inproc = in_process_child(args, world)
        dist.barrier(group=cpu_group)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = peak_file()
    # DRAM bytes per launch from the committed ncu captures of this kernel: the capture whose launch
    # (points x modes) is the one timed here (N = 1: the whole workload; N ranks: one shard)
    traffic, traffic_src = None, None
    try:
        import glob
        for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_r2_*_traffic.json"))):
            with open(path) as f:
                t = json.load(f)
            if t.get("modes") == n and abs(t.get("points_per_launch", -1) - m) <= 0.01 * m:
                traffic, traffic_src = t["dram_bytes_read"] + t["dram_bytes_write"], os.path.relpath(path, ROOT)
    except Exception:
        traffic = None
    w_exec = variant["fp64_slots"]
    pm_per_s_kernel = pm_local / (kernel_ms * 1e-3)
    ach_tflops = pm_per_s_kernel * w_exec * 2 / 1e12
    peak_tflops = dfma_rate * 2 / 1e12
    line = {
        "metric": METRIC, "value": pm_total / (dev_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_for(args, m_total),
        "e2e": {"value": pm_total / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms, "steps": Ke,
                "h2d_bytes_per_step": st["h2d_bytes"] * world, "d2h_bytes_per_step": st["d2h_bytes"] * world,
                "h2d_bytes_per_step_per_rank": st["h2d_bytes"], "d2h_bytes_per_step_per_rank": st["d2h_bytes"],
                "api": "gstools_core.%s(plain pageable numpy arrays) -> host ndarray, one call per rank on its shard" % kind,
                "chunks_per_step": st["n_chunks"], "staging_threads": st["staging_threads"], "checksum": checksum,
                "repetitions_ms_per_step": reps,
                "transfer": "pageable positions staged through a pinned ring by the library's host crew, chunked "
                            "H2D / kernel / D2H pipeline; result lands in a pinned-pool ndarray",
                "pinned_input": {"value": pm_total / (e2e_pinned_ms * 1e-3) / 1e9, "unit": UNIT,
                                 "ms_per_step": e2e_pinned_ms, "steps": Kp, "chunks_per_step": st_pin["n_chunks"],
                                 "note": "same call, positions page-locked by the caller (gstools_core.pinned)"}},
        "gpu_launches": launches + e2e_launches,
        "roofline": {
            "bound": "fp64", "achieved": ach_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": ach_tflops / peak_tflops, "traffic": traffic,
            "traffic_unit": "bytes of DRAM per launch (ncu dram__bytes_read+write); algorithmic bytes: %d" % (pos_bytes + out_bytes),
            "traffic_source": traffic_src,
            "kernel": "gsf_sum_kernel<D=%d,NC=%d,P=%d,L=%d,DEG=%d>" % (d, nc, variant["points_per_thread"],
                                                                      variant["lanes_per_point"], variant["poly_degree"]),
            "kernel_ms": kernel_ms, "timed": "library CUDA events around every launch, same loop as ms_per_step",
            "fp64_slots_per_point_mode": w_exec, "poly_degree": variant["poly_degree"],
            "peak_source": "gsf_dfma_peak measured in this run: %.2f T DFMA/s over %.0f ms (of measured)"
                           % (dfma_rate / 1e12, dfma_ms),
            "survey_slots_per_point_mode": workloads.W_SURVEY[args.workload],
            "hbm_gbs_needed": (pos_bytes + out_bytes) / (kernel_ms * 1e-3) / 1e9,
            "hbm_gbs_measured_peak": peaks.get("hbm_gbs"),
        },
        "clocks": clocks,
        "detail": {
            "points_per_gpu": m, "point_modes_per_gpu": pm_local,
            "device_sets": n_sets,
            "grid_detection": "off for value / e2e / roofline (general point x mode kernel); the default behaviour "
                              "on this gridded input is reported under structured_grid / in_process.default_api",
        },
    }
    if c2 is not None:
        line["c2"] = c2
    if grid is not None:
        line["structured_grid"] = grid
    if inproc is not None:
        line["in_process"] = inproc
    if world == 1 and not args.no_cpu_baseline:
        _, info, _ = cpu_reference_rate(args.workload, args.scale, args.cpu_seconds)
        line["cpu_baseline"] = info
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def c2_block(gc, workloads, torch, timed_calls_factory, world):
    """configs[1] on every rank at once (weak): device-timed kernel, pageable and pinned e2e."""
    barrier, max_over_ranks = timed_calls_factory
    w = workloads.make("c2")
    k, z1, z2, pos = w["args"]
    pm = w["n"] * w["m"]
    dk, dz1, dz2 = (torch.from_numpy(a).cuda() for a in (k, z1, z2))
    sets = 9
    dpos = [torch.from_numpy(pos).cuda() for _ in range(sets)]
    dout = [torch.empty(w["m"], dtype=torch.float64, device="cuda") for _ in range(sets)]
    stream = torch.cuda.current_stream()
    gc.set_grid_detection(False)
    for i in range(5):
        gc.summate_device(dk, dz1, dz2, dpos[i % sets], dout[i % sets], stream=stream.cuda_stream)
    barrier()
    gc.set_profiling(2)
    K = 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        gc.summate_device(dk, dz1, dz2, dpos[i % sets], dout[i % sets], stream=stream.cuda_stream)
    e1.record(stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1)) / K
    stv = gc.last_stats()
    kernel_ms = max_over_ranks(stv["kernel_ms"]) / K
    gc.set_profiling(False)
    del dpos, dout

    def host_ms(arrays, k_steps):
        r = None
        for i in range(5):
            r = gc.summate(k, z1, z2, arrays[i % len(arrays)])
        reps = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            for i in range(k_steps):
                r = gc.summate(k, z1, z2, arrays[i % len(arrays)])
            reps.append(max_over_ranks((time.perf_counter() - t0) * 1e3) / k_steps)
        return sorted(reps)[1], reps

    pages = [pos, pos.copy()]
    page_ms, page_reps = host_ms(pages, 50)
    st = gc.last_stats()
    h0, h1 = gc.pinned(pages[0]), gc.pinned(pages[1])
    pin_ms, pin_reps = host_ms(pages, 50)
    h0.release(); h1.release()
    # opt-in auto-pin (gstools_core.set_auto_pin): the plain call on the same pageable arrays; the
    # library page-locks a position array the second time it sees it (registration happens in the
    # warm-up calls) and reads it in place from then on
    gc.set_auto_pin(256)
    auto_ms, auto_reps = host_ms(pages, 50)
    auto_mem = gc.last_stats()["pos_memory"]
    gc.set_auto_pin(0)
    out = {
        "workload": WORKLOADS["c2"][0] + " per rank (weak)",
        "kernel": {"value": world * pm / (dev_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": dev_ms, "kernel_ms": kernel_ms,
                   "poly_degree": stv["poly_degree"], "fp64_slots_per_point_mode": stv["fp64_slots"]},
        "e2e_pageable": {"value": world * pm / (page_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": page_ms,
                         "repetitions_ms_per_step": page_reps, "chunks_per_step": st["n_chunks"],
                         "staging_threads": st["staging_threads"]},
        "e2e_pinned": {"value": world * pm / (pin_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": pin_ms,
                       "repetitions_ms_per_step": pin_reps},
        "e2e_pageable_auto_pin": {"value": world * pm / (auto_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": auto_ms,
                                  "repetitions_ms_per_step": auto_reps, "pos_memory_seen_by_library": auto_mem,
                                  "note": "opt-in set_auto_pin(256 MB): repeated position arrays are page-locked in place"},
    }
    if world > 1:
        out["note"] = ("%d ranks x (24 MB in + 8 MB out) per %.2f ms of kernel = %.0f GB/s of host DRAM traffic wanted from ONE host: "
                       "this weak block is bound by the host's memory system, not by the GPUs (profiles/scaling_r2.md)"
                       % (world, kernel_ms, world * 32e6 / (kernel_ms * 1e-3) / 1e9))
    # default API on this gridded input (exact grid detection on): the structured-grid GEMM path,
    # which never moves the positions to the GPU (every rank at once when world > 1)
    gc.set_grid_detection(True)
    r = None
    for i in range(5):
        r = gc.summate(k, z1, z2, pages[i % 2])
    gp = gc.last_stats()["grid_path"]
    barrier()
    t0 = time.perf_counter()
    for i in range(50):
        r = gc.summate(k, z1, z2, pages[i % 2])
    grid_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / 50
    out["default_api_grid_path"] = {"ms_per_step": grid_ms, "value": world * pm / (grid_ms * 1e-3) / 1e9, "unit": UNIT,
                                    "grid_path": gp, "note": "gstools_core.summate(pageable pos), detection on"}
    gc.set_grid_detection(False)
    return out


def grid_block(gc, torch, kind, margs, pos_host, axes, nc, m, pm, W, K, local):
    host_fn = getattr(gc, kind)
    gc.set_grid_detection(True)
    for i in range(min(W, 2)):
        resg = host_fn(*margs, pos_host)
    assert gc.last_stats()["grid_path"] == 1
    t0 = time.perf_counter()
    for i in range(K):
        resg = host_fn(*margs, pos_host)
    torch.cuda.synchronize()
    g_e2e_ms = (time.perf_counter() - t0) * 1e3 / K
    del resg
    # kernel path: explicit axes, device-resident result, library CUDA events around the GEMM
    grid_fn = getattr(gc, kind + "_grid")
    gout = torch.empty((m, nc) if nc > 1 else (m,), dtype=torch.float64, device="cuda")
    gout = gout.t() if nc > 1 else gout
    gc.set_profiling(True)
    gk = []
    for i in range(2 + min(K, 10)):
        grid_fn(*margs, axes, out=gout)
        torch.cuda.synchronize()
        if i >= 2:
            gk.append(gc.last_stats()["kernel_ms"])
    mode_group = max(1, gc.last_stats().get("mode_group", 1))
    gc.set_profiling(False)
    g_kernel_ms = statistics.mean(gk)
    dmma_rate, dmma_ms = gc.dmma_peak(local, 300.0)
    # tensor-structured modes (Fourier lattice) are summed per group before the GEMM: the contraction,
    # and with it the algorithmic work per point*mode, shrinks by the group size
    fma_per_pm = 2 * nc / mode_group
    return {
        "note": "points form a rectilinear grid: the sum factorises per axis into an FP64 GEMM "
                "(2*NC FMA per point*mode + O(1/n_last)); same results within 1e-9 sigma",
        "e2e": {"value": pm / (g_e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": g_e2e_ms,
                "api": "gstools_core.%s(pageable host arrays) with automatic exact grid detection (default)" % kind},
        "kernel": {"value": pm / (g_kernel_ms * 1e-3) / 1e9, "unit": UNIT, "kernel_ms": g_kernel_ms,
                   "name": "gsf_grid_gemm (DMMA.8x8x4)"},
        "roofline": {"bound": "fp64 tensor", "achieved": pm * fma_per_pm * 2 / (g_kernel_ms * 1e-3) / 1e12,
                     "peak": dmma_rate * 2 / 1e12, "unit": "TFLOP/s",
                     "frac": pm * fma_per_pm / (g_kernel_ms * 1e-3) / dmma_rate,
                     "fma_per_point_mode": fma_per_pm, "mode_group": mode_group,
                     "peak_source": "gsf_dmma_peak measured in this run: %.2f T FMA/s over %.0f ms (of measured)"
                                    % (dmma_rate / 1e12, dmma_ms)},
    }


def in_process_child(args, world):
    """Run in_process_block in a child of rank 0 WITHOUT the launcher's rank environment: the library
    sizes its host crews (staging, grid verification) from LOCAL_WORLD_SIZE, and the in-process design
    is one process that owns the host and drives all the GPUs.  The other ranks idle at a host barrier."""
    env = {k: v for k, v in os.environ.items()
           if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "ROLE_WORLD_SIZE",
                        "GROUP_WORLD_SIZE", "ROLE_NAME", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS")
           and not k.startswith("TORCHELASTIC") and not k.startswith("NCCL_")}
    cmd = [sys.executable, os.path.abspath(__file__), "--in-process-child", str(world), "--workload", args.workload,
           "--scale", repr(args.scale)]
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1200)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": "child rc=%d: %s" % (r.returncode, (r.stderr or r.stdout)[-400:])}
        return json.loads(lines[-1])
    except Exception as exc:   # never lose the headline line over the extra block
        return {"error": "%s: %s" % (type(exc).__name__, exc)}


def run_in_process_child(args):
    import gstools_core as gc
    from gstools_core import workloads
    world = args.in_process_child
    if gc.device_count() < world:
        raise RuntimeError("in-process child: %d devices visible, %d wanted" % (gc.device_count(), world))
    emit(in_process_block(gc, workloads, args, world))


def in_process_block(gc, workloads, args, world):
    """ONE gstools_core call on rank 0 sharding the whole workload over all `world` devices inside
    the process (gsf_set_devices; one host thread + staging crew per device), from pageable memory;
    parity against the single-device result at every shard boundary."""
    w = workloads.make(args.workload, args.scale)
    kind, n, m = w["kind"], w["n"], w["m"]
    fn = getattr(gc, kind)
    a = w["args"]
    gc.set_grid_detection(False)
    gc.set_devices(list(range(world)))
    many = fn(*a)                                     # warm-up (allocates the per-device rings)
    many = fn(*a)                                     # ... and the second result block of the steady state
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        many = fn(*a)
        times.append((time.perf_counter() - t0) * 1e3)
    stm = gc.last_stats()
    # parity: 1024 points either side of every shard boundary + first/last 1024 + a strided sample,
    # recomputed on ONE device with the same polynomial degree; the per-point mode order does not
    # depend on the shard, so the results must be bit-identical
    idx = [np.arange(0, min(m, 1024)), np.arange(max(0, m - 1024), m), np.arange(0, m, max(1, m // 65536))]
    for r in range(1, world):
        b, _ = gc.shard_bounds(m, world, r)
        idx.append(np.arange(max(0, b - 1024), min(m, b + 1024)))
    idx = np.unique(np.concatenate(idx))
    sub = list(a)
    sub[-1] = np.ascontiguousarray(a[-1][:, idx])
    gc.set_devices([0])
    gc.set_poly_degree(stm["poly_degree"])
    one = fn(*sub)
    gc.set_poly_degree(0)
    got = many[:, idx] if many.ndim == 2 else many[idx]
    parity = bool(np.array_equal(got, one))
    max_diff = float(np.max(np.abs(got - one)))
    del many
    out = {
        "api": "one gstools_core.%s(pageable host arrays) call, gsf_set_devices(range(%d))" % (kind, world),
        "value": n * m / (min(times) * 1e-3) / 1e9, "unit": UNIT, "ms_per_call_min": min(times), "ms_per_call": times,
        "n_devices": stm["n_devices"], "chunks": stm["n_chunks"], "poly_degree": stm["poly_degree"],
        "parity_bit_identical_to_one_device": parity, "parity_points": int(idx.size), "max_abs_diff": max_diff,
    }
    if w.get("axes") is not None:
        # what the default API does on this gridded input across the devices (structured-grid path)
        gc.set_devices(list(range(world)))
        gc.set_grid_detection(True)
        fn(*a)
        t0 = time.perf_counter()
        g = fn(*a)
        ms = (time.perf_counter() - t0) * 1e3
        sg = gc.last_stats()
        gg = g[:, idx] if g.ndim == 2 else g[idx]
        out["default_api"] = {"ms_per_call": ms, "value": n * m / (ms * 1e-3) / 1e9, "grid_path": sg["grid_path"],
                              "n_devices": sg["n_devices"],
                              "max_abs_diff_vs_general_over_sigma": float(np.max(np.abs(gg - one)) / np.std(one)),
                              "note": "exact grid detection reads every position on the host (%.1f GB): at many "
                                      "devices that pass, not the GPUs, bounds the call" % (a[-1].nbytes / 1e9)}
        del g, gg
        # the same field from the axis vectors (GSTools mesh_type='structured'): no positions to verify
        grid_fn = getattr(gc, kind + "_grid")
        ga = a[:-1] + (w["axes"],)
        g2 = grid_fn(*ga)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            g2 = grid_fn(*ga)
            ts.append((time.perf_counter() - t0) * 1e3)
        g2i = g2[:, idx] if g2.ndim == 2 else g2[idx]
        out["grid_axes_api"] = {"api": "gstools_core.%s_grid(modes, axes) -> host ndarray" % kind, "ms_per_call_min": min(ts),
                                "ms_per_call": ts, "value": n * m / (min(ts) * 1e-3) / 1e9,
                                "n_devices": gc.last_stats()["n_devices"],
                                "max_abs_diff_vs_general_over_sigma": float(np.max(np.abs(g2i - one)) / np.std(one))}
        gc.set_grid_detection(False)
    return out


def claim_stdout():
    """Keep stdout to the ONE JSON line: libraries (NCCL prints its version banner on stdout) write
    to fd 1 behind Python's back, so fd 1 is pointed at stderr for the run and the line goes to a
    private duplicate of the original stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched directly with --gpus N: re-launch one process per GPU like the driver does
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                  "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                                  "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    claim_stdout()
    if args.in_process_child > 0:
        return run_in_process_child(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
