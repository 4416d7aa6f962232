"""GPU probe of the structured-grid path: kernel time (library CUDA events) and warm e2e per config."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch
import gstools_core as gc
from gstools_core import workloads

rate, _ = gc.dfma_peak(0, 200.0)
print("dfma peak %.2f T/s" % (rate / 1e12))
for cfg in ("c2", "c3", "c4", "c5"):
    w = workloads.make(cfg)
    fn = getattr(gc, w["kind"])
    pm = w["m"] * w["n"]
    args = list(w["args"])
    pin = torch.from_numpy(args[-1]).pin_memory().numpy()
    for label, pos in (("pageable", args[-1]), ("pinned", pin)):
        a = args[:-1] + [pos]
        for det in (True, False):
            if not det and cfg in ("c4", "c5") and label == "pageable":
                continue
            gc.set_grid_detection(det)
            gc.set_profiling(True)
            fn(*a); fn(*a)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); out = fn(*a); ts.append(time.perf_counter() - t0)
            st = gc.last_stats()
            t = min(ts)
            print("%s %-8s grid=%d: e2e %.2f ms (%.0f Gpm/s)  kernel_ms(sum over chunks) %.3f  launches=%d chunks=%d"
                  % (cfg, label, st["grid_path"], t * 1e3, pm / t / 1e9, st["kernel_ms"], st["kernel_launches"], st["n_chunks"]), flush=True)
            gc.set_profiling(False)
            del out
gc.set_grid_detection(None)

# ---- kernel-only: explicit axes, device-resident output, one launch
def axes_of(cfg):
    if cfg in ("c2", "c3"): return [np.arange(100.0)] * 3
    if cfg == "c4": return [np.arange(4096) * (100.0 / 4096)] * 2
    if cfg == "c5": return [np.arange(1000) * 0.1, np.arange(1000) * 0.1, np.arange(100) * 0.1]
print("kernel-only (explicit axes, device output):")
for cfg in ("c2", "c3", "c4", "c5"):
    w = workloads.make(cfg, 0.001)          # modes only; positions come from the axes
    axes = axes_of(cfg)
    m = int(np.prod([len(a) for a in axes]))
    pm = m * w["n"]
    kind = w["kind"]
    fn = getattr(gc, kind + "_grid")
    margs = w["args"][:-1]
    nc = 3 if kind == "summate_incompr" else 1
    out = torch.empty((m, nc) if nc > 1 else (m,), dtype=torch.float64, device="cuda")
    o = out.t() if nc > 1 else out
    gc.set_profiling(True)
    fn(*margs, axes, out=o); fn(*margs, axes, out=o)
    ts, ks = [], []
    for _ in range(5):
        t0 = time.perf_counter(); fn(*margs, axes, out=o); ts.append(time.perf_counter() - t0)
        ks.append(gc.last_stats()["kernel_ms"])
    gc.set_profiling(False)
    km = min(ks)
    print("%s: call %.3f ms, gemm kernel %.3f ms -> %.0f Gpm/s, %.2f T FMA/s (2*NC FMA/pm)  launches=%d"
          % (cfg, min(ts) * 1e3, km, pm / km / 1e6, 2 * (3 if kind == "summate_incompr" else 1) * pm / km / 1e9, gc.last_stats()["kernel_launches"]), flush=True)
