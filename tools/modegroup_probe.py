"""C4 (summate_fourier, 1e4 lattice modes x 4096^2 grid points): structured-grid path with and without
summing the mode groups (GSF_MODE_GROUPS=0): GEMM kernel time (device-resident result), explicit-axes
call and the default API from pageable host arrays."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, torch, gstools_core as gc
from gstools_core import workloads
w = workloads.make("c4")
sf, k, z1, z2, pos = w["args"]
axes, m = w["axes"], w["m"]
print("GSF_MODE_GROUPS=%s" % os.environ.get("GSF_MODE_GROUPS", "-"))
out = torch.empty(m, dtype=torch.float64, device="cuda")
gc.set_profiling(True)
ks, calls = [], []
for i in range(8):
    t0 = time.perf_counter(); gc.summate_fourier_grid(sf, k, z1, z2, axes, out=out); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if i >= 2: ks.append(gc.last_stats()["kernel_ms"]); calls.append(dt * 1e3)
st = gc.last_stats()
print("device-resident result: call %.3f ms, GEMM kernel %.3f ms, mode_group %d, launches %d" % (min(calls), min(ks), st["mode_group"], st["kernel_launches"]))
gc.set_profiling(False)
r = None
for _ in range(2): r = gc.summate_fourier_grid(sf, k, z1, z2, axes)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); r = gc.summate_fourier_grid(sf, k, z1, z2, axes); ts.append((time.perf_counter() - t0) * 1e3)
print("explicit axes -> host ndarray: %.3f ms (min of 5)" % min(ts))
for _ in range(2): r = gc.summate_fourier(sf, k, z1, z2, pos)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); r = gc.summate_fourier(sf, k, z1, z2, pos); ts.append((time.perf_counter() - t0) * 1e3)
st = gc.last_stats()
print("default API (pageable pos, exact detection) -> host ndarray: %.3f ms (min of 5), grid_path %d mode_group %d chunks %d" % (min(ts), st["grid_path"], st["mode_group"], st["n_chunks"]))
