#!/bin/bash
# round-2 GPU session 27 (2 GPUs): mode groups + multi-device grid path; multi-GPU tests on the final tree
set -x
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/s27_modegroup_2gpu.log 2>&1 <<'PY'
import sys, time
sys.path[:0] = [".", "gstools-core_b200"]
import numpy as np, gstools_core as gc, oracle
from gstools_core import workloads
w = workloads.make("c4", 1.0 / 16)
a = w["args"]
gc.set_devices([0]); one = gc.summate_fourier(*a); s1 = gc.last_stats()
gc.set_devices([0, 1]); two = gc.summate_fourier(*a); s2 = gc.last_stats()
t0 = time.perf_counter(); two = gc.summate_fourier(*a); dt = (time.perf_counter() - t0) * 1e3
print("one device:", s1["grid_path"], s1["mode_group"], s1["n_devices"], "| two devices:", s2["grid_path"], s2["mode_group"], s2["n_devices"], "%.3f ms" % dt)
idx = np.arange(0, w["m"], 97)
ref = oracle.summate_fourier(a[0], a[1], a[2], a[3], np.ascontiguousarray(a[4][:, idx]), oracle.max_threads())
print("max|two - one| =", float(np.max(np.abs(two - one))), " max|two - oracle|/sigma =", float(np.max(np.abs(two[idx] - ref)) / np.std(ref)))
assert s2["n_devices"] == 2 and s2["mode_group"] == 100 and np.max(np.abs(two[idx] - ref)) <= 1e-9 * np.std(ref)
print("OK")
PY
cat gpurun_out/s27_modegroup_2gpu.log
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_multigpu_gpu.py tests/test_grid_gpu.py -m gpu -x -q -k "multi or peer or two_ranks or devices" > gpurun_out/pytest_multigpu.log 2>&1; tail -4 gpurun_out/pytest_multigpu.log
