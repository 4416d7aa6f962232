#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/e2e_size_sweep.py > gpurun_out/s13_size_sweep.log 2>&1; cat gpurun_out/s13_size_sweep.log
timeout 600 python -m pytest tests/test_small_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "not full_size" > gpurun_out/s13_pytest.log 2>&1; tail -3 gpurun_out/s13_pytest.log
timeout 120 python tools/latency_c1.py > gpurun_out/s13_latency_c1.log 2>&1; cat gpurun_out/s13_latency_c1.log
