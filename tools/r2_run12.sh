#!/bin/bash
# round-2 GPU session 12 (1 GPU): launch floor with streaming-store gather, C1 latency after the lean single-launch path
set -x
mkdir -p gpurun_out
timeout 120 tools/micro/launch_floor > gpurun_out/s12_launch_floor.txt 2>&1; cat gpurun_out/s12_launch_floor.txt
timeout 600 python -m pytest tests/test_small_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "not full_size" > gpurun_out/s12_pytest.log 2>&1; tail -3 gpurun_out/s12_pytest.log
timeout 120 python tools/latency_c1.py > gpurun_out/s12_latency_c1.log 2>&1; cat gpurun_out/s12_latency_c1.log
timeout 120 python tools/trace_probe.py > gpurun_out/s12_trace.log 2>&1; grep -A9 "== C1 call 4$" gpurun_out/s12_trace.log
