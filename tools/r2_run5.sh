#!/bin/bash
# round-2 GPU session 5 (1 GPU): new one-launch small path + native binding, grid zero-copy output, full gpu suite, benches
set -x
mkdir -p gpurun_out
nproc > gpurun_out/s5_host.txt; lscpu | grep -E "Model name|^CPU\(s\)|NUMA|L3" >> gpurun_out/s5_host.txt
timeout 600 python -m pytest tests/test_small_gpu.py tests/test_grid_gpu.py -m gpu -x -q > gpurun_out/s5_pytest_new.log 2>&1; tail -3 gpurun_out/s5_pytest_new.log
timeout 120 python tools/latency_c1.py > gpurun_out/s5_latency_c1.log 2>&1; cat gpurun_out/s5_latency_c1.log
GSF_NATIVE_BINDING=0 timeout 120 python tools/latency_c1.py > gpurun_out/s5_latency_c1_ctypes.log 2>&1
GSF_SMALL_FUSED=0 timeout 120 python tools/latency_c1.py > gpurun_out/s5_latency_c1_twolaunch.log 2>&1
timeout 120 python tools/trace_probe.py > gpurun_out/s5_trace.log 2>&1
GSF_GRID_ZC_OUT_MB=0 timeout 120 python tools/trace_probe.py > gpurun_out/s5_trace_chunked.log 2>&1
timeout 300 python bench.py --workload c2 --steps 100 --warmup 5 > gpurun_out/s5_bench_c2.json 2> gpurun_out/s5_bench_c2.err
GSF_GRID_ZC_OUT_MB=0 timeout 300 python bench.py --workload c2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/s5_bench_c2_chunked.json 2> gpurun_out/s5_bench_c2_chunked.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s5_pytest_gpu.log 2>&1; tail -3 gpurun_out/s5_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s5_bench_c5.json 2> gpurun_out/s5_bench_c5.err; echo "bench rc=$?"
