#!/bin/bash
# round-2 GPU session 6 (1 GPU): launch floor, C1 latency with promotion, grid e2e sweep, P variants, full suite, benches
set -x
mkdir -p gpurun_out
timeout 120 tools/micro/launch_floor > gpurun_out/s6_launch_floor.txt 2>&1; cat gpurun_out/s6_launch_floor.txt
timeout 600 python -m pytest tests/test_small_gpu.py -m gpu -x -q > gpurun_out/s6_pytest_small.log 2>&1; tail -3 gpurun_out/s6_pytest_small.log
timeout 120 python tools/latency_c1.py > gpurun_out/s6_latency_c1.log 2>&1; cat gpurun_out/s6_latency_c1.log
GSF_SMALL_PROMOTE=0 timeout 120 python tools/latency_c1.py > gpurun_out/s6_latency_c1_nopromote.log 2>&1
timeout 120 python tools/trace_probe.py > gpurun_out/s6_trace.log 2>&1
timeout 900 python tools/grid_e2e_probe.py > gpurun_out/s6_grid_e2e.log 2>&1
timeout 300 python tools/variant_p_probe.py > gpurun_out/s6_variant_p.log 2>&1; cat gpurun_out/s6_variant_p.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s6_pytest_gpu.log 2>&1; tail -3 gpurun_out/s6_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s6_bench_c5.json 2> gpurun_out/s6_bench_c5.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/s6_bench_ref.json 2> gpurun_out/s6_bench_ref.err
