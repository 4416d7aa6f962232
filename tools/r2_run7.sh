#!/bin/bash
# round-2 GPU session 7 (1 GPU): stream-K grid GEMM
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_grid_gpu.py -m gpu -x -q > gpurun_out/s7_pytest_grid.log 2>&1; tail -5 gpurun_out/s7_pytest_grid.log
timeout 900 python tools/grid_sk_probe.py > gpurun_out/s7_grid_sk.log 2>&1; cat gpurun_out/s7_grid_sk.log
timeout 300 python tools/grid_probe.py > gpurun_out/s7_grid_probe.log 2>&1; tail -12 gpurun_out/s7_grid_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gsf_grid_gemm -s 2 -c 1 -o gpurun_out/ncu_r2_grid_c2 python tools/ncu_grid_target.py c2 3 > gpurun_out/s7_ncu_grid.log 2>&1; tail -3 gpurun_out/s7_ncu_grid.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s7_pytest_gpu.log 2>&1; tail -3 gpurun_out/s7_pytest_gpu.log
timeout 300 python bench.py --workload c2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/s7_bench_c2.json 2> gpurun_out/s7_bench_c2.err
