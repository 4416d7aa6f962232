#!/bin/bash
# round-2 GPU session 28 (1 GPU): last check of the final tree
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s28_pytest_gpu.log 2>&1; tail -3 gpurun_out/s28_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s28_smoke.log 2>&1; tail -1 gpurun_out/s28_smoke.log
timeout 600 python bench.py > gpurun_out/s28_bench_default.json 2> gpurun_out/s28_bench_default.err; echo "bench rc=$?"
