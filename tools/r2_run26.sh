#!/bin/bash
# round-2 GPU session 26 (1 GPU): after the degree-5 schedule retune -- full suite, smoke, default bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s26_pytest_gpu.log 2>&1; tail -3 gpurun_out/s26_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s26_smoke.log 2>&1; tail -1 gpurun_out/s26_smoke.log
timeout 300 python tools/variant_p_probe.py > gpurun_out/s26_variant_p.log 2>&1; cat gpurun_out/s26_variant_p.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s26_bench_c5.json 2> gpurun_out/s26_bench_c5.err; echo "bench rc=$?"
