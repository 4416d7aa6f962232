#!/bin/bash
# round-2 GPU session 19 (1 GPU): final check of the committed tree -- full suite, smoke, initcheck, default bench, reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s19_pytest_gpu.log 2>&1; tail -3 gpurun_out/s19_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s19_smoke.log 2>&1; tail -1 gpurun_out/s19_smoke.log
timeout 900 compute-sanitizer --tool initcheck python tools/sanitize_target.py > gpurun_out/s19_sanitizer_initcheck.log 2>&1; tail -3 gpurun_out/s19_sanitizer_initcheck.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/s19_bench_ref.json 2> gpurun_out/s19_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s19_bench_c5.json 2> gpurun_out/s19_bench_c5.err; echo "bench rc=$?"
