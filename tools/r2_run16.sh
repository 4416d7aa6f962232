#!/bin/bash
# round-2 GPU session 16 (1 GPU): checkpoint -- full suite, smoke, per-config report, default bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s16_pytest_gpu.log 2>&1; tail -3 gpurun_out/s16_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s16_smoke.log 2>&1; tail -2 gpurun_out/s16_smoke.log | cut -c1-200
timeout 120 python tools/latency_c1.py > gpurun_out/s16_latency_c1.log 2>&1; head -2 gpurun_out/s16_latency_c1.log
timeout 1500 python tests/measure/report_configs.py > gpurun_out/s16_report_configs.log 2>&1; tail -3 gpurun_out/s16_report_configs.log | cut -c1-300
cp profiles/configs_r2.md profiles/configs_r2.jsonl gpurun_out/ 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s16_bench_c5.json 2> gpurun_out/s16_bench_c5.err; echo "bench rc=$?"
