#!/bin/bash
# round-2 GPU session 11 (1 GPU): short-tile tail sweep with the degree-5 kernels, full suite, final benches
set -x
mkdir -p gpurun_out
for tw in 0.1 0.25 0.35 0.5 0.75 1.0; do GSF_TAIL_WAVES=$tw timeout 200 python tools/tail_probe.py >> gpurun_out/s11_tail.log 2>&1; done; cat gpurun_out/s11_tail.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s11_pytest_gpu.log 2>&1; tail -3 gpurun_out/s11_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s11_bench_c5.json 2> gpurun_out/s11_bench_c5.err; echo "bench rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s11_smoke.log 2>&1; tail -3 gpurun_out/s11_smoke.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py > gpurun_out/s11_sanitizer_$tool.log 2>&1
  tail -3 gpurun_out/s11_sanitizer_$tool.log
done
