#!/bin/bash
# round-2 GPU session 2: traces, latency, pinned policy, sanitizers, ncu counters
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_grid_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "grid or c2_full or kat or degree" > gpurun_out/pytest2.log 2>&1; tail -3 gpurun_out/pytest2.log
timeout 120 python tools/trace_probe.py > gpurun_out/trace.log 2>&1
timeout 120 python tools/latency_c1.py > gpurun_out/latency_c1.log 2>&1
timeout 400 python tools/pinned_probe.py > gpurun_out/pinned_probe.log 2>&1
timeout 300 python bench.py --workload c2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_b.json 2> gpurun_out/bench_c2_b.err
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_$tool.log
done
GSF_LIB=$PWD/gstools-core_b200/gstools_core/libgsfield_asan.so LD_PRELOAD="/usr/lib/x86_64-linux-gnu/libasan.so.8 /usr/lib/x86_64-linux-gnu/libubsan.so.1" ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 UBSAN_OPTIONS=print_stacktrace=1 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_grid_gpu.py tests/test_fuzz_gpu.py -m gpu -x -q -k "not full_size" > gpurun_out/asan_gpu.log 2>&1
tail -5 gpurun_out/asan_gpu.log
# ncu: op-level FP64 counters + DRAM traffic of the headline kernel at C2 and at a C5-sized launch
timeout 600 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
  --clock-control none --import-source on -k regex:gsf_sum_kernel -c 2 -o gpurun_out/ncu_r2_sum python tools/ncu_target.py > gpurun_out/ncu_r2.log 2>&1
tail -3 gpurun_out/ncu_r2.log
timeout 600 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum \
  --clock-control none -k regex:gsf_sum_kernel -c 1 -o gpurun_out/ncu_r2_sum_c5shard python tools/ncu_target.py c5 0 0 1 0.125 > gpurun_out/ncu_r2_c5.log 2>&1
tail -3 gpurun_out/ncu_r2_c5.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
