#!/bin/bash
# round-2 GPU session 8 (1 GPU): ncu traffic of the full C5 launch, launch list of the default bench command, per-config report, full suite
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:gsf_sum_kernel -c 1 --csv --log-file gpurun_out/ncu_r2_c5_full.csv python tools/ncu_target.py c5 0 0 1 1.0 > gpurun_out/s8_ncu_c5.log 2>&1; tail -3 gpurun_out/s8_ncu_c5.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2_c5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/s8_bench_under_ncu.log 2>&1; tail -2 gpurun_out/s8_bench_under_ncu.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s8_pytest_gpu.log 2>&1; tail -3 gpurun_out/s8_pytest_gpu.log
timeout 1500 python tests/measure/report_configs.py > gpurun_out/s8_report_configs.log 2>&1; tail -5 gpurun_out/s8_report_configs.log | cut -c1-400
cp profiles/configs_r2.md profiles/configs_r2.jsonl gpurun_out/ 2>/dev/null
