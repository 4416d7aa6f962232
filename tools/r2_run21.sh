#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s21_launches_c4grid.csv python tools/ncu_grid_target.py c4 3 > gpurun_out/s21_ncu.log 2>&1; tail -2 gpurun_out/s21_ncu.log
python tools/launch_summary.py gpurun_out/s21_launches_c4grid.csv "python tools/ncu_grid_target.py c4 3"
GSF_TRACE=1 timeout 120 python tools/ncu_grid_target.py c4 4 2>&1 | tail -12
