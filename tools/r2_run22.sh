#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_grid_gpu.py -m gpu -x -q > gpurun_out/s22_pytest_grid.log 2>&1; tail -3 gpurun_out/s22_pytest_grid.log
timeout 300 python tools/modegroup_probe.py > gpurun_out/s22_modegroup.log 2>&1; cat gpurun_out/s22_modegroup.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s22_launches_c4grid.csv python tools/ncu_grid_target.py c4 3 > gpurun_out/s22_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s22_launches_c4grid.csv "python tools/ncu_grid_target.py c4 3"
