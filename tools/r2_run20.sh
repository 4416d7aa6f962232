#!/bin/bash
# round-2 GPU session 20 (1 GPU): tensor-structured modes (mode groups) on the structured-grid path
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_grid_gpu.py -m gpu -x -q > gpurun_out/s20_pytest_grid.log 2>&1; tail -5 gpurun_out/s20_pytest_grid.log
timeout 300 python tools/modegroup_probe.py > gpurun_out/s20_modegroup.log 2>&1; cat gpurun_out/s20_modegroup.log
GSF_MODE_GROUPS=0 timeout 300 python tools/modegroup_probe.py >> gpurun_out/s20_modegroup.log 2>&1; tail -6 gpurun_out/s20_modegroup.log
