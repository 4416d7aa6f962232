#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/e2e_size_sweep.py > gpurun_out/s14_sweep_default.log 2>&1; cat gpurun_out/s14_sweep_default.log
for kb in 800 1600 3200; do GSF_SMALL_KB=$kb timeout 300 python tools/e2e_size_sweep.py 16000,20000,30000,45000,60000,100000 >> gpurun_out/s14_sweep_small.log 2>&1; done; cat gpurun_out/s14_sweep_small.log
for t in 2 4 6 8 12; do GSF_STAGING_THREADS=$t timeout 300 python tools/e2e_size_sweep.py 300000,1000000,3000000,10000000 >> gpurun_out/s14_sweep_threads.log 2>&1; done; cat gpurun_out/s14_sweep_threads.log
