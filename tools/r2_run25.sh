#!/bin/bash
set -x
mkdir -p gpurun_out
for s in 0 1; do for u in 2 4 8 16; do timeout 120 tools/micro/tune_s${s}_u${u} >> gpurun_out/s25_tune.log 2>&1; done; done
cat gpurun_out/s25_tune.log
