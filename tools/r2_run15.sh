#!/bin/bash
set -x
mkdir -p gpurun_out
for f in 32768 16384 8192 4096; do GSF_CHUNK_FIRST=$f timeout 300 python tools/e2e_size_sweep.py 45000,60000,100000,200000,300000,600000,1000000,3000000 >> gpurun_out/s15_sweep_first.log 2>&1; echo "^^ GSF_CHUNK_FIRST=$f" >> gpurun_out/s15_sweep_first.log; done; grep -v "^d=3" gpurun_out/s15_sweep_first.log
timeout 600 python -m pytest tests/test_small_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "not full_size" > gpurun_out/s15_pytest.log 2>&1; tail -3 gpurun_out/s15_pytest.log
