#!/bin/bash
# round-2 GPU session 3 (2 GPUs): multi-device tests + the driver's N=2 command
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/mg_devices.txt; nproc >> gpurun_out/mg_devices.txt
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_multigpu_gpu.py tests/test_grid_gpu.py -m gpu -x -q -s -k "multi or peer or two_ranks or devices" > gpurun_out/pytest_multigpu.log 2>&1; tail -5 gpurun_out/pytest_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"
tail -c 3000 gpurun_out/bench_n2.json
