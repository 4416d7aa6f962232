#!/bin/bash
# round-2 GPU session 29 (4 GPUs): the driver's N = 4 command on the final tree
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "rc=$?"
tail -c 600 gpurun_out/bench_n4.json
