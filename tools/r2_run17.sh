#!/bin/bash
# round-2 GPU session 17 (8 GPUs): final N=8 driver command + weak-C2 probe (zero-copy vs copy engines, staging threads)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?"
timeout 200 $TR --master-port 29513 tools/c2_weak_probe.py > gpurun_out/s17_weak.log 2>gpurun_out/s17_weak.err
GSF_ZERO_COPY=0 timeout 200 $TR --master-port 29514 tools/c2_weak_probe.py >> gpurun_out/s17_weak.log 2>>gpurun_out/s17_weak.err
GSF_STAGING_THREADS=2 timeout 200 $TR --master-port 29515 tools/c2_weak_probe.py >> gpurun_out/s17_weak.log 2>>gpurun_out/s17_weak.err
GSF_PAGEABLE_DIRECT=1 timeout 200 $TR --master-port 29516 tools/c2_weak_probe.py >> gpurun_out/s17_weak.log 2>>gpurun_out/s17_weak.err
cat gpurun_out/s17_weak.log
