#!/bin/bash
# round-2 GPU session 18 (8 GPUs): weak C2 from pageable memory -- cached staging copies + ring small enough for the LLC
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29513 tools/c2_weak_probe.py pageable > gpurun_out/s18_weak.log 2>gpurun_out/s18_weak.err
for v in "GSF_STAGE_CACHED=1" "GSF_STAGE_CACHED=1 GSF_CHUNK_CAP=65536 GSF_RING_DEPTH=3" "GSF_STAGE_CACHED=1 GSF_CHUNK_CAP=32768 GSF_RING_DEPTH=4" "GSF_CHUNK_CAP=65536 GSF_RING_DEPTH=3" "GSF_STAGE_CACHED=1 GSF_CHUNK_CAP=65536 GSF_RING_DEPTH=3 GSF_STAGING_THREADS=2"; do
  echo "## $v" >> gpurun_out/s18_weak.log
  env $v timeout 200 $TR --master-port 29514 tools/c2_weak_probe.py pageable >> gpurun_out/s18_weak.log 2>>gpurun_out/s18_weak.err
done
cat gpurun_out/s18_weak.log
