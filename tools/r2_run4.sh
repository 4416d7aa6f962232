#!/bin/bash
# round-2 GPU session 4 (8 GPUs): host topology, the driver's N=8 command (C5 strong), weak C2 block inside it
set -x
mkdir -p gpurun_out
{ nvidia-smi -L; nproc; free -g; lscpu | grep -E "Model name|^CPU\(s\)|NUMA|L3|Socket|Thread"; nvidia-smi topo -m; numactl -H 2>/dev/null; 
  for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; 
  cat /proc/self/status | grep -i allowed; } > gpurun_out/topo8.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?"
tail -c 1500 gpurun_out/bench_n8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "rc=$?"
