#!/bin/bash
# round-2 GPU session 1: host copy probe, full gpu test-suite, bench (C5 default + C2), degree probe, staging sweep
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)|NUMA|L3" >> gpurun_out/nproc.txt
nvidia-smi topo -m >> gpurun_out/nproc.txt 2>&1
timeout 300 tools/micro/host_copy_probe > gpurun_out/host_copy_probe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench rc=$?"
timeout 300 python bench.py --workload c2 --steps 200 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python tests/measure/poly_degree_probe.py > gpurun_out/poly_degree.log 2>&1
for t in 1 2 3 4 6 8; do GSF_STAGING_THREADS=$t timeout 120 python tools/pageable_probe.py >> gpurun_out/pageable_sweep.log 2>&1; done
cp profiles/poly_degree_r2.md gpurun_out/ 2>/dev/null
tail -3 gpurun_out/bench_c5.json | cut -c1-1500
