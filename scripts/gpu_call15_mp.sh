#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 380 ) > gpurun_out/c15_pytest_mp.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c15_pytest_mp.log
grep -v "^  File\|site-packages" gpurun_out/c15_pytest_mp.log | tail -n 25
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 > gpurun_out/c15_bench_2gpu.json 2> gpurun_out/c15_bench_2gpu.err
echo "bench exit $?"
cut -c1-1200 gpurun_out/c15_bench_2gpu.json; tail -n 3 gpurun_out/c15_bench_2gpu.err
