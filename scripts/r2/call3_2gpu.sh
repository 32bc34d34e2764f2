#!/bin/bash
# round 2, call 3 (2 GPUs): first run of the peer-memory path: parity worker (all 2-rank grids) + fallback variants + bench at 2 GPUs
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
nvidia-smi topo -m > gpurun_out/r2c3_topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 -x ) > gpurun_out/r2c3_pytest_mp.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c3_pytest_mp.log
grep -v "^  File\|site-packages" gpurun_out/r2c3_pytest_mp.log | tail -n 40
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --no-e2e > gpurun_out/r2c3_bench_2gpu.json 2> gpurun_out/r2c3_bench_2gpu.err
echo "bench exit $?"
cut -c1-1500 gpurun_out/r2c3_bench_2gpu.json; tail -n 5 gpurun_out/r2c3_bench_2gpu.err
NTB_P2P=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --no-e2e > gpurun_out/r2c3_bench_2gpu_nccl.json 2> gpurun_out/r2c3_bench_2gpu_nccl.err
echo "bench (nccl halo) exit $?"
cut -c1-400 gpurun_out/r2c3_bench_2gpu_nccl.json
