#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/mpb${N}_bench.json 2> gpurun_out/mpb${N}_bench.err
echo "exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/mpb${N}_ref.json 2> gpurun_out/mpb${N}_ref.err
echo "exit $?"
cat gpurun_out/mpb${N}_bench.json | cut -c1-400; tail -n 3 gpurun_out/mpb${N}_bench.err; cat gpurun_out/mpb${N}_ref.json | cut -c1-200
