#!/bin/bash
# multi-GPU check: parity worker on every grid of this world size, then the bench at N ranks
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 900 python -m pytest tests -m gpu -q --timeout 800 ) > gpurun_out/mp${N}_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/mp${N}_pytest.log
timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/mp${N}_bench_1.json 2> gpurun_out/mp${N}_bench_1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/mp${N}_bench_${N}.json 2> gpurun_out/mp${N}_bench_${N}.err
NTB_HALO_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e > gpurun_out/mp${N}_halo_timing.json 2> gpurun_out/mp${N}_halo_timing.err
tail -n 7 gpurun_out/mp${N}_pytest.log
for f in gpurun_out/mp${N}_bench_1.json gpurun_out/mp${N}_bench_${N}.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(" n_gpus %d ms/step %.3f value %.0f numeric_share %.3f e2e %s" % (d["n_gpus"], d["ms_per_step"], d["value"], r["numeric_share_of_step"], d.get("e2e")))
except Exception as e: print(" failed", e)
PY
done
tail -n 5 gpurun_out/mp${N}_bench_${N}.err
