#!/bin/bash
# round 2, call 27 (4 GPUs): bench c4 on 1x4x1 with the final kernel (scaling table)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --no-cpu-baseline --no-e2e --no-peaks > gpurun_out/r2c27_bench_1x4x1.json 2> gpurun_out/r2c27_bench_1x4x1.err
echo "bench exit $?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2c27_bench_1x4x1.json") if l.startswith('{')][0]; r=d["roofline"]
    print("1x4x1 ms/step %.3f value %.0f parity %s" % (d["ms_per_step"], d["value"], d["parity_checked"] and d["parity_checked"]["ok"]), d["clocks"])
    print("   phases", r.get("step_ms_by_phase"))
except Exception as e: print("failed", e)
PY
