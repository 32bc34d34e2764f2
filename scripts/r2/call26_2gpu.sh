#!/bin/bash
# round 2, call 26 (2 GPUs): device-memory part list of the stacked panel (forced), clocks window of a short timed region
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 400 python -m pytest "tests/test_gpu_multi.py::test_column_split_fallbacks" -m gpu -q --timeout 300 -k device_part_list ) > gpurun_out/r2c26_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c26_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c26_pytest.log | grep -v "^$" | tail -n 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-e2e --no-peaks > gpurun_out/r2c26_bench_1x2x1.json 2> gpurun_out/r2c26_bench_1x2x1.err
echo "bench exit $?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2c26_bench_1x2x1.json") if l.startswith('{')][0]; r=d["roofline"]
    print("1x2x1 ms/step %.3f value %.0f parity %s" % (d["ms_per_step"], d["value"], d["parity_checked"] and d["parity_checked"]["ok"]), d["clocks"])
    print("   phases", r.get("step_ms_by_phase"), "alg GB/s", r["achieved"], "launches", d["gpu_launches"], "waits", r.get("host_waits_per_step"))
except Exception as e: print("failed", e)
PY
grep -v "NCCL\|^$\|Warning\|warn\|OMP_NUM\|\*\*\*" gpurun_out/r2c26_bench_1x2x1.err | tail -n 5 | cut -c1-300
