#!/bin/bash
# round 2, call 5 (2 GPUs): lazy completion + batched read-backs on the peer path: parity, then bench at 1 and 2 GPUs
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multiply.py tests/test_gpu_solvers.py -m gpu -q --timeout 500 -x ) > gpurun_out/r2c5_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c5_pytest.log
grep -v "^  File\|site-packages" gpurun_out/r2c5_pytest.log | tail -n 30
timeout 300 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r2c5_bench_1gpu.json 2> gpurun_out/r2c5_bench_1gpu.err; echo "bench1 exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-e2e > gpurun_out/r2c5_bench_2gpu.json 2> gpurun_out/r2c5_bench_2gpu.err
echo "bench2 exit $?"; grep -c "NCCL INFO" gpurun_out/r2c5_bench_2gpu.err; grep -i "nranks" gpurun_out/r2c5_bench_2gpu.err | head -3
python - <<'PY'
import json
for c in ("1gpu","2gpu"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c5_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f numeric_share %.3f fp64_frac %s launches %d waits/step %s parity %s peer %s" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r.get("fp64_frac"), d["gpu_launches"], r.get("host_waits_per_step"), d["parity_checked"] and d["parity_checked"]["ok"], r.get("peer")))
    except Exception as e: print(c, "failed", e)
PY
tail -n 5 gpurun_out/r2c5_bench_2gpu.err | cut -c1-300
