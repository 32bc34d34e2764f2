#!/bin/bash
# round 2, call 13 (2 GPUs): plan kernel + ends-first order + batched kmeta loads: parity at 2 ranks, hash-bin test, bench at 2
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multiply.py tests/test_gpu_tile_space.py tests/test_gpu_solvers.py -m gpu -q --timeout 500 ) > gpurun_out/r2c13_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c13_pytest.log
grep -v "^  File\|site-packages" gpurun_out/r2c13_pytest.log | grep -v "^$" | tail -n 30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-e2e --no-check > gpurun_out/r2c13_bench_2gpu.json 2> gpurun_out/r2c13_bench_2gpu.err
echo "bench2 exit $?"
timeout 300 python bench.py --no-e2e --no-check --no-cpu-baseline > gpurun_out/r2c13_bench_1gpu.json 2> gpurun_out/r2c13_bench_1gpu.err
python - <<'PY'
import json
for c in ("1gpu","2gpu"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c13_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f launches %d waits/step %s" % (d["ms_per_step"], d["value"], d["gpu_launches"], r.get("host_waits_per_step")))
        print("   phases", r.get("step_ms_by_phase"))
    except Exception as e: print(c, "failed", e)
PY
