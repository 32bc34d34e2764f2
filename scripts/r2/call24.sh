#!/bin/bash
# round 2, call 24 (1 GPU): Chebyshev recurrence in tile space, copy path of the sparse add, fast magnitude in the sweep: parity
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 600 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_helpers.py tests/test_gpu_multiply.py tests/test_gpu_drivers_more.py -m gpu -q --timeout 300 ) > gpurun_out/r2c24_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c24_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c24_pytest.log | grep -v "^$" | tail -n 40
timeout 300 python bench.py --config c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peaks > gpurun_out/r2c24_bench_c5.json 2> gpurun_out/r2c24_bench_c5.err; echo "c5 exit $?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2c24_bench_c5.json") if l.startswith('{')][0]; r=d["roofline"]
    print("c5 ms/step %.3f value %.0f parity %s" % (d["ms_per_step"], d["value"], d["parity_checked"] and d["parity_checked"]["ok"]))
except Exception as e: print("c5 failed", e)
PY
