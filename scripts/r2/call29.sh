#!/bin/bash
# round 2, call 29 (1 GPU): difference norm fused into the epilogue of the second product of a sign step: parity, bench c4
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 600 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_solvers.py tests/test_gpu_tile_space.py -m gpu -q --timeout 300 ) > gpurun_out/r2c29_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c29_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c29_pytest.log | grep -v "^$" | tail -n 30
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-peaks > gpurun_out/r2c29_bench_c4.json 2> gpurun_out/r2c29_bench_c4.err; echo "c4 exit $?"
NTB_FUSED_NORM=0 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-peaks --no-check > gpurun_out/r2c29_bench_c4_unfused.json 2> gpurun_out/r2c29_bench_c4_unfused.err; echo "c4 unfused exit $?"
python - <<'PY'
import json
for c in ("c4","c4_unfused"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c29_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f launches %s parity %s" % (d["ms_per_step"], d["value"], d.get("gpu_launches"), d.get("parity_checked")))
        print("   phases", r.get("step_ms_by_phase"))
    except Exception as e: print(c, "failed", e)
PY
grep -v "NCCL\|^$" gpurun_out/r2c29_bench_c4.err | tail -n 3
