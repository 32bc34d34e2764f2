#!/bin/bash
# round 2, call 19 (1 GPU): stage descriptors precomputed and TMA'd (copy warp computes nothing): parity, bench c4, c3, c5 (oracle check, scale 0.001)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 400 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_tile_space.py tests/test_gpu_solvers.py -m gpu -q --timeout 300 -x ) > gpurun_out/r2c19_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c19_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c19_pytest.log | grep -v "^$" | tail -n 8
timeout 200 python bench.py --no-e2e --no-check --no-cpu-baseline > gpurun_out/r2c19_bench_c4.json 2> gpurun_out/r2c19_bench_c4.err; echo "c4 exit $?"
timeout 300 python bench.py --config c3 --steps 2 --warmup 1 --no-e2e --no-check --no-cpu-baseline > gpurun_out/r2c19_bench_c3.json 2> gpurun_out/r2c19_bench_c3.err; echo "c3 exit $?"
timeout 300 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c19_bench_c5.json 2> gpurun_out/r2c19_bench_c5.err; echo "c5 exit $?"
python - <<'PY'
import json
for c in ("c4","c3","c5"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c19_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f fp64_frac %s hbm_frac %.3f launches %s parity %s" % (d["ms_per_step"], d["value"], r.get("fp64_frac"), r["frac"], d.get("gpu_launches"), d.get("parity_checked")))
        print("   phases", r.get("step_ms_by_phase"), d.get("details"))
        if d.get("e2e"): print("   e2e", d["e2e"].get("value"), d["e2e"].get("ms_per_step"))
    except Exception as e: print(c, "failed", e)
PY
grep -v "NCCL\|^$" gpurun_out/r2c19_bench_c5.err | tail -n 5
