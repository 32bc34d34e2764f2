#!/bin/bash
# round 2, call 14 (1 GPU): byte-granular ring + interior epilogue: parity tests, then A/B bench c4 (NTB_RING=0 vs 1), c3
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 600 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_tile_space.py tests/test_gpu_solvers.py -m gpu -q --timeout 300 -x ) > gpurun_out/r2c14_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c14_pytest.log
grep -v "^  File\|site-packages" gpurun_out/r2c14_pytest.log | grep -v "^$" | tail -n 15
NTB_RING=0 timeout 300 python bench.py --no-e2e --no-check --no-cpu-baseline > gpurun_out/r2c14_bench_ring0.json 2> gpurun_out/r2c14_bench_ring0.err; echo "ring0 exit $?"
NTB_RING=1 timeout 300 python bench.py --no-e2e --no-check --no-cpu-baseline > gpurun_out/r2c14_bench_ring1.json 2> gpurun_out/r2c14_bench_ring1.err; echo "ring1 exit $?"
NTB_RING=1 timeout 300 python bench.py --config c3 --steps 2 --warmup 1 --no-e2e --no-check --no-cpu-baseline > gpurun_out/r2c14_bench_c3.json 2> gpurun_out/r2c14_bench_c3.err; echo "c3 exit $?"
python - <<'PY'
import json
for c in ("ring0","ring1","c3"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c14_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f fp64_frac %s hbm_frac %.3f" % (d["ms_per_step"], d["value"], r.get("fp64_frac"), r["frac"]))
        print("   phases", r.get("step_ms_by_phase"))
    except Exception as e: print(c, "failed", e)
PY
