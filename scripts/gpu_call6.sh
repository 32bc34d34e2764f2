#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/c6_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c6_pytest.log
timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
NTB_TILE_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c6_phase.json 2> gpurun_out/c6_phase.err
tail -n 30 gpurun_out/c6_pytest.log
python - gpurun_out/c6_bench.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(" ms/step %.3f value %.0f numeric_share %.3f fp64_frac %.3f" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r["fp64_frac"]))
except Exception as e: print(" failed", e)
PY
tail -n 5 gpurun_out/c6_bench.err; tail -n 4 gpurun_out/c6_phase.err
