#!/bin/bash
# round 2, call 37 (1 GPU): bench.py after the last (python-only) edit: the line still prints
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 80 python bench.py --no-e2e --no-cpu-baseline --no-peaks --no-check --steps 10 > gpurun_out/r2c37_bench_c4.json 2> gpurun_out/r2c37_bench_c4.err; echo "c4 exit $?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2c37_bench_c4.json") if l.startswith('{')][0]; r=d["roofline"]
    print("c4 ms/step %.3f value %.0f" % (d["ms_per_step"], d["value"]), r.get("fused_norms_in_timed_region"), r.get("numeric_time_includes")[:60])
except Exception as e: print("failed", e)
PY
grep -v "NCCL\|^$" gpurun_out/r2c37_bench_c4.err | tail -n 3
