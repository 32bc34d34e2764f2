#!/bin/bash
# round 2, call 36 (1 GPU): the shipped binary once more: smoke, c5 line with its oracle check
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v NCCL | tail -n 1
timeout 90 python bench.py --config c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peaks > gpurun_out/r2c36_bench_c5.json 2> gpurun_out/r2c36_bench_c5.err; echo "c5 exit $?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2c36_bench_c5.json") if l.startswith('{')][0]
    print("c5 ms/step %.1f value %.1f parity %s" % (d["ms_per_step"], d["value"], d["parity_checked"]["ok"]))
except Exception as e: print("failed", e)
PY
