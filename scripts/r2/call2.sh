#!/bin/bash
# round 2, call 2 (1 GPU): bisect the ComplexMatrix example failure; measure the A-complete block (NTB_DENSE_STAGE=2)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python scripts/r2/debug_complex.py > gpurun_out/r2c2_debug_complex.log 2>&1
echo "debug exit $?"; tail -n 30 gpurun_out/r2c2_debug_complex.log
NTB_DENSE_STAGE=2 timeout 200 python bench.py --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/r2c2_bench_dense2.json 2> gpurun_out/r2c2_bench_dense2.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r2c2_bench_dense2.json") if l.startswith('{')][0]; r=d["roofline"]
print("dense2: ms/step %.3f value %.0f numeric_share %.3f fp64_frac %.3f" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r["fp64_frac"]))
PY
