#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( timeout 40 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_solvers.py -m gpu -q -x --timeout 35 ) > gpurun_out/c17_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c17_pytest.log
grep -v "^  File\|site-packages" gpurun_out/c17_pytest.log | tail -n 6
timeout 40 python bench.py --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err
echo "bench exit $?"; tail -n 2 gpurun_out/c17_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/c17_bench.json") if l.startswith('{')][0]; r=d["roofline"]
print("ms/step %.3f value %.0f numeric_share %.3f fp64_frac %.3f hbm_frac %.3f" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r["fp64_frac"], r["frac"]))
PY
