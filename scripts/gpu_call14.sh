#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 600 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/c14_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c14_pytest.log
grep -v "^  File\|site-packages" gpurun_out/c14_pytest.log | tail -n 30
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
tail -n 5 gpurun_out/c14_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/c14_bench.json") if l.startswith('{')][0]
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"])
PY
