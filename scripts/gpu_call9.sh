#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/c9_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c9_pytest.log
timeout 600 python bench.py > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c9_bench_ref.json 2> gpurun_out/c9_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c9_smoke.log 2>&1
grep -v "^  File\|site-packages" gpurun_out/c9_pytest.log | tail -n 12
python - gpurun_out/c9_bench.json <<'PY'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][0]; r=d["roofline"]
    print(sys.argv[1], " ms/step %.3f value %.0f numeric_share %.3f fp64_frac %.3f hbm_frac %.3f e2e %s" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r["fp64_frac"], r["frac"], d.get("e2e")))
    print(d["cpu_baseline"])
except Exception as e: print(" failed", e)
PY
tail -n 3 gpurun_out/c9_bench.err; cat gpurun_out/c9_bench_ref.json | cut -c1-300; tail -n 2 gpurun_out/c9_smoke.log
