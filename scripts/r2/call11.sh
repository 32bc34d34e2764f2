#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 600 python -m pytest tests -m gpu -q --timeout 400 ) > gpurun_out/r2c11_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c11_pytest.log
grep -v "^  File\|site-packages" gpurun_out/r2c11_pytest.log | grep -v "^$" | tail -n 40
timeout 600 python bench.py --config c3 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2c11_bench_c3.json 2> gpurun_out/r2c11_bench_c3.err; echo "c3 exit $?"; tail -n 5 gpurun_out/r2c11_bench_c3.err
python - <<'PY'
import json
for c in ("c3",):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c11_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f numeric_share %.3f fp64_frac %s launches %d parity %s" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r.get("fp64_frac"), d["gpu_launches"], d["parity_checked"]))
        print("   phases", r.get("step_ms_by_phase"), "details", {k:v for k,v in d.get("details",{}).items() if k!="flops"})
    except Exception as e: print(c, "failed", e)
PY
