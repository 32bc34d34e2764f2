#!/bin/bash
# round 2, call 4 (1 GPU): full GPU test suite after the peer/LeftView refactor; new bench.py on all four configs
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 500 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/r2c4_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c4_pytest.log
grep -v "^  File\|site-packages" gpurun_out/r2c4_pytest.log | tail -n 25
timeout 400 python bench.py > gpurun_out/r2c4_bench_c4.json 2> gpurun_out/r2c4_bench_c4.err; echo "c4 exit $?"; tail -n 3 gpurun_out/r2c4_bench_c4.err
timeout 200 python bench.py --config c1 > gpurun_out/r2c4_bench_c1.json 2> gpurun_out/r2c4_bench_c1.err; echo "c1 exit $?"; tail -n 3 gpurun_out/r2c4_bench_c1.err
timeout 600 python bench.py --config c3 --steps 3 --warmup 1 > gpurun_out/r2c4_bench_c3.json 2> gpurun_out/r2c4_bench_c3.err; echo "c3 exit $?"; tail -n 5 gpurun_out/r2c4_bench_c3.err
timeout 400 python bench.py --config c5 --steps 2 --warmup 1 > gpurun_out/r2c4_bench_c5.json 2> gpurun_out/r2c4_bench_c5.err; echo "c5 exit $?"; tail -n 5 gpurun_out/r2c4_bench_c5.err
python - <<'PY'
import json
for c in ("c4","c1","c3","c5"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c4_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f numeric_share %.3f fp64_frac %s hbm_frac %.3f launches %d" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r.get("fp64_frac"), r["frac"], d["gpu_launches"]))
        print("   e2e", d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "parity", d["parity_checked"] and d["parity_checked"]["ok"], "peaks", r.get("fp64_peaks_measured_in_this_run"))
        print("   details", d.get("details"))
    except Exception as e: print(c, "failed", e)
PY
