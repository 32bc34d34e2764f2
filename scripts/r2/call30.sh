#!/bin/bash
# round 2, call 30 (1 GPU): fused difference norm with L1 prefetch: parity, then the FINAL c4 line (full bench) and a launch list
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 600 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_solvers.py -m gpu -q --timeout 300 ) > gpurun_out/r2c30_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c30_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c30_pytest.log | grep -v "^$" | tail -n 8
timeout 400 python bench.py > gpurun_out/r2c30_bench_c4.json 2> gpurun_out/r2c30_bench_c4.err; echo "c4 exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c30_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check --no-peaks > gpurun_out/r2c30_ncu_list.out 2>&1; echo "ncu list exit $?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2c30_bench_c4.json") if l.startswith('{')][0]; r=d["roofline"]
    print("c4 ms/step %.3f value %.0f fp64_frac %s hbm_frac %s launches %s parity %s" % (d["ms_per_step"], d["value"], r.get("fp64_frac"), r.get("frac"), d.get("gpu_launches"), d.get("parity_checked")))
    print("   phases", r.get("step_ms_by_phase"))
    print("   e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["streaming"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
except Exception as e: print("failed", e)
PY
