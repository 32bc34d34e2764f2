#!/bin/bash
# round 2, call 32 (1 GPU): FINAL c4 line of the shipped build (fused difference norm), ncu of the numeric launches, launch list
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 400 python bench.py > gpurun_out/r2c32_bench_c4.json 2> gpurun_out/r2c32_bench_c4.err; echo "c4 exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile_numeric -s 12 -c 4 -f -o gpurun_out/r2c32_numeric \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check --no-peaks > gpurun_out/r2c32_ncu_numeric.out 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/r2c32_numeric.ncu-rep --page source --csv > gpurun_out/r2c32_numeric_source.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c32_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check --no-peaks > gpurun_out/r2c32_ncu_list.out 2>&1; echo "ncu list exit $?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2c32_bench_c4.json") if l.startswith('{')][0]; r=d["roofline"]
    print("c4 ms/step %.3f value %.0f fp64_frac %s hbm_frac %s parity %s" % (d["ms_per_step"], d["value"], r.get("fp64_frac"), r.get("frac"), d["parity_checked"]["ok"]))
    print("   phases", r.get("step_ms_by_phase"))
    print("   e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["streaming"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"], d["clocks"]["sm_mhz"])
except Exception as e: print("failed", e)
PY
