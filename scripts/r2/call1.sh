#!/bin/bash
# round 2, call 1 (1 GPU): first hardware run of the pending driver tests with their output, the reference front-end
# examples, and an ncu --set full capture of the SHIPPED numeric kernel (dense-stage path)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( timeout 300 python -m pytest tests/test_gpu_drivers_pending.py tests/test_gpu_reference_frontend.py -m gpu -q --runxfail -rA --timeout 200 ) > gpurun_out/r2c1_pending.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c1_pending.log
grep -v "^  File\|site-packages" gpurun_out/r2c1_pending.log | tail -n 60
timeout 300 python bench.py --steps 20 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_numeric -s 14 -c 2 -f -o gpurun_out/r2c1_numeric \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c1_ncu_numeric.out 2>&1
echo "ncu exit $?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r2c1_bench.json") if l.startswith('{')][0]; r=d["roofline"]
print("ms/step %.3f value %.0f numeric_share %.3f fp64_frac %.3f hbm_frac %.3f launches %d" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r["fp64_frac"], r["frac"], d["gpu_launches"]))
print("e2e", d["e2e"]["ms_per_step"]); print("cpu", d["cpu_baseline"]["value"])
PY
