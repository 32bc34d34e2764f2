#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/c8_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c8_pytest.log
timeout 600 python bench.py > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err

timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c8_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c8_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_numeric -s 14 -c 2 -f -o gpurun_out/c8_numeric \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c8_ncu_numeric.out 2>&1
grep -v "^  File\|site-packages" gpurun_out/c8_pytest.log | tail -n 12
for f in gpurun_out/c8_bench.json; do python - "$f" <<'PY'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][0]; r=d["roofline"]
    print(sys.argv[1], " ms/step %.3f value %.0f numeric_share %.3f fp64_frac %.3f hbm_frac %.3f e2e %s" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r["fp64_frac"], r["frac"], d.get("e2e")))
except Exception as e: print(" failed", e)
PY
done
