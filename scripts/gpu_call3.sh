#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time NTB_NUMERIC_VER=9 timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/c3_pytest_v9.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c3_pytest_v9.log
for v in 8 9; do
  NTB_NUMERIC_VER=$v timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c3_bench_v${v}.json 2> gpurun_out/c3_bench_v${v}.err
done
NTB_NUMERIC_VER=9 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_numeric -s 14 -c 2 -f -o gpurun_out/c3_numeric_v9 \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c3_ncu_numeric.out 2>&1
tail -n 6 gpurun_out/c3_pytest_v9.log
for f in gpurun_out/c3_bench_v*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(" ms/step %.3f value %.0f numeric_share %.3f fp64_frac %.3f" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r["fp64_frac"]))
except Exception as e: print(" failed", e)
PY
done
