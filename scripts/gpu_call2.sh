#!/bin/bash
# GPU call 2: parity suite under both numeric-kernel versions, bench for versions x pipeline shapes, ncu of v9
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/c2_pytest_v8.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c2_pytest_v8.log
( time NTB_NUMERIC_VER=9 timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/c2_pytest_v9.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c2_pytest_v9.log
( time NTB_NUMERIC_VER=9 NTB_NUMERIC_SHAPE=23 timeout 900 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_solvers.py -m gpu -q --timeout 600 ) > gpurun_out/c2_pytest_v9s23.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c2_pytest_v9s23.log
for v in 8 9; do for s in 32 23; do
  NTB_NUMERIC_VER=$v NTB_NUMERIC_SHAPE=$s timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c2_bench_v${v}_s${s}.json 2> gpurun_out/c2_bench_v${v}_s${s}.err
done; done
NTB_NUMERIC_VER=9 NTB_TILE_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c2_phase.json 2> gpurun_out/c2_phase.err
NTB_NUMERIC_VER=9 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c2_launches_v9.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c2_launches.out 2>&1
NTB_NUMERIC_VER=9 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_numeric -s 14 -c 2 -f -o gpurun_out/c2_numeric_v9 \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c2_ncu_numeric.out 2>&1
tail -4 gpurun_out/c2_pytest_v8.log gpurun_out/c2_pytest_v9.log gpurun_out/c2_pytest_v9s23.log
for f in gpurun_out/c2_bench_v*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(" ms/step %.3f value %.0f numeric_share %.3f fp64_frac %.3f deferred %s/%s" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r["fp64_frac"], r.get("deferred_csc_products_in_timed_region"), r.get("deferred_csc_materialized_in_timed_region")))
except Exception as e: print(" failed", e)
PY
done
