#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 200 python -m pytest tests -m gpu -q --timeout 150 --deselect tests/test_gpu_container.py ) > gpurun_out/c16_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c16_pytest.log
grep -v "^  File\|site-packages" gpurun_out/c16_pytest.log | tail -n 8
( timeout 100 python -m pytest tests/test_gpu_container.py -m gpu -q --timeout 80 ) > gpurun_out/c16_pytest_container.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c16_pytest_container.log
grep -v "^  File\|site-packages" gpurun_out/c16_pytest_container.log | tail -n 30
timeout 200 python bench.py > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err
echo "bench exit $?"; tail -n 3 gpurun_out/c16_bench.err
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c16_bench_ref.json 2> gpurun_out/c16_bench_ref.err
echo "ref exit $?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c16_smoke.log 2>&1; tail -n 2 gpurun_out/c16_smoke.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c16_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c16_launches.out 2>&1
echo "ncu exit $?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/c16_bench.json") if l.startswith('{')][0]; r=d["roofline"]
print("ms/step %.3f value %.0f numeric_share %.3f fp64_frac %.3f hbm_frac %.3f launches %d" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r["fp64_frac"], r["frac"], d["gpu_launches"]))
print("e2e", d["e2e"]); print("cpu", d["cpu_baseline"])
PY
cut -c1-260 gpurun_out/c16_bench_ref.json
