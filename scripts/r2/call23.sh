#!/bin/bash
# round 2, call 23 (1 GPU): atomic slab kernel with shared-k mode and 4-way loads: parity of the bins, c5 products probe, c5 bench line
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 300 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_smatrix.py -m gpu -q --timeout 200 ) > gpurun_out/r2c23_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c23_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c23_pytest.log | grep -v "^$" | tail -n 8
timeout 100 python scripts/r2/probe_c5.py products 32768 2>&1 | grep -v NCCL > gpurun_out/r2c23_probe_products.log; echo "products exit $?"; cat gpurun_out/r2c23_probe_products.log
timeout 400 python bench.py --config c5 --steps 2 --warmup 1 > gpurun_out/r2c23_bench_c5.json 2> gpurun_out/r2c23_bench_c5.err; echo "c5 exit $?"
python - <<'PY'
import json
for c in ("c5",):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c23_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f hbm_frac %.4f launches %s parity %s" % (d["ms_per_step"], d["value"], r["frac"], d.get("gpu_launches"), d.get("parity_checked")))
        print("   phases", r.get("step_ms_by_phase"), d.get("details"))
        print("   e2e", d["e2e"].get("value"), d["e2e"].get("ms_per_step"), "cpu", d.get("cpu_baseline"))
    except Exception as e: print(c, "failed", e)
PY
grep -v "NCCL\|^$" gpurun_out/r2c23_bench_c5.err | tail -n 5
