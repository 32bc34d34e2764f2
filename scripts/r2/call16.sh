#!/bin/bash
# round 2, call 16 (1 GPU): lane-parallel producer (ring 0/1 A/B), c5 probes with per-call timing
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 300 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_tile_space.py -m gpu -q --timeout 200 -x ) > gpurun_out/r2c16_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c16_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c16_pytest.log | grep -v "^$" | tail -n 8
NTB_RING=0 timeout 200 python bench.py --no-e2e --no-check --no-cpu-baseline --no-peaks > gpurun_out/r2c16_bench_ring0.json 2> gpurun_out/r2c16_bench_ring0.err; echo "ring0 exit $?"
NTB_RING=1 timeout 200 python bench.py --no-e2e --no-check --no-cpu-baseline --no-peaks > gpurun_out/r2c16_bench_ring1.json 2> gpurun_out/r2c16_bench_ring1.err; echo "ring1 exit $?"
python - <<'PY'
import json
for c in ("ring0","ring1"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c16_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f fp64_frac %s hbm_frac %.3f" % (d["ms_per_step"], d["value"], r.get("fp64_frac"), r["frac"]), r.get("step_ms_by_phase"))
    except Exception as e: print(c, "failed", e)
PY
timeout 100 python scripts/r2/probe_c5.py products 32768 2>&1 | grep -v NCCL > gpurun_out/r2c16_probe_products.log; echo "products exit $?"; cat gpurun_out/r2c16_probe_products.log
timeout 60 python scripts/r2/probe_c5.py inv 4096 0.5 2>&1 | grep -v NCCL > gpurun_out/r2c16_probe_inv4096.log; echo "inv4096 exit $?"; tail -n 30 gpurun_out/r2c16_probe_inv4096.log
timeout 60 python scripts/r2/probe_c5.py exp 4096 0.5 2>&1 | grep -v NCCL > gpurun_out/r2c16_probe_exp4096.log; echo "exp4096 exit $?"; tail -n 30 gpurun_out/r2c16_probe_exp4096.log
timeout 120 python scripts/r2/probe_c5.py inv 32768 0.0125 2>&1 | grep -v NCCL > gpurun_out/r2c16_probe_inv32768.log; echo "inv32768 exit $?"; tail -n 30 gpurun_out/r2c16_probe_inv32768.log
