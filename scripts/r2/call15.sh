#!/bin/bash
# round 2, call 15 (1 GPU): byte-granular ring (fixed wrap) + interior epilogue + complex tile path: parity tests, A/B bench c4, c3, c5 probes
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 600 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_tile_space.py tests/test_gpu_solvers.py -m gpu -q --timeout 300 ) > gpurun_out/r2c15_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c15_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c15_pytest.log | grep -v "^$" | tail -n 25
NTB_RING=1 timeout 300 python bench.py --no-e2e --no-check --no-cpu-baseline > gpurun_out/r2c15_bench_ring1.json 2> gpurun_out/r2c15_bench_ring1.err; echo "ring1 exit $?"
NTB_RING=1 timeout 300 python bench.py --config c3 --steps 2 --warmup 1 --no-e2e --no-check --no-cpu-baseline > gpurun_out/r2c15_bench_c3.json 2> gpurun_out/r2c15_bench_c3.err; echo "c3 exit $?"
timeout 200 python bench.py --config c5 --n 32768 --c5-scale 0.005 --steps 1 --warmup 0 --no-e2e --no-check --no-cpu-baseline --no-peaks > gpurun_out/r2c15_bench_c5_s005.json 2> gpurun_out/r2c15_bench_c5_s005.err; echo "c5 s0.005 exit $?"
timeout 300 python bench.py --config c5 --n 32768 --c5-scale 0.0125 --steps 1 --warmup 0 --no-e2e --no-check --no-cpu-baseline --no-peaks > gpurun_out/r2c15_bench_c5_s0125.json 2> gpurun_out/r2c15_bench_c5_s0125.err; echo "c5 s0.0125 exit $?"
timeout 200 python bench.py --config c5 --n 4096 --c5-scale 0.5 --steps 1 --warmup 0 --no-e2e --no-check --no-cpu-baseline --no-peaks > gpurun_out/r2c15_bench_c5_dense4096.json 2> gpurun_out/r2c15_bench_c5_dense4096.err; echo "c5 dense 4096 exit $?"
python - <<'PY'
import json
for c in ("ring1","c3","c5_s005","c5_s0125","c5_dense4096"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c15_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f fp64_frac %s hbm_frac %.3f launches %s" % (d["ms_per_step"], d["value"], r.get("fp64_frac"), r["frac"], d.get("gpu_launches")))
        print("   phases", r.get("step_ms_by_phase"), {k:v for k,v in d["config"].items() if k not in ("workload","l2")})
    except Exception as e: print(c, "failed", e)
PY
for f in c5_s005 c5_s0125 c5_dense4096; do grep -v "NCCL\|^$" gpurun_out/r2c15_bench_$f.err | tail -n 3; done
