#!/bin/bash
# round 2, call 31 (2 GPUs): fused difference norm on the peer path: parity worker on the 2-rank grids, bench 1x2x1
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 600 python -m pytest "tests/test_gpu_multi.py::test_process_grids" -m gpu -q --timeout 500 ) > gpurun_out/r2c31_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c31_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c31_pytest.log | grep -v "^$" | tail -n 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-peaks > gpurun_out/r2c31_bench_1x2x1.json 2> gpurun_out/r2c31_bench_1x2x1.err
echo "bench exit $?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2c31_bench_1x2x1.json") if l.startswith('{')][0]; r=d["roofline"]
    print("1x2x1 ms/step %.3f value %.0f parity %s" % (d["ms_per_step"], d["value"], d["parity_checked"]), d["clocks"]["sm_mhz"])
    print("   phases", r.get("step_ms_by_phase"), "e2e", d["e2e"]["ms_per_step"])
except Exception as e: print("failed", e)
PY
grep -v "NCCL\|^$\|Warning\|warn\|OMP_NUM\|\*\*\*" gpurun_out/r2c31_bench_1x2x1.err | tail -n 4 | cut -c1-300
timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-peaks --no-check > gpurun_out/r2c31_bench_1gpu.json 2> gpurun_out/r2c31_bench_1gpu.err; echo "1gpu exit $?"
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2c31_bench_1gpu.json") if l.startswith('{')][0]; r=d["roofline"]
    print("1 GPU ms/step %.3f value %.0f" % (d["ms_per_step"], d["value"]), r.get("step_ms_by_phase"))
except Exception as e: print("failed", e)
PY
