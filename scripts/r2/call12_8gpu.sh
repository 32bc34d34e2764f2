#!/bin/bash
# round 2, call 12 (8 GPUs): parity worker on all four 8-rank grids (2x2x2, 4x2x1, 1x2x4, 1x8x1); bench on 1x8x1 with e2e,
# 1x4x1, and north_star's 3D grid 2x2x2 through the general path
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 900 python -m pytest "tests/test_gpu_multi.py::test_process_grids[8]" -m gpu -q --timeout 800 -x ) > gpurun_out/r2c12_pytest_8gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c12_pytest_8gpu.log
grep -v "^  File\|site-packages" gpurun_out/r2c12_pytest_8gpu.log | tail -n 30
run() { # name nproc extra-args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 ${@:3} > gpurun_out/r2c12_bench_$1.json 2> gpurun_out/r2c12_bench_$1.err
  echo "bench $1 exit $?"
}
run 8gpu_1x8x1 8
run 4gpu_1x4x1 4 --no-e2e
run 8gpu_2x2x2 8 --grid 2x2x2 --no-e2e --steps 3 --warmup 1
python - <<'PY'
import json
for c in ("8gpu_1x8x1","4gpu_1x4x1","8gpu_2x2x2"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c12_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f numeric_share %.3f fp64_frac %s launches %d waits/step %s parity %s" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r.get("fp64_frac"), d["gpu_launches"], r.get("host_waits_per_step"), d["parity_checked"] and d["parity_checked"]["ok"]))
        print("   phases", r.get("step_ms_by_phase"), "peer", r.get("peer"))
        if d["e2e"]: print("   e2e solve", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["solve"], "streaming", d["e2e"]["streaming"]["ms_per_step"])
    except Exception as e: print(c, "failed", e)
PY
tail -n 4 gpurun_out/r2c12_bench_8gpu_2x2x2.err | cut -c1-300
