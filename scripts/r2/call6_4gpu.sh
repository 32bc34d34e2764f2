#!/bin/bash
# round 2, call 6 (4 GPUs): parity worker on all four 4-rank grids; bench on 1x4x1 (peer path) and 2x2x1 (north_star's 2D grid, general path)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 700 python -m pytest "tests/test_gpu_multi.py::test_process_grids[4]" -m gpu -q --timeout 600 -x ) > gpurun_out/r2c6_pytest_4gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c6_pytest_4gpu.log
grep -v "^  File\|site-packages" gpurun_out/r2c6_pytest_4gpu.log | tail -n 30
run() { # name nproc extra-args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 ${@:3} > gpurun_out/r2c6_bench_$1.json 2> gpurun_out/r2c6_bench_$1.err
  echo "bench $1 exit $?"
}
run 4gpu_1x4x1 4
run 4gpu_2x2x1 4 --grid 2x2x1 --no-e2e --steps 5 --warmup 2
run 2gpu_1x2x1 2 --no-e2e
grep -i "nranks" gpurun_out/r2c6_bench_4gpu_1x4x1.err | head -3
python - <<'PY'
import json
for c in ("4gpu_1x4x1","4gpu_2x2x1","2gpu_1x2x1"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c6_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f numeric_share %.3f fp64_frac %s launches %d waits/step %s parity %s peer %s" % (d["ms_per_step"], d["value"], r["numeric_share_of_step"], r.get("fp64_frac"), d["gpu_launches"], r.get("host_waits_per_step"), d["parity_checked"], r.get("peer")))
        if d["e2e"]: print("   e2e solve", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["solve"], "streaming", d["e2e"]["streaming"]["ms_per_step"])
    except Exception as e: print(c, "failed", e)
PY
tail -n 4 gpurun_out/r2c6_bench_4gpu_2x2x1.err | cut -c1-300
