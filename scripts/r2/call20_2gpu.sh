#!/bin/bash
# round 2, call 20 (2 GPUs): parity at 2 ranks with the precomputed stage plan; bench 1x2x1 (peer path), 2x1x1 and 1x1x2 (general path with phases), c5 at 2 ranks
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 700 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 ) > gpurun_out/r2c20_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c20_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c20_pytest.log | grep -v "^$" | tail -n 12
run() { # name extra-args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 ${@:2} > gpurun_out/r2c20_bench_$1.json 2> gpurun_out/r2c20_bench_$1.err
  echo "bench $1 exit $?"
}
run 1x2x1 --no-cpu-baseline
run 2x1x1 --grid 2x1x1 --no-e2e --no-cpu-baseline --no-peaks --steps 5 --warmup 3
run 1x1x2 --grid 1x1x2 --no-e2e --no-cpu-baseline --no-peaks --steps 5 --warmup 3
run c5 --config c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peaks
python - <<'PY'
import json
for c in ("1x2x1","2x1x1","1x1x2","c5"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c20_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f fp64_frac %s launches %s waits/step %s parity %s" % (d["ms_per_step"], d["value"], r.get("fp64_frac"), d.get("gpu_launches"), r.get("host_waits_per_step"), d["parity_checked"] and d["parity_checked"]["ok"]))
        print("   phases", r.get("step_ms_by_phase"))
        if d.get("e2e"): print("   e2e", d["e2e"].get("value"), d["e2e"].get("ms_per_step"))
    except Exception as e: print(c, "failed", e)
PY
for f in 2x1x1 1x1x2 c5; do grep -v "NCCL\|^$\|Warning\|warn" gpurun_out/r2c20_bench_$f.err | tail -n 3 | cut -c1-300; done
