#!/bin/bash
# round 2, call 25 (8 GPUs): parity worker on 2x2x2 and 1x8x1, bench c4 on 1x8x1 (peer path, with e2e), c5 on 8 ranks
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time NTB_WORKER_GRIDS=2x2x2,1x8x1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29628 tests/mp_gpu_worker.py ) > gpurun_out/r2c25_worker_8gpu.log 2>&1
echo "worker exit: $?" >> gpurun_out/r2c25_worker_8gpu.log
grep -v "NCCL\|^$\|OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r2c25_worker_8gpu.log | tail -n 8
run() { # name extra-args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 ${@:2} > gpurun_out/r2c25_bench_$1.json 2> gpurun_out/r2c25_bench_$1.err
  echo "bench $1 exit $?"
}
run 1x8x1 --no-cpu-baseline
run c5 --config c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peaks
python - <<'PY'
import json
for c in ("1x8x1","c5"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c25_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f fp64_frac %s launches %s waits/step %s parity %s" % (d["ms_per_step"], d["value"], r.get("fp64_frac"), d.get("gpu_launches"), r.get("host_waits_per_step"), d["parity_checked"] and d["parity_checked"]["ok"]))
        print("   phases", r.get("step_ms_by_phase"))
        if d.get("e2e"): print("   e2e", d["e2e"].get("value"), d["e2e"].get("ms_per_step"), d["e2e"].get("streaming",{}).get("ms_per_step"))
    except Exception as e: print(c, "failed", e)
PY
grep -i "nranks" gpurun_out/r2c25_bench_1x8x1.err | head -2 | cut -c1-200
for f in 1x8x1 c5; do grep -v "NCCL\|^$\|Warning\|warn\|OMP_NUM\|\*\*\*" gpurun_out/r2c25_bench_$f.err | tail -n 3 | cut -c1-300; done
