#!/bin/bash
# round 2, call 35 (1 GPU): c5, how many accumulator slabs (CTAs) the heavy-column kernel should keep in flight (L2 budget)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for mb in 288 512; do
  NTB_SLAB_L2_MB=$mb timeout 100 python bench.py --config c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peaks --no-check > gpurun_out/r2c35_bench_c5_$mb.json 2> gpurun_out/r2c35_bench_c5_$mb.err; echo "$mb exit $?"
done
python - <<'PY'
import json
for mb in (288, 512):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c35_bench_c5_{mb}.json") if l.startswith('{')][0]
        print(mb, "MB: ms/step %.1f value %.1f" % (d["ms_per_step"], d["value"]))
    except Exception as e: print(mb, "failed", e)
PY
