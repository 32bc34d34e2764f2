#!/bin/bash
# round 2, call 8 (1 GPU): direct parity tests of the tile-space helpers; numeric kernel time vs matrix size
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( timeout 300 python -m pytest tests/test_gpu_tile_space.py -m gpu -q --timeout 200 ) > gpurun_out/r2c8_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c8_pytest.log
grep -v "^  File\|site-packages" gpurun_out/r2c8_pytest.log | grep -v "^$" | tail -n 60
for n in 131072 65536; do
  timeout 200 python bench.py --n $n --no-e2e --no-cpu-baseline --no-check --no-peaks > gpurun_out/r2c8_bench_n$n.json 2> gpurun_out/r2c8_bench_n$n.err
  python - "$n" <<'PY'
import json,sys
d=[json.loads(l) for l in open(f"gpurun_out/r2c8_bench_n{sys.argv[1]}.json") if l.startswith('{')][0]; r=d["roofline"]
print("n", sys.argv[1], "ms/step %.3f" % d["ms_per_step"], r["step_ms_by_phase"])
PY
done
