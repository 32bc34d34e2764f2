#!/bin/bash
# round 2, call 28 (1 GPU): FINAL state - smoke, whole GPU suite, all four bench lines + reference arm, ncu of the numeric kernel + launch list
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c28_smoke.log 2>&1; echo "smoke exit $?"; grep -v NCCL gpurun_out/r2c28_smoke.log | tail -n 2
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c28_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c28_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c28_pytest.log | grep -v "^$" | tail -n 8
timeout 400 python bench.py > gpurun_out/r2c28_bench_c4.json 2> gpurun_out/r2c28_bench_c4.err; echo "c4 exit $?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c28_bench_reference_c4.json 2> gpurun_out/r2c28_bench_reference_c4.err; echo "ref exit $?"
timeout 200 python bench.py --config c1 > gpurun_out/r2c28_bench_c1.json 2> gpurun_out/r2c28_bench_c1.err; echo "c1 exit $?"
timeout 400 python bench.py --config c3 --steps 3 --warmup 1 > gpurun_out/r2c28_bench_c3.json 2> gpurun_out/r2c28_bench_c3.err; echo "c3 exit $?"
timeout 400 python bench.py --config c5 --steps 3 --warmup 1 > gpurun_out/r2c28_bench_c5.json 2> gpurun_out/r2c28_bench_c5.err; echo "c5 exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile_numeric -s 12 -c 4 -f -o gpurun_out/r2c28_numeric \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check --no-peaks > gpurun_out/r2c28_ncu_numeric.out 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/r2c28_numeric.ncu-rep --page source --csv > gpurun_out/r2c28_numeric_source.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c28_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check --no-peaks > gpurun_out/r2c28_ncu_list.out 2>&1; echo "ncu list exit $?"
python - <<'PY'
import json
for c in ("c4","reference_c4","c1","c3","c5"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c28_bench_{c}.json") if l.startswith('{')][0]; r=d.get("roofline") or {}
        print(c, "ms/step %.3f value %.1f fp64_frac %s hbm_frac %s launches %s parity %s" % (d["ms_per_step"], d["value"], r.get("fp64_frac"), r.get("frac"), d.get("gpu_launches"), (d.get("parity_checked") or {}).get("ok")))
        if d.get("e2e"): print("   e2e", d["e2e"].get("value"), d["e2e"].get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d.get("clocks"))
    except Exception as e: print(c, "failed", e)
PY
for f in c4 reference_c4 c1 c3 c5; do grep -v "NCCL\|^$" gpurun_out/r2c28_bench_$f.err | tail -n 2 | cut -c1-300; done
