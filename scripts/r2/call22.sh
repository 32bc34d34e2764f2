#!/bin/bash
# round 2, call 22 (1 GPU): whole GPU suite on the current build; c5 launch list + ncu --set full of its top kernel (scalar path evidence); c4 quick bench
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c22_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c22_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c22_pytest.log | grep -v "^$" | tail -n 12
timeout 200 python bench.py --no-e2e --no-check --no-cpu-baseline --no-peaks > gpurun_out/r2c22_bench_c4.json 2> gpurun_out/r2c22_bench_c4.err; echo "c4 exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c22_launches_c5.csv python bench.py --config c5 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-check --no-peaks > gpurun_out/r2c22_ncu_c5_list.out 2>&1; echo "ncu list c5 exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_numeric_cta_atomic -s 40 -c 2 -f -o gpurun_out/r2c22_c5_atomic python bench.py --config c5 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-check --no-peaks > gpurun_out/r2c22_ncu_c5_full.out 2>&1; echo "ncu full c5 exit $?"
ncu -i gpurun_out/r2c22_c5_atomic.ncu-rep --page raw --csv > gpurun_out/r2c22_c5_atomic_raw.csv 2>/dev/null
python - <<'PY'
import json, csv, collections
try:
    d=[json.loads(l) for l in open("gpurun_out/r2c22_bench_c4.json") if l.startswith('{')][0]; r=d["roofline"]
    print("c4 ms/step %.3f value %.0f" % (d["ms_per_step"], d["value"]), r.get("step_ms_by_phase"))
except Exception as e: print("c4 failed", e)
try:
    rows=[r for r in csv.reader(open("gpurun_out/r2c22_launches_c5.csv")) if len(r)>5]
    hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); rows=rows[1:]
    agg=collections.OrderedDict(); tot=0
    for r in rows:
        k=r[ik].split("(")[0][-50:]; v=float(r[iv].replace(",","")); a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
    print("c5 kernels total %.1f ms in %d launches" % (tot/1e6, len(rows)))
    for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print("   %-52s n=%4d total %9.1f us share %4.1f %%" % (k,n,t/1e3,100*t/tot))
except Exception as e: print("c5 launch list failed", e)
PY
