#!/bin/bash
# round 2, call 21 (1 GPU): A/B of the copy-warp variants with per-kernel times (ncu launch list)
#   A = look-ups precomputed, digest in the copy warp (libntpoly_b200_planhd.so)   B = descriptors precomputed and TMA'd, with L2 prefetch
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for v in A B; do
  if [ $v = A ]; then export NTB_LIB=$PWD/ntpoly_b200/lib/libntpoly_b200_planhd.so; else unset NTB_LIB; fi
  timeout 200 python bench.py --no-e2e --no-check --no-cpu-baseline --no-peaks > gpurun_out/r2c21_bench_$v.json 2> gpurun_out/r2c21_bench_$v.err; echo "$v exit $?"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c21_launches_$v.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check --no-peaks > gpurun_out/r2c21_ncu_$v.out 2>&1; echo "ncu $v exit $?"
done
python - <<'PY'
import json, csv, collections
for c in ("A","B"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2c21_bench_{c}.json") if l.startswith('{')][0]; r=d["roofline"]
        print(c, "ms/step %.3f value %.0f" % (d["ms_per_step"], d["value"]), r.get("step_ms_by_phase"))
    except Exception as e: print(c, "failed", e)
    try:
        rows=[r for r in csv.reader(open(f"gpurun_out/r2c21_launches_{c}.csv")) if len(r)>5]
        hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
        rows=rows[1:]
        # last two steps = last N launches: find the last 4 numeric launches
        idx=[i for i,r in enumerate(rows) if "k_tile_numeric" in r[ik]]
        start=idx[-4]-8 if len(idx)>=4 else 0
        agg=collections.OrderedDict()
        for r in rows[max(start,0):]:
            k=r[ik].split("(")[0][-40:]; v=float(r[iv].replace(",",""))
            a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
        for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:12]: print("   %-42s n=%3d total %.1f us avg %.1f us" % (k,n,t/1e3,t/1e3/n))
    except Exception as e: print(c, "launch list failed", e)
PY
