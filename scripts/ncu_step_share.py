"""Per-kernel device time of the LAST `steps` bench steps in an ncu launch list
(ncu --metrics gpu__time_duration.sum --csv ... python bench.py ...). A step holds two k_tile_numeric launches.
   python scripts/ncu_step_share.py launches.csv [steps]"""
import csv, collections, sys
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
L = []
for r in rows:
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = r["Metric Unit"]
    L.append((r["Kernel Name"].split("(")[0], v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)))
idx = [i for i, x in enumerate(L) if "k_tile_numeric" in x[0]]
# a step starts with the launches that precede its first numeric kernel (useful-product count, bounds, scans, task table)
first = idx[-2 * steps]
start = first
while start > 0 and "k_forms_fill" not in L[start - 1][0] and "k_diff_col_abs" not in L[start - 1][0] and "k_reduce" not in L[start - 1][0]:
    start -= 1
agg = collections.OrderedDict()
for name, us in L[start:]:
    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"last {steps} steps: {len(L) - start} launches, {tot / steps:.1f} us of kernel time per step (cold-cache, serialised under ncu)")
print(f"{'us/step':>10} {'launches/step':>14} {'share':>7}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / steps:10.1f} {v[0] / steps:14.1f} {100 * v[1] / tot:6.1f}%  {k[:90]}")
