"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0]
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'total us':>12} {'launches':>8} {'share':>6}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{v[1]:12.1f} {v[0]:8d} {100 * v[1] / tot:5.1f}%  {k[:100]}")
print(f"{tot:12.1f} total")
