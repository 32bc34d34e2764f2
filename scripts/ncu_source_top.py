"""Summarise `ncu --page source --csv` output: stall-reason totals and the hottest SASS instructions of a kernel.
usage: python scripts/ncu_source_top.py source.csv [kernel_index] [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 60
secs, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        secs.append(cur)
    elif cur is not None and cur["hdr"] is None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r:
        cur["rows"].append(r)
s = secs[which]
h = s["hdr"]
ix = {n: i for i, n in enumerate(h)}
tot = sum(int(r[ix["# Samples"]]) for r in s["rows"])
print(f"kernel {which} of {len(secs)}: {s['name'][:80]}")
print("total samples", tot, "instructions", len(s["rows"]),
      "executed", sum(int(r[ix["Instructions Executed"]]) for r in s["rows"]))
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[ix[n]]) for r in s["rows"]) for n in stalls}
for n, v in sorted(agg.items(), key=lambda x: -x[1]):
    if v:
        print(f"  {n:26s} {v:8d} {100 * v / tot:5.1f}%")
base = int(s["rows"][0][0], 16)
top = sorted(s["rows"], key=lambda r: -int(r[ix["# Samples"]]))[:topn]
print()
for r in sorted(top, key=lambda r: int(r[0], 16)):
    off = int(r[0], 16) - base
    st = {n[6:]: int(r[ix[n]]) for n in stalls if int(r[ix[n]]) > 0}
    st = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{off:5x} {r[1].strip()[:58]:58s} {int(r[ix['# Samples']]):6d} {100 * int(r[ix['# Samples']]) / tot:4.1f}% "
          f"exec={r[ix['Instructions Executed']]:>9s} {st}")
