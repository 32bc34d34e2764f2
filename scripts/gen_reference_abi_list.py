"""Generates tests/golden/reference_c_abi.json: every `*_wrp` symbol NTPoly's own C headers declare, per header
(read from /root/reference/Source/C; run in the build container, the fixture is committed)."""
import glob, json, os, re, sys
src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/Source/C"
out = {}
for path in sorted(glob.glob(os.path.join(src, "*_c.h"))):
    text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    syms = sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*_wrp)\s*\(", text)))
    if syms:
        out[os.path.basename(path)] = syms
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_c_abi.json")
json.dump(out, open(dst, "w"), indent=1)
print(dst, sum(len(v) for v in out.values()), "symbols in", len(out), "headers")
