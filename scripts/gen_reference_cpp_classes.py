"""Generates tests/golden/reference_cpp_classes.json: the public method names of NTPoly's C++ classes (the names its SWIG
Python module exposes), read from /root/reference/Source/CPlusPlus/*.h. Run in the build container; the fixture is committed."""
import glob, json, os, re, sys
src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/Source/CPlusPlus"
out = {}
for path in sorted(glob.glob(os.path.join(src, "*.h"))):
    text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    for m in re.finditer(r"\bclass\s+(\w+)\s*(?::[^{]*)?\{(.*?)\n\};", text, flags=re.S):
        name, body = m.group(1), m.group(2)
        methods = set()
        access = "private"
        for line in re.split(r"\n", body):
            a = re.match(r"\s*(public|private|protected)\s*:", line)
            if a:
                access = a.group(1)
                continue
            if access != "public":
                continue
            f = re.match(r"\s*(?:static\s+|virtual\s+)?[\w:<>,\s\*&]+?[\s\*&](\w+)\s*\(", line)
            if f and f.group(1) not in (name, "operator"):
                methods.add(f.group(1))
        if methods:
            out[name] = sorted(methods)
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_cpp_classes.json")
json.dump(out, open(dst, "w"), indent=1)
print(dst, {k: len(v) for k, v in out.items()})
