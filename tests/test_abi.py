"""CPU-only: the C-ABI library loads and exports every symbol include/ntpoly_b200.h declares
(no compute calls: there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ntpoly_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)) - {"defined"})


def test_library_builds_and_exports_header():
    from ntpoly_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) > 120
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_python_mirror_lists_every_header_symbol():
    import ntpoly_b200.api as api
    assert sorted(api.EXPORTED_SYMBOLS) == header_symbols()


def test_reference_hot_path_symbols_present():
    """the `_wrp` names NTPoly's C++ layer binds for this path (reference Source/C/PSMatrix_c.h:51-68,
    PMatrixMemoryPool_c.h:4-5) must be exported with C linkage"""
    from ntpoly_b200 import build
    out = subprocess.run(["nm", "-D", "--defined-only", build.build()], capture_output=True, text=True).stdout
    for s in ["MatrixMultiply_ps_wrp", "IncrementMatrix_ps_wrp", "ScaleMatrix_ps_wrp", "MatrixTrace_ps_wrp",
              "MatrixNorm_ps_wrp", "DotMatrix_psr_wrp", "DotMatrix_psc_wrp", "ConstructMatrixMemoryPool_p_wrp",
              "DestructMatrixMemoryPool_p_wrp", "TRS2_wrp", "TRS4_wrp", "PM_wrp", "SignFunction_wrp", "Invert_wrp",
              "InverseSquareRoot_wrp", "ComputeExponential_wrp"]:
        assert re.search(rf"\bT {s}\b", out), s


def test_no_oracle_in_product():
    """the shipped package must not import, link or call anything under oracle/"""
    pkg = os.path.join(ROOT, "ntpoly_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in text.lower(), os.path.join(dp, f)


def test_kernels_are_sm100a():
    from ntpoly_b200 import build
    out = subprocess.run(["cuobjdump", "--list-elf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
