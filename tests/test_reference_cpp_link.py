"""CPU-only, needs the reference checkout (skipped on the GPU box): NTPoly's OWN C++ layer (Source/CPlusPlus/*.cc, the
classes its SWIG module wraps) is compiled from where it lies and linked against libntpoly_b200.so instead of
libNTPolyWrapper + libNTPoly. Every `*_wrp` symbol those classes call must be resolved by this library, except the
documented exclusions; and a small program written against the reference's TripletList class runs on top of it
(triplet lists are host objects: no GPU needed). Nothing from the reference is copied into the repository."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Source"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "CPlusPlus")), reason="reference checkout not present")

# the C++ translation units of the path, its containers and its drivers
UNITS = ["PSMatrix", "SMatrix", "TripletList", "MatrixMemoryPool", "PMatrixMemoryPool", "SolverParameters", "Permutation",
         "LoadBalancer", "DensityMatrixSolvers", "SignSolvers", "InverseSolvers", "SquareRootSolvers", "EigenBounds",
         "MatrixConversion", "ProcessGrid", "SolverBase", "ExponentialSolvers"]
EXPECTED_UNRESOLVED = {
    "ConstructMatrixFromBinary_ps_wrp", "ConstructMatrixFromBinaryPG_ps_wrp", "WriteMatrixToBinary_ps_wrp",   # MPI-IO
    "DenseDensity_wrp", "DenseSignFunction_wrp", "DenseInvert_wrp", "DenseSquareRoot_wrp", "DenseInverseSquareRoot_wrp",
    "ComputeDenseExponential_wrp", "ComputeDenseLogarithm_wrp",                                              # eigensolver based
    "ComputeExponentialPade_wrp", "ComputeLogarithm_wrp",                                                    # other solver modules
}
MPI_STUB = """#pragma once
typedef int MPI_Comm;
typedef int MPI_Fint;
#define MPI_COMM_WORLD 0
static inline MPI_Fint MPI_Comm_c2f(MPI_Comm c) { return c; }
"""
PROGRAM = r"""
#include "TripletList.h"
#include "Triplet.h"
#include <cstdio>
using namespace NTPoly;
int main() {
  TripletList_r list, sorted;
  const int cols[4] = {3, 1, 2, 1}, rows[4] = {1, 2, 2, 1};
  for (int i = 0; i < 4; ++i) { Triplet_r t; t.index_column = cols[i]; t.index_row = rows[i]; t.point_value = 10.0 * i; list.Append(t); }
  TripletList_r::SortTripletList(list, 3, sorted);
  if (sorted.GetSize() != 4) return 1;
  const int want_c[4] = {1, 1, 2, 3}, want_r[4] = {1, 2, 2, 1};
  const double want_v[4] = {30.0, 10.0, 20.0, 0.0};
  for (int i = 0; i < 4; ++i) {
    Triplet_r t = sorted.GetTripletAt(i);
    if (t.index_column != want_c[i] || t.index_row != want_r[i] || t.point_value != want_v[i]) return 2 + i;
  }
  std::printf("REFERENCE_CPP_OK\n");
  return 0;
}
"""


def test_reference_cpp_layer_links_and_runs(tmp_path):
    from ntpoly_b200 import build
    lib = build.build()
    inc = tmp_path / "inc"
    inc.mkdir()
    (inc / "mpi.h").write_text(MPI_STUB)
    flags = ["-std=c++11", "-fPIC", f"-I{REF}/CPlusPlus", f"-I{REF}/C", f"-I{inc}"]
    objs = []
    for u in UNITS:
        obj = str(tmp_path / f"{u}.o")
        subprocess.run(["g++", *flags, "-c", f"{REF}/CPlusPlus/{u}.cc", "-o", obj], check=True, capture_output=True)
        objs.append(obj)
    so = str(tmp_path / "libNTPolyCPP_on_b200.so")
    libdir = os.path.dirname(lib)
    subprocess.run(["g++", "-shared", "-o", so, *objs, f"-L{libdir}", "-lntpoly_b200"], check=True, capture_output=True)
    undefined = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True).stdout.splitlines()
                 if l.strip().endswith("_wrp")}
    ours = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout.splitlines() if l.strip()}
    assert undefined - ours == EXPECTED_UNRESOLVED
    assert len(undefined & ours) >= 100                      # the classes really call into this library
    # a program against the reference's own TripletList class, running on this library (host-only objects)
    src = tmp_path / "prog.cc"
    src.write_text(PROGRAM)
    exe = str(tmp_path / "prog")
    subprocess.run(["g++", *flags, str(src), str(tmp_path / "TripletList.o"), "-o", exe, f"-L{libdir}", "-lntpoly_b200",
                    f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "REFERENCE_CPP_OK" in r.stdout, (r.returncode, r.stdout, r.stderr)
