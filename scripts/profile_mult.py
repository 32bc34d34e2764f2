"""X*X on the banded N matrix a few times (target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ntpoly_b200.api as nt
from ntpoly_b200.workloads import banded
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nt.ConstructGlobalProcessGrid(1, 1, 1)
A = nt.Matrix_ps(n); A.fill_from_scipy(banded(n))
C = nt.Matrix_ps(n)
for _ in range(reps):
    C.Gemm(A, A, None, threshold=1e-8)
nt.synchronize()
print("done", C.GetSize(), nt.tile_counters())
