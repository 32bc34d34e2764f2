"""Device time of the numeric SpGEMM kernel alone and of the whole product (developer probe).
   python scripts/numeric_time.py [n] [reps]   (X*X on the banded matrix and on the sign iterate X_3)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ntpoly_b200.api as nt
from ntpoly_b200.workloads import banded
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nt.ConstructGlobalProcessGrid(1, 1, 1)
A = nt.Matrix_ps(n); A.fill_from_scipy(banded(n))
C = nt.Matrix_ps(n)
for thr in (1e-8,):
    for _ in range(3):
        C.Gemm(A, A, None, threshold=thr)
    nt.synchronize()
    nt.profile_enable(True); nt.profile_read()
    t0 = time.perf_counter()
    for _ in range(reps):
        C.Gemm(A, A, None, threshold=thr)
    nt.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    pr = nt.profile_read(); nt.profile_enable(False)
    print(f"pipe={os.environ.get('NTB_TILE_PIPE','default')} n={n} thr={thr}: numeric kernel {pr['numeric_ms']/max(pr['products'],1):.3f} ms, "
          f"product wall {wall:.3f} ms, nnzC={C.GetSize()} {nt.tile_counters()}", flush=True)
