"""wall-clock per C-ABI call of one sign iteration (developer probe)"""
import sys, os, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ntpoly_b200.api as nt
from ntpoly_b200.workloads import banded_sign_input
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
thr = 1e-6
nt.ConstructGlobalProcessGrid(1, 1, 1)
M = nt.Matrix_ps(n); M.fill_from_scipy(banded_sign_input(n))
I = nt.Matrix_ps(n); I.FillIdentity()
emin, emax = nt.EigenBounds.GershgorinBounds(M)
X = nt.Matrix_ps(M); X.Scale(1.0 / abs(emax))
T1, T2, D = nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n)
def timed(name, fn):
    nt.synchronize(); t0 = time.perf_counter(); r = fn(); nt.synchronize()
    print(f"  {name:28s} {(time.perf_counter()-t0)*1e3:9.3f} ms", flush=True); return r
pool = None
for it in range(3):
    nt.synchronize(); t0 = time.perf_counter()
    nv = nt.sign_iteration(X, I, T1, T2, 1.2, thr)
    nt.synchronize()
    print(f"driver sign_iteration {it}: {(time.perf_counter()-t0)*1e3:.3f} ms norm={nv:.4e} nnz(X)={X.GetSize()} builds={nt.tile_builds()}", flush=True)
for it in range(3):
    print("iteration", it, "nnz(X)", X.GetSize(), flush=True)
    ak = 1.2
    timed("gemm X*X", lambda: T1.Gemm(X, X, None, alpha=-ak*ak, threshold=thr))
    timed("increment 3I", lambda: T1.Increment(I, 3.0))
    timed("gemm X*T1", lambda: T2.Gemm(X, T1, None, alpha=0.5*ak, threshold=thr))
    timed("copy", lambda: nt.lib().CopyMatrix_ps_wrp(X.ih, D.ih))
    timed("increment", lambda: D.Increment(T2, -1.0))
    timed("norm", lambda: D.Norm())
    if it < 2:
        nt.lib().CopyMatrix_ps_wrp(T2.ih, X.ih)
    print("   nnz T1", T1.GetSize(), "T2", T2.GetSize(), nt.tile_counters(), flush=True)
