"""c5 probe: where does the time of the complex graph workload go? Every product / solve is timed and printed at once
(flush), so that a run that has to be cut off still tells how far it came."""
import sys, time, os
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ntpoly_b200.api as nt
from ntpoly_b200.workloads import complex_hermitian_graph

def T(label, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    print(f"{label}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True); return r

def mat(m):
    M = nt.Matrix_ps(m.shape[0], is_complex=True); M.fill_from_scipy(sp.csc_matrix(m)); return M

def products(n):
    g = complex_hermitian_graph(n)
    G = mat(g)
    X, Y, Z = nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n)
    for rep in range(2):
        nt.reset_counters()
        T(f"N={n} G*G (25x25/col)", lambda: X.Gemm(G, G, None, threshold=1e-6))
        print("   nnz/row", X.GetSize() / n, "hash cols", nt.hash_columns(), flush=True)
        T(f"N={n} X*G ({X.GetSize() / n:.0f}x25/col)", lambda: Y.Gemm(X, G, None, threshold=1e-6))
        print("   nnz/row", Y.GetSize() / n, flush=True)
        T(f"N={n} G*X (25x{X.GetSize() / n:.0f}/col)", lambda: Y.Gemm(G, X, None, threshold=1e-6))
        T(f"N={n} X*X ({X.GetSize() / n:.0f}^2/col)", lambda: Z.Gemm(X, X, None, threshold=1e-3))
        print("   nnz/row", Z.GetSize() / n, flush=True)

def solves(n, scale, what):
    g = complex_hermitian_graph(n)
    shift = 8.0 * float(np.asarray(abs(g).sum(axis=0)).max()) + 1.0
    p = nt.SolverParameters(); p.SetThreshold(1e-6); p.SetConvergeDiff(1e-5); p.SetVerbosity(True)
    if "inv" in what:
        A = mat(g + sp.identity(n) * shift); Ai = nt.Matrix_ps(n)
        for rep in range(2):
            nt.reset_counters()
            T(f"N={n} Invert", lambda: nt.InverseSolvers.Invert(A, Ai, p))
            c = nt.counters(); print("   iterations", nt.last_solve()["loop_counter"], "nnz/row", Ai.GetSize() / n, "multiplies", c["multiplies"], "complex tile products", nt.complex_tile_products(), "hash cols", nt.hash_columns(), flush=True)
    if "exp" in what:
        E = mat(g * scale + sp.identity(n) * 0.01); Ee = nt.Matrix_ps(n)
        for rep in range(2):
            nt.reset_counters()
            T(f"N={n} Exponential scale {scale}", lambda: nt.ExponentialSolvers.ComputeExponential(E, Ee, p))
            c = nt.counters(); print("   sigma_counter", nt.last_solve()["loop_counter"], "nnz/row", Ee.GetSize() / n, "multiplies", c["multiplies"], "complex tile products", nt.complex_tile_products(), "hash cols", nt.hash_columns(), flush=True)

if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "products": products(int(sys.argv[2]))
    else: solves(int(sys.argv[2]), float(sys.argv[3]), mode)
