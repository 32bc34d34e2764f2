"""Bisect the ComplexMatrix example failure: front-end construction vs the exponential driver."""
import os, sys
import numpy as np, scipy.io as sio, scipy.linalg as la, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ntpoly_b200.api as nt
from ntpoly_b200.workloads import guo_transform
nt.ConstructGlobalProcessGrid(1, 1, 1)
a = sio.mmread(os.path.join(ROOT, "tests/golden/complex_input.mtx"))
g = guo_transform(a)
n = g.shape[0]
def togpu(m):
    m = sp.coo_matrix(m)
    M = nt.Matrix_ps(n, is_complex=np.iscomplexobj(m.data))
    M.fill_from_arrays(m.row + 1, m.col + 1, m.data)
    return M
for f in (0.05, 0.2, 0.5):
    for thr in (1e-9, 1e-6):
        G = togpu(sp.csc_matrix(f * g))
        p = nt.SolverParameters(); p.SetThreshold(thr)
        pb = nt.EigenBounds.PowerBounds(G, p)
        E = nt.Matrix_ps(n)
        nt.ExponentialSolvers.ComputeExponential(G, E, p)
        want = la.expm(f * g.toarray())
        got = E.to_scipy().toarray()
        print(f"scale {f} thr {thr}: power bound {pb:.6f} (true {np.linalg.eigvalsh(f*g.toarray()).max():.6f}) "
              f"sigma_counter {nt.last_solve()['loop_counter']} rel err {np.linalg.norm(got-want)/np.linalg.norm(want):.3e}", flush=True)
# manual Chebyshev + squaring with plain products to localise
f, thr = 0.5, 1e-6
G = togpu(sp.csc_matrix(f * g / 32.0))
X = G.to_scipy().toarray()
T = nt.Matrix_ps(n)
T.Gemm(G, G, None, alpha=2.0, threshold=thr / 32)
print("G*G err", abs(T.to_scipy().toarray() - 2 * X @ X).max(), flush=True)
R = togpu(sp.csc_matrix(la.expm(X)))
Rn = la.expm(X)
for k in range(5):
    T = nt.Matrix_ps(n)
    T.Gemm(R, R, None, threshold=thr)
    Rn = Rn @ Rn
    print("squaring", k, "rel err", np.linalg.norm(T.to_scipy().toarray() - Rn) / np.linalg.norm(Rn), "max", abs(Rn).max(), flush=True)
    R = T
