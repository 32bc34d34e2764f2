"""Replicate the fused TRS4 loop from Python (tile_scalars / tile_combine / Gemm) and compare every quantity with dense numpy."""
import os, sys
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ntpoly_b200.api as nt
from ntpoly_b200.workloads import block_sparse_hamiltonian
nt.ConstructGlobalProcessGrid(1, 1, 1)
n, thr = 2048, 1e-6
h = block_sparse_hamiltonian(n)
H = nt.Matrix_ps(n); H.fill_from_scipy(h)
e_min, e_max = nt.EigenBounds.GershgorinBounds(H)
x = (sp.identity(n) * e_max - h) / (e_max - e_min)
x = sp.csc_matrix(x)
X = nt.Matrix_ps(n); X.fill_from_scipy(x)
xd = x.toarray()
hd = h.toarray()
trace = n // 2
for it in range(1, 8):
    X2 = nt.Matrix_ps(n)
    X2.Gemm(X, X, None, threshold=thr)
    x2d = xd @ xd; x2d[abs(x2d) <= thr] = 0
    print(it, "X2 err", abs(X2.to_scipy().toarray() - x2d).max(), flush=True)
    x2d = X2.to_scipy().toarray()
    fx = 4 * xd - 3 * x2d; gx = np.eye(n) - 2 * xd + x2d
    sc = nt.tile_scalars(1, X2, X)
    tfx, tgx = (x2d * fx).sum(), (x2d * gx).sum()
    print("   scalars", sc, (tfx, tgx), flush=True)
    sigma = (trace - tfx) / tgx
    print("   sigma", sigma)
    if sigma > 6.0:
        T = nt.Matrix_ps(n); ok = nt.tile_combine(X2, X, T, mode=0, alpha=-1.0, beta=2.0, threshold=0.0)
        td = 2 * xd - x2d
    elif sigma < 0.0:
        T = nt.Matrix_ps(X2); td = x2d; ok = True
    else:
        FG = nt.Matrix_ps(n); ok = nt.tile_combine(X2, X, FG, mode=1, sigma=sigma)
        fgd = fx + sigma * gx
        print("   FG err", abs(FG.to_scipy().toarray() - fgd).max(), "ok", ok, flush=True)
        T = nt.Matrix_ps(n); T.Gemm(X2, FG, None, threshold=thr)
        td = x2d @ fgd; td[abs(td) <= thr] = 0
    print("   T err", abs(T.to_scipy().toarray() - td).max(), "energy", T.Dot(H), (T.to_scipy().toarray() * hd).sum(), "trace", T.Trace(), flush=True)
    X = T
    xd = T.to_scipy().toarray()
# now the drivers
for fused in (True, False):
    nt.set_fused_steps(fused)
    ISQ = nt.Matrix_ps(n); ISQ.FillIdentity(); K = nt.Matrix_ps(n)
    p = nt.SolverParameters(); p.SetThreshold(thr); p.SetConvergeDiff(1e-5); p.SetVerbosity(True)
    e, mu = nt.DensityMatrixSolvers.TRS4(H, ISQ, trace, K, p)
    print("driver fused", fused, "energy", e, "mu", mu, "its", nt.last_solve()["loop_counter"], "trace", K.Trace(), "nnz", K.GetSize(), flush=True)
