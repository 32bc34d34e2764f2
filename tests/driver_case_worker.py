"""WORKER of tests/test_gpu_drivers_more.py (one case per process, so that an abort inside the library cannot take
the test session down). Cases: HPCP, PolarDecomposition, PowerBounds, McWeenyStep(S), EnergyDensityMatrix. The tests
follow the reference's own
(UnitTests/test_chemistry.py: test_hpcp; test_solvers.py: test_polarfunction :880, test_powermethod :826;
test_chemistry.py: test_mcweeny_step, test_energy_density) at its tolerance (helpers.py THRESHOLD = 1e-4).
"""
import os

import numpy as np
import scipy.io as sio
import scipy.linalg as la
import scipy.sparse as sp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
THRESHOLD = 1e-4


def to_gpu(nt, m):
    m = sp.csc_matrix(m)
    M = nt.Matrix_ps(m.shape[0], is_complex=np.iscomplexobj(m.data))
    M.fill_from_scipy(m)
    return M


def params(nt, conv=1e-6, thr=0.0, monitor=False):
    p = nt.SolverParameters()
    p.SetConvergeDiff(conv)
    p.SetThreshold(thr)
    p.SetMonitorConvergence(monitor)
    return p


def dense_symmetric(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.uniform(0.0, 1.0, (n, n))
    return a + a.T


def test_hpcp_premade_density(nt):
    H = nt.Matrix_ps(os.path.join(GOLD, "premade_Hamiltonian.mtx"))
    S = nt.Matrix_ps(os.path.join(GOLD, "premade_Overlap.mtx"))
    D = sio.mmread(os.path.join(GOLD, "premade_Density-Reference.mtx")).toarray()
    Hd = sio.mmread(os.path.join(GOLD, "premade_Hamiltonian.mtx")).toarray()
    Sd = sio.mmread(os.path.join(GOLD, "premade_Overlap.mtx")).toarray()
    ISQ, K = nt.Matrix_ps(7), nt.Matrix_ps(7)
    p = params(nt, 1e-3, 1e-6, monitor=True)
    nt.SquareRootSolvers.InverseSquareRoot(S, ISQ, p)
    p.SetConvergeDiff(1e-5)
    e, mu = nt.DensityMatrixSolvers.HPCP(H, ISQ, 5.0, K, p)
    k = K.to_scipy().toarray()
    assert np.linalg.norm(k - D) <= THRESHOLD
    assert abs(np.trace(k @ Sd) - 5.0) <= THRESHOLD
    assert abs(e - np.trace(k @ Hd)) <= 1e-3
    w = la.eigh(Hd, Sd, eigvals_only=True)
    assert w[4] < mu < w[5]
    # same iteration count and energy as the CPU restatement (oracle.hpcp, pinned on the golden density on the CPU side)
    from oracle import oracle as O
    rec = nt.last_solve()
    po = O.SolverParameters(converge_diff=1e-3, threshold=1e-6)
    Ho, So = O.PSMatrix.from_scipy(sp.csc_matrix(Hd)), O.PSMatrix.from_scipy(sp.csc_matrix(Sd))
    ISQo, _ = O.inverse_square_root(So, po)
    po.converge_diff = 1e-5
    Ko, info = O.hpcp(Ho, ISQo, 5.0, po)
    assert rec["loop_counter"] == info.iterations, (rec, info.iterations)
    assert abs(e - info.energy) <= 1e-8 * abs(info.energy)
    assert abs(mu - info.chemical_potential) <= 1e-6 * abs(info.chemical_potential)


def test_polar_decomposition(nt):
    n = 31
    a = dense_symmetric(n, 3) + np.random.default_rng(4).uniform(0.0, 1.0, (n, n))      # not symmetric
    u_ref, h_ref = la.polar(a)
    A, U, Hm = to_gpu(nt, a), nt.Matrix_ps(n), nt.Matrix_ps(n)
    nt.SignSolvers.ComputePolarDecomposition(A, U, Hm, params(nt))
    assert np.linalg.norm(U.to_scipy().toarray() - u_ref) <= THRESHOLD
    assert np.linalg.norm(Hm.to_scipy().toarray() - h_ref) <= THRESHOLD * np.linalg.norm(h_ref)


def test_power_bounds(nt):
    n = 31
    a = dense_symmetric(n, 5)
    got = nt.EigenBounds.PowerBounds(to_gpu(nt, a), params(nt))
    want = np.abs(la.eigvalsh(a)).max()
    assert abs(got - want) <= THRESHOLD


def test_mcweeny_step_and_energy_density(nt):
    n = 40
    rng = np.random.default_rng(6)
    d = sp.random(n, n, 0.3, random_state=rng, format="csc")
    d = sp.csc_matrix((d + d.T) * 0.1)
    s = sp.csc_matrix(sp.identity(n) + 0.05 * sp.csc_matrix(dense_symmetric(n, 7)) / n)
    h = sp.csc_matrix(dense_symmetric(n, 8))
    D, S, Hm, Out = to_gpu(nt, d), to_gpu(nt, s), to_gpu(nt, h), nt.Matrix_ps(n)
    dd, sd, hd = d.toarray(), s.toarray(), h.toarray()
    nt.DensityMatrixSolvers.McWeenyStep(D, Out)                         # 3 D^2 - 2 D^3
    assert np.linalg.norm(Out.to_scipy().toarray() - (3 * dd @ dd - 2 * dd @ dd @ dd)) <= 1e-12 * max(1.0, np.linalg.norm(dd))
    nt.DensityMatrixSolvers.McWeenyStep(D, Out, S)                      # 3 DSD - 2 DSDSD
    dsd = dd @ sd @ dd
    assert np.linalg.norm(Out.to_scipy().toarray() - (3 * dsd - 2 * dd @ sd @ dsd)) <= 1e-12 * max(1.0, np.linalg.norm(dd))
    nt.DensityMatrixSolvers.EnergyDensityMatrix(Hm, D, Out)             # D H D
    assert np.linalg.norm(Out.to_scipy().toarray() - dd @ hd @ dd) <= 1e-12 * np.linalg.norm(dd @ hd @ dd)


if __name__ == "__main__":
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import ntpoly_b200.api as api
    api.ConstructGlobalProcessGrid(1, 1, 1)
    globals()["test_" + sys.argv[1]](api)
    print("DRIVER_CASE_OK", sys.argv[1], flush=True)
