"""Solver drivers on the CUDA path vs the oracle and vs the reference's golden density:
identical iteration counts, energy within 1e-8 relative (north_star tolerances)."""
import json
import os

import numpy as np
import pytest
import scipy.io as sio
import scipy.sparse as sp

from util import banded, compare_sparse, random_sparse

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def to_gpu(nt, m):
    M = nt.Matrix_ps(m.shape[0], is_complex=np.iscomplexobj(m.data))
    M.fill_from_scipy(m)
    return M


def params(nt, conv, thr, monitor=True, maxit=None):
    sp_ = nt.SolverParameters()
    sp_.SetConvergeDiff(conv)
    sp_.SetThreshold(thr)
    sp_.SetMonitorConvergence(monitor)
    if maxit:
        sp_.SetMaxIterations(maxit)
    return sp_


def test_premade_matrix_density_golden(nt, oracle):
    """BASELINE config 2: Examples/PremadeMatrix — NS inverse square root then TRS2; the density must
    reproduce Density-Reference.mtx, Tr(KH) within 1e-8 relative of the oracle, iteration counts equal."""
    H = nt.Matrix_ps(os.path.join(GOLD, "premade_Hamiltonian.mtx"))
    S = nt.Matrix_ps(os.path.join(GOLD, "premade_Overlap.mtx"))
    D = sio.mmread(os.path.join(GOLD, "premade_Density-Reference.mtx")).toarray()
    trace = json.load(open(os.path.join(GOLD, "premade_oracle_trace.json")))
    n = H.GetActualDimension()
    ISQ, K = nt.Matrix_ps(n), nt.Matrix_ps(n)
    p = params(nt, 1e-3, 1e-6)
    nt.reset_counters()
    nt.SquareRootSolvers.InverseSquareRoot(S, ISQ, p)
    assert nt.last_solve()["loop_counter"] == trace["isq_loop_counter"]
    assert nt.counters()["dense_rule_blocks"] > 0            # 7x7 dense inputs take the dense-branch rule
    p.SetConvergeDiff(1e-5)
    for name, fn in (("trs2", nt.DensityMatrixSolvers.TRS2), ("trs4", nt.DensityMatrixSolvers.TRS4),
                     ("pm", nt.DensityMatrixSolvers.PM)):
        e, mu = fn(H, ISQ, 5.0, K, p)
        rec = nt.last_solve()
        assert rec["loop_counter"] == trace[name]["loop_counter"], name
        assert e == pytest.approx(trace[name]["energy"], rel=1e-8), name
        assert mu == pytest.approx(trace[name]["mu"], rel=1e-8), name
        assert np.linalg.norm(K.to_scipy().toarray() - D) <= 1e-4, name


def test_premade_scale_and_fold(nt, oracle):
    """ScaleAndFold_wrp (DensityMatrixSolversModule.F90:953-1117) fed with the HOMO/LUMO of the generalised
    eigenproblem like the reference's own test (UnitTests/test_chemistry.py:236-264): golden density, and the same
    iteration count and energy as the oracle's restatement"""
    import scipy.linalg as la
    O = oracle
    Hs, Ss = (sp.csc_matrix(sio.mmread(os.path.join(GOLD, f))) for f in ("premade_Hamiltonian.mtx", "premade_Overlap.mtx"))
    D = sio.mmread(os.path.join(GOLD, "premade_Density-Reference.mtx")).toarray()
    w = la.eigh(Hs.toarray(), Ss.toarray(), eigvals_only=True)
    H, S = to_gpu(nt, Hs), to_gpu(nt, Ss)
    ISQ, K = nt.Matrix_ps(7), nt.Matrix_ps(7)
    p = params(nt, 1e-3, 1e-6)
    nt.SquareRootSolvers.InverseSquareRoot(S, ISQ, p)
    p.SetConvergeDiff(1e-5)
    e = nt.DensityMatrixSolvers.ScaleAndFold(H, ISQ, 5.0, K, w[4], w[5], p)
    rec = nt.last_solve()
    po = O.SolverParameters(converge_diff=1e-3, threshold=1e-6)
    ISQo, _ = O.inverse_square_root(O.PSMatrix.from_scipy(Ss), po)
    po.converge_diff = 1e-5
    Ko, info = O.scale_and_fold(O.PSMatrix.from_scipy(Hs), ISQo, 5.0, w[4], w[5], po)
    assert rec["loop_counter"] == info.iterations
    assert e == pytest.approx(info.energy, rel=1e-8)
    assert np.linalg.norm(K.to_scipy().toarray() - D) <= 1e-4
    assert np.linalg.norm(K.to_scipy().toarray() - Ko.todense()) <= 1e-8


def test_pseudo_inverse_is_the_hotelling_iteration(nt, oracle):
    """PseudoInverse (InverseSolversModule.F90:187-298) runs the same iteration as Invert"""
    n = 200
    m = spd_banded(n, seed=4)
    M, P, Inv = to_gpu(nt, m), nt.Matrix_ps(n), nt.Matrix_ps(n)
    p = params(nt, 1e-8, 1e-10)
    nt.InverseSolvers.PseudoInverse(M, P, p)
    it = nt.last_solve()["loop_counter"]
    nt.InverseSolvers.Invert(M, Inv, p)
    assert it == nt.last_solve()["loop_counter"]
    assert abs(P.to_scipy() - Inv.to_scipy()).sum() == 0.0
    assert abs(P.to_scipy() @ m - sp.identity(n)).max() < 1e-6


def test_premade_with_load_balancing_permutation(nt):
    """every reference example runs with a random permutation: results are permutation invariant up to
    threshold effects"""
    H = nt.Matrix_ps(os.path.join(GOLD, "premade_Hamiltonian.mtx"))
    S = nt.Matrix_ps(os.path.join(GOLD, "premade_Overlap.mtx"))
    D = sio.mmread(os.path.join(GOLD, "premade_Density-Reference.mtx")).toarray()
    n = H.GetActualDimension()
    perm = nt.Permutation(H.GetLogicalDimension())
    perm.SetRandomPermutation(seed=3)
    p = params(nt, 1e-3, 1e-6)
    p.SetLoadBalance(perm)
    ISQ, K = nt.Matrix_ps(n), nt.Matrix_ps(n)
    nt.SquareRootSolvers.InverseSquareRoot(S, ISQ, p)
    p.SetConvergeDiff(1e-5)
    e, mu = nt.DensityMatrixSolvers.TRS2(H, ISQ, 5.0, K, p)
    assert np.linalg.norm(K.to_scipy().toarray() - D) <= 1e-4
    assert e == pytest.approx(-22.97196, abs=1e-4)


def spd_banded(n, w=6, seed=0):
    rng = np.random.default_rng(seed)
    a = sp.diags([rng.uniform(0.05, 0.2, n - d) for d in range(1, w + 1)], list(range(1, w + 1)), shape=(n, n))
    m = a + a.T + sp.identity(n) * 2.0
    return sp.csc_matrix(m)


@pytest.mark.parametrize("thr", [0.0, 1e-7])
def test_sign_function_iterations_and_result(nt, oracle, thr):
    n = 600
    m = sp.csc_matrix(spd_banded(n) - sp.identity(n) * 2.0 + sp.diags(np.where(np.arange(n) % 2 == 0, 1.5, -1.5)))
    M, Out = to_gpu(nt, m), nt.Matrix_ps(n)
    p = params(nt, 1e-8, thr)
    nt.SignSolvers.ComputeSign(M, Out, p)
    rec = nt.last_solve()
    ref, info = oracle.sign_function(oracle.PSMatrix.from_scipy(m), oracle.SolverParameters(converge_diff=1e-8, threshold=thr))
    assert rec["loop_counter"] == info.iterations
    compare_sparse(Out.to_scipy(), ref.to_scipy(), thr, tol=1e-9)
    s2 = Out.to_scipy() @ Out.to_scipy()
    assert abs(s2 - sp.identity(n)).max() < 1e-5               # sign(M)^2 = I


@pytest.mark.parametrize("order", [2, 3, 5])
def test_inverse_square_root_orders(nt, oracle, order):
    n = 400
    m = spd_banded(n, seed=1)
    M, Out = to_gpu(nt, m), nt.Matrix_ps(n)
    thr = 1e-9
    p = params(nt, 1e-7, thr)
    nt.SquareRootSolvers.InverseSquareRoot(M, Out, p, order=order)
    rec = nt.last_solve()
    ref, info = oracle.inverse_square_root(oracle.PSMatrix.from_scipy(m), oracle.SolverParameters(converge_diff=1e-7, threshold=thr), order=order)
    assert rec["loop_counter"] == info.iterations
    compare_sparse(Out.to_scipy(), ref.to_scipy(), thr, tol=1e-8)
    z = Out.to_scipy()
    assert abs(z @ m @ z - sp.identity(n)).max() < 1e-5


def test_square_root_and_invert(nt, oracle):
    n = 300
    m = spd_banded(n, seed=2)
    M, R, Inv = to_gpu(nt, m), nt.Matrix_ps(n), nt.Matrix_ps(n)
    p = params(nt, 1e-8, 1e-10)
    nt.SquareRootSolvers.SquareRoot(M, R, p)
    r = R.to_scipy()
    assert abs(r @ r - m).max() < 1e-6
    nt.InverseSolvers.Invert(M, Inv, p)
    rec = nt.last_solve()
    ref, info = oracle.invert(oracle.PSMatrix.from_scipy(m), oracle.SolverParameters(converge_diff=1e-8, threshold=1e-10))
    assert rec["loop_counter"] == info.iterations
    compare_sparse(Inv.to_scipy(), ref.to_scipy(), 1e-10, tol=1e-8)
    assert abs(Inv.to_scipy() @ m - sp.identity(n)).max() < 1e-6


def test_complex_hermitian_invert_and_exponential(nt):
    """config 5 at the shipped size: Examples/ComplexMatrix/input.mtx (512x512 Hermitian, complex path)"""
    import scipy.linalg as la
    from ntpoly_b200.workloads import guo_transform
    g = guo_transform(sio.mmread(os.path.join(GOLD, "complex_input.mtx")))   # main.f90:109-160
    n = g.shape[0]
    assert abs(g - g.conj().T).max() < 1e-14
    shift = float(np.asarray(abs(g).sum(axis=0)).max()) + 1.0
    m = sp.csc_matrix(g + sp.identity(n) * shift)
    M, Inv, E = to_gpu(nt, m), nt.Matrix_ps(n), nt.Matrix_ps(n)
    p = params(nt, 1e-9, 1e-12)
    nt.InverseSolvers.Invert(M, Inv, p)
    assert Inv.IsComplex()
    assert abs(Inv.to_scipy() @ m - sp.identity(n)).max() < 1e-7
    G = to_gpu(nt, sp.csc_matrix(0.05 * g))
    pe = params(nt, 1e-9, 1e-10)
    nt.ExponentialSolvers.ComputeExponential(G, E, pe)
    expect = la.expm(0.05 * g.toarray())
    assert np.linalg.norm(E.to_scipy().toarray() - expect) / np.linalg.norm(expect) < 1e-6


def test_trs2_banded_medium(nt, oracle):
    """a sparse-branch purification (banded Hamiltonian, identity overlap): iteration count + energy vs oracle"""
    n = 1024
    h = banded(n, half_bandwidth=12, scale=0.2)
    thr = 1e-7
    H, ISQ, K = to_gpu(nt, h), nt.Matrix_ps(n), nt.Matrix_ps(n)
    ISQ.FillIdentity()
    p = params(nt, 1e-6, thr)
    e, mu = nt.DensityMatrixSolvers.TRS2(H, ISQ, n // 2, K, p)
    rec = nt.last_solve()
    OH = oracle.PSMatrix.from_scipy(h)
    Kref, info = oracle.trs2(OH, oracle.identity(OH), n // 2, oracle.SolverParameters(converge_diff=1e-6, threshold=thr))
    assert rec["loop_counter"] == info.iterations
    assert e == pytest.approx(info.energy, rel=1e-8)
    assert K.Trace() == pytest.approx(n // 2, abs=1e-3)
    compare_sparse(K.to_scipy(), Kref.to_scipy(), thr, tol=1e-7)


@pytest.mark.parametrize("solver", ["TRS4", "PM"])
def test_trs4_pm_banded_medium(nt, oracle, solver):
    """TRS4 (config c3's driver) and PM on a banded Hamiltonian: iteration count, energy, chemical potential, density"""
    n = 1024
    h = banded(n, half_bandwidth=12, scale=0.2)
    thr = 1e-7
    H, ISQ, K = to_gpu(nt, h), nt.Matrix_ps(n), nt.Matrix_ps(n)
    ISQ.FillIdentity()
    p = params(nt, 1e-6, thr)
    e, mu = getattr(nt.DensityMatrixSolvers, solver)(H, ISQ, n // 2, K, p)
    rec = nt.last_solve()
    OH = oracle.PSMatrix.from_scipy(h)
    fn = oracle.trs4 if solver == "TRS4" else oracle.pm
    Kref, info = fn(OH, oracle.identity(OH), n // 2, oracle.SolverParameters(converge_diff=1e-6, threshold=thr))
    assert rec["loop_counter"] == info.iterations
    assert e == pytest.approx(info.energy, rel=1e-8)
    assert K.Trace() == pytest.approx(n // 2, abs=1e-3)
    compare_sparse(K.to_scipy(), Kref.to_scipy(), thr, tol=1e-7)


def test_trs4_block_sparse_tile_path(nt, oracle):
    """config c3 in small: block-sparse Hamiltonian (32x32 blocks) through TRS4 on the tile path, with the tile
    forms of every iterate emitted by the products"""
    from ntpoly_b200.workloads import block_sparse
    n = 2048
    # gapped (alternating on-site energies) so that the density matrix stays block-sparse
    h = block_sparse(n, block=32, neighbours=4, band_blocks=3, seed=1234) * 0.3
    h = sp.csc_matrix(h + sp.diags(np.where(np.arange(n) % 2 == 0, 0.5, -0.5)))
    thr = 1e-6
    H, ISQ, K = to_gpu(nt, h), nt.Matrix_ps(n), nt.Matrix_ps(n)
    ISQ.FillIdentity()
    p = params(nt, 1e-6, thr)
    nt.reset_counters()
    e, mu = nt.DensityMatrixSolvers.TRS4(H, ISQ, n // 2, K, p)
    rec = nt.last_solve()
    assert nt.tile_counters()["tile_products"] > 0
    OH = oracle.PSMatrix.from_scipy(h)
    Kref, info = oracle.trs4(OH, oracle.identity(OH), n // 2, oracle.SolverParameters(converge_diff=1e-6, threshold=thr))
    assert rec["loop_counter"] == info.iterations
    assert e == pytest.approx(info.energy, rel=1e-8)
    compare_sparse(K.to_scipy(), Kref.to_scipy(), thr, tol=1e-7)


# ---- full-size style checks through size-independent properties (the oracle does not finish these in seconds) -------
def test_sign_function_c4_shape_properties(nt):
    """config-4 matrix at N=32768 (one eighth of the bench size; the whole iteration runs in tile space):
    sign(M)^2 = I, sign(M) symmetric, integer trace (= #positive - #negative eigenvalues), entries read back once"""
    from ntpoly_b200.workloads import banded_sign_input
    n, thr = 32768, 1e-6
    m = banded_sign_input(n)
    M, S, I, P = to_gpu(nt, m), nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n)
    I.FillIdentity()
    nt.reset_counters()
    nt.SignSolvers.ComputeSign(M, S, params(nt, 1e-5, thr))
    rec = nt.last_solve()
    assert 3 <= rec["loop_counter"] <= 40
    dc = nt.deferred_counters()
    assert dc["products"] >= 2 * rec["loop_counter"] - 1 and dc["materialized"] == 0, dc   # nothing left tile space
    assert nt.tile_counters()["tile_products"] == 2 * rec["loop_counter"]
    tr = S.Trace()                                                  # from the right tile form: still no CSC entries
    assert nt.deferred_counters()["materialized"] == 0
    assert S.GetSize() > n
    assert abs(tr - round(tr)) < 1e-2 and abs(tr) < n
    assert S.MeasureAsymmetry() < 5e-3                              # thresholded products are not symmetric to the bit
    P.Gemm(S, S, None, threshold=thr)
    P.Increment(I, -1.0)
    assert P.Norm() < 1e-3


def test_trs4_c3_shape_properties(nt):
    """config-3 matrix (32x32 dense blocks, block band) at N=8192: TRS4 density is idempotent, has the requested
    trace and commutes with H"""
    from ntpoly_b200.workloads import block_sparse
    n, thr, nel = 8192, 1e-6, 4096
    # gapped (alternating on-site energies), so that the density matrix is well defined and stays block-sparse
    h = block_sparse(n, neighbours=12, band_blocks=24) * 0.3
    h = sp.csc_matrix(h + sp.diags(np.where(np.arange(n) % 2 == 0, 0.5, -0.5)))
    H, ISQ, K, K2, C1, C2 = (to_gpu(nt, h), nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n),
                             nt.Matrix_ps(n))
    ISQ.FillIdentity()
    nt.reset_counters()
    e, mu = nt.DensityMatrixSolvers.TRS4(H, ISQ, nel, K, params(nt, 1e-6, thr))
    assert nt.tile_counters()["tile_products"] > 0
    assert K.Trace() == pytest.approx(nel, abs=1e-2)                # the requested trace
    K2.Gemm(K, K, None, threshold=thr)
    K2.Increment(K, -1.0)
    assert K2.Norm() < 1e-2                                         # idempotent
    C1.Gemm(K, H, None, threshold=thr)
    C2.Gemm(H, K, None, threshold=thr)
    C1.Increment(C2, -1.0)
    assert C1.Norm() < 1e-2                                         # [K, H] = 0
    assert e == pytest.approx(K.Dot(H), rel=1e-5)                   # E = Tr(K H) (identity overlap)


# ---- fused driver steps in tile space -------------------------------------------------------------
@pytest.mark.parametrize("solver", ["TRS4", "TRS2"])
def test_fused_steps_equal_the_reference_call_sequence(nt, oracle, solver):
    """TRS4 / TRS2 with their per-iteration helpers evaluated straight from the tile forms (Fx + sigma*Gx, Tr(X2 Fx),
    Tr(X2 Gx), 2X - X^2 with threshold, Tr(XH), Tr(X)) against the same driver issuing the reference's call sequence on
    CSC entries (NTB fused steps off) and against the oracle: identical iteration counts, energies to 1e-10 / 1e-8,
    densities within the parity bar."""
    from ntpoly_b200.workloads import block_sparse_hamiltonian
    n, thr = 2048, 1e-6
    h = block_sparse_hamiltonian(n)
    H, ISQ = to_gpu(nt, h), nt.Matrix_ps(n)
    ISQ.FillIdentity()
    fn = getattr(nt.DensityMatrixSolvers, solver)
    res = {}
    for fused in (True, False):
        nt.set_fused_steps(fused)
        nt.reset_counters()
        K = nt.Matrix_ps(n)
        e, mu = fn(H, ISQ, n // 2, K, params(nt, 1e-5, thr))
        res[fused] = (e, mu, nt.last_solve()["loop_counter"], K.to_scipy(), nt.tile_combines(), K.Trace())
    nt.set_fused_steps(True)
    assert res[True][4] > 0 and res[False][4] == 0          # the fused run really combined in tile space
    assert res[True][2] == res[False][2]
    assert res[True][0] == pytest.approx(res[False][0], rel=1e-10)
    assert res[True][1] == pytest.approx(res[False][1], rel=1e-8, abs=1e-10)
    compare_sparse(res[True][3], res[False][3], thr, tol=1e-7)
    OH = oracle.PSMatrix.from_scipy(h)
    ofn = oracle.trs4 if solver == "TRS4" else oracle.trs2
    Kref, info = ofn(OH, oracle.identity(OH), n // 2, oracle.SolverParameters(converge_diff=1e-5, threshold=thr))
    assert res[True][2] == info.iterations
    assert res[True][0] == pytest.approx(info.energy, rel=1e-8)
    assert res[True][5] == pytest.approx(n // 2, abs=1e-2)
    compare_sparse(res[True][3], Kref.to_scipy(), thr, tol=1e-6)


def test_chebyshev_recurrence_in_tile_space_equals_the_reference_call_sequence(nt, oracle):
    """ComputeExponential of a real banded matrix (ChebyshevSolversModule.F90:146-163): with the fused steps the
    recurrence T_k = 2*Bal*T_{k-1} - T_{k-2}, Res += c_k*T_k is combined straight from the tile forms (the two sparse
    adds per step never see CSC); against the same driver issuing the reference's calls, the oracle and scipy's expm"""
    import scipy.linalg as la
    n, thr = 1536, 1e-9
    m = banded(n, half_bandwidth=24, scale=0.2)
    m = sp.csc_matrix(m * 0.4)                      # spectral radius < 1: PowerBounds does not scale it away
    M = to_gpu(nt, m)
    res = {}
    for fused in (True, False):
        nt.set_fused_steps(fused)
        nt.reset_counters()
        E = nt.Matrix_ps(n)
        nt.ExponentialSolvers.ComputeExponential(M, E, params(nt, 1e-9, thr))
        res[fused] = (E.to_scipy(), nt.tile_combines(), nt.last_solve()["loop_counter"])
    nt.set_fused_steps(True)
    assert res[True][1] >= 20 and res[False][1] == 0            # 14 steps x 2 combines (the first ones may decline)
    assert res[True][2] == res[False][2]
    compare_sparse(res[True][0], res[False][0], thr, tol=1e-9)
    ref, info = oracle.compute_exponential(oracle.PSMatrix.from_scipy(m), oracle.SolverParameters(converge_diff=1e-9, threshold=thr))
    assert res[True][2] == info.iterations
    compare_sparse(res[True][0], ref.to_scipy(), thr, tol=1e-8)
    want = la.expm(m.toarray())
    assert np.linalg.norm(res[True][0].toarray() - want) / np.linalg.norm(want) < 1e-6
