"""CPU-only: pins the oracle (a) to the reference's shipped golden vector, (b) to SciPy the way the
reference's own unit tests are pinned (UnitTests/test_psmatrixalgebra.py, THRESHOLD=1e-4 in helpers.py:13),
and (c) checks the restated threshold rules on hand-made cases."""
import json
import os
import warnings

import numpy as np
import pytest
import scipy.io as sio
import scipy.linalg as la
import scipy.sparse as sp

warnings.filterwarnings("ignore", category=DeprecationWarning)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def mm(name):
    return sp.csc_matrix(sio.mmread(os.path.join(GOLD, name)))


def test_premade_density_matches_reference_golden(oracle):
    """Examples/PremadeMatrix: NS inverse square root (order 5) then TRS2, threshold 1e-6,
    converge 1e-3 / 1e-5, trace 5 -> Density-Reference.mtx."""
    O = oracle
    H, S = O.PSMatrix.from_scipy(mm("premade_Hamiltonian.mtx")), O.PSMatrix.from_scipy(mm("premade_Overlap.mtx"))
    D = mm("premade_Density-Reference.mtx").toarray()
    p = O.SolverParameters(converge_diff=1e-3, threshold=1e-6)
    ISQ, i1 = O.inverse_square_root(S, p)
    p.converge_diff = 1e-5
    K, i2 = O.trs2(H, ISQ, 5.0, p)
    assert np.linalg.norm(K.todense() - D) <= 1e-4            # reference tolerance (helpers.py:13)
    trace = json.load(open(os.path.join(GOLD, "premade_oracle_trace.json")))
    assert i1.iterations == trace["isq_loop_counter"] == 4
    assert i2.iterations == trace["trs2"]["loop_counter"] == 19
    assert i2.energy == pytest.approx(trace["trs2"]["energy"], rel=1e-12)
    # known answers established with SciPy in the survey: Tr(DS)=5, Tr(DH)=-22.971971
    Sd, Hd = S.todense(), H.todense()
    assert np.trace(K.todense() @ Sd) == pytest.approx(5.0, abs=1e-4)
    assert np.trace(K.todense() @ Hd) == pytest.approx(-22.971971, abs=1e-4)
    # the chemical potential must sit in the HOMO-LUMO gap (test_chemistry.py:184-191)
    w = la.eigh(Hd, Sd, eigvals_only=True)
    assert w[4] < i2.chemical_potential < w[5]


@pytest.mark.parametrize("solver", ["trs4", "pm"])
def test_premade_other_density_solvers(oracle, solver):
    O = oracle
    H, S = O.PSMatrix.from_scipy(mm("premade_Hamiltonian.mtx")), O.PSMatrix.from_scipy(mm("premade_Overlap.mtx"))
    D = mm("premade_Density-Reference.mtx").toarray()
    p = O.SolverParameters(converge_diff=1e-3, threshold=1e-6)
    ISQ, _ = O.inverse_square_root(S, p)
    p.converge_diff = 1e-5
    K, info = getattr(O, solver)(H, ISQ, 5.0, p)
    assert np.linalg.norm(K.todense() - D) <= 1e-4


def test_premade_scale_and_fold(oracle):
    """Scale and Fold with the HOMO/LUMO of the generalised eigenproblem, as the reference's own test feeds it
    (UnitTests/test_chemistry.py:236-264), must land on the shipped Density-Reference.mtx too."""
    O = oracle
    H, S = O.PSMatrix.from_scipy(mm("premade_Hamiltonian.mtx")), O.PSMatrix.from_scipy(mm("premade_Overlap.mtx"))
    D = mm("premade_Density-Reference.mtx").toarray()
    w = la.eigh(H.todense(), S.todense(), eigvals_only=True)
    p = O.SolverParameters(converge_diff=1e-3, threshold=1e-6)
    ISQ, _ = O.inverse_square_root(S, p)
    p.converge_diff = 1e-5
    K, info = O.scale_and_fold(H, ISQ, 5.0, w[4], w[5], p)
    assert np.linalg.norm(K.todense() - D) <= 1e-4
    assert np.trace(K.todense() @ S.todense()) == pytest.approx(5.0, abs=1e-4)
    assert info.iterations < 19                                  # the acceleration must beat plain TRS2 (19)


def test_premade_hpcp(oracle):
    """HPCP (DensityMatrixSolversModule.F90:720-950) on the shipped example: golden density, chemical potential in the
    HOMO-LUMO gap (UnitTests/test_chemistry.py basic_solver / check_cp)"""
    O = oracle
    H, S = O.PSMatrix.from_scipy(mm("premade_Hamiltonian.mtx")), O.PSMatrix.from_scipy(mm("premade_Overlap.mtx"))
    D = mm("premade_Density-Reference.mtx").toarray()
    p = O.SolverParameters(converge_diff=1e-3, threshold=1e-6)
    ISQ, _ = O.inverse_square_root(S, p)
    p.converge_diff = 1e-5
    K, info = O.hpcp(H, ISQ, 5.0, p)
    assert np.linalg.norm(K.todense() - D) <= 1e-4
    assert info.energy == pytest.approx(np.trace(K.todense() @ H.todense()), abs=1e-3)
    w = la.eigh(H.todense(), S.todense(), eigvals_only=True)
    assert w[4] < info.chemical_potential < w[5]


def test_power_bounds_mcweeny_energy_density(oracle):
    """PowerBounds vs the dominant eigenvalue (UnitTests/test_solvers.py:826-843), McWeenyStep and EnergyDensityMatrix vs
    their dense definitions (test_chemistry.py)"""
    O = oracle
    rng = np.random.default_rng(5)
    n = 31
    a = rng.uniform(0.0, 1.0, (n, n))
    a = a + a.T
    val, info = O.power_bounds(O.PSMatrix.from_scipy(sp.csc_matrix(a)), O.SolverParameters(monitor_convergence=False))
    assert abs(val - np.abs(la.eigvalsh(a)).max()) <= 1e-4
    d = sp.random(n, n, 0.3, random_state=rng, format="csc")
    d = sp.csc_matrix((d + d.T) * 0.1)
    s = sp.csc_matrix(sp.identity(n) + 0.01 * sp.csc_matrix(a))
    Dm, Sm, Hm = (O.PSMatrix.from_scipy(x) for x in (d, s, sp.csc_matrix(a)))
    dd, sd = d.toarray(), s.toarray()
    assert np.linalg.norm(O.mcweeny_step(Dm).todense() - (3 * dd @ dd - 2 * dd @ dd @ dd)) <= 1e-13
    dsd = dd @ sd @ dd
    assert np.linalg.norm(O.mcweeny_step(Dm, Sm).todense() - (3 * dsd - 2 * dd @ sd @ dsd)) <= 1e-13
    assert np.linalg.norm(O.energy_density_matrix(Hm, Dm).todense() - dd @ a @ dd) <= 1e-12


GRIDS = [(1, 1, 1, 1), (2, 1, 1, 1), (1, 2, 1, 1), (2, 2, 1, 1), (1, 1, 2, 1), (2, 1, 2, 1), (1, 2, 2, 1), (2, 2, 2, 1),
         (3, 2, 1, 1), (2, 1, 3, 1), (6, 1, 1, 1), (1, 1, 1, 8)]


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("fill", [1.0, 0.2, 0.0])
def test_multiply_vs_scipy_all_reference_grids(oracle, grid, fill):
    """process grids of UnitTests/CMakeLists.txt:42-69 (111 211 121 112 221 212 122 222 321 213 611), size 33"""
    O = oracle
    g = O.Grid(grid[0], grid[1], grid[2], threads=grid[3])
    n = 33
    a = sp.random(n, n, fill, random_state=1, format="csc")
    b = sp.random(n, n, fill, random_state=2, format="csc")
    C = O.multiply(O.PSMatrix.from_scipy(a, g), O.PSMatrix.from_scipy(b, g))
    assert abs(C.to_scipy() - a @ b).sum() < 1e-10
    Dm = O.increment(O.PSMatrix.from_scipy(a, g), O.PSMatrix.from_scipy(b, g), alpha=2.0)
    assert abs(Dm.to_scipy() - (2 * a + b)).sum() < 1e-12
    assert O.dot(O.PSMatrix.from_scipy(a, g), O.PSMatrix.from_scipy(b, g)) == pytest.approx(a.multiply(b).sum(), abs=1e-12)


def test_complex_and_mixed(oracle):
    O = oracle
    g = O.Grid(2, 2, 2)
    a = sp.random(33, 33, 0.2, random_state=1, format="csc") + 1j * sp.random(33, 33, 0.2, random_state=3, format="csc")
    b = sp.random(33, 33, 0.2, random_state=2, format="csc")
    A, B = O.PSMatrix.from_scipy(a, g), O.PSMatrix.from_scipy(b, g)
    assert abs(O.multiply(A, B).to_scipy() - a @ b).sum() < 1e-12
    assert abs(O.multiply(B, A).to_scipy() - b @ a).sum() < 1e-12
    assert O.dot(A, A) == pytest.approx((a.conj().multiply(a)).sum(), abs=1e-12)   # test_psmatrixalgebra.py:406-435


def test_padding_rule(oracle):
    """CalculateScaledDimension (PSMatrixModule.F90:1596-1618)"""
    O = oracle
    assert O.Grid(1, 1, 1).padded(33) == 33
    assert O.Grid(2, 2, 1).padded(33) == 36
    assert O.Grid(2, 2, 2).padded(33) == 40
    assert O.Grid(1, 1, 1, threads=8).padded(33) == 36      # block multiplier 4
    assert O.Grid.default_for(8).slices == 4 and (O.Grid.default_for(8).rows, O.Grid.default_for(8).cols) == (1, 2)
    assert (O.Grid.default_for(2).rows, O.Grid.default_for(2).cols, O.Grid.default_for(2).slices) == (1, 1, 2)


def test_threshold_rules_of_the_local_product(oracle):
    """sparse branch tests |alpha*v| > thr, dense branch tests |v| > thr and scales afterwards
    (PruneList.f90:27, DenseBranch.f90:14-15); strict inequality."""
    O = oracle
    a = sp.csc_matrix(np.array([[1.0, 0.0], [0.0, 2.0]]))
    b = sp.csc_matrix(np.array([[0.5, 0.0], [0.0, 0.5]]))
    AT, BT = sp.csc_matrix(a.T), sp.csc_matrix(b.T)
    c, br = O.local_gemm(AT, BT, alpha=0.5, thr=0.25, force_branch=1)      # values 0.5,1.0 ; alpha*v = .25,.5
    assert br == 1 and c.nnz == 1 and c[1, 1] == 0.5
    c, br = O.local_gemm(AT, BT, alpha=0.5, thr=0.25, force_branch=2)      # |v| = .5, 1 both > .25
    assert br == 2 and c.nnz == 2 and c[0, 0] == 0.25
    c, _ = O.local_gemm(AT, BT, alpha=1.0, thr=0.5, force_branch=1)        # strict: 0.5 > 0.5 is false
    assert c.nnz == 1


def test_increment_tail_rule(oracle):
    """AddSparseVectors.f90:57-68: once one list is exhausted the rest of the other is copied untested"""
    O = oracle
    a = sp.csc_matrix(np.array([[1e-9], [0.0], [1e-9], [1e-9]]))   # column 0: rows 0,2,3
    b = sp.csc_matrix(np.array([[0.0], [1.0], [0.0], [0.0]]))     # column 0: row 1
    c = O.local_increment(a, b, alpha=1.0, thr=1e-6)
    # row 0 of A is merged while B still has entries -> dropped; rows 2,3 come after B is exhausted -> kept
    assert sorted(c.indices.tolist()) == [1, 2, 3]
    c2 = O.local_increment(b, a, alpha=1.0, thr=1e-6)              # roles swapped: same merge
    assert sorted(c2.indices.tolist()) == [1, 2, 3]


def test_slice_threshold_rule(oracle):
    """S>1: local products at thr/(1000*S), the user threshold only in the last pairwise add
    (MatrixMultiply.f90:25-29, ReduceAndSumMatrixCleanup.f90:23-29)."""
    O = oracle
    n = 48
    rng = np.random.default_rng(3)
    a = sp.random(n, n, 0.08, random_state=rng, format="csc")
    thr = 5e-2
    C1 = O.multiply(O.PSMatrix.from_scipy(a, O.Grid(1, 1, 1)), O.PSMatrix.from_scipy(a, O.Grid(1, 1, 1)), thr=thr)
    C2 = O.multiply(O.PSMatrix.from_scipy(a, O.Grid(1, 1, 2)), O.PSMatrix.from_scipy(a, O.Grid(1, 1, 2)), thr=thr)
    full = (a @ a).toarray()
    kept1 = C1.todense()
    assert np.all((np.abs(full) > thr) == (kept1 != 0))
    # with slices, partial sums below thr survive until the last add; entries unmatched in the tail of
    # that add are never tested, so the sliced result keeps a superset of the 1-slice pattern
    assert np.all((C2.todense() != 0) >= (kept1 != 0))
    assert np.abs(C2.todense() - full)[C2.todense() != 0].max() < 1e-12


def test_monitor(oracle):
    """ConvergenceMonitorModule.F90:122-191"""
    O = oracle
    m = O.Monitor(automatic=True, tight=1e-8)
    for v in [1e-3, 1e-4, 1e-5, 1e-6]:
        m.append(v)
        assert not m.converged()
    m.append(1e-9)
    assert m.converged()                       # tight criterion
    m = O.Monitor(automatic=True, tight=1e-12)
    for v in [3e-6, 2e-6, 2.5e-6, 2.2e-6, 2.1e-6, 2.3e-6]:
        m.append(v)
    assert m.converged()                       # automatic plateau detection after 6 values
    m2 = O.Monitor(automatic=False, tight=1e-12)
    for v in [3e-6, 2e-6, 2.5e-6, 2.2e-6, 2.1e-6, 2.3e-6]:
        m2.append(v)
    assert not m2.converged()


@pytest.mark.parametrize("fn,ref", [("sign", None), ("invert", None), ("isr", None), ("sqrt", None)])
def test_solvers_vs_scipy(oracle, fn, ref):
    """UnitTests/test_solvers.py: 31x31 symmetric inputs, relative error <= 1e-4, monitor off (:58-59)"""
    O = oracle
    rng = np.random.default_rng(7)
    n = 31
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    p = O.SolverParameters(converge_diff=1e-10, threshold=0.0, monitor_convergence=False)
    if fn == "sign":
        w = np.concatenate([rng.uniform(0.5, 2.0, 16), -rng.uniform(0.5, 2.0, 15)])
        m = (q * w) @ q.T
        out, _ = O.sign_function(O.PSMatrix.from_scipy(sp.csc_matrix(m)), p)
        expect = (q * np.sign(w)) @ q.T
    else:
        w = rng.uniform(0.5, 2.0, n)
        m = (q * w) @ q.T
        M = O.PSMatrix.from_scipy(sp.csc_matrix(m))
        if fn == "invert":
            out, _ = O.invert(M, p)
            expect = np.linalg.inv(m)
        elif fn == "isr":
            out, _ = O.inverse_square_root(M, p)
            expect = (q * w ** -0.5) @ q.T
        else:
            out, _ = O.square_root(M, p)
            expect = (q * w ** 0.5) @ q.T
    assert np.linalg.norm(out.todense() - expect) / np.linalg.norm(expect) <= 1e-4


def test_exponential_restatement(oracle):
    """ComputeExponential (ExponentialSolversModule.F90:37-148), restated: pinned against scipy.linalg.expm the way the
    reference's own test is (UnitTests/test_solvers.py test_exponential: relative error <= 1e-4), on a real symmetric
    31x31 input and on the shipped complex example with a self loop at node 1 (radius 25.8 -> sigma 32, 5 squarings).
    Also pins the reference quirk the shipped example runs into: M(1,1) = 0 -> PowerBounds returns 0 in iteration 1
    -> no scaling."""
    import scipy.linalg as la
    import scipy.io as sio
    import os
    from ntpoly_b200.workloads import guo_transform
    O = oracle
    rng = np.random.default_rng(11)
    n = 31
    m = rng.uniform(0.0, 1.0, (n, n))
    m = 0.5 * (m + m.T)
    out, info = O.compute_exponential(O.PSMatrix.from_scipy(sp.csc_matrix(m)), O.SolverParameters(threshold=0.0))
    expect = la.expm(m)
    assert info.iterations > 1
    assert np.linalg.norm(out.todense() - expect) / np.linalg.norm(expect) <= 1e-8
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "complex_input.mtx")
    g = guo_transform(sio.mmread(gold))
    shifted = sp.csc_matrix(0.5 * g + 0.3 * sp.identity(g.shape[0]))
    out, info = O.compute_exponential(O.PSMatrix.from_scipy(shifted, is_complex=True), O.SolverParameters(threshold=1e-6))
    expect = la.expm(shifted.toarray())
    assert info.iterations == 6 and info.sigmas == [32.0]
    assert np.linalg.norm(out.todense() - expect) / np.linalg.norm(expect) <= 1e-8
    radius, pinfo = O.power_bounds(O.PSMatrix.from_scipy(sp.csc_matrix(0.5 * g), is_complex=True),
                                   O.SolverParameters(max_iterations=10, threshold=1e-6))
    assert radius == 0.0 and pinfo.iterations == 1


def test_hotelling_on_the_c5_graph_needs_the_larger_shift(oracle):
    """Why bench.py's c5 inverts G + (8*||G||_1 + 1)*I and not G + (||G||_1 + 1)*I: at thr 1e-6
    the dropped terms of the smaller shift keep the residual ||I - X*A||_1 above the monitor's loose cutoff 1e-2, the
    automatic exit never fires and the solve runs into max_iterations; with the larger shift it leaves through the
    stagnation rule after 8-10 iterations. The reference's monitor (restated in oracle.Monitor) decides both."""
    from ntpoly_b200.workloads import complex_hermitian_graph
    O = oracle
    n = 768
    g = complex_hermitian_graph(n)
    n1 = float(np.asarray(abs(g).sum(axis=0)).max())
    p = O.SolverParameters(converge_diff=1e-5, threshold=1e-6, max_iterations=14)
    _, slow = O.invert(O.PSMatrix.from_scipy(sp.csc_matrix(g + sp.identity(n) * (n1 + 1.0)), is_complex=True), p)
    assert slow.iterations == 15 and slow.history[-1] > 1e-2          # ran into max_iterations, residual stagnating
    assert abs(slow.history[-1] - slow.history[-2]) < 1e-3
    inv, fast = O.invert(O.PSMatrix.from_scipy(sp.csc_matrix(g + sp.identity(n) * (8.0 * n1 + 1.0)), is_complex=True), p)
    assert fast.iterations <= 10 and 1e-5 < fast.history[-1] < 1e-2   # left through the stagnation rule
    a = (g + sp.identity(n) * (8.0 * n1 + 1.0)).toarray()
    assert np.abs(inv.todense() @ a - np.eye(n)).sum(axis=0).max() < 1e-2


def test_chebyshev_low_order_coefficients_behind_the_c5_scale():
    """the reference evaluates exp through T_0..T_15, and T_k has LARGE low-order coefficients
    (x^3 in T_15: 560, x^4 in T_14: 1568), so the iterates T_k(scale*G) of a graph matrix fill in unless
    560*scale^3 stays below the threshold - the reason for bench.py's --c5-scale 0.001 at N = 32768."""
    from numpy.polynomial import chebyshev as C
    t15 = C.cheb2poly([0] * 15 + [1])
    t14 = C.cheb2poly([0] * 14 + [1])
    assert abs(t15[3]) == 560 and abs(t14[4]) == 1568
    assert 560 * 0.001 ** 3 < 1e-6 < 560 * 0.005 ** 3
