"""Per-iteration helpers (IncrementMatrix, ScaleMatrix, MatrixTrace, MatrixNorm, DotMatrix, Gershgorin,
Transpose, Filter ...) on the GPU vs the oracle. Mirrors reference UnitTests/test_psmatrixalgebra.py
test_addition (:127-163), test_pairwisemultiply (:165-191), test_dot (:368-435), test_asymmetry (:288-308),
test_symmetrize (:310-329) and adds the threshold cases."""
import numpy as np
import pytest
import scipy.sparse as sp

from util import banded, compare_sparse, random_sparse

pytestmark = pytest.mark.gpu


def to_gpu(nt, m):
    M = nt.Matrix_ps(m.shape[0], is_complex=np.iscomplexobj(m.data))
    M.fill_from_scipy(m)
    return M


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("fills", [(1.0, 1.0), (0.2, 0.2), (0.0, 0.0), (1.0, 0.0), (0.0, 1.0)])
def test_addition(nt, oracle, cplx, fills):
    n = 33
    a, b = random_sparse(n, fills[0], 1, cplx), random_sparse(n, fills[1], 2, cplx)
    A, B = to_gpu(nt, a), to_gpu(nt, b)
    B.Increment(A, 1.7, 0.0)
    ref = oracle.increment(oracle.PSMatrix.from_scipy(a, is_complex=cplx), oracle.PSMatrix.from_scipy(b, is_complex=cplx), alpha=1.7)
    got = B.to_scipy()
    assert got.nnz == ref.nnz()
    assert abs(got - ref.to_scipy()).sum() == 0.0          # same arithmetic, bit exact


@pytest.mark.parametrize("thr", [1e-3, 0.3])
@pytest.mark.parametrize("n,fill", [(200, 0.05), (1000, 0.01)])
def test_addition_threshold_and_tail_rule_bit_exact(nt, oracle, n, fill, thr):
    a, b = random_sparse(n, fill, 5), random_sparse(n, fill, 6)
    A, B = to_gpu(nt, a), to_gpu(nt, b)
    B.Increment(A, -0.9, thr)
    ref = oracle.increment(oracle.PSMatrix.from_scipy(a), oracle.PSMatrix.from_scipy(b), alpha=-0.9, thr=thr).to_scipy()
    got = B.to_scipy()
    assert got.nnz == ref.nnz
    assert (got != ref).nnz == 0                              # pattern (incl. untested tails) and values identical


def test_mixed_addition(nt, oracle):
    n = 50
    a, b = random_sparse(n, 0.2, 7, True), random_sparse(n, 0.2, 8, False)
    A, B = to_gpu(nt, a), to_gpu(nt, b)
    B.Increment(A, 2.0)                                       # real += complex -> upcast (PSMatrixAlgebraModule.F90:436-439)
    assert B.IsComplex()
    assert abs(B.to_scipy() - (2 * a + b)).sum() < 1e-12
    A2, B2 = to_gpu(nt, a), to_gpu(nt, b)
    A2.Increment(B2, 2.0)
    assert abs(A2.to_scipy() - (a + 2 * b)).sum() < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
def test_scalars(nt, oracle, cplx):
    n = 300
    a, b = random_sparse(n, 0.05, 9, cplx), random_sparse(n, 0.05, 10, cplx)
    a = sp.csc_matrix(a + sp.identity(n) * 0.3)
    A, B = to_gpu(nt, a), to_gpu(nt, b)
    OA, OB = oracle.PSMatrix.from_scipy(a, is_complex=cplx), oracle.PSMatrix.from_scipy(b, is_complex=cplx)
    assert A.Trace() == pytest.approx(oracle.trace(OA), rel=1e-13)
    assert A.Norm() == pytest.approx(oracle.norm(OA), rel=1e-13)
    emin, emax = nt.EigenBounds.GershgorinBounds(A)
    omin, omax = oracle.gershgorin(OA)
    assert emin == pytest.approx(omin, rel=1e-12) and emax == pytest.approx(omax, rel=1e-12)
    if cplx:
        d = A.Dot_c(B)
        o = oracle.dot(OA, OB)
        assert d.real == pytest.approx(o.real, rel=1e-12, abs=1e-13) and d.imag == pytest.approx(o.imag, rel=1e-12, abs=1e-13)
    assert A.Dot(B) == pytest.approx(np.real(oracle.dot(OA, OB)), rel=1e-12, abs=1e-13)
    A.Scale(-2.5)
    assert abs(A.to_scipy() - (-2.5) * a).sum() < 1e-12


def test_pairwise_transpose_conjugate_filter_symmetrize(nt, oracle):
    n = 120
    a, b = random_sparse(n, 0.1, 11, True), random_sparse(n, 0.1, 12, True)
    A, B, C, T = to_gpu(nt, a), to_gpu(nt, b), nt.Matrix_ps(n), nt.Matrix_ps(n)
    C.PairwiseMultiply(A, B)
    assert abs(C.to_scipy() - a.multiply(b)).sum() < 1e-13
    T.Transpose(A)
    assert abs(T.to_scipy() - a.T).sum() == 0.0
    T.Conjugate()
    assert abs(T.to_scipy() - a.conj().T).sum() == 0.0
    assert A.MeasureAsymmetry() == pytest.approx(float(np.asarray(abs(a - a.conj().T).sum(axis=0)).max()), rel=1e-12)
    A.Symmetrize()
    assert abs(A.to_scipy() - 0.5 * (a + a.conj().T)).sum() < 1e-12
    F = to_gpu(nt, b)
    F.Filter(0.5)
    ref = b.copy(); ref.data[abs(ref.data) <= 0.5] = 0; ref.eliminate_zeros()
    assert (F.to_scipy() != ref).nnz == 0


def test_identity_permutation_and_triplet_roundtrip(nt):
    n = 77
    I = nt.Matrix_ps(n)
    I.FillIdentity()
    assert I.IsIdentity() and I.GetSize() == n and I.Trace() == n
    p = nt.Permutation(n)
    p.SetRandomPermutation(seed=5)
    a = random_sparse(n, 0.1, 13)
    A, P, U = to_gpu(nt, a), nt.Matrix_ps(n), nt.Matrix_ps(n)
    nt.LoadBalancer.PermuteMatrix(A, P, p)
    assert not np.allclose(P.to_scipy().toarray(), a.toarray())
    nt.LoadBalancer.UndoPermuteMatrix(P, U, p)
    assert abs(U.to_scipy() - a).sum() == 0.0
    # element-wise triplet API (reference TripletList / Matrix_ps::FillFromTripletList)
    tl = nt.TripletList_r()
    coo = a.tocoo()
    for r, c, v in zip(coo.row, coo.col, coo.data):
        tl.Append(nt.Triplet_r(int(c) + 1, int(r) + 1, float(v)))
    B = nt.Matrix_ps(n)
    B.FillFromTripletList(tl)
    out = nt.TripletList_r()
    B.GetTripletList(out)
    assert out.GetSize() == a.nnz
    t0 = out.GetTripletAt(0)
    assert a[t0.index_row - 1, t0.index_column - 1] == t0.point_value
    assert abs(B.to_scipy() - a).sum() == 0.0


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("kind", ["random", "reverse", "default"])
def test_permutation_by_relabelling_equals_the_two_products(nt, cplx, kind):
    """PermuteMatrix / UndoPermuteMatrix as an index relabelling on the device against the
    reference's form, two products by permutation matrices (LoadBalancerModule.F90:38-47, 77-86): same matrix bit
    for bit, also with a padded logical dimension and with an explicit zero among the entries"""
    n = 123
    a = random_sparse(n, 0.08, 21, cplx).tolil()
    a[3, 7] = 0.0
    a = sp.csc_matrix(a)
    p = nt.Permutation(n)
    if kind == "random":
        p.SetRandomPermutation(seed=11)
    elif kind == "reverse":
        p.SetReversePermutation()
    A = to_gpu(nt, a)
    res = {}
    for gemm in (True, False):
        nt.set_permute_gemm(gemm)
        P, U = nt.Matrix_ps(n), nt.Matrix_ps(n)
        nt.reset_counters()
        nt.LoadBalancer.PermuteMatrix(A, P, p)
        nt.LoadBalancer.UndoPermuteMatrix(P, U, p)
        assert nt.counters()["multiplies"] == (4 if gemm else 0)
        res[gemm] = (P.get_arrays(), U.get_arrays())
    nt.set_permute_gemm(False)
    for k in range(2):
        for x, y in zip(res[True][k], res[False][k]):
            assert np.array_equal(x, y)
    rows, cols, vals = res[False][1]
    back = sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(n, n)).tocsc()
    assert abs(back - a).sum() == 0.0
    if kind != "default":
        prow, pcol, _ = res[False][0]
        assert not (np.array_equal(prow, rows) and np.array_equal(pcol, cols))


def test_matrix_market_roundtrip(nt, tmp_path):
    import os
    import scipy.io as sio
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    H = nt.Matrix_ps(os.path.join(gold, "premade_Hamiltonian.mtx"))
    ref = sp.csc_matrix(sio.mmread(os.path.join(gold, "premade_Hamiltonian.mtx")))
    assert abs(H.to_scipy() - ref).sum() < 1e-15
    out = str(tmp_path / "out.mtx")
    H.WriteToMatrixMarket(out)
    assert abs(sp.csc_matrix(sio.mmread(out)) - ref).sum() < 1e-15
    from ntpoly_b200.workloads import guo_transform
    hc = guo_transform(sio.mmread(os.path.join(gold, "complex_input.mtx")))
    cpath = str(tmp_path / "herm.mtx")
    sio.mmwrite(cpath, hc)
    G = nt.Matrix_ps(cpath)
    assert G.IsComplex() and abs(G.to_scipy() - hc).sum() < 1e-12
    G.WriteToMatrixMarket(out)
    assert abs(sp.csc_matrix(sio.mmread(out)) - hc).sum() < 1e-12


def test_large_banded_helpers_properties(nt):
    """size-independent properties at a config-1 sized input: linearity of the add, Tr(aX)=a Tr(X),
    X - X == 0 under threshold 0 keeps only cancellation-free pattern"""
    n = 8192
    a = banded(n)
    A = to_gpu(nt, a)
    B = nt.Matrix_ps(A)
    t = A.Trace()
    B.Scale(3.0)
    assert B.Trace() == pytest.approx(3.0 * t, rel=1e-13)
    B.Increment(A, -3.0)            # 3A - 3A: matched entries cancel exactly and are dropped (|0| > 0 false)
    assert B.GetSize() == 0
    C = nt.Matrix_ps(n)
    C.Gemm(A, A, None, threshold=1e-8)
    assert C.Dot(A) == pytest.approx(A.Dot(C), rel=1e-13)
    assert C.MeasureAsymmetry() < 1e-12    # square of a symmetric matrix stays symmetric
