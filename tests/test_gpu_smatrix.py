"""The LOCAL sparse matrix layer on the GPU (Matrix_lsr / Matrix_lsc, the C entry points of Source/C/SMatrix_c.h)
against SciPy and the oracle's local kernels. Mirrors reference UnitTests/test_matrix.py (same shapes and fills,
:77-88; test_addition :125, test_dot :171, test_transpose :186, test_pairwise :202, test_multiply* :224-362,
test_get_row/column :364-400, test_scalediag :402) and adds thresholded and tile-path products."""
import numpy as np
import pytest
import scipy.io as sio
import scipy.sparse as sp

from util import banded, compare_sparse

pytestmark = pytest.mark.gpu

# (rows, columns, sparsity) of the reference suite
SHAPES = [(2, 4, 0.0), (8, 8, 0.0), (2, 2, 1.0), (4, 4, 1.0), (19, 19, 1.0), (4, 2, 1.0), (2, 4, 1.0), (4, 4, 0.2),
          (8, 8, 1.0), (57, 31, 0.3), (31, 57, 0.05)]


def rnd(rows, cols, fill, seed, cplx):
    rng = np.random.default_rng(seed)
    m = sp.random(rows, cols, fill, random_state=rng, format="csc")
    if cplx:
        m = m + 1j * sp.random(rows, cols, fill, random_state=rng, format="csc")
    return sp.csc_matrix(m)


def cls(nt, cplx):
    return (nt.Matrix_lsc, nt.MatrixMemoryPool_c) if cplx else (nt.Matrix_lsr, nt.MatrixMemoryPool_r)


def dense_diff(got, ref):
    return float(abs(sp.csc_matrix(got) - sp.csc_matrix(ref)).sum())


@pytest.mark.parametrize("cplx", [False, True])
def test_construct_roundtrip_copy_and_files(nt, cplx, tmp_path):
    M, _ = cls(nt, cplx)
    for k, (r, c, f) in enumerate(SHAPES):
        a = rnd(r, c, f, 100 + k, cplx)
        A = M.from_scipy(a)
        assert (A.GetRows(), A.GetColumns()) == (r, c)
        assert dense_diff(A.to_scipy(), a) == 0.0
        B = M(A)                                              # copy
        A.Scale(2.0)
        assert dense_diff(B.to_scipy(), a) == 0.0
        assert dense_diff(A.to_scipy(), 2.0 * a) == 0.0
        # MatrixMarket: written by SciPy -> read -> written by us -> read by SciPy (test_read, test_readcircular)
        f1, f2 = str(tmp_path / "m1.mtx"), str(tmp_path / "m2.mtx")
        sio.mmwrite(f1, a)
        C = M(f1)
        C.WriteToMatrixMarket(f2)
        back = sp.csc_matrix(sio.mmread(f2))
        assert back.shape == a.shape
        assert dense_diff(back, a) < 1e-14 * max(1, a.nnz)
    # symmetric / hermitian storage is expanded on reading (test_readsymmetric)
    a = rnd(19, 19, 0.3, 7, cplx)
    a = sp.csc_matrix(a + a.conj().T)
    f1 = str(tmp_path / "sym.mtx")
    sio.mmwrite(f1, a)
    assert dense_diff(M(f1).to_scipy(), a) < 1e-13
    # sorted triplet list (TripletListModule SortTripletList)
    tl = nt.TripletList_c() if cplx else nt.TripletList_r()
    coo = sp.coo_matrix(rnd(31, 57, 0.2, 8, cplx))
    perm = np.random.default_rng(0).permutation(coo.nnz)
    tl.set_arrays(coo.row[perm] + 1, coo.col[perm] + 1, coo.data[perm])
    rows, cols, _ = tl.Sort(57).get_arrays()
    key = cols.astype(np.int64) * 100 + rows
    assert np.all(np.diff(key) > 0)


@pytest.mark.parametrize("cplx", [False, True])
def test_addition_dot_pairwise_transpose(nt, oracle, cplx):
    M, _ = cls(nt, cplx)
    for k, (r, c, f) in enumerate(SHAPES):
        a, b = rnd(r, c, f, 200 + k, cplx), rnd(r, c, f, 300 + k, cplx)
        A, B = M.from_scipy(a), M.from_scipy(b)
        # addition: bit exact against the oracle's merge (same arithmetic), incl. a threshold
        for alpha, thr in ((1.3, 0.0), (-0.8, 0.2)):
            C = M(B)
            C.Increment(A, alpha, thr)
            ref = oracle.local_increment(a, b, alpha=alpha, thr=thr, is_complex=cplx)
            got = C.to_scipy()
            assert got.nnz == ref.nnz
            assert dense_diff(got, ref) == 0.0
        # zero matrices on either side (test_addzero, test_addzeroreverse)
        Z = M(c, r)
        Z.Increment(A, 1.0, 0.0)
        assert dense_diff(Z.to_scipy(), a) == 0.0
        A2 = M(A)
        A2.Increment(M(c, r), 1.0, 0.0)
        assert dense_diff(A2.to_scipy(), a) == 0.0
        # dot: sum conj(a) b
        want = complex((a.conj().multiply(b)).sum())
        got = A.Dot(B)
        assert abs(got - (want if cplx else want.real)) <= 1e-13 * max(1.0, abs(want))
        # pairwise
        P = M(c, r)
        P.PairwiseMultiply(A, B)
        assert dense_diff(P.to_scipy(), a.multiply(b)) <= 1e-15 * max(1, a.nnz)
        # transpose (+ conjugate)
        T = M(r, c)
        T.Transpose(A)
        assert (T.GetRows(), T.GetColumns()) == (c, r)
        assert dense_diff(T.to_scipy(), a.T) == 0.0
        if cplx:
            T.Conjugate()
            assert dense_diff(T.to_scipy(), a.conj().T) == 0.0


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_multiply_all_transpose_flags(nt, oracle, cplx, ta, tb):
    """C = alpha*op(A)*op(B) (+ beta*C) for every flag pair at the reference's shapes, against SciPy at the
    reference's tolerance and against the oracle's local kernel at 1e-10"""
    M, Pool = cls(nt, cplx)
    rng = np.random.default_rng(5)
    for k, (r, c, f) in enumerate(SHAPES):
        a = rnd(r, c, f, 400 + k, cplx)
        b = rnd(r, c, f, 500 + k, cplx).T                     # c x r, like the reference's getH() minus the conjugation
        b = sp.csc_matrix(b)
        opa, opb = a, b
        sa, sb = (sp.csc_matrix(a.T) if ta else a), (sp.csc_matrix(b.T) if tb else b)   # what is stored and passed
        alpha = float(rng.uniform(1.0, 2.0))
        A, B = M.from_scipy(sa), M.from_scipy(sb)
        C = M(r, r)
        pool = Pool(r, r)
        C.Gemm(A, B, ta, tb, alpha, 0.0, 0.0, pool)
        want = alpha * (opa @ opb)
        got = C.to_scipy()
        assert (C.GetRows(), C.GetColumns()) == (r, r)
        assert abs(got - want).sum() <= 1e-12 * max(1.0, abs(want).sum())
        # oracle's local kernel takes op(A)^T and op(B)^T in CSC
        ref, _ = oracle.local_gemm(sp.csc_matrix(opa.T), sp.csc_matrix(opb.T), alpha=alpha, thr=0.0, is_complex=cplx)
        compare_sparse(got, ref, thr=0.0)
        # beta: C = alpha*A*B + beta*C
        C.Gemm(A, B, ta, tb, alpha, 0.5, 0.0, pool)
        assert abs(C.to_scipy() - 1.5 * want).sum() <= 1e-12 * max(1.0, abs(want).sum())


def test_multiply_zero_operand(nt):
    M, Pool = cls(nt, False)
    a = rnd(19, 19, 1.0, 1, False)
    A, Z = M.from_scipy(a), M(19, 19)
    C = M(19, 19)
    C.Gemm(A, Z, False, False, 1.0, 0.0, 0.0, Pool(19, 19))
    assert C.to_scipy().nnz == 0
    C.Gemm(Z, A, False, False, 1.0, 0.0, 0.0, None)
    assert C.to_scipy().nnz == 0


@pytest.mark.parametrize("n,fill,thr,alpha", [(64, 0.5, 0.3, 0.25), (300, 0.03, 1e-3, -1.7), (128, 1.0, 5.0, 0.1)])
def test_multiply_threshold_rules_match_oracle(nt, oracle, n, fill, thr, alpha):
    """dense rule (|v| > thr before alpha) when both operands are > 10 % full, sparse rule (|alpha*v| > thr) otherwise
    (GemmMatrix.f90:59-61, DenseBranch.f90:14-15, PruneList.f90:27)"""
    M, Pool = cls(nt, False)
    a, b = rnd(n, n, fill, 11, False), rnd(n, n, fill, 12, False)
    A, B, C = M.from_scipy(a), M.from_scipy(b), M(n, n)
    C.Gemm(A, B, False, False, alpha, 0.0, thr, Pool(n, n))
    ref, branch = oracle.local_gemm(sp.csc_matrix(a.T), sp.csc_matrix(b.T), alpha=alpha, thr=thr)
    assert branch == (2 if fill > 0.1 else 1)
    got = C.to_scipy()
    compare_sparse(got, ref, thr=thr if branch == 1 else thr * abs(alpha))   # |kept value| at the edge of the rule
    assert abs(got.nnz - ref.nnz) <= max(2, 1e-5 * ref.nnz)


def test_rectangular_banded_product_runs_on_the_tile_path(nt, oracle):
    """a tall slab of a banded matrix times a wide one: the shapes the distributed multiply hands to the local
    kernel on column-split grids; long columns -> FP64 tensor-core tile path"""
    M, Pool = cls(nt, False)
    full = banded(1536, 40) * 0.5
    a = sp.csc_matrix(full[:, :1024])                         # 1536 x 1024
    b = sp.csc_matrix(full[:1024, :768])                      # 1024 x 768
    A, B, C = M.from_scipy(a), M.from_scipy(b), M(768, 1536)
    nt.reset_counters()
    C.Gemm(A, B, False, False, 1.0, 0.0, 1e-7, Pool(768, 1536))
    assert (C.GetRows(), C.GetColumns()) == (1536, 768)
    assert nt.tile_counters()["tile_products"] == 1
    ref, _ = oracle.local_gemm(sp.csc_matrix(a.T), sp.csc_matrix(b.T), alpha=1.0, thr=1e-7)
    compare_sparse(C.to_scipy(), ref, thr=1e-7)
    # the product is a valid operand of the next local call (its entries may be deferred)
    D = M(768, 1536)
    D.Transpose(C)
    assert dense_diff(D.to_scipy(), C.to_scipy().T) == 0.0


@pytest.mark.parametrize("cplx", [False, True])
def test_get_row_column_and_scalediag(nt, cplx):
    M, _ = cls(nt, cplx)
    rng = np.random.default_rng(3)
    for k, (r, c, f) in enumerate(SHAPES):
        a = rnd(r, c, f, 600 + k, cplx)
        A = M.from_scipy(a)
        i, j = int(rng.integers(0, r)), int(rng.integers(0, c))
        R = M(c, 1)
        A.ExtractRow(i, R)
        assert (R.GetRows(), R.GetColumns()) == (1, c)
        assert dense_diff(R.to_scipy(), a[i, :]) == 0.0
        Cc = M(1, r)
        A.ExtractColumn(j, Cc)
        assert (Cc.GetRows(), Cc.GetColumns()) == (r, 1)
        assert dense_diff(Cc.to_scipy(), a[:, j]) == 0.0
        # column scaling by a triplet list (test_scalediag): column i scaled by i
        tl = nt.TripletList_c() if cplx else nt.TripletList_r()
        Trip = nt.Triplet_c if cplx else nt.Triplet_r
        want = sp.lil_matrix(a)
        for col in range(c):
            tl.Append(Trip(col + 1, col + 1, (col + 0.5j) if cplx else float(col)))
            want[:, col] = want[:, col] * ((col + 0.5j) if cplx else float(col))
        A.DiagonalScale(tl)
        assert dense_diff(A.to_scipy(), sp.csc_matrix(want)) <= 1e-14 * max(1.0, abs(sp.csc_matrix(want)).sum())


def test_distributed_diagonal_scale(nt):
    """MatrixDiagonalScale_ps (PSMatrixAlgebraModule.F90:507-532): column j of the distributed matrix scaled by the
    listed value; repeated columns compose"""
    n = 203
    a = sp.csc_matrix(banded(n, 9))
    A = nt.Matrix_ps(n)
    A.fill_from_scipy(a)
    tl = nt.TripletList_r()
    d = np.ones(n)
    for col in (0, 5, 5, 17, n - 1):
        tl.Append(nt.Triplet_r(col + 1, col + 1, 1.5 + col))
        d[col] *= 1.5 + col
    A.DiagonalScale(tl)
    want = a @ sp.diags(d)
    assert abs(A.to_scipy() - want).sum() <= 1e-14 * abs(want).sum()
