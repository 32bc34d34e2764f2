"""Container utilities either side of the path (FillMatrixDense, ResizeMatrix, GetMatrixSlice, GetMatrixBlock,
SnapMatrixToSparsityPattern) against SciPy. Mirrors reference UnitTests/test_psmatrix.py (test_slice, test_resize,
test_getblock, test_snap semantics: PSMatrixModule.F90:958-990, 1153-1225, 1718-1741; MatrixConversionModule.F90:21-61)."""
import numpy as np
import pytest
import scipy.sparse as sp

from util import random_sparse

pytestmark = pytest.mark.gpu


def to_gpu(nt, m):
    M = nt.Matrix_ps(m.shape[0], is_complex=np.iscomplexobj(m.data))
    M.fill_from_scipy(m)
    return M


def test_fill_dense(nt):
    n = 37
    M = nt.Matrix_ps(n)
    M.FillDense()
    assert M.GetSize() == n * n
    assert abs(M.to_scipy().toarray() - np.ones((n, n))).sum() == 0.0


@pytest.mark.parametrize("cplx", [False, True])
def test_resize_and_slice(nt, cplx):
    n = 50
    a = random_sparse(n, 0.15, 31, cplx)
    A = to_gpu(nt, a)
    A.Resize(30)
    assert A.GetActualDimension() == 30
    assert abs(A.to_scipy() - a[:30, :30]).sum() == 0.0
    A.Resize(70)
    assert A.GetActualDimension() == 70
    big = sp.lil_matrix((70, 70), dtype=a.dtype)
    big[:30, :30] = a[:30, :30]
    assert abs(A.to_scipy() - sp.csc_matrix(big)).sum() == 0.0
    # slice: rows 5..24, columns 10..19 (0-based inclusive) -> dimension max(20, 10)
    B, S = to_gpu(nt, a), nt.Matrix_ps(1)
    B.GetMatrixSlice(S, 5, 24, 10, 19)
    assert S.GetActualDimension() == 20 and S.IsComplex() == cplx
    want = sp.lil_matrix((20, 20), dtype=a.dtype)
    want[:20, :10] = a[5:25, 10:20]
    assert abs(S.to_scipy() - sp.csc_matrix(want)).sum() == 0.0


@pytest.mark.parametrize("cplx", [False, True])
def test_get_block(nt, cplx):
    n = 64
    a = random_sparse(n, 0.1, 32, cplx)
    A = to_gpu(nt, a)
    tl = nt.TripletList_c() if cplx else nt.TripletList_r()
    A.GetMatrixBlock(tl, 3, 20, 0, 50)                        # rows [3, 20), columns [0, 50), 0-based
    rows, cols, vals = tl.get_arrays()
    got = sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(n, n)).tocsc()
    want = sp.lil_matrix((n, n), dtype=a.dtype)
    want[3:20, 0:50] = a[3:20, 0:50]
    assert abs(got - sp.csc_matrix(want)).sum() == 0.0
    assert len(rows) == a[3:20, 0:50].nnz


@pytest.mark.parametrize("cplx", [False, True])
def test_snap_to_sparsity_pattern(nt, cplx):
    n = 80
    a = random_sparse(n, 0.1, 33, cplx)
    pat = random_sparse(n, 0.1, 34, False)
    A, P = to_gpu(nt, a), to_gpu(nt, pat)
    A.SnapToSparsityPattern(P)
    rows, cols, vals = A.get_arrays()
    # the stored pattern is exactly the pattern's; positions the matrix lacked are explicit zeros
    got_pat = sp.coo_matrix((np.ones(len(rows)), (rows - 1, cols - 1)), shape=(n, n)).tocsc()
    ref_pat = sp.csc_matrix((np.ones(pat.nnz), pat.indices, pat.indptr), shape=(n, n))
    assert (got_pat != ref_pat).nnz == 0 and len(rows) == pat.nnz
    got = sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(n, n)).tocsc()
    assert abs(got - a.multiply(ref_pat)).sum() == 0.0
