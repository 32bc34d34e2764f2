"""Tile-space helpers of the fused TRS2 / TRS4 steps against plain scipy arithmetic of the reference's call sequence
(ScaleMatrix / IncrementMatrix / DotMatrix / MatrixTrace on the same matrices)."""
import numpy as np
import pytest
import scipy.sparse as sp

from util import banded, compare_sparse

pytestmark = pytest.mark.gpu


def thresholded_add(a, p, b, q, thr):
    """IncrementMatrix(P, Q', a, thr) with Q' = ScaleMatrix(Q, b), on dense arrays, one local row block: matched entries
    kept iff |v| > thr; an unmatched entry kept iff |w| > thr or no entry of the other operand follows further down
    in the column (the reference's untested tail, AddSparseVectors.f90:21-70)"""
    p, q = np.asarray(p.todense()), np.asarray(q.todense())
    n = p.shape[0]
    v = a * p + b * q
    hasp, hasq = p != 0, q != 0
    rows = np.arange(n)[:, None]
    lastp = np.where(hasp, rows, -1).max(axis=0)[None, :]
    lastq = np.where(hasq, rows, -1).max(axis=0)[None, :]
    keep = (hasp & hasq & (abs(v) > thr)) | (hasp & ~hasq & ((abs(v) > thr) | (lastq < rows))) \
        | (~hasp & hasq & ((abs(v) > thr) | (lastp < rows)))
    keep &= v != 0
    return sp.csc_matrix(np.where(keep, v, 0.0))


def to_gpu(nt, m):
    M = nt.Matrix_ps(m.shape[0])
    M.fill_from_scipy(m)
    return M


@pytest.fixture()
def iterates(nt):
    """X (built from CSC) and X2 = X*X (a tile product: lives as tile forms), banded so that the tile path is taken"""
    n = 1536
    x = sp.csc_matrix(banded(n, half_bandwidth=40) * 0.4 + sp.identity(n) * 0.5)
    X, X2 = to_gpu(nt, x), nt.Matrix_ps(n)
    nt.reset_counters()
    X2.Gemm(X, X, None, threshold=1e-7)
    assert nt.tile_counters()["tile_products"] == 1
    return n, x, X, X2, X2.to_scipy()


def test_tile_scalars(nt, iterates):
    n, x, X, X2, x2 = iterates
    eye = sp.identity(n, format="csc")
    fx, gx = 4.0 * x - 3.0 * x2, eye - 2.0 * x + x2
    got = nt.tile_scalars(1, X2, X)
    assert got is not None
    assert got[0] == pytest.approx(x2.multiply(fx).sum(), rel=1e-12)
    assert got[1] == pytest.approx(x2.multiply(gx).sum(), rel=1e-12)
    assert nt.tile_scalars(0, X2, X)[0] == pytest.approx(x2.multiply(x).sum(), rel=1e-12)
    assert nt.tile_scalars(2, X2)[0] == pytest.approx(x2.diagonal().sum(), rel=1e-12)


@pytest.mark.parametrize("mode", [0, 1])
def test_tile_combine(nt, iterates, mode):
    n, x, X, X2, x2 = iterates
    eye = sp.identity(n, format="csc")
    Out = nt.Matrix_ps(n)
    if mode == 0:
        thr = 1e-7
        assert nt.tile_combine(X2, X, Out, mode=0, alpha=-1.0, beta=2.0, threshold=thr)
        want = thresholded_add(-1.0, x2, 2.0, x, thr)
    else:
        thr, sigma = 0.0, 1.7
        assert nt.tile_combine(X2, X, Out, mode=1, sigma=sigma)
        want = sp.csc_matrix((4.0 * x - 3.0 * x2) + sigma * (eye - 2.0 * x + x2))
    got = Out.to_scipy()
    want.eliminate_zeros()
    compare_sparse(got, want, thr, tol=1e-14)
    assert abs(got.nnz - want.nnz) <= 2
    # the result is a tile-space matrix like a product's: it can be both operands of the next product
    P, Pref = nt.Matrix_ps(n), nt.Matrix_ps(n)
    P.Gemm(Out, Out, None, threshold=1e-7)
    W = to_gpu(nt, got)
    Pref.Gemm(W, W, None, threshold=1e-7)
    compare_sparse(P.to_scipy(), Pref.to_scipy(), 1e-7, tol=1e-12)
    assert Out.Trace() == pytest.approx(got.diagonal().sum(), rel=1e-12)


def test_tile_space_helpers_on_product_written_operands(nt, iterates):
    """second iteration of a fused TRS2 / TRS4 loop: BOTH operands of the helpers are tile-space results (X1 from a
    combine, X1^2 from a product of it), nothing was ever built from CSC"""
    n, x, X, X2, x2 = iterates
    eye = sp.identity(n, format="csc")
    thr = 1e-7
    X1 = nt.Matrix_ps(n)
    assert nt.tile_combine(X2, X, X1, mode=0, alpha=-1.0, beta=2.0, threshold=thr)
    x1 = X1.to_scipy()
    X1sq = nt.Matrix_ps(n)
    X1sq.Gemm(X1, X1, None, threshold=thr)
    x1sq = X1sq.to_scipy()
    fx, gx = 4.0 * x1 - 3.0 * x1sq, eye - 2.0 * x1 + x1sq
    got = nt.tile_scalars(1, X1sq, X1)
    assert got[0] == pytest.approx(x1sq.multiply(fx).sum(), rel=1e-12)
    assert got[1] == pytest.approx(x1sq.multiply(gx).sum(), rel=1e-12)
    assert nt.tile_scalars(0, X1sq, X1)[0] == pytest.approx(x1sq.multiply(x1).sum(), rel=1e-12)
    assert nt.tile_scalars(2, X1)[0] == pytest.approx(x1.diagonal().sum(), rel=1e-12)
    for mode, sigma in ((0, 0.0), (1, 2.3)):
        Out = nt.Matrix_ps(n)
        if mode == 0:
            assert nt.tile_combine(X1sq, X1, Out, mode=0, alpha=-1.0, beta=2.0, threshold=thr)
            want = thresholded_add(-1.0, x1sq, 2.0, x1, thr)
            t = thr
        else:
            assert nt.tile_combine(X1sq, X1, Out, mode=1, sigma=sigma)
            want = sp.csc_matrix(fx + sigma * gx)
            t = 0.0
        want.eliminate_zeros()
        compare_sparse(Out.to_scipy(), want, t, tol=1e-14)
