"""Asynchronous egress (ntb_GetMatrixArraysAsync_ps / ntb_EgressWait) returns exactly what the blocking call returns."""
import numpy as np
import pytest
import torch

from util import banded

pytestmark = pytest.mark.gpu


def test_async_egress_matches_blocking(nt):
    n = 3000
    a = banded(n, half_bandwidth=20)
    A, C = nt.Matrix_ps(n), nt.Matrix_ps(n)
    A.fill_from_scipy(a)
    C.Gemm(A, A, None, threshold=1e-9)
    ref = C.get_arrays()
    cap = len(ref[0]) + 16
    bufs = [(torch.empty(cap, dtype=torch.int32).pin_memory().numpy(),
             torch.empty(cap, dtype=torch.int32).pin_memory().numpy(),
             torch.empty(cap, dtype=torch.float64).pin_memory().numpy()) for _ in range(2)]
    out0 = C.get_arrays_async(bufs[0])
    C.Scale(2.0)                                  # the matrix changes while the first copy may still be in flight
    out1 = C.get_arrays_async(bufs[1])
    nt.egress_wait()
    for k in range(2):
        assert np.array_equal(out0[k], ref[k]) and np.array_equal(out1[k], ref[k])
    assert np.array_equal(out0[2], ref[2]) and np.array_equal(out1[2], 2.0 * ref[2])
    nt.egress_wait()                              # idempotent
    # buffers that cannot hold the block: nothing is written, the required count comes back (the call used to
    # trust the caller's buffers)
    small = tuple(x[: len(ref[0]) - 1] for x in bufs[0])
    before = [x.copy() for x in bufs[0]]
    with pytest.raises(ValueError, match=str(len(ref[0]))):
        C.get_arrays_async(small)
    nt.egress_wait()
    assert all(np.array_equal(x, y) for x, y in zip(before, bufs[0]))


def test_sorted_list_ingest_equals_sorting_ingest(nt):
    """a list that already is the rank's block in column-major order is taken as it comes (no sort, no gather);
    the same entries in any other order, or with duplicates, go through the sort - both give the same matrix"""
    n = 2500
    a = banded(n, half_bandwidth=15).tocsc()
    coo = a.tocoo()                                # csc -> coo keeps the column-major order
    rows, cols, vals = coo.row.astype(np.int32) + 1, coo.col.astype(np.int32) + 1, coo.data
    A, B, C = nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n)
    nt.reset_counters()
    A.fill_from_arrays(rows, cols, vals)
    assert nt.sorted_ingests() == 1
    perm = np.random.default_rng(1).permutation(len(rows))
    B.fill_from_arrays(rows[perm], cols[perm], vals[perm])
    assert nt.sorted_ingests() == 1               # shuffled: sorted on the device
    # duplicates are summed by the sorting path: split every value into two halves
    C.fill_from_arrays(np.concatenate([rows, rows]), np.concatenate([cols, cols]), np.concatenate([0.5 * vals, 0.5 * vals]))
    assert nt.sorted_ingests() == 1
    ra, rb, rc = A.get_arrays(), B.get_arrays(), C.get_arrays()
    for k in range(3):
        assert np.array_equal(ra[k], rb[k]) and np.array_equal(ra[k], rc[k])
    assert np.array_equal(ra[0], rows) and np.array_equal(ra[1], cols) and np.array_equal(ra[2], vals)
    # the matrix built without a sort is a full citizen: product against the sorted one
    P, Q = nt.Matrix_ps(n), nt.Matrix_ps(n)
    P.Gemm(A, A, None, threshold=1e-9)
    Q.Gemm(B, B, None, threshold=1e-9)
    assert all(np.array_equal(x, y) for x, y in zip(P.get_arrays(), Q.get_arrays()))


def test_staged_ingest_pipeline(nt):
    """ntb_StageArrays + ntb_FillMatrixFromStaged_ps: the copies of the next input are in flight while the current
    matrix is being worked on; results equal the blocking ingest"""
    n = 4000
    mats = [banded(n, half_bandwidth=10 + 3 * k, seed=k).tocsc() for k in range(3)]
    pins = []
    for m in mats:
        coo = m.tocoo()
        pins.append([torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()
                     for x in (coo.row.astype(np.int32) + 1, coo.col.astype(np.int32) + 1, coo.data)])
    X, C, R = nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n)
    st = nt.stage_arrays(*pins[0])
    got = []
    for k in range(3):
        nxt = nt.stage_arrays(*pins[k + 1]) if k + 1 < 3 else None
        X.fill_from_staged(st)
        C.Gemm(X, X, None, threshold=1e-9)
        got.append(C.get_arrays())
        st = nxt
    for k in range(3):
        R.fill_from_arrays(*pins[k])
        C.Gemm(R, R, None, threshold=1e-9)
        ref = C.get_arrays()
        assert all(np.array_equal(x, y) for x, y in zip(got[k], ref))
    # a stage that is never consumed is released cleanly
    unused = nt.stage_arrays(*pins[0])
    del unused
    nt.synchronize()
