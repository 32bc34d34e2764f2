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
