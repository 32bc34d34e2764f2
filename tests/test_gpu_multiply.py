"""Parity of the CUDA hot path (through the C ABI) with the CPU oracle.
Mirrors reference UnitTests/test_psmatrixalgebra.py (sizes/fills :74,:103-107) and adds the
threshold / alpha / beta cases the reference never tests."""
import numpy as np
import pytest
import scipy.sparse as sp

from util import banded, compare_sparse, random_sparse

pytestmark = pytest.mark.gpu


def to_gpu(nt, m, n=None):
    n = n or m.shape[0]
    M = nt.Matrix_ps(n, is_complex=np.iscomplexobj(m.data))
    M.fill_from_scipy(m)
    return M


@pytest.mark.parametrize("fill", [1.0, 0.2, 0.0])
@pytest.mark.parametrize("cplx", [False, True])
def test_multiply_reference_sizes(nt, oracle, fill, cplx):
    n = 33
    a = random_sparse(n, fill, 1, cplx)
    b = random_sparse(n, fill, 2, cplx)
    A, B, C = to_gpu(nt, a), to_gpu(nt, b), nt.Matrix_ps(n)
    pool = nt.PMatrixMemoryPool(A)
    C.Gemm(A, B, pool)
    ref = oracle.multiply(oracle.PSMatrix.from_scipy(a, is_complex=cplx), oracle.PSMatrix.from_scipy(b, is_complex=cplx))
    compare_sparse(C.to_scipy(), ref.to_scipy())
    # and against scipy the way the reference tests do (THRESHOLD 1e-4, helpers.py:13)
    assert abs(C.to_scipy() - a @ b).sum() < 1e-4


@pytest.mark.parametrize("pair", [(1.0, 0.0), (0.0, 1.0)])
def test_multiply_empty_operand(nt, oracle, pair):
    n = 33
    a = random_sparse(n, pair[0], 3)
    b = random_sparse(n, pair[1], 4)
    A, B, C = to_gpu(nt, a), to_gpu(nt, b), nt.Matrix_ps(n)
    C.Gemm(A, B)
    assert C.GetSize() == 0


@pytest.mark.parametrize("n,fill,thr,alpha", [
    (257, 0.03, 0.0, 1.0), (257, 0.03, 1e-3, 1.0), (512, 0.05, 1e-2, 0.5), (512, 0.3, 5e-2, -2.0),
    (1000, 0.004, 1e-4, 1.7)])
def test_multiply_threshold_alpha(nt, oracle, n, fill, thr, alpha):
    a = random_sparse(n, fill, 11)
    b = random_sparse(n, fill, 12)
    A, B, C = to_gpu(nt, a), to_gpu(nt, b), nt.Matrix_ps(n)
    C.Gemm(A, B, None, alpha=alpha, threshold=thr)
    st = oracle.MultiplyStats()
    ref = oracle.multiply(oracle.PSMatrix.from_scipy(a), oracle.PSMatrix.from_scipy(b), alpha=alpha, thr=thr, stats=st)
    compare_sparse(C.to_scipy(), ref.to_scipy(), thr)
    assert C.GetSize() == pytest.approx(ref.nnz(), abs=max(3, ref.nnz() * 1e-4))


def test_multiply_beta(nt, oracle):
    n = 300
    a, b, c = random_sparse(n, 0.02, 21), random_sparse(n, 0.02, 22), random_sparse(n, 0.02, 23)
    A, B, C = to_gpu(nt, a), to_gpu(nt, b), to_gpu(nt, c)
    C.Gemm(A, B, None, alpha=0.7, beta=-1.3, threshold=1e-3)
    ref = oracle.multiply(oracle.PSMatrix.from_scipy(a), oracle.PSMatrix.from_scipy(b),
                          C=oracle.PSMatrix.from_scipy(c), alpha=0.7, beta=-1.3, thr=1e-3)
    compare_sparse(C.to_scipy(), ref.to_scipy(), 1e-3)


def test_multiply_mixed_real_complex(nt, oracle):
    n = 120
    a = random_sparse(n, 0.05, 31, True)
    b = random_sparse(n, 0.05, 32, False)
    A, B, C, D = to_gpu(nt, a), to_gpu(nt, b), nt.Matrix_ps(n), nt.Matrix_ps(n)
    C.Gemm(A, B)
    D.Gemm(B, A)
    compare_sparse(C.to_scipy(), oracle.multiply(oracle.PSMatrix.from_scipy(a), oracle.PSMatrix.from_scipy(b)).to_scipy())
    compare_sparse(D.to_scipy(), oracle.multiply(oracle.PSMatrix.from_scipy(b), oracle.PSMatrix.from_scipy(a)).to_scipy())


def test_multiply_banded_config1_small(nt, oracle):
    """config 1 shape (banded, 165 nnz/row, thr 1e-8) at a size the oracle finishes in seconds"""
    n = 2048
    a = banded(n)
    A, C = to_gpu(nt, a), nt.Matrix_ps(n)
    nt.reset_counters()
    nt.set_flop_counting(True)
    C.Gemm(A, A, None, threshold=1e-8)
    nt.set_flop_counting(False)
    st = oracle.MultiplyStats()
    ref = oracle.multiply(oracle.PSMatrix.from_scipy(a), oracle.PSMatrix.from_scipy(a), thr=1e-8, stats=st)
    compare_sparse(C.to_scipy(), ref.to_scipy(), 1e-8)
    assert nt.counters()["flops"] == pytest.approx(st.flops, rel=1e-12)
    assert nt.counters()["launches"] > 0


def test_multiply_wide_rows_use_cta_and_slab_bins(nt, oracle):
    """windows wider than a warp window / than shared memory exercise bins 5 and 6"""
    n = 40000
    rng = np.random.default_rng(5)
    rows = rng.integers(0, n, 6000)
    cols = rng.integers(0, n, 6000)
    a = sp.coo_matrix((rng.standard_normal(6000), (rows, cols)), shape=(n, n)).tocsc()
    a.sum_duplicates()
    A, C = to_gpu(nt, a), nt.Matrix_ps(n)
    C.Gemm(A, A)
    compare_sparse(C.to_scipy(), a @ a)


def test_dense_rule_is_applied_before_alpha(nt, oracle):
    """dense branch keeps |v|>thr then scales; sparse branch tests |alpha*v| (DenseBranch.f90:14-15, PruneList.f90:27)"""
    n = 64
    a = random_sparse(n, 0.5, 41)
    b = random_sparse(n, 0.5, 42)
    A, B, C = to_gpu(nt, a), to_gpu(nt, b), nt.Matrix_ps(n)
    thr, alpha = 0.5, 0.25
    C.Gemm(A, B, None, alpha=alpha, threshold=thr)
    st = oracle.MultiplyStats()
    ref = oracle.multiply(oracle.PSMatrix.from_scipy(a), oracle.PSMatrix.from_scipy(b), alpha=alpha, thr=thr, stats=st)
    assert set(st.branches) == {2}
    compare_sparse(C.to_scipy(), ref.to_scipy(), thr)
    assert C.GetSize() == ref.nnz()
    # the sparse rule would have kept far fewer entries
    full = (a @ b).toarray()
    assert ref.nnz() == int((abs(full) > thr).sum()) > int((abs(alpha * full) > thr).sum())


@pytest.mark.parametrize("n,hb,thr,alpha", [(2048, 82, 1e-8, 1.0), (3000, 20, 1e-6, -0.7), (1531, 9, 0.0, 1.0)])
def test_tile_path_matches_scalar_path_and_oracle(nt, oracle, n, hb, thr, alpha):
    """the FP64 tensor-core tile path (DMMA) and the scalar window kernels must agree with the oracle and
    with each other on locally dense (banded) operands, including sizes that are not tile multiples"""
    a = banded(n, half_bandwidth=hb)
    A, C1, C2 = to_gpu(nt, a), nt.Matrix_ps(n), nt.Matrix_ps(n)
    nt.set_tile_path(True)
    nt.reset_counters()
    C1.Gemm(A, A, None, alpha=alpha, threshold=thr)
    assert nt.tile_counters()["tile_products"] == 1, "banded product did not take the tile path"
    nt.set_tile_path(False)
    nt.reset_counters()
    C2.Gemm(A, A, None, alpha=alpha, threshold=thr)
    assert nt.tile_counters()["tile_products"] == 0
    nt.set_tile_path(True)
    ref = oracle.multiply(oracle.PSMatrix.from_scipy(a), oracle.PSMatrix.from_scipy(a), alpha=alpha, thr=thr)
    compare_sparse(C1.to_scipy(), ref.to_scipy(), thr)
    compare_sparse(C2.to_scipy(), ref.to_scipy(), thr)
    compare_sparse(C1.to_scipy(), C2.to_scipy(), thr)


def test_tile_path_block_sparse(nt, oracle):
    from ntpoly_b200.workloads import block_sparse
    n = 2048
    a = block_sparse(n, block=32, neighbours=6, band_blocks=8, seed=3)
    A, C = to_gpu(nt, a), nt.Matrix_ps(n)
    nt.reset_counters()
    C.Gemm(A, A, None, threshold=1e-9)
    assert nt.tile_counters()["tile_products"] == 1
    ref = oracle.multiply(oracle.PSMatrix.from_scipy(a), oracle.PSMatrix.from_scipy(a), thr=1e-9)
    compare_sparse(C.to_scipy(), ref.to_scipy(), 1e-9)


def test_scattered_product_stays_on_scalar_path(nt):
    n = 4096
    a = random_sparse(n, 0.002, 77)
    A, C = to_gpu(nt, a), nt.Matrix_ps(n)
    nt.reset_counters()
    C.Gemm(A, A)
    assert nt.tile_counters()["tile_products"] == 0
    compare_sparse(C.to_scipy(), a @ a)


# ---- tile forms carried by the operands (emitted with a product's result, shared by copies) ----------------------
def test_tile_forms_chain_matches_oracle(nt, oracle):
    """C = A*A, D = C*A, E = A*C, F = C*C: the products after the first read the tile forms emitted with C
    instead of rebuilding them from CSC; results must not depend on where the forms came from."""
    n, thr = 1500, 1e-7
    a = banded(n, half_bandwidth=40)
    A, C, D, E, F = to_gpu(nt, a), nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n)
    nt.reset_counters()
    C.Gemm(A, A, None, threshold=thr)
    assert nt.tile_counters()["tile_products"] == 1
    builds_first = nt.tile_builds()
    assert builds_first == 2                      # left + right form of A
    D.Gemm(C, A, None, threshold=thr)
    E.Gemm(A, C, None, threshold=thr)
    F.Gemm(C, C, None, alpha=-0.5, threshold=thr)
    assert nt.tile_builds() == builds_first       # nothing was rebuilt from CSC
    oa = oracle.PSMatrix.from_scipy(a)
    oc = oracle.multiply(oa, oa, thr=thr)
    compare_sparse(C.to_scipy(), oc.to_scipy(), thr)
    # feed the oracle the GPU's C so that only the product under test differs
    ocg = oracle.PSMatrix.from_scipy(C.to_scipy())
    compare_sparse(D.to_scipy(), oracle.multiply(ocg, oa, thr=thr).to_scipy(), thr)
    compare_sparse(E.to_scipy(), oracle.multiply(oa, ocg, thr=thr).to_scipy(), thr)
    compare_sparse(F.to_scipy(), oracle.multiply(ocg, ocg, alpha=-0.5, thr=thr).to_scipy(), thr)
    # a copy shares the forms; a changed matrix drops them
    G = nt.Matrix_ps(C)
    H = nt.Matrix_ps(n)
    H.Gemm(G, A, None, threshold=thr)
    assert nt.tile_builds() == builds_first
    compare_sparse(H.to_scipy(), D.to_scipy(), thr)
    G.Scale(2.0)                                  # G shares C's forms: they must not be scaled under C's feet
    H.Gemm(G, A, None, threshold=thr)
    assert nt.tile_builds() == builds_first + 1
    compare_sparse(H.to_scipy(), oracle.multiply(oracle.PSMatrix.from_scipy(G.to_scipy()), oa, thr=thr).to_scipy(), thr)
    H.Gemm(C, A, None, threshold=thr)             # C is unchanged
    compare_sparse(H.to_scipy(), D.to_scipy(), thr)
    F.Scale(-3.0)                                 # sole owner: forms are scaled in place
    H.Gemm(F, A, None, threshold=thr)
    assert nt.tile_builds() == builds_first + 1
    compare_sparse(H.to_scipy(), oracle.multiply(oracle.PSMatrix.from_scipy(F.to_scipy()), oa, thr=thr).to_scipy(), thr)


@pytest.mark.parametrize("n,hb,thr,alpha,sigma", [(1500, 40, 1e-7, -1.3, 3.0), (777, 30, 0.0, 1.0, -1.0),
                                                   (2048, 12, 1e-3, 0.5, 2.5)])
def test_fused_identity_shift_is_bit_exact(nt, n, hb, thr, alpha, sigma):
    """MatrixMultiplyShift == MatrixMultiply followed by IncrementMatrix(Identity, C, sigma): same pattern, same bits."""
    a = banded(n, half_bandwidth=hb)
    A, I = to_gpu(nt, a), nt.Matrix_ps(n)
    I.FillIdentity()
    ref, fused, twice = nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n)
    ref.Gemm(A, A, None, alpha=alpha, threshold=thr)
    ref.Increment(I, sigma)
    nt.set_fused_shift(True)
    fused.GemmShift(A, A, I, sigma, None, alpha=alpha, threshold=thr)
    r, f = ref.to_scipy().tocsc(), fused.to_scipy().tocsc()
    r.sort_indices(); f.sort_indices()
    assert np.array_equal(r.indptr, f.indptr) and np.array_equal(r.indices, f.indices)
    assert np.array_equal(r.data, f.data)
    # the emitted tile forms carry the shift: use the result as an operand
    nt.reset_counters()
    twice.Gemm(fused, fused, None, threshold=thr)
    assert nt.tile_builds() == 0
    chk = nt.Matrix_ps(n)
    nt.set_fused_shift(False)
    chk.Gemm(ref, ref, None, threshold=thr)
    nt.set_fused_shift(True)
    compare_sparse(twice.to_scipy(), chk.to_scipy(), thr)


def test_fused_shift_empty_product_columns(nt):
    """diagonal entries must appear even where the product has no entry at all (window extension)"""
    n = 600
    a = banded(n, half_bandwidth=20).tolil()
    a[:, 100:164] = 0.0          # 64 empty columns -> a whole group of output columns is empty
    a[100:164, :] = 0.0
    a = sp.csc_matrix(a); a.eliminate_zeros()
    A, I = to_gpu(nt, a), nt.Matrix_ps(n)
    I.FillIdentity()
    ref, fused = nt.Matrix_ps(n), nt.Matrix_ps(n)
    ref.Gemm(A, A, None, threshold=1e-9)
    ref.Increment(I, 3.0)
    fused.GemmShift(A, A, I, 3.0, None, threshold=1e-9)
    r, f = ref.to_scipy().tocsc(), fused.to_scipy().tocsc()
    r.sort_indices(); f.sort_indices()
    assert np.array_equal(r.indptr, f.indptr) and np.array_equal(r.indices, f.indices)
    assert np.array_equal(r.data, f.data)


# ---- driver intermediates: right tile form + deferred CSC entries -------------------------------------------------
def _bits_equal(a, b):
    a, b = a.to_scipy().tocsc(), b.to_scipy().tocsc()
    a.sort_indices(); b.sort_indices()
    return (np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            and np.array_equal(a.data, b.data))


@pytest.mark.parametrize("n,hb,thr", [(2048, 40, 1e-7), (1111, 25, 0.0), (4096, 82, 1e-6)])
def test_sign_step_deferred_intermediate_is_bit_exact(nt, n, hb, thr):
    """ntb_SignStep emits T1 = 3I - a^2 X^2 as outer index + right tile form only (its CSC entries are deferred).
    The next iterate, the convergence norm and T1 itself (materialized on demand) must equal, bit for bit, what the
    separate MatrixMultiply / IncrementMatrix / MatrixNorm calls of the reference loop body give."""
    ak = 1.3
    x = banded(n, half_bandwidth=hb) * 0.3
    X, I = to_gpu(nt, x), nt.Matrix_ps(n)
    I.FillIdentity()
    T1r, X1r = nt.Matrix_ps(n), nt.Matrix_ps(n)
    T1r.GemmShift(X, X, I, 3.0, None, alpha=-ak * ak, threshold=thr)
    X1r.Gemm(X, T1r, None, alpha=0.5 * ak, threshold=thr)
    D = nt.Matrix_ps(X)
    D.Increment(X1r, -1.0)
    norm_ref = D.Norm()

    T1, X1 = nt.Matrix_ps(n), nt.Matrix_ps(n)
    nt.reset_counters()
    nv = nt.sign_step(X, I, T1, X1, ak, thr)
    dc = nt.deferred_counters()
    assert dc["products"] == 2 and dc["materialized"] == 0, dc     # T1 and the next iterate live as tile forms
    assert nt.tile_counters()["tile_products"] == 2
    assert nv == pytest.approx(norm_ref, rel=1e-12)                # norm taken from the right tile forms
    assert T1.GetSize() == T1r.GetSize() and X1.GetSize() == X1r.GetSize()   # nnz is known without the entries
    assert nt.deferred_counters()["materialized"] == 0
    assert _bits_equal(X1, X1r)                                # reading a matrix materializes its entries
    assert nt.deferred_counters()["materialized"] == 1
    assert _bits_equal(T1, T1r)
    assert nt.deferred_counters()["materialized"] == 2
    assert _bits_equal(X, to_gpu(nt, x))                       # X untouched
    assert T1.Trace() == T1r.Trace() and T1.Norm() == T1r.Norm()
    D2 = nt.Matrix_ps(X)
    D2.Increment(X1, -1.0)
    assert D2.Norm() == norm_ref                               # CSC route on the materialized entries

    # the in-place driver form: X advances, T2 is scratch
    X2, T2 = nt.Matrix_ps(X), nt.Matrix_ps(n)
    nv2 = nt.sign_iteration(X2, I, T1, T2, ak, thr)
    assert nv2 == nv and _bits_equal(X2, X1r)
    # a deferred matrix survives copy / scale / use as a LEFT operand (its left form is rebuilt from the entries)
    nt.sign_step(X, I, T1, X1, ak, thr)
    C1, C2, P, Pr = nt.Matrix_ps(T1), nt.Matrix_ps(T1r), nt.Matrix_ps(n), nt.Matrix_ps(n)
    C1.Scale(0.5); C2.Scale(0.5)
    assert _bits_equal(C1, C2)
    P.Gemm(T1, X, None, threshold=thr)
    Pr.Gemm(T1r, X, None, threshold=thr)
    assert _bits_equal(P, Pr)


def test_flop_counting_is_optional_instrumentation(nt, oracle):
    """useful-product counts are exact when switched on and never change a result"""
    n, thr = 1536, 1e-7
    a = banded(n, half_bandwidth=30)
    A, C, D, E = to_gpu(nt, a), nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n)
    C.Gemm(A, A, None, threshold=thr)            # C and A now carry tile forms
    try:
        nt.set_flop_counting(True)
        nt.reset_counters()
        D.Gemm(C, A, None, threshold=thr)
        st = oracle.MultiplyStats()
        oracle.multiply(oracle.PSMatrix.from_scipy(C.to_scipy()), oracle.PSMatrix.from_scipy(a), thr=thr, stats=st)
        assert nt.counters()["flops"] == pytest.approx(st.flops, rel=1e-12)
    finally:
        nt.set_flop_counting(False)
    nt.reset_counters()
    E.Gemm(C, A, None, threshold=thr)
    assert nt.tile_counters()["tile_products"] == 1
    assert _bits_equal(D, E)


@pytest.mark.parametrize("cplx", [False, True])
def test_scattered_columns_use_the_hash_accumulator(nt, oracle, cplx):
    """graph-like pattern (c5 in small: directed ER graph, ~25 entries per row, rows anywhere in 0..N): every output
    column has a row window as wide as the matrix but only ~600 products - served by the shared-memory hash accumulator
    (bin 7), result identical to the oracle (same k-ascending summation order per entry)"""
    from ntpoly_b200.workloads import complex_hermitian_graph
    n = 16384
    g = complex_hermitian_graph(n)
    if not cplx:
        g = sp.csc_matrix(g.real + g.imag)
    A = nt.Matrix_ps(n, is_complex=cplx)
    A.fill_from_scipy(g)
    C = nt.Matrix_ps(n)
    nt.reset_counters()
    C.Gemm(A, A, None, alpha=0.7, threshold=1e-6)
    assert nt.hash_columns() > 0.9 * n
    OA = oracle.PSMatrix.from_scipy(g, is_complex=cplx)
    ref = oracle.multiply(OA, OA, alpha=0.7, thr=1e-6)
    compare_sparse(C.to_scipy(), ref.to_scipy(), 1e-6, tol=1e-13)
    assert C.GetSize() == ref.nnz()


# ---- complex128 on the FP64 tensor cores (real embedding, spgemm.cu: spgemm_complex_tiles) ---------------------------
def _complex_banded(n, hb, seed):
    rng = np.random.default_rng(seed)
    m = banded(n, half_bandwidth=hb).astype(np.complex128)
    ph = sp.csc_matrix(m)
    ph.data = ph.data * np.exp(1j * rng.uniform(0, 2 * np.pi, ph.nnz))
    return sp.csc_matrix((ph + ph.conj().T) * 0.5)           # Hermitian band


@pytest.mark.parametrize("n,hb,thr,alpha", [(1536, 40, 1e-8, 1.0), (1001, 13, 1e-5, -0.6), (640, 100, 0.0, 1.0)])
def test_complex_tile_path_matches_scalar_path_and_oracle(nt, oracle, n, hb, thr, alpha):
    """locally dense complex operands (Hermitian band): one real DMMA tile product of the embeddings
    [[Re A, -Im A], [Im A, Re A]] x (Re B; Im B), threshold on the complex magnitude when the pairs are zipped back -
    against the oracle and against the scalar complex kernels, sizes that are not tile multiples included"""
    a = _complex_banded(n, hb, 5)
    b = _complex_banded(n, hb, 6)
    A, B, C1, C2 = to_gpu(nt, a), to_gpu(nt, b), nt.Matrix_ps(n), nt.Matrix_ps(n)
    nt.set_tile_path(True)
    nt.reset_counters()
    C1.Gemm(A, B, None, alpha=alpha, threshold=thr)
    assert nt.complex_tile_products() == 1, "the complex banded product did not take the tile path"
    nt.set_tile_path(False)
    nt.reset_counters()
    C2.Gemm(A, B, None, alpha=alpha, threshold=thr)
    assert nt.complex_tile_products() == 0
    nt.set_tile_path(True)
    ref = oracle.multiply(oracle.PSMatrix.from_scipy(a, is_complex=True), oracle.PSMatrix.from_scipy(b, is_complex=True),
                          alpha=alpha, thr=thr)
    compare_sparse(C1.to_scipy(), ref.to_scipy(), thr)
    compare_sparse(C2.to_scipy(), ref.to_scipy(), thr)
    compare_sparse(C1.to_scipy(), C2.to_scipy(), thr)


def test_complex_dense_product_takes_the_tile_path(nt, oracle):
    """the filled-in regime of the ComplexMatrix example (a dense exponential times itself): complex dense x dense"""
    n = 300
    rng = np.random.default_rng(9)
    a = sp.csc_matrix(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    A, C = to_gpu(nt, a), nt.Matrix_ps(n)
    nt.reset_counters()
    C.Gemm(A, A, None, threshold=1e-9)
    assert nt.complex_tile_products() == 1
    compare_sparse(C.to_scipy(), a @ a, 1e-9)


def test_complex_scattered_product_stays_on_scalar_kernels(nt):
    n = 4096
    a = random_sparse(n, 0.002, 78, True)
    A, C = to_gpu(nt, a), nt.Matrix_ps(n)
    nt.reset_counters()
    C.Gemm(A, A)
    assert nt.complex_tile_products() == 0
    compare_sparse(C.to_scipy(), a @ a)


@pytest.mark.parametrize("n,hb,thr", [(2048, 40, 1e-7), (1111, 25, 0.0), (3000, 82, 1e-6)])
def test_difference_norm_from_the_product_epilogue(nt, n, hb, thr):
    """||X_new - X|| of the sign iteration comes out of the epilogue of the product that computes X_new (the left
    operand's own tiles are fetched for the finished strip): same value as the separate pass over the two iterates,
    same X_new bit for bit, deterministic; sizes that are not multiples of 64 included"""
    ak = 1.3
    rng = np.random.default_rng(3)
    x = banded(n, half_bandwidth=hb) * 0.3
    # an unsymmetric perturbation, so that rows and columns cannot be confused anywhere
    x = sp.csc_matrix(x + sp.diags([rng.uniform(-0.05, 0.05, n - 3)], [3], shape=(n, n)))
    X, I = to_gpu(nt, x), nt.Matrix_ps(n)
    I.FillIdentity()
    res = {}
    for fused in (True, False):
        nt.set_fused_norm(fused)
        T1, X1 = nt.Matrix_ps(n), nt.Matrix_ps(n)
        nt.reset_counters()
        nv = nt.sign_step(X, I, T1, X1, ak, thr)
        nv_again = nt.sign_step(X, I, T1, X1, ak, thr)
        res[fused] = (nv, nv_again, nt.fused_norms(), X1)
    nt.set_fused_norm(True)
    assert res[True][2] == 2 and res[False][2] == 0
    assert res[True][0] == res[True][1]                                   # deterministic
    assert res[True][0] == pytest.approx(res[False][0], rel=1e-13)
    assert _bits_equal(res[True][3], res[False][3])
    D = nt.Matrix_ps(X)
    D.Increment(res[True][3], -1.0)
    assert res[True][0] == pytest.approx(D.Norm(), rel=1e-13)            # the reference's two calls on CSC entries
