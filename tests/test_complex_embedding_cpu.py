"""The index arithmetic of the complex tile path (csrc/spgemm.cu: k_embed_left / k_embed_right / k_zip_complex),
restated in numpy and checked against scipy on the CPU: C = A*B for complex CSC blocks equals the pairs of ONE real
product A^ * B^ with
    B^ : rows (2k, 2k+1) <- (Re, Im) of row k of B
    A^ : column 2k = (Re, Im) of column k of A at rows (2i, 2i+1); column 2k+1 = (-Im, Re)
and the real flop count is exactly 4x the complex useful products (no redundancy)."""
import numpy as np
import scipy.sparse as sp


def embed_right(b):
    b = sp.csc_matrix(b)
    outer = 2 * b.indptr
    inner = np.empty(2 * b.nnz, np.int64)
    val = np.empty(2 * b.nnz)
    inner[0::2], inner[1::2] = 2 * b.indices, 2 * b.indices + 1
    val[0::2], val[1::2] = b.data.real, b.data.imag
    return sp.csc_matrix((val, inner, outer), shape=(2 * b.shape[0], b.shape[1]))


def embed_left(a):
    a = sp.csc_matrix(a)
    n, m = a.shape
    outer = np.zeros(2 * m + 1, np.int64)
    inner = np.empty(4 * a.nnz, np.int64)
    val = np.empty(4 * a.nnz)
    for k in range(m):
        s, e = a.indptr[k], a.indptr[k + 1]
        b0, b1 = 4 * s, 2 * s + 2 * e                     # the offsets k_embed_left computes
        outer[2 * k], outer[2 * k + 1] = b0, b1
        i, v = a.indices[s:e], a.data[s:e]
        q = 2 * np.arange(e - s)
        inner[b0 + q], inner[b0 + q + 1] = 2 * i, 2 * i + 1
        val[b0 + q], val[b0 + q + 1] = v.real, v.imag
        inner[b1 + q], inner[b1 + q + 1] = 2 * i, 2 * i + 1
        val[b1 + q], val[b1 + q + 1] = -v.imag, v.real
    outer[2 * m] = 4 * a.nnz
    return sp.csc_matrix((val, inner, outer), shape=(2 * n, 2 * m))


def zip_complex(ch, alpha, thr):
    ch = sp.csc_matrix(ch)
    ch.sort_indices()
    n2, m = ch.shape
    rows, cols, vals = [], [], []
    for j in range(m):
        s, e = ch.indptr[j], ch.indptr[j + 1]
        p = s
        while p < e:                                      # an entry opens a complex entry; its partner follows directly
            r = ch.indices[p]
            re = im = 0.0
            if r & 1:
                im = ch.data[p]; p += 1
            else:
                re = ch.data[p]; p += 1
                if p < e and ch.indices[p] == r + 1:
                    im = ch.data[p]; p += 1
            v = alpha * complex(re, im)
            if abs(v) > thr:
                rows.append(r >> 1); cols.append(j); vals.append(v)
    return sp.csc_matrix((vals, (rows, cols)), shape=(n2 // 2, m), dtype=np.complex128)


def test_one_real_product_of_the_embeddings_is_the_complex_product():
    rng = np.random.default_rng(7)
    n, k, m = 37, 29, 41
    a = sp.random(n, k, 0.2, random_state=rng) + 1j * sp.random(n, k, 0.2, random_state=rng)
    b = sp.random(k, m, 0.2, random_state=rng) + 1j * sp.random(k, m, 0.2, random_state=rng)
    ah, bh = embed_left(a), embed_right(b)
    assert ah.has_sorted_indices or (ah.sort_indices() is None)
    ch = ah @ bh
    alpha, thr = -0.7, 1e-3
    got = zip_complex(ch, alpha, thr)
    full = (alpha * (sp.csc_matrix(a) @ sp.csc_matrix(b))).toarray()
    want = np.where(abs(full) > thr, full, 0.0)
    assert np.allclose(got.toarray(), want, rtol=1e-14, atol=1e-15)
    # flops: real useful products of A^ * B^ = 4 x the complex useful products - the 8 real flops per complex
    # multiply-add that the arithmetic needs, nothing more
    acsc, bcsr = sp.csc_matrix(a), sp.csr_matrix(b)
    complex_products = int(sum((acsc.indptr[kk + 1] - acsc.indptr[kk]) * (bcsr.indptr[kk + 1] - bcsr.indptr[kk]) for kk in range(k)))
    ahc, bhr = sp.csc_matrix(ah), sp.csr_matrix(bh)
    real_products = int(sum((ahc.indptr[kk + 1] - ahc.indptr[kk]) * (bhr.indptr[kk + 1] - bhr.indptr[kk]) for kk in range(2 * k)))
    assert real_products == 4 * complex_products


def test_a_complex_4x2_block_fills_one_real_8x4_tile():
    a = sp.csc_matrix(np.arange(1, 9).reshape(4, 2) * (1 + 2j))
    ah = embed_left(a).toarray()
    assert ah.shape == (8, 4) and np.count_nonzero(ah) == 32
