"""Deterministic synthetic inputs of the named configurations.

Host-side generators only (numpy/scipy): they produce the matrices that are then
handed to the CUDA path through the C ABI (the tests feed the same matrices to their CPU checker).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def banded(n: int, half_bandwidth: int = 82, seed: int = 20240617, lam: float = 20.0,
           scale: float = 0.05) -> sp.csc_matrix:
    """c1 / c4: symmetric band, a_ij = s*exp(-|i-j|/lam)/(1+|i-j|), random diagonal in [-1,1].
    half_bandwidth 82 -> 165 nnz per row (~2 % fill at n=8192)."""
    rng = np.random.default_rng(seed)
    diags, offs = [rng.uniform(-1.0, 1.0, n)], [0]
    for d in range(1, half_bandwidth + 1):
        v = scale * np.exp(-d / lam) / (1.0 + d)
        diags += [np.full(n - d, v), np.full(n - d, v)]
        offs += [d, -d]
    return sp.diags(diags, offs, shape=(n, n), format="csc", dtype=np.float64)


def banded_sign_input(n: int, **kw) -> sp.csc_matrix:
    """c4: the band shifted so that the spectrum straddles zero (sign function input)."""
    m = banded(n, **kw)
    d = m.diagonal()
    return sp.csc_matrix(m - sp.identity(n, format="csc") * float(np.median(d)))


def block_sparse(n: int = 65536, block: int = 32, neighbours: int = 20, band_blocks: int = 64,
                 seed: int = 1234) -> sp.csc_matrix:
    """c3: symmetric block-sparse Hamiltonian, dense block x block tiles, ~neighbours tiles per
    block row inside a block band, values N(0,1)*exp(-dist/8), Gershgorin-scaled into [-1,1]."""
    rng = np.random.default_rng(seed)
    nb = n // block
    bi, bj = [], []
    half = max(1, neighbours // 2)
    for i in range(nb):
        lo, hi = max(0, i - band_blocks), min(nb - 1, i + band_blocks)
        cand = np.arange(i + 1, hi + 1)
        take = min(half, len(cand))
        if take:
            sel = rng.choice(cand, size=take, replace=False)
            bi += [i] * take
            bj += list(sel)
    bi = np.asarray(bi, np.int64)
    bj = np.asarray(bj, np.int64)
    nblk = len(bi)
    vals = rng.standard_normal((nblk, block, block)) * np.exp(-np.abs(bi - bj) / 8.0)[:, None, None]
    ii = (bi[:, None, None] * block + np.arange(block)[None, :, None]).repeat(block, axis=2)
    jj = (bj[:, None, None] * block + np.arange(block)[None, None, :]).repeat(block, axis=1)
    upper = sp.coo_matrix((vals.ravel(), (ii.ravel(), jj.ravel())), shape=(n, n))
    dvals = rng.standard_normal((nb, block, block))
    dvals = 0.5 * (dvals + dvals.transpose(0, 2, 1))
    di = (np.arange(nb)[:, None, None] * block + np.arange(block)[None, :, None]).repeat(block, axis=2)
    dj = (np.arange(nb)[:, None, None] * block + np.arange(block)[None, None, :]).repeat(block, axis=1)
    diag = sp.coo_matrix((dvals.ravel(), (di.ravel(), dj.ravel())), shape=(n, n))
    m = (upper + upper.T + diag).tocsc()
    radius = float(np.asarray(abs(m).sum(axis=0)).max())
    m = m * (1.0 / radius)
    m.sort_indices()
    return sp.csc_matrix(m)


def block_sparse_hamiltonian(n: int = 65536, block: int = 32, neighbours: int = 20, band_blocks: int = 64,
                             seed: int = 1234, coupling: float = 0.35, decay: float = 8.0) -> sp.csc_matrix:
    """c3: block-sparse LINEAR-SCALING Hamiltonian (an insulator): dense block x block tiles, ~neighbours tiles per
    block row inside a block band (1 % block fill at n = 65536), Gaussian couplings that decay with the block
    distance, and on-site energies -1 / +1 for the first / second half of every block's orbitals, so that the
    spectrum has a gap of ~1.4 at half filling (n/2 electrons) and the density matrix decays: ~1800 entries per row
    above 1e-6 against ~640 of H (measured at n = 4096). `block_sparse` above has no gap - its density matrix is
    dense, purification of it is not a linear-scaling workload - and stays for the pattern tests."""
    rng = np.random.default_rng(seed)
    nb = n // block
    bi, bj = [], []
    half = max(1, neighbours // 2)
    for i in range(nb):
        hi = min(nb - 1, i + band_blocks)
        cand = np.arange(i + 1, hi + 1)
        take = min(half, len(cand))
        if take:
            sel = rng.choice(cand, size=take, replace=False)
            bi += [i] * take
            bj += list(sel)
    bi = np.asarray(bi, np.int64)
    bj = np.asarray(bj, np.int64)
    nblk = len(bi)
    vals = rng.standard_normal((nblk, block, block)) * (coupling / np.sqrt(block * neighbours))
    vals *= np.exp(-np.abs(bi - bj) / decay)[:, None, None]
    ii = (bi[:, None, None] * block + np.arange(block)[None, :, None]).repeat(block, axis=2)
    jj = (bj[:, None, None] * block + np.arange(block)[None, None, :]).repeat(block, axis=1)
    upper = sp.coo_matrix((vals.ravel(), (ii.ravel(), jj.ravel())), shape=(n, n))
    dvals = rng.standard_normal((nb, block, block)) * (0.5 * coupling / np.sqrt(block))
    dvals = 0.5 * (dvals + dvals.transpose(0, 2, 1))
    dvals += np.diag(np.where(np.arange(block) < block // 2, -1.0, 1.0))[None, :, :]
    di = (np.arange(nb)[:, None, None] * block + np.arange(block)[None, :, None]).repeat(block, axis=2)
    dj = (np.arange(nb)[:, None, None] * block + np.arange(block)[None, None, :]).repeat(block, axis=1)
    diag = sp.coo_matrix((dvals.ravel(), (di.ravel(), dj.ravel())), shape=(n, n))
    m = (upper + upper.T + diag).tocsc()
    m.sort_indices()
    return sp.csc_matrix(m)


def guo_transform(a: sp.spmatrix) -> sp.csc_matrix:
    """Hermitian matrix of a directed graph as built by the reference example
    (Examples/ComplexMatrix/main.f90:109-160): S = symmetrised adjacency pattern,
    G = S - A (the missing reverse edges), H = S + i*G - i*G^T."""
    a = sp.csr_matrix(a)
    a = sp.csr_matrix((np.ones(a.nnz), a.indices, a.indptr), shape=a.shape)
    s = a.maximum(a.T)
    g = sp.csr_matrix(s - a)
    g.eliminate_zeros()
    h = s.astype(np.complex128) + 1j * g - 1j * g.T
    h = sp.csc_matrix(h)
    h.sort_indices()
    return h


def complex_hermitian_graph(n: int = 32768, avg_degree: float = 25.0, seed: int = 99) -> sp.csc_matrix:
    """c5: Guo-transformed directed Erdos-Renyi graph with ~avg_degree nnz per row (the shipped
    512-node example has 24.8)."""
    rng = np.random.default_rng(seed)
    m_edges = int(n * avg_degree / 2)
    src = rng.integers(0, n, m_edges)
    dst = rng.integers(0, n, m_edges)
    keep = src != dst
    a = sp.coo_matrix((np.ones(int(keep.sum())), (src[keep], dst[keep])), shape=(n, n))
    return guo_transform(a)


def useful_flops(a: sp.spmatrix, b: sp.spmatrix) -> float:
    """F = 2 * sum_{(i,k) in pattern(A)} nnz(B(k,:)) (x4 for complex128)  — SURVEY 8(d)."""
    brow = np.diff(sp.csr_matrix(b).indptr)
    f = 2.0 * float(brow[sp.coo_matrix(a).col].sum())
    if np.iscomplexobj(a.data) or np.iscomplexobj(b.data):
        f *= 4.0
    return f
