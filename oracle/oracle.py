"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the shipped package).

numpy/ctypes host layer of the CPU restatement of NTPoly's distributed sparse
multiply and of the solver drivers that iterate on it.  Arithmetic with the
reference's threshold quirks lives in C (oracle/ntpoly_oracle.c); this file
restates the *distributed* structure (process grid, padding, local blocks,
slices) and the drivers.  Every function cites the reference lines it follows
(paths relative to /root/reference).

The distributed matrix is simulated in ONE process: it is held as a global
padded CSC matrix plus the process grid; everything the reference does "per
rank, per local block" is done here per global block, which is the same set of
block operations (Source/Fortran/distributed_algebra_includes/MatrixMultiply.f90).

Parity status: the Fortran reference cannot be built in this image, so the
restatement is pinned by the reference's shipped golden vector
(Examples/PremadeMatrix) and by SciPy in the way the reference's own unit tests
are; threshold>0 / alpha,beta / slices>1 rules are "parity unpinned" by any
reference test and follow the cited source lines only.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/ntpoly_oracle.c (gcc -O3 -fopenmp)."""
    out = os.path.join(_HERE, "_build", "libntpoly_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ntpoly_oracle.c", "local_kernels.inc.h")]
    if force or not os.path.exists(out) or any(
            os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return out


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libntpoly_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_gemm_r.restype = ctypes.c_long
        _LIB.orc_gemm_c.restype = ctypes.c_long
    return _LIB


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _csc(m, dtype):
    m = sp.csc_matrix(m, dtype=dtype)
    if not m.has_sorted_indices:
        m.sort_indices()
    return m


def _arrays(m, dtype):
    return (np.ascontiguousarray(m.indptr, dtype=np.int32),
            np.ascontiguousarray(m.indices, dtype=np.int32),
            np.ascontiguousarray(m.data, dtype=dtype))


# --------------------------------------------------------------------------
# local kernels (thin ctypes wrappers)
# --------------------------------------------------------------------------
def local_gemm(AT, BT, alpha=1.0, thr=0.0, is_complex=False, pool_mode=0,
               force_branch=0):
    """C = alpha*op(A)*op(B); AT/BT are CSC of op(A)^T / op(B)^T.
    Returns (C as CSC [AT.cols x BT.rows], branch) with branch 1=sparse 2=dense.
    (SMatrixAlgebraModule.F90:221-287, sparse_includes/GemmMatrix.f90)"""
    dt = np.complex128 if is_complex else np.float64
    AT = _csc(AT, dt)
    BT = _csc(BT, dt)
    assert AT.shape[0] == BT.shape[1]
    ao, ai, av = _arrays(AT, dt)
    bo, bi, bv = _arrays(BT, dt)
    co = ctypes.POINTER(ctypes.c_int)()
    ci = ctypes.POINTER(ctypes.c_int)()
    cv = ctypes.c_void_p()
    br = ctypes.c_int(0)
    fn = lib().orc_gemm_c if is_complex else lib().orc_gemm_r
    nnz = fn(ctypes.c_int(AT.shape[0]), ctypes.c_int(AT.shape[1]), _ip(ao), _ip(ai), _vp(av),
             ctypes.c_int(BT.shape[0]), ctypes.c_int(BT.shape[1]), _ip(bo), _ip(bi), _vp(bv),
             ctypes.c_double(alpha), ctypes.c_double(thr), ctypes.c_int(pool_mode),
             ctypes.c_int(force_branch), ctypes.byref(co), ctypes.byref(ci),
             ctypes.byref(cv), ctypes.byref(br))
    c_rows, c_cols = AT.shape[1], BT.shape[0]
    outer = np.ctypeslib.as_array(co, shape=(c_cols + 1,)).copy()
    if nnz > 0:
        inner = np.ctypeslib.as_array(ci, shape=(nnz,)).copy()
        vals = np.ctypeslib.as_array(
            ctypes.cast(cv, ctypes.POINTER(ctypes.c_double)),
            shape=(nnz * (2 if is_complex else 1),)).copy()
        if is_complex:
            vals = vals.view(np.complex128)
    else:
        inner = np.zeros(0, np.int32)
        vals = np.zeros(0, dt)
    lib().orc_free(co)
    lib().orc_free(ci)
    lib().orc_free(cv)
    return sp.csc_matrix((vals, inner, outer), shape=(c_rows, c_cols)), br.value


def local_increment(A, B, alpha=1.0, thr=0.0, is_complex=False):
    """B <- alpha*A + B column-wise with the reference's tail rule
    (sparse_includes/IncrementMatrix.f90, AddSparseVectors.f90)."""
    dt = np.complex128 if is_complex else np.float64
    A = _csc(A, dt)
    B = _csc(B, dt)
    ao, ai, av = _arrays(A, dt)
    bo, bi, bv = _arrays(B, dt)
    cap = max(1, len(ai) + len(bi))
    oc = np.zeros(A.shape[1] + 1, np.int32)
    ic = np.zeros(cap, np.int32)
    vc = np.zeros(cap, dt)
    fn = lib().orc_increment_c if is_complex else lib().orc_increment_r
    n = fn(ctypes.c_int(A.shape[1]), _ip(ao), _ip(ai), _vp(av), _ip(bo), _ip(bi), _vp(bv),
           _ip(oc), _ip(ic), _vp(vc), ctypes.c_double(alpha), ctypes.c_double(thr))
    return sp.csc_matrix((vc[:n].copy(), ic[:n].copy(), oc), shape=A.shape)


def local_pairwise(A, B, is_complex=False):
    dt = np.complex128 if is_complex else np.float64
    A = _csc(A, dt)
    B = _csc(B, dt)
    ao, ai, av = _arrays(A, dt)
    bo, bi, bv = _arrays(B, dt)
    cap = max(1, min(len(ai), len(bi)))
    oc = np.zeros(A.shape[1] + 1, np.int32)
    ic = np.zeros(cap, np.int32)
    vc = np.zeros(cap, dt)
    fn = lib().orc_pairwise_c if is_complex else lib().orc_pairwise_r
    n = fn(ctypes.c_int(A.shape[1]), _ip(ao), _ip(ai), _vp(av), _ip(bo), _ip(bi), _vp(bv),
           _ip(oc), _ip(ic), _vp(vc))
    return sp.csc_matrix((vc[:n].copy(), ic[:n].copy(), oc), shape=A.shape)


# --------------------------------------------------------------------------
# process grid (ProcessGridModule.F90:130-264, 576-638)
# --------------------------------------------------------------------------
@dataclass
class Grid:
    rows: int = 1
    cols: int = 1
    slices: int = 1
    threads: int = 1  # omp thread count the block multiplier is derived from

    def __post_init__(self):
        R, C, S = self.rows, self.cols, self.slices
        if S > 1 and max(R, C) % min(R, C) != 0:
            raise ValueError("if slices >1, either rows or columns must be a multiple of the other")
        cbm = (R // C) * S
        if cbm == 0:
            cbm = S
        rbm = (C // R) * S
        if rbm == 0:
            rbm = S
        bm = self.threads // (cbm + rbm)
        if bm == 0:
            bm = 1
        self.block_multiplier = bm
        self.nbc = cbm * bm  # number_of_blocks_columns (per rank)
        self.nbr = rbm * bm  # number_of_blocks_rows (per rank)

    @property
    def size(self):
        return self.rows * self.cols * self.slices

    def coords(self, rank):
        """rank -> (slice,row,col)  (ProcessGridModule.F90:180-183)"""
        ss = self.rows * self.cols
        return rank // ss, (rank % ss) // self.cols, rank % self.cols

    def padded(self, n):
        """CalculateScaledDimension (PSMatrixModule.F90:1596-1618)"""
        lcm = self.block_multiplier * self.slices * self.cols * self.rows
        q = n // lcm
        return n if q * lcm == n else (q + 1) * lcm

    @staticmethod
    def default_for(nprocs):
        """ComputeNumSlices + ComputeGridSize (ProcessGridModule.F90:576-638)"""
        slices = 1
        for s in range(min(4, nprocs), 1, -1):
            ss = nprocs // s
            if ss * s != nprocs:
                continue
            d = int(math.floor(math.sqrt(ss)))
            if d * d == ss:
                slices = s
                break
            d = int(math.floor(math.sqrt(ss // 2)))
            if d * d * 2 == ss:
                slices = s
                break
        return Grid(*Grid.grid_size(nprocs, slices), slices)

    @staticmethod
    def grid_size(nprocs, slices):
        ss = nprocs // slices
        rows, cols = 1, 1
        for ii in range(int(math.floor(math.sqrt(ss))), 0, -1):
            if ss % ii == 0:
                rows, cols = ii, ss // ii
                break
        return rows, cols


# --------------------------------------------------------------------------
# distributed matrix (PSMatrixModule.F90:33-51, 190-252)
# --------------------------------------------------------------------------
class PSMatrix:
    def __init__(self, n, grid=None, is_complex=False, mat=None):
        self.grid = grid or Grid()
        self.n = int(n)
        self.N = self.grid.padded(self.n)
        self.is_complex = bool(is_complex)
        dt = np.complex128 if self.is_complex else np.float64
        if mat is None:
            self.mat = sp.csc_matrix((self.N, self.N), dtype=dt)
        else:
            self.mat = _csc(mat, dt)
            assert self.mat.shape == (self.N, self.N)

    # -- construction ------------------------------------------------------
    @staticmethod
    def from_scipy(m, grid=None, is_complex=None):
        m = sp.coo_matrix(m)
        if is_complex is None:
            is_complex = np.iscomplexobj(m.data)
        out = PSMatrix(m.shape[0], grid, is_complex)
        dt = np.complex128 if is_complex else np.float64
        out.mat = _csc(sp.coo_matrix((m.data.astype(dt), (m.row, m.col)),
                                     shape=(out.N, out.N)), dt)
        return out

    def like(self):
        return PSMatrix(self.n, self.grid, self.is_complex)

    def copy(self):
        return PSMatrix(self.n, self.grid, self.is_complex, self.mat.copy())

    def to_scipy(self):
        return self.mat[: self.n, : self.n].copy()

    def todense(self):
        return np.asarray(self.to_scipy().todense())

    @property
    def dtype(self):
        return np.complex128 if self.is_complex else np.float64

    def nnz(self):
        return int(self.mat.nnz)

    def to_complex(self):
        return PSMatrix(self.n, self.grid, True, self.mat.astype(np.complex128))

    # -- blocks ------------------------------------------------------------
    def row_block(self):
        return self.N // (self.grid.rows * self.grid.nbr)

    def col_block(self):
        return self.N // (self.grid.cols * self.grid.nbc)

    def local_triplets(self, rank):
        """GetMatrixTripletList for one rank: global 1-based (row, col, value),
        column-major sorted (distributed_includes/GetMatrixTripletList.f90)."""
        _, r, c = self.grid.coords(rank)
        lr, lc = self.N // self.grid.rows, self.N // self.grid.cols
        blk = sp.coo_matrix(self.mat[r * lr:(r + 1) * lr, c * lc:(c + 1) * lc].tocsc())
        order = np.lexsort((blk.row, blk.col))
        return (blk.row[order] + r * lr + 1, blk.col[order] + c * lc + 1, blk.data[order])


def identity(like: PSMatrix) -> PSMatrix:
    """FillMatrixIdentity: ones only for indices <= actual dimension
    (distributed_includes/FillMatrixIdentity.f90:9-22)."""
    out = like.like()
    d = np.zeros(out.N)
    d[: out.n] = 1.0
    idx = np.arange(out.n)
    out.mat = _csc(sp.coo_matrix((np.ones(out.n, out.dtype), (idx, idx)),
                                 shape=(out.N, out.N)), out.dtype)
    return out


def transpose(A: PSMatrix) -> PSMatrix:
    """TransposeMatrix_ps (distributed_includes/TransposeMatrix.f90)."""
    return PSMatrix(A.n, A.grid, A.is_complex, A.mat.T.tocsc())


def conjugate(A: PSMatrix) -> PSMatrix:
    return PSMatrix(A.n, A.grid, A.is_complex, A.mat.conj())


def scale(A: PSMatrix, c) -> PSMatrix:
    """ScaleMatrix_psr/psc (sparse_includes/ScaleMatrix.f90:1)."""
    if isinstance(c, complex) and not A.is_complex:
        A = A.to_complex()
    m = A.mat.copy()
    m.data = m.data * c
    return PSMatrix(A.n, A.grid, A.is_complex, m)


def _blocks(M: PSMatrix):
    rb, cb = M.row_block(), M.col_block()
    for J in range(M.N // cb):
        for I in range(M.N // rb):
            yield I, J, rb, cb


def _blockwise(A: PSMatrix, B: PSMatrix, fn):
    """apply fn(blockA, blockB) per local block and reassemble"""
    rb, cb = A.row_block(), A.col_block()
    nI, nJ = A.N // rb, A.N // cb
    if nI == 1 and nJ == 1:
        return fn(A.mat, B.mat)
    grid = [[None] * nJ for _ in range(nI)]
    for I in range(nI):
        for J in range(nJ):
            grid[I][J] = fn(A.mat[I * rb:(I + 1) * rb, J * cb:(J + 1) * cb].tocsc(),
                            B.mat[I * rb:(I + 1) * rb, J * cb:(J + 1) * cb].tocsc())
    return sp.bmat(grid, format="csc")


def increment(A: PSMatrix, B: PSMatrix, alpha=1.0, thr=0.0) -> PSMatrix:
    """B <- alpha*A + B  (PSMatrixAlgebraModule.F90:414-460; per local block:
    distributed_algebra_includes/IncrementMatrix.f90). Returns the new B."""
    cplx = A.is_complex or B.is_complex
    if cplx and not A.is_complex:
        A = A.to_complex()
    if cplx and not B.is_complex:
        B = B.to_complex()
    m = _blockwise(A, B, lambda a, b: local_increment(a, b, alpha, thr, cplx))
    return PSMatrix(B.n, B.grid, cplx, m)


def pairwise(A: PSMatrix, B: PSMatrix) -> PSMatrix:
    cplx = A.is_complex or B.is_complex
    if cplx and not A.is_complex:
        A = A.to_complex()
    if cplx and not B.is_complex:
        B = B.to_complex()
    m = _blockwise(A, B, lambda a, b: local_pairwise(a, b, cplx))
    return PSMatrix(A.n, A.grid, cplx, m)


def dot(A: PSMatrix, B: PSMatrix):
    """DotMatrix: sum conj(a_ij) b_ij (distributed_algebra_includes/DotMatrix.f90).
    Returns a complex when either is complex, else a float."""
    if A.is_complex:
        A = conjugate(A)
    C = pairwise(A, B)
    s = C.mat.data.sum()
    return complex(s) if C.is_complex else float(s)


def dot_real(A, B):
    return float(np.real(dot(A, B)))


def trace(A: PSMatrix) -> float:
    """MatrixTrace_psr: sum of Re(a_ii) (distributed_algebra_includes/MatrixTrace.f90)."""
    return float(np.real(A.mat.diagonal()).sum())


def norm(A: PSMatrix) -> float:
    """MatrixNorm_ps: max column sum of |a_ij| (distributed_algebra_includes/MatrixNorm.f90)."""
    if A.mat.nnz == 0:
        return 0.0
    return float(np.asarray(abs(A.mat).sum(axis=0)).max())


def sigma(A: PSMatrix) -> float:
    """MatrixSigma: 1/norm^2 (distributed_algebra_includes/MatrixSigma.f90:15-17)."""
    return 1.0 / (norm(A) ** 2)


def gershgorin(A: PSMatrix):
    """GershgorinBounds (solver_includes/GershgorinBounds.f90): column-wise,
    over all logical (padded) columns. Returns (e_min, e_max)."""
    d = np.real(A.mat.diagonal())
    absum = np.asarray(abs(A.mat).sum(axis=0)).ravel()
    off = absum - np.abs(A.mat.diagonal())
    return float((d - off).min()), float((d + off).max())


def filter_matrix(A: PSMatrix, thr: float) -> PSMatrix:
    """FilterMatrix_ps: keep |v| > thr (distributed_includes/FilterMatrix.f90)."""
    m = sp.coo_matrix(A.mat)
    keep = np.abs(m.data) > thr
    return PSMatrix(A.n, A.grid, A.is_complex,
                    sp.coo_matrix((m.data[keep], (m.row[keep], m.col[keep])), shape=m.shape))


def is_identity(A: PSMatrix) -> bool:
    """IsIdentity (distributed_includes/IsIdentity.f90)."""
    m = sp.coo_matrix(A.mat)
    if np.any(m.row != m.col):
        return False
    tiny = np.finfo(np.float64).tiny
    ok = np.abs(m.data - 1.0) <= tiny
    return bool(ok.all() and ok.sum() == A.n)


def permutation_matrices(like: PSMatrix, perm):
    """FillMatrixPermutation rows/columns (distributed_includes/FillMatrixPermutation.f90).
    perm is the 0-based index_lookup over the LOGICAL dimension."""
    N = like.N
    perm = np.asarray(perm)
    ones = np.ones(N, like.dtype)
    rows = sp.coo_matrix((ones, (np.arange(N), perm)), shape=(N, N))
    cols = sp.coo_matrix((ones, (perm, np.arange(N))), shape=(N, N))
    return (PSMatrix(like.n, like.grid, like.is_complex, rows),
            PSMatrix(like.n, like.grid, like.is_complex, cols))


# --------------------------------------------------------------------------
# the distributed multiply
# --------------------------------------------------------------------------
@dataclass
class MultiplyStats:
    flops: float = 0.0           # useful flops = 2*sum_{(i,k) in A} nnz(B(k,:))  (x4 complex)
    branches: list = field(default_factory=list)


TOTAL_STATS = None          # set to a MultiplyStats to count every multiply (bench.py, cpu_baseline of whole solves)


def useful_flops(A: PSMatrix, B: PSMatrix) -> float:
    """F = 2 * sum_{(i,k) in pattern(A)} nnz(B(k,:)); x4 if complex."""
    brow = np.diff(sp.csr_matrix(B.mat).indptr)
    f = 2.0 * float(brow[sp.coo_matrix(A.mat).col].sum())
    return f * (4.0 if (A.is_complex or B.is_complex) else 1.0)


def multiply(A: PSMatrix, B: PSMatrix, C: PSMatrix | None = None, alpha=1.0, beta=0.0,
             thr=0.0, stats: MultiplyStats | None = None, pool_mode=0) -> PSMatrix:
    """C := alpha*A*B + beta*C with drop threshold
    (PSMatrixAlgebraModule.F90:108-211; distributed_algebra_includes/MatrixMultiply.f90)."""
    g = A.grid
    cplx = A.is_complex or B.is_complex
    if cplx and not A.is_complex:
        A = A.to_complex()          # PSMatrixAlgebraModule.F90:171-188
    if cplx and not B.is_complex:
        B = B.to_complex()
    S = g.slices
    wthr = thr / (S * 1000) if S > 1 else thr          # MatrixMultiply.f90:25-29
    rb, cb = A.row_block(), A.col_block()
    nI, nJ = A.N // rb, A.N // cb
    AT = A.mat.T.tocsc()      # column i = row i of A   (GatheredRowContributionT)
    BT = B.mat.T.tocsc()      # column k = row k of B   (GatheredColumnContribution)
    AT.sort_indices()
    BT.sort_indices()
    if S > 1:
        kb = rb                # inner block size (== cb when S>1)
        nK = A.N // kb
        ksel = [np.concatenate([np.arange(gk * kb, (gk + 1) * kb)
                                for gk in range(nK) if gk % S == s]) for s in range(S)]
    grid = [[None] * nJ for _ in range(nI)]
    for I in range(nI):
        for J in range(nJ):
            if S == 1:
                blk, br = local_gemm(AT[:, I * rb:(I + 1) * rb], BT[J * cb:(J + 1) * cb, :],
                                     alpha, wthr, cplx, pool_mode)
                if stats is not None:
                    stats.branches.append(br)
            else:
                acc = sp.csc_matrix((rb, cb), dtype=A.dtype)
                for s in range(S):                     # MatrixMultiply.f90:75-80,98-105,158-164
                    contrib, br = local_gemm(AT[ksel[s], I * rb:(I + 1) * rb],
                                             BT[J * cb:(J + 1) * cb, :][:, ksel[s]],
                                             alpha, wthr, cplx, pool_mode)
                    if stats is not None:
                        stats.branches.append(br)
                    # ReduceAndSumMatrixCleanup.f90:11-32: last add carries the threshold
                    acc = local_increment(contrib, acc, 1.0, thr if s == S - 1 else 0.0, cplx)
                blk = acc
            grid[I][J] = blk
    AB = PSMatrix(A.n, g, cplx, grid[0][0] if (nI == 1 and nJ == 1) else sp.bmat(grid, format="csc"))
    if stats is not None:
        stats.flops += useful_flops(A, B)
    if TOTAL_STATS is not None:          # bench.py: the useful flops of a whole driver (its multiplies take no stats)
        TOTAL_STATS.flops += useful_flops(A, B)
    if C is None or abs(beta) < np.finfo(np.float64).tiny:     # MatrixMultiply.f90:324-329
        return AB
    return increment(AB, scale(C, beta))


def similarity_transform(A, P, PInv, thr=0.0):
    """SimilarityTransform (PSMatrixAlgebraModule.F90:603-654)."""
    if is_identity(P):
        return A.copy()
    T = multiply(P, A, thr=thr)
    return multiply(T, PInv, thr=thr)


def permute(M, perm):
    """PermuteMatrix (LoadBalancerModule.F90:16-52)."""
    PR, PC = permutation_matrices(M, perm)
    return multiply(multiply(PR, M), PC)


def undo_permute(M, perm):
    """UndoPermuteMatrix (LoadBalancerModule.F90:55-92)."""
    PR, PC = permutation_matrices(M, perm)
    return multiply(multiply(PC, M), PR)


# --------------------------------------------------------------------------
# solver infrastructure
# --------------------------------------------------------------------------
class Monitor:
    """ConvergenceMonitorModule.F90:35-191."""

    def __init__(self, automatic=True, tight=1e-8, loose=1e-2, short_len=3, long_len=6):
        self.short = [0.0] * short_len
        self.long = [0.0] * long_len
        self.loose, self.tight, self.automatic = loose, tight, automatic
        self.nval = 0

    def append(self, v):
        self.short = self.short[1:] + [v]
        self.long = self.long[1:] + [v]
        self.nval += 1

    def converged(self):
        last, last2 = self.short[-1], self.short[-2]
        conv = not (abs(last) > self.tight)
        if not self.automatic or conv:
            return conv
        conv = True
        if self.nval < len(self.long):
            conv = False
        avs = sum(self.short) / len(self.short)
        avl = sum(self.long) / len(self.long)
        if not (10 * avs > avl and avs / 10 < avl):
            conv = False
        if not (10 * last > avl and last / 10 < avl):
            conv = False
        if last < 0:
            conv = False
        if abs(last) < abs(last2):
            conv = False
        if avl > self.loose:
            conv = False
        return conv


@dataclass
class SolverParameters:
    """SolverParametersModule.F90:14-33, 48-112 (defaults)."""
    converge_diff: float = 1e-6
    max_iterations: int = 1000
    threshold: float = 0.0
    be_verbose: bool = False
    permutation: np.ndarray | None = None   # 0-based index_lookup over the logical dim
    step_thresh: float = 1e-2
    monitor_convergence: bool = True

    @property
    def do_load_balancing(self):
        return self.permutation is not None

    def monitor(self):
        return Monitor(self.monitor_convergence, self.converge_diff)


@dataclass
class SolveInfo:
    iterations: int = 0          # value of the loop counter II at exit
    energy: float = 0.0
    chemical_potential: float = 0.0
    history: list = field(default_factory=list)
    sigmas: list = field(default_factory=list)
    flops: float = 0.0


def _loop_counter(ii, maxit, broke):
    """Fortran DO II=1,max ... EXIT leaves II=max+1 when the loop runs out."""
    return ii if broke else maxit + 1


# --------------------------------------------------------------------------
# density matrix solvers (DensityMatrixSolversModule.F90)
# --------------------------------------------------------------------------
def _density_setup(H, ISQ, p):
    I = identity(H)
    ISQT = transpose(ISQ)
    WH = similarity_transform(H, ISQ, ISQT, p.threshold)
    if p.do_load_balancing:
        WH = permute(WH, p.permutation)
        I = permute(I, p.permutation)
    return I, ISQT, WH


def _density_finish(X, ISQT, ISQ, p):
    if p.do_load_balancing:
        X = undo_permute(X, p.permutation)
    return similarity_transform(X, ISQT, ISQ, p.threshold)


def trs2(H, ISQ, trace_target, p: SolverParameters | None = None):
    """TRS2 (DensityMatrixSolversModule.F90:285-481). Returns (K, SolveInfo)."""
    p = p or SolverParameters()
    mon = p.monitor()
    info = SolveInfo()
    I, ISQT, WH = _density_setup(H, ISQ, p)
    e_min, e_max = gershgorin(WH)
    X = scale(WH, -1.0)
    X = increment(I, X, alpha=e_max)
    X = scale(X, 1.0 / (e_max - e_min))
    energy = 0.0
    broke = False
    ii = 0
    for ii in range(1, p.max_iterations + 1):
        tv = trace(X)
        sg = -1.0 if trace_target - tv < 0.0 else 1.0
        info.sigmas.append(sg)
        X2 = multiply(X, X, thr=p.threshold)
        if sg > 0.0:
            X = scale(X, 2.0)
            X = increment(X2, X, alpha=-1.0, thr=p.threshold)
        else:
            X = X2.copy()
        old = energy
        energy = dot_real(X, WH)
        mon.append(energy - old)
        info.history.append(energy - old)
        if mon.converged():
            broke = True
            break
    info.iterations = _loop_counter(ii, p.max_iterations, broke)
    total = info.iterations - 1
    info.energy = energy
    K = _density_finish(X, ISQT, ISQ, p)
    a, b, mid = 0.0, 1.0, 0.0
    for _ in range(p.max_iterations):
        mid = (b - a) / 2.0 + a
        z = mid
        for jj in range(total):
            z = z * z if info.sigmas[jj] < 0.0 else 2.0 * z - z * z
        if z < 0.5:
            a = mid
        else:
            b = mid
        if abs(z - 0.5) < p.converge_diff:
            break
    info.chemical_potential = e_max + (e_min - e_max) * mid
    return K, info


def scale_and_fold(H, ISQ, trace_target, homo, lumo, p: SolverParameters | None = None):
    """Scale and Fold (DensityMatrixSolversModule.F90:953-1117). Returns (K, SolveInfo)."""
    p = p or SolverParameters()
    mon = p.monitor()
    info = SolveInfo()
    I, ISQT, WH = _density_setup(H, ISQ, p)
    e_min, e_max = gershgorin(WH)
    X = scale(WH, -1.0)
    X = increment(I, X, alpha=e_max)
    X = scale(X, 1.0 / (e_max - e_min))
    beta = (e_max - lumo) / (e_max - e_min)
    beta_bar = (e_max - homo) / (e_max - e_min)
    energy = 0.0
    broke = False
    ii = 0
    for ii in range(1, p.max_iterations + 1):
        tv = trace(X)
        if tv > trace_target:                                   # :1051-1059
            alpha = 2.0 / (2.0 - beta)
            X = scale(X, alpha)
            X = increment(I, X, alpha=1.0 - alpha)
            X = multiply(X, X, thr=p.threshold)
            beta = (alpha * beta + 1 - alpha) ** 2
            beta_bar = (alpha * beta_bar + 1 - alpha) ** 2
        else:                                                   # :1060-1068
            alpha = 2.0 / (1.0 + beta_bar)
            X2 = multiply(X, X, thr=p.threshold)
            X = scale(X, 2 * alpha)
            X = increment(X2, X, alpha=-1.0 * alpha ** 2)
            beta = 2.0 * alpha * beta - alpha ** 2 * beta ** 2
            beta_bar = 2.0 * alpha * beta_bar - alpha ** 2 * beta_bar ** 2
        old = energy
        energy = 2.0 * dot_real(X, WH)
        mon.append(energy - old)
        info.history.append(energy - old)
        if mon.converged():
            broke = True
            break
    info.iterations = _loop_counter(ii, p.max_iterations, broke)
    info.energy = energy
    K = _density_finish(X, ISQT, ISQ, p)
    return K, info


def hpcp(H, ISQ, trace_target, p: SolverParameters | None = None):
    """HPCP (DensityMatrixSolversModule.F90:720-950). Returns (K, SolveInfo)."""
    p = p or SolverParameters()
    mon = p.monitor()
    info = SolveInfo()
    I, ISQT, WH = _density_setup(H, ISQ, p)
    n = float(H.n)
    e_min, e_max = gershgorin(WH)
    mu = trace(WH) / n
    sigma_bar = (n - trace_target) / n
    sigma = 1.0 - sigma_bar
    beta = sigma / (e_max - mu)
    beta_bar = sigma_bar / (mu - e_min)
    beta_1, beta_2 = sigma, min(beta, beta_bar)
    D1 = scale(I, beta_1)                                        # :812-818
    T = scale(I, mu)
    T = increment(WH, T, alpha=-1.0)
    T = scale(T, beta_2)
    D1 = increment(T, D1)
    energy = 0.0
    broke = False
    ii = 0
    for ii in range(1, p.max_iterations + 1):
        DH = increment(I, D1.copy(), alpha=-1.0)                 # :830-832
        DH = scale(DH, -1.0)
        DDH = multiply(D1, DH, thr=p.threshold)
        tv = trace(DDH)
        D2DH = multiply(D1, DDH, thr=p.threshold)
        sg = trace(D2DH) / tv
        info.sigmas.append(sg)
        D1 = increment(D2DH, D1, alpha=2.0)
        D1 = increment(DDH, D1, alpha=-1.0 * 2.0 * sg)
        old = energy
        energy = dot_real(D1, WH)
        mon.append(energy - old)
        info.history.append(energy - old)
        if mon.converged():
            broke = True
            break
    info.iterations = _loop_counter(ii, p.max_iterations, broke)
    total = info.iterations - 1
    info.energy = energy
    K = _density_finish(D1, ISQT, ISQ, p)
    a, b, mid = 0.0, 1.0, 0.0
    for _ in range(p.max_iterations):                            # :905-925
        mid = (b - a) / 2.0 + a
        z = mid
        for jj in range(total):
            z = z + 2.0 * ((z * z) * (1.0 - z) - info.sigmas[jj] * z * (1.0 - z))
        if z < 0.5:
            a = mid
        else:
            b = mid
        if abs(z - 0.5) < p.converge_diff:
            break
    info.chemical_potential = mu + (beta_1 - mid) / beta_2
    return K, info


def mcweeny_step(D, S=None, thr=0.0):
    """McWeenyStep (DensityMatrixSolversModule.F90:1190-1231): 3 DSD - 2 DSDSD."""
    DS = multiply(D, S, thr=thr) if S is not None else D.copy()
    DSD = multiply(DS, D, thr=thr)
    out = multiply(DS, DSD, alpha=-2.0, thr=thr)
    return increment(DSD, out, alpha=3.0)


def energy_density_matrix(H, D, thr=0.0):
    """EnergyDensityMatrix (DensityMatrixSolversModule.F90:1165-1187): D H D via SimilarityTransform(H, D, D)."""
    return similarity_transform(H, D, D, thr)


def power_bounds(M, p: SolverParameters | None = None):
    """PowerBounds (EigenBoundsModule.F90:60-189): power iteration with Aitken extrapolation. Returns (value, SolveInfo)."""
    default = p is None
    p = p or SolverParameters()
    maxit = 10 if default else p.max_iterations
    mon = p.monitor()
    info = SolveInfo()
    n = M.n
    v = sp.csc_matrix((np.full(n, 1.0 / n), (np.zeros(n, int), np.arange(n))), shape=(n, n))   # :97-108: first row
    vec = PSMatrix.from_scipy(v, M.grid)
    ritz, ait = [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
    broke = False
    ii = 0
    for ii in range(1, maxit + 1):
        vec2 = multiply(M, vec, thr=p.threshold)
        sv = dot_real(vec, vec)
        mv = dot_real(vec, vec2) / sv
        vec = scale(vec2, 1.0 / norm(vec2))
        ritz = [ritz[1], ritz[2], mv]
        ait = [ait[1], ait[2], 0.0]
        if ii >= 3:
            num = ritz[2] * ritz[0] - ritz[1] ** 2
            den = ritz[2] - 2 * ritz[1] + ritz[0]
            ait[2] = num / den if abs(den) > 1e-14 else ritz[2]
        else:
            ait[2] = ritz[2]
        mon.append(-(ait[2] - ait[1]))
        info.history.append(-(ait[2] - ait[1]))
        if mon.converged() and abs(ait[2] - ritz[2]) < mon.loose:
            broke = True
            break
    info.iterations = _loop_counter(ii, maxit, broke)
    return ait[2], info


# --------------------------------------------------------------------------
# Chebyshev polynomials and the exponential (ChebyshevSolversModule.F90, ExponentialSolversModule.F90)
# --------------------------------------------------------------------------
EXP_CHEBYSHEV_COEFFICIENTS = [                 # ExponentialSolversModule.F90:98-113
    1.266065877752007e+00, 1.130318207984970e+00, 2.714953395340771e-01, 4.433684984866504e-02,
    5.474240442092110e-03, 5.429263119148932e-04, 4.497732295351912e-05, 3.198436462630565e-06,
    1.992124801999838e-07, 1.103677287249654e-08, 5.505891628277851e-10, 2.498021534339559e-11,
    1.038827668772902e-12, 4.032447357431817e-14, 2.127980007794583e-15, -1.629151584468762e-16]


def chebyshev_compute(M, coefficients, p: SolverParameters | None = None):
    """Compute_cheby (ChebyshevSolversModule.F90:83-186): three-term recurrence T_k = 2 M T_{k-1} - T_{k-2},
    the products thresholded, the adds not. Returns (matrix, SolveInfo)."""
    p = p or SolverParameters()
    info = SolveInfo()
    stats = MultiplyStats()
    degree = len(coefficients)
    ident = identity(M)
    bal = M
    if p.do_load_balancing:
        ident = permute(ident, p.permutation)
        bal = permute(bal, p.permutation)
    tkm2 = ident
    if degree == 1:
        out = scale(tkm2, coefficients[0])
    else:
        tkm1 = bal
        out = scale(tkm2, coefficients[0])
        out = increment(tkm1, out, alpha=coefficients[1])
        if degree > 2:
            tk = multiply(bal, tkm1, alpha=2.0, thr=p.threshold, stats=stats)
            tk = increment(tkm2, tk, alpha=-1.0)
            out = increment(tk, out, alpha=coefficients[2])
            for ii in range(4, degree + 1):
                tkm2, tkm1 = tkm1, tk
                tk = multiply(bal, tkm1, alpha=2.0, thr=p.threshold, stats=stats)
                tk = increment(tkm2, tk, alpha=-1.0)
                out = increment(tk, out, alpha=coefficients[ii - 1])
    if p.do_load_balancing:
        out = undo_permute(out, p.permutation)
    info.flops = stats.flops
    return out, info


def compute_exponential(M, p: SolverParameters | None = None):
    """ComputeExponential (ExponentialSolversModule.F90:37-148): spectral radius from PowerBounds with at most 10
    iterations, scaling by a power of two, Chebyshev series of degree 15 with threshold/sigma, repeated squaring.
    info.iterations = sigma_counter (squarings + 1). NOTE (reference behaviour, kept): PowerBounds leaves its loop in
    iteration 1 with the value 0 when M(1,1) = 0 - the first Ritz value is M(1,1), the monitor's tight criterion
    |0| <= converge_diff fires (ConvergenceMonitorModule.F90:121-129) and |aitken - ritz| = 0 < loose_cutoff
    (EigenBoundsModule.F90:156-165) - so such a matrix is NOT scaled, whatever its spectral radius."""
    p = p or SolverParameters()
    psub = SolverParameters(**{**p.__dict__, "max_iterations": 10})
    radius, _ = power_bounds(M, psub)
    sigma_val, sigma_counter = 1.0, 1
    while radius / sigma_val > 1.0:
        sigma_val *= 2
        sigma_counter += 1
    sub = SolverParameters(**{**p.__dict__, "threshold": p.threshold / sigma_val})
    out, info = chebyshev_compute(scale(M, 1.0 / sigma_val), EXP_CHEBYSHEV_COEFFICIENTS, sub)
    stats = MultiplyStats()
    if p.do_load_balancing:
        out = permute(out, p.permutation)
    for _ in range(1, sigma_counter):
        out = multiply(out, out, thr=p.threshold, stats=stats)
    if p.do_load_balancing:
        out = undo_permute(out, p.permutation)
    info.iterations = sigma_counter
    info.flops += stats.flops
    info.sigmas = [sigma_val]
    return out, info


def trs4(H, ISQ, trace_target, p: SolverParameters | None = None):
    """TRS4 (DensityMatrixSolversModule.F90:485-716)."""
    p = p or SolverParameters()
    smin, smax = 0.0, 6.0
    mon = p.monitor()
    info = SolveInfo()
    I, ISQT, WH = _density_setup(H, ISQ, p)
    e_min, e_max = gershgorin(WH)
    X = scale(WH, -1.0)
    X = increment(I, X, alpha=e_max)
    X = scale(X, 1.0 / (e_max - e_min))
    energy = 0.0
    broke = False
    ii = 0
    for ii in range(1, p.max_iterations + 1):
        X2 = multiply(X, X, thr=p.threshold)
        Fx = scale(X2, -3.0)
        Fx = increment(X, Fx, alpha=4.0)
        Gx = I.copy()
        Gx = increment(X, Gx, alpha=-2.0)
        Gx = increment(X2, Gx)
        tfx = dot_real(X2, Fx)
        tgx = dot_real(X2, Gx)
        sg = 0.5 * (smax - smin) if abs(tgx) < 1.0e-14 else (trace_target - tfx) / tgx
        info.sigmas.append(sg)
        if sg > smax:
            T = scale(X, 2.0)
            T = increment(X2, T, alpha=-1.0)
        elif sg < smin:
            T = X2.copy()
        else:
            Gx = scale(Gx, sg)
            Gx = increment(Fx, Gx)
            T = multiply(X2, Gx, thr=p.threshold)
        X = T
        old = energy
        energy = dot_real(X, WH)
        mon.append(energy - old)
        info.history.append(energy - old)
        if mon.converged():
            broke = True
            break
    info.iterations = _loop_counter(ii, p.max_iterations, broke)
    total = info.iterations - 1
    info.energy = energy
    K = _density_finish(X, ISQT, ISQ, p)
    a, b, mid = 0.0, 1.0, 0.0
    for _ in range(p.max_iterations):
        mid = (b - a) / 2.0 + a
        z = mid
        for jj in range(total):
            s = info.sigmas[jj]
            if s > smax:
                z = 2.0 * z - z * z
            elif s < smin:
                z = z * z
            else:
                fx = (z * z) * (4.0 * z - 3.0 * z * z)
                gx = (z * z) * (1.0 - z) * (1.0 - z)
                z = fx + s * gx
        if z < 0.5:
            a = mid
        else:
            b = mid
        if abs(z - 0.5) < p.converge_diff:
            break
    info.chemical_potential = e_max + (e_min - e_max) * mid
    return K, info


def pm(H, ISQ, trace_target, p: SolverParameters | None = None):
    """PM (DensityMatrixSolversModule.F90:37-281)."""
    p = p or SolverParameters()
    mon = p.monitor()
    info = SolveInfo()
    I, ISQT, WH = _density_setup(H, ISQ, p)
    e_min, e_max = gershgorin(WH)
    n = H.n
    X = WH.copy()
    lam = trace(X) / n
    alpha = min(trace_target / (e_max - lam), (n - trace_target) / (lam - e_min))
    X = scale(X, -alpha / n)
    X = increment(I, X, alpha=(alpha * lam + trace_target) / n)
    energy = 0.0
    broke = False
    ii = 0
    tiny = np.finfo(np.float64).tiny
    for ii in range(1, p.max_iterations + 1):
        X2 = multiply(X, X, thr=p.threshold)
        X3 = multiply(X, X2, thr=p.threshold)
        T = increment(X2, X.copy(), alpha=-1.0, thr=p.threshold)
        tv = trace(T)
        tv2 = dot_real(T, X)
        sg = 1.0 if tv <= tiny else tv2 / tv
        info.sigmas.append(sg)
        if sg > 0.5:
            a1, a2, a3 = 0.0, 1.0 + 1.0 / sg, -1.0 / sg
        else:
            a1 = (1.0 - 2.0 * sg) / (1.0 - sg)
            a2 = (1.0 + sg) / (1.0 - sg)
            a3 = -1.0 / (1.0 - sg)
        X = scale(X, a1)
        X = increment(X2, X, alpha=a2, thr=p.threshold)
        X = increment(X3, X, alpha=a3, thr=p.threshold)
        old = energy
        energy = dot_real(X, WH)
        mon.append(energy - old)
        info.history.append(energy - old)
        if mon.converged():
            broke = True
            break
    info.iterations = _loop_counter(ii, p.max_iterations, broke)
    total = info.iterations - 1
    info.energy = energy
    K = _density_finish(X, ISQT, ISQ, p)
    a, b, mid = 0.0, 1.0, 0.0
    for _ in range(p.max_iterations):
        mid = (b - a) / 2.0 + a
        z = mid
        for jj in range(total):
            s = info.sigmas[jj]
            if s > 0.5:
                z = ((1.0 + s) * z ** 2) - (z ** 3)
                z = z / s
            else:
                z = ((1.0 - 2.0 * s) * z) + ((1.0 + s) * z ** 2) - (z ** 3)
                z = z / (1.0 - s)
        if z < 0.5:
            a = mid
        else:
            b = mid
        if abs(z - 0.5) < p.converge_diff:
            break
    info.chemical_potential = lam - (n * mid - trace_target) / alpha
    return K, info


# --------------------------------------------------------------------------
# sign / inverse / square root
# --------------------------------------------------------------------------
def sign_function(M, p: SolverParameters | None = None, polar=False):
    """SignFunction / PolarDecomposition core (SignSolversModule.F90:150-258)."""
    p = p or SolverParameters()
    mon = p.monitor()
    info = SolveInfo()
    alpha = 1.69770248526
    I = identity(M)
    if p.do_load_balancing:
        I = permute(I, p.permutation)
        Out = permute(M, p.permutation)
    else:
        Out = M.copy()
    e_min, e_max = gershgorin(M)
    xk = abs(e_min / e_max)
    Out = scale(Out, 1.0 / abs(e_max))
    broke = False
    ii = 0
    for ii in range(1, p.max_iterations + 1):
        ak = min(math.sqrt(3.0 / (1.0 + xk + xk ** 2)), alpha)
        xk = 0.5 * ak * xk * (3.0 - (ak ** 2) * xk ** 2)
        if polar:
            OT = transpose(Out)
            if OT.is_complex:
                OT = conjugate(OT)
            T1 = multiply(OT, Out, alpha=-1.0 * ak ** 2, thr=p.threshold)
        else:
            T1 = multiply(Out, Out, alpha=-1.0 * ak ** 2, thr=p.threshold)
        T1 = increment(I, T1, alpha=3.0)
        T2 = multiply(Out, T1, alpha=0.5 * ak, thr=p.threshold)
        Out = increment(T2, Out, alpha=-1.0)
        nv = norm(Out)
        Out = T2
        mon.append(nv)
        info.history.append(nv)
        if mon.converged():
            broke = True
            break
    info.iterations = _loop_counter(ii, p.max_iterations, broke)
    if p.do_load_balancing:
        Out = undo_permute(Out, p.permutation)
    return Out, info


def invert(M, p: SolverParameters | None = None):
    """Hotelling Invert (InverseSolversModule.F90:29-149)."""
    p = p or SolverParameters()
    mon = p.monitor()
    info = SolveInfo()
    I = identity(M)
    if p.do_load_balancing:
        I = permute(I, p.permutation)
        Bal = permute(M, p.permutation)
    else:
        Bal = M.copy()
    sg = sigma(Bal)
    Out = scale(Bal, sg)
    broke = False
    ii = 0
    for ii in range(1, p.max_iterations + 1):
        T1 = multiply(Out, Bal, thr=p.threshold)
        T2 = increment(T1, I.copy(), alpha=-1.0)
        nv = norm(T2)
        T2 = multiply(T1, Out, alpha=-1.0, thr=p.threshold)
        Out = scale(Out, 2.0)
        Out = increment(T2, Out, thr=p.threshold)
        mon.append(nv)
        info.history.append(nv)
        if mon.converged():
            broke = True
            break
    info.iterations = _loop_counter(ii, p.max_iterations, broke)
    if p.do_load_balancing:
        Out = undo_permute(Out, p.permutation)
    return Out, info


def _ns_isr_taylor(M, p, order, compute_inverse):
    """NewtonSchultzISRTaylor (SquareRootSolversModule.F90:340-531)."""
    mon = p.monitor()
    info = SolveInfo()
    I = identity(M)
    e_min, e_max = gershgorin(M)
    lam = 1.0 / max(abs(e_min), abs(e_max))
    Z = identity(M)                       # InverseSquareRootMat
    Y = scale(M, lam)                     # SquareRootMat
    if p.do_load_balancing:
        Y = permute(Y, p.permutation)
        I = permute(I, p.permutation)
        Z = permute(Z, p.permutation)
    broke = False
    ii = 0
    for ii in range(1, p.max_iterations + 1):
        X = multiply(Z, Y, thr=p.threshold)
        X = increment(I, X, alpha=-1.0)
        nv = norm(X)
        if order == 3:
            T = multiply(X, X, thr=p.threshold)
            X = scale(X, -0.5)
            X = increment(I, X)
            X = increment(T, X, alpha=0.375)
        elif order == 5:
            aa, bb, cc, dd = -40.0 / 35.0, 48.0 / 35.0, -64.0 / 35.0, 128.0 / 35.0
            a = (aa - 1.0) / 2.0
            b = bb * (a + 1.0) - cc - a * (a + 1.0) ** 2
            c = bb - b - a * (a + 1.0)
            d = dd - b * c
            T = multiply(X, X, thr=p.threshold)
            T = increment(X, T, alpha=a)
            T2 = scale(I, b)
            T2 = increment(X, T2)
            T2 = increment(T, T2)
            T = increment(I, T, alpha=c)
            X = multiply(T2, T, thr=p.threshold)
            X = increment(I, X, alpha=d)
            X = scale(X, 35.0 / 128.0)
        else:
            raise ValueError("order must be 3 or 5 here")
        Z = multiply(X, Z, thr=p.threshold)
        Y = multiply(Y, X, thr=p.threshold)
        mon.append(nv)
        info.history.append(nv)
        if mon.converged():
            broke = True
            break
    info.iterations = _loop_counter(ii, p.max_iterations, broke)
    Out = scale(Z, math.sqrt(lam)) if compute_inverse else scale(Y, 1.0 / math.sqrt(lam))
    if p.do_load_balancing:
        Out = undo_permute(Out, p.permutation)
    return Out, info


def _ns_isr_order2(M, p, compute_inverse):
    """NewtonSchultzISROrder2 (SquareRootSolversModule.F90:201-337)."""
    mon = p.monitor()
    info = SolveInfo()
    I = identity(M)
    Z = identity(M)                       # InverseSquareRootMat
    Y = M.copy()                          # SquareRootMat (NOT pre-scaled in this variant)
    if p.do_load_balancing:
        Y = permute(Y, p.permutation)
        I = permute(I, p.permutation)
        Z = permute(Z, p.permutation)
    broke = False
    ii = 0
    for ii in range(1, p.max_iterations + 1):
        X = multiply(Y, Z, thr=p.threshold)
        e_min, e_max = gershgorin(X)
        lam = 1.0 / max(abs(e_min), abs(e_max))
        X = scale(X, lam)
        T = increment(X, I.copy(), alpha=-1.0)
        nv = norm(T)
        Tk = scale(I, 3.0)
        Tk = increment(X, Tk, alpha=-1.0)
        Tk = scale(Tk, 0.5)
        Z = scale(multiply(Z, Tk, thr=p.threshold), math.sqrt(lam))
        Y = scale(multiply(Tk, Y, thr=p.threshold), math.sqrt(lam))
        mon.append(nv)
        info.history.append(nv)
        if mon.converged():
            broke = True
            break
    info.iterations = _loop_counter(ii, p.max_iterations, broke)
    Out = Z.copy() if compute_inverse else Y.copy()
    if p.do_load_balancing:
        Out = undo_permute(Out, p.permutation)
    return Out, info


def inverse_square_root(M, p: SolverParameters | None = None, order=5):
    p = p or SolverParameters()
    if order == 2:
        return _ns_isr_order2(M, p, True)
    return _ns_isr_taylor(M, p, order, True)


def square_root(M, p: SolverParameters | None = None, order=5):
    p = p or SolverParameters()
    if order == 2:
        return _ns_isr_order2(M, p, False)
    return _ns_isr_taylor(M, p, order, False)
