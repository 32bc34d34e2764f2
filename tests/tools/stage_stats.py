"""Offline analysis (CPU, TEST INFRASTRUCTURE: lives under tests/ because it runs the CPU checker): how the DMMAs of the bench step's two products distribute over stage kinds.
Builds the bench iterate X_3 of a banded N (default 8192; the band structure is translation invariant) with the CPU
restatement, lays the tile masks of the numeric kernel over A and B (8x4 / 4x8 tiles, 64x32 / 32x64 super-tiles) and
counts, per (stage, DMMA warp): DMMAs in 'dense' stages (A super-tile complete and the warp's B tile column complete:
the fast path of k_tile_numeric9), in 'A complete, B column partial' stages, and the rest."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
thr, iterate = 1e-6, 3
from oracle import oracle as O
from ntpoly_b200.workloads import banded_sign_input
O.build()
M = O.PSMatrix.from_scipy(banded_sign_input(n)); I = O.identity(M)
e_min, e_max = O.gershgorin(M)
al = bench.alpha_sequence(e_min, e_max, iterate)
X = O.scale(M, 1.0 / abs(e_max))
def step(X, ak):
    T1 = O.increment(I, O.multiply(X, X, alpha=-ak * ak, thr=thr), alpha=3.0)
    return T1, O.multiply(X, T1, alpha=0.5 * ak, thr=thr)
for k in range(iterate - 1):
    _, X = step(X, al[k])
T1, Xn = step(X, al[iterate - 1])
Xs, T1s = X.to_scipy().tocsc(), T1.to_scipy().tocsc()
print(f"n={n} nnz/col X={Xs.nnz / n:.1f} T1={T1s.nnz / n:.1f} Xnext={Xn.to_scipy().nnz / n:.1f}")

def tile_grid(m, th, tw):
    """boolean grid of th x tw tiles holding at least one entry"""
    c = m.tocoo()
    g = np.zeros((-(-m.shape[0] // th), -(-m.shape[1] // tw)), bool)
    g[c.row // th, c.col // tw] = True
    return g

def stats(Am, Bm, name):
    A = tile_grid(Am, 8, 4)          # A tiles: 8 rows x 4 inner
    B = tile_grid(Bm, 4, 8)          # B tiles: 4 inner x 8 cols
    nIb, nc, nJ = A.shape[0] // 8, A.shape[1] // 8, B.shape[1]
    tot = dense = afull = 0
    for c in range(nc):
        Ac = A[:, 8 * c:8 * c + 8].reshape(nIb, 8, 8)          # [Ib, ii, kk]
        Bc = B[8 * c:8 * c + 8, :]                               # [kk, J]
        a_cnt = Ac.sum(axis=1)                                   # [Ib, kk] tiles per inner tile
        a_full = Ac.reshape(nIb, 64).all(axis=1)                 # [Ib]
        b_full = Bc.all(axis=0)                                  # [J]
        d = a_cnt.astype(np.int64) @ Bc.astype(np.int64)         # [Ib, J] DMMAs of (row block, tile column) in this stage
        tot += d.sum()
        dense += d[np.ix_(a_full, b_full)].sum()
        afull += d[a_full, :].sum()
    print(f"{name}: DMMAs {tot}  dense stages {dense / tot:.3f}  A complete (any B column) {afull / tot:.3f}")

stats(Xs, Xs, "X * X ")
stats(Xs, T1s, "X * T1")


def ragged_histogram(Am, Bm, name):
    """DMMAs per (stage, warp) pair outside the dense stages: how much work a warp finds per stage it has to visit"""
    A = tile_grid(Am, 8, 4); B = tile_grid(Bm, 4, 8)
    nIb, nc = A.shape[0] // 8, A.shape[1] // 8
    hist = np.zeros(65, np.int64)
    visited_empty = 0
    for c in range(nc):
        Ac = A[:, 8 * c:8 * c + 8].reshape(nIb, 8, 8)
        Bc = B[8 * c:8 * c + 8, :]
        a_cnt = Ac.sum(axis=1)
        a_full = Ac.reshape(nIb, 64).all(axis=1)
        b_full = Bc.all(axis=0)
        d = a_cnt.astype(np.int64) @ Bc.astype(np.int64)                 # [Ib, J]
        # a stage exists for task (g, Ib) when the A super-tile and the B super-tile of the group share an inner tile
        Bg = Bc.reshape(8, -1, 8).any(axis=2)                            # [kk, g]
        stage = (Ac.any(axis=1).astype(np.int64) @ Bg.astype(np.int64)) > 0   # [Ib, g]
        stage_J = np.repeat(stage, 8, axis=1)[:, :d.shape[1]]
        dense = np.outer(a_full, b_full)
        sel = stage_J & ~dense
        np.add.at(hist, d[sel], 1)
    tot_pairs = hist.sum()
    work = (hist * np.arange(65)).sum()
    print(f"{name}: ragged (stage, warp) pairs {tot_pairs}, mean DMMAs per pair {work / tot_pairs:.1f} of 64; "
          f"pairs with 0 DMMAs {hist[0] / tot_pairs:.2f}, 1-16: {hist[1:17].sum() / tot_pairs:.2f}, "
          f"17-48: {hist[17:49].sum() / tot_pairs:.2f}, 49-64: {hist[49:].sum() / tot_pairs:.2f}")


ragged_histogram(Xs, Xs, "X * X ")
ragged_histogram(Xs, T1s, "X * T1")


def ownership_balance(Am, Bm, name):
    """Sum over stages of the busiest warp's DMMA count (the stage advances at that warp's pace) for different ways of
    giving the 8x8 tiles of a 64x64 block to 8 warps; the ideal is total/8."""
    A = tile_grid(Am, 8, 4); B = tile_grid(Bm, 4, 8)
    nIb, nc, nG = A.shape[0] // 8, A.shape[1] // 8, B.shape[1] // 8
    schemes = {"8x1 (tile column per warp, today)": (8, 1), "4x2": (4, 2), "2x4": (2, 4), "1x8 (row tile per warp)": (1, 8)}
    busiest = {k: 0 for k in schemes}
    busiest_diag = [0]
    total = 0
    for c in range(nc):
        Ac = A[:, 8 * c:8 * c + 8].reshape(nIb, 8, 8).astype(np.int64)      # [Ib, ii, kk]
        Bc = B[8 * c:8 * c + 8, :nG * 8].reshape(8, nG, 8).astype(np.int64)  # [kk, g, jj]
        # d[Ib, g, ii, jj] = number of inner tiles kk with A(ii,kk) and B(kk,jj) present
        d = np.einsum("aik,kgj->agij", Ac, Bc)
        total += d.sum()
        for k, (hr, wc) in schemes.items():
            per_warp = d.reshape(nIb, nG, 8 // hr, hr, 8 // wc, wc).sum(axis=(3, 5))   # [Ib, g, row group, col group]
            busiest[k] += per_warp.reshape(nIb, nG, -1).max(axis=2).sum()
        # wrapped diagonals: warp w owns the tiles (ii, (ii + w) mod 8)
        ii = np.arange(8)
        diag = np.stack([d[:, :, ii, (ii + w) % 8].sum(axis=2) for w in range(8)], axis=2)   # [Ib, g, w]
        busiest_diag[0] += diag.max(axis=2).sum()
    print(f"{name}: DMMAs {total}, ideal busiest-warp sum {total / 8:.0f}; " +
          "; ".join(f"{k}: {v / (total / 8):.3f}x" for k, v in busiest.items()) +
          f"; wrapped diagonals: {busiest_diag[0] / (total / 8):.3f}x")


ownership_balance(Xs, Xs, "X * X ")
ownership_balance(Xs, T1s, "X * T1")


if len(sys.argv) > 2 and sys.argv[2] == "block":
    # the block-sparse workload (config 3: 32x32 dense blocks) under the same masks, first product of a purification
    from ntpoly_b200.workloads import block_sparse
    Hb = block_sparse(n=n, block=32, neighbours=20, band_blocks=64, seed=1234).tocsc()
    print(f"block-sparse n={n} nnz/col {Hb.nnz / n:.1f}")
    stats(Hb, Hb, "H * H ")
    ragged_histogram(Hb, Hb, "H * H ")
    ownership_balance(Hb, Hb, "H * H ")
