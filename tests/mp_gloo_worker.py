"""world_size-2 CPU worker (gloo): host-side distribution logic of the N>1 path.
Each rank asks the C ABI what it owns, the ranks exchange that over gloo and check that the blocks
tile the padded matrix exactly once per slice, that communicator colours group the right ranks and that
the per-rank triplet ownership matches the oracle's simulation of the same grid."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    dist.init_process_group("gloo")
    rank, size = dist.get_rank(), dist.get_world_size()
    import ntpoly_b200.api as nt
    from oracle import oracle as O
    import scipy.sparse as sp
    n = 33
    for (R, C, S) in [(size, 1, 1), (1, size, 1), (1, 1, size)]:
        lay = nt.grid_layout(rank, size, R, C, S, n)
        mine = torch.tensor([lay[k] for k in ("my_slice", "my_row", "my_col", "logical_dim", "local_rows", "local_cols",
                                               "start_row", "start_col", "row_comm_colour", "col_comm_colour")])
        allv = [torch.zeros_like(mine) for _ in range(size)]
        dist.all_gather(allv, mine)
        allv = torch.stack(allv).numpy()
        g = O.Grid(R, C, S)
        N = g.padded(n)
        assert np.all(allv[:, 3] == N)
        for s in range(S):
            cover = np.zeros((N, N), int)
            for q in range(size):
                sl, r, c, _, lr, lc, r0, c0 = allv[q, :8]
                assert (sl, r, c) == g.coords(q)
                if sl == s:
                    cover[r0:r0 + lr, c0:c0 + lc] += 1
            assert np.all(cover == 1), "blocks of one slice must tile the padded matrix exactly once"
        # ranks sharing a row-communicator colour share (slice,row); same for columns
        for q in range(size):
            for p in range(size):
                assert (allv[q, 8] == allv[p, 8]) == (tuple(allv[q, :2]) == tuple(allv[p, :2]))
                assert (allv[q, 9] == allv[p, 9]) == ((allv[q, 0], allv[q, 2]) == (allv[p, 0], allv[p, 2]))
        # ownership of a random matrix's entries: this rank's block vs the oracle's per-rank triplets
        a = sp.random(n, n, 0.3, random_state=5, format="coo")
        M = O.PSMatrix.from_scipy(a, g)
        rows, cols, vals = M.local_triplets(rank)
        r0, c0, lr, lc = lay["start_row"], lay["start_col"], lay["local_rows"], lay["local_cols"]
        sel = (a.row >= r0) & (a.row < r0 + lr) & (a.col >= c0) & (a.col < c0 + lc)
        assert len(rows) == int(sel.sum())
        assert np.all((rows - 1 >= r0) & (rows - 1 < r0 + lr) & (cols - 1 >= c0) & (cols - 1 < c0 + lc))
        # the distributed product on this grid equals the 1x1x1 product (checked on the owned block)
        P = O.multiply(M, M)
        P1 = O.multiply(O.PSMatrix.from_scipy(a), O.PSMatrix.from_scipy(a))
        pr, pc, pv = P.local_triplets(rank)
        ref = P1.to_scipy().tocsr()
        for i, j, v in zip(pr, pc, pv):
            assert abs(ref[i - 1, j - 1] - v) < 1e-12
    assert nt.default_grid(8) == (1, 2, 4) and nt.default_grid(2) == (1, 1, 2) and nt.default_grid(4) == (1, 1, 4)
    dist.barrier()
    if rank == 0:
        print("MP_GLOO_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
