"""Multi-GPU worker (torchrun, one process per GPU): the distributed multiply, helpers and a solver on
an R x C x S process grid, each rank checking ITS block against the oracle's simulation of that grid."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def local_block(nt_mat):
    rows, cols, vals = nt_mat.get_arrays()
    n = nt_mat.GetLogicalDimension()
    return sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(n, n)).tocsc()


def oracle_block(O, M, rank):
    rows, cols, vals = M.local_triplets(rank)
    return sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(M.N, M.N)).tocsc()


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ntpoly_b200.api as nt
    from oracle import oracle as O
    from util import compare_sparse, banded
    nt.init_world_from_torch()
    grids = {2: [(2, 1, 1), (1, 2, 1), (1, 1, 2)], 4: [(2, 2, 1), (1, 2, 2), (4, 1, 1), (1, 4, 1)],
             8: [(2, 2, 2), (4, 2, 1), (1, 2, 4), (1, 8, 1)]}[world]
    if os.environ.get("NTB_WORKER_GRIDS"):            # e.g. "1x2x1": the fallback variants of tests/test_gpu_multi.py
        grids = [tuple(int(x) for x in gs.split("x")) for gs in os.environ["NTB_WORKER_GRIDS"].split(",")]
    # what a column-split grid is expected to do: "peer" (default: operand tiles of the neighbours read in place over
    # NVLink), "nccl" (NTB_P2P=0: tile halo copied with ncclSend/Recv), "gather" (the tile path declines collectively,
    # the reference-style CSC panel gather runs)
    expect = os.environ.get("NTB_WORKER_EXPECT", "peer")
    for (R, C, S) in grids:
        nt.ConstructGlobalProcessGrid(R, C, S)
        g = O.Grid(R, C, S)
        assert (nt.GetGlobalMySlice(), nt.GetGlobalMyRow(), nt.GetGlobalMyColumn()) == g.coords(rank)
        for n, fill, thr, cplx in [(33, 0.2, 0.0, False), (33, 1.0, 0.0, False), (257, 0.05, 1e-3, False),
                                   (120, 0.1, 1e-4, True)]:
            rng = np.random.default_rng(100 + n)
            a = sp.random(n, n, fill, random_state=rng, format="coo")
            b = sp.random(n, n, fill, random_state=rng, format="coo")
            if cplx:
                a = (a + 1j * sp.random(n, n, fill, random_state=rng, format="coo")).tocoo()
            A, B, Cm = nt.Matrix_ps(n, is_complex=cplx), nt.Matrix_ps(n), nt.Matrix_ps(n)
            # every entry is contributed by exactly one rank (round robin), like a user program would
            A.fill_from_arrays(a.row[rank::world] + 1, a.col[rank::world] + 1, a.data[rank::world])
            B.fill_from_arrays(b.row[rank::world] + 1, b.col[rank::world] + 1, b.data[rank::world])
            OA, OB = O.PSMatrix.from_scipy(a, g, is_complex=cplx), O.PSMatrix.from_scipy(b, g)
            assert abs(local_block(A) - oracle_block(O, OA, rank)).sum() < 1e-14, "ingest / ownership"
            assert A.GetSize() == OA.nnz()
            Cm.Gemm(A, B, None, alpha=0.8, threshold=thr)
            ref = O.multiply(OA, OB, alpha=0.8, thr=thr)
            compare_sparse(local_block(Cm), oracle_block(O, ref, rank), thr)
            assert abs(Cm.GetSize() - ref.nnz()) <= max(2, 1e-3 * ref.nnz())
            # helpers (collective scalars)
            assert abs(A.Trace() - O.trace(OA)) < 1e-12
            assert abs(A.Norm() - O.norm(OA)) < 1e-12
            assert abs(A.Dot(B) - np.real(O.dot(OA, OB))) < 1e-12
            B.Increment(A, 0.5, thr)
            refB = O.increment(OA, OB, alpha=0.5, thr=thr)
            assert abs(local_block(B) - oracle_block(O, refB, rank)).sum() < 1e-13
            T = nt.Matrix_ps(n)
            T.Transpose(A)
            assert abs(local_block(T) - oracle_block(O, O.transpose(OA), rank)).sum() < 1e-14
            # a list that is the rank's own block in column-major order is taken without sort or gather (collective
            # decision inside the slice); the ranks of the other slices contribute nothing and receive the replica
            rows, cols, vals = T.get_arrays()
            if nt.GetGlobalMySlice() != 0:
                rows, cols, vals = rows[:0], cols[:0], vals[:0]
            T2 = nt.Matrix_ps(n, is_complex=cplx)
            before = nt.sorted_ingests()
            T2.fill_from_arrays(rows, cols, vals)
            assert nt.sorted_ingests() == before + 1
            assert abs(local_block(T2) - local_block(T)).sum() == 0.0 and T2.GetSize() == T.GetSize()
            # load balancing by relabelling == by two products, on this grid
            perm = nt.Permutation(T.GetLogicalDimension()); perm.SetRandomPermutation(seed=3)
            blocks = []
            for gemm in (True, False):
                nt.set_permute_gemm(gemm)
                P1, U1 = nt.Matrix_ps(n), nt.Matrix_ps(n)
                nt.LoadBalancer.PermuteMatrix(T, P1, perm)
                nt.LoadBalancer.UndoPermuteMatrix(P1, U1, perm)
                blocks.append((local_block(P1), local_block(U1)))
            nt.set_permute_gemm(False)
            assert abs(blocks[0][0] - blocks[1][0]).sum() == 0.0 and abs(blocks[0][1] - blocks[1][1]).sum() == 0.0
            assert abs(blocks[1][1] - local_block(T)).sum() == 0.0
        # banded product on the tile path + a solver with identical iteration count
        n = 2048
        a = banded(n, half_bandwidth=24).tocoo()
        A, Cm = nt.Matrix_ps(n), nt.Matrix_ps(n)
        A.fill_from_arrays(a.row[rank::world] + 1, a.col[rank::world] + 1, a.data[rank::world])
        OA = O.PSMatrix.from_scipy(a, g)
        nt.reset_counters()
        Cm.Gemm(A, A, None, threshold=1e-9)
        compare_sparse(local_block(Cm), oracle_block(O, O.multiply(OA, OA, thr=1e-9), rank), 1e-9)
        column_split = (R == 1 and S == 1 and C > 1)
        # column-split grids fetch the left operand as a tile halo; the other grids gather CSC panels
        halo = column_split and expect != "gather"
        assert nt.halo_counters()["products"] == (1 if halo else 0), nt.halo_counters()
        if column_split:
            pc = nt.peer_counters()
            assert pc["ok"] == (os.environ.get("NTB_P2P") != "0") and pc["products"] == (1 if expect == "peer" else 0), pc
        # a product of products reads the tile forms emitted with Cm; check it against the oracle fed with the GPU's Cm
        parts = [None] * world
        dist.all_gather_object(parts, local_block(Cm))
        # slices hold replicas: assemble the global matrix from slice 0 only
        OC = O.PSMatrix.from_scipy(sum(pb for r_, pb in enumerate(parts) if g.coords(r_)[0] == 0).tocoo(), g)
        D = nt.Matrix_ps(n)
        D.Gemm(Cm, A, None, alpha=-0.5, threshold=1e-9)
        compare_sparse(local_block(D), oracle_block(O, O.multiply(OC, OA, alpha=-0.5, thr=1e-9), rank), 1e-9)
        D.Gemm(A, Cm, None, threshold=1e-9)
        compare_sparse(local_block(D), oracle_block(O, O.multiply(OA, OC, thr=1e-9), rank), 1e-9)
        if halo:
            assert nt.halo_counters()["products"] == 3 and nt.tile_builds() == 2
        # fused identity shift across the grid == the two reference calls, bit for bit
        I = nt.Matrix_ps(n); I.FillIdentity()
        F, G2 = nt.Matrix_ps(n), nt.Matrix_ps(n)
        F.GemmShift(A, A, I, 3.0, None, alpha=-1.1, threshold=1e-9)
        G2.Gemm(A, A, None, alpha=-1.1, threshold=1e-9)
        G2.Increment(I, 3.0)
        fb, gb = local_block(F), local_block(G2)
        fb.sort_indices(); gb.sort_indices()
        assert np.array_equal(fb.indptr, gb.indptr) and np.array_equal(fb.indices, gb.indices) and np.array_equal(fb.data, gb.data)
        # sign function: iteration count and result against the oracle
        sgn = (banded(1024, half_bandwidth=10, scale=0.2) - 0.05 * sp.identity(1024)).tocoo()
        Sg, So = nt.Matrix_ps(1024), nt.Matrix_ps(1024)
        Sg.fill_from_arrays(sgn.row[rank::world] + 1, sgn.col[rank::world] + 1, sgn.data[rank::world])
        sps = nt.SolverParameters(); sps.SetConvergeDiff(1e-6); sps.SetThreshold(1e-8)
        nt.SignSolvers.ComputeSign(Sg, So, sps)
        OS = O.PSMatrix.from_scipy(sgn, g)
        Sref, sinfo = O.sign_function(OS, O.SolverParameters(converge_diff=1e-6, threshold=1e-8))
        assert nt.last_solve()["loop_counter"] == sinfo.iterations, (nt.last_solve(), sinfo.iterations)
        compare_sparse(local_block(So), oracle_block(O, Sref, rank), 1e-8, tol=1e-7)
        h = (banded(512, half_bandwidth=6, scale=0.2)).tocoo()
        H, ISQ, K = nt.Matrix_ps(512), nt.Matrix_ps(512), nt.Matrix_ps(512)
        H.fill_from_arrays(h.row[rank::world] + 1, h.col[rank::world] + 1, h.data[rank::world])
        ISQ.FillIdentity()
        sp_ = nt.SolverParameters(); sp_.SetConvergeDiff(1e-6); sp_.SetThreshold(1e-8)
        e, mu = nt.DensityMatrixSolvers.TRS2(H, ISQ, 256, K, sp_)
        OH = O.PSMatrix.from_scipy(h, g)
        Kref, info = O.trs2(OH, O.identity(OH), 256, O.SolverParameters(converge_diff=1e-6, threshold=1e-8))
        assert nt.last_solve()["loop_counter"] == info.iterations, (nt.last_solve(), info.iterations)
        assert abs(e - info.energy) <= 1e-8 * abs(info.energy)
        compare_sparse(local_block(K), oracle_block(O, Kref, rank), 1e-8, tol=1e-7)
        nt.DestructGlobalProcessGrid()
        dist.barrier()
        if rank == 0:
            print(f"grid {R}x{C}x{S} ok", flush=True)
    if rank == 0:
        print("MP_GPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
