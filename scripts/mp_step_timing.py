"""torchrun worker: wall-clock per C-ABI call of one sign iteration on a 1 x C x 1 grid (developer probe)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import ntpoly_b200.api as nt
from ntpoly_b200.workloads import banded_sign_input
nt.init_world_from_torch()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
thr = 1e-6
nt.ConstructGlobalProcessGrid(1, world, 1)
m = banded_sign_input(n).tocoo()
M = nt.Matrix_ps(n); M.fill_from_arrays(m.row[rank::world] + 1, m.col[rank::world] + 1, m.data[rank::world])
I = nt.Matrix_ps(n); I.FillIdentity()
emin, emax = nt.EigenBounds.GershgorinBounds(M)
X = nt.Matrix_ps(M); X.Scale(1.0 / abs(emax))
T1, T2, W = nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n)
def timed(name, fn):
    nt.synchronize(); dist.barrier(); t0 = time.perf_counter(); r = fn(); nt.synchronize()
    if rank == 0: print(f"  {name:28s} {(time.perf_counter()-t0)*1e3:9.3f} ms", flush=True)
    return r
for it in range(4):
    nt.reset_counters()
    timed("sign_iteration", lambda: nt.sign_iteration(X, I, T1, T2, 1.2, thr))
    if rank == 0: print("   halo", nt.halo_counters(), "builds", nt.tile_builds(), "nnz local", X.get_arrays()[0].size, flush=True)
ak = 1.2
nt.profile_enable(True)
for it in range(2):
    nt.reset_counters(); nt.profile_read()
    timed("gemm-shift X*X", lambda: T1.GemmShift(X, X, I, 3.0, None, alpha=-ak*ak, threshold=thr))
    if rank == 0: print("     numeric", nt.profile_read(), nt.tile_counters(), "nnz out", T1.get_arrays()[0].size, flush=True)
    nt.reset_counters()
    timed("gemm X*T1", lambda: T2.Gemm(X, T1, None, alpha=0.5*ak, threshold=thr))
    if rank == 0: print("     numeric", nt.profile_read(), nt.tile_counters(), "nnz out", T2.get_arrays()[0].size, flush=True)
    timed("copy", lambda: nt.lib().CopyMatrix_ps_wrp(X.ih, W.ih))
    timed("norm", lambda: X.Norm())
dist.destroy_process_group()
