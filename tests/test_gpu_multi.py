"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a 1-GPU box): tests/mp_gpu_worker.py under torchrun, every
rank checking its block against the oracle's simulation of the same process grid."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run(nproc, extra_env, port):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mp_gpu_worker.py")]
    env = dict(os.environ)
    env.update(extra_env)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-4000:]
    assert "MP_GPU_OK" in r.stdout


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_process_grids(nproc):
    _run(nproc, {}, 29620 + nproc)


@pytest.mark.parametrize("name,env", [
    # without a peer space the tile halo is copied with NCCL (round-1 path, kept as fallback)
    ("nccl_halo", {"NTB_P2P": "0", "NTB_WORKER_GRIDS": "1x2x1", "NTB_WORKER_EXPECT": "nccl"}),
    # a result whose tile slots exceed the budget: every rank declines together (the verdicts ride on a peer exchange)
    # and the reference-style panel gather runs - degrade, do not abort
    ("peer_declines", {"NTB_TILE_TASK_LIMIT": "4", "NTB_WORKER_GRIDS": "1x2x1", "NTB_WORKER_EXPECT": "gather"}),
    # the same for the NCCL halo: a halo beyond the buffer's index range declines collectively
    ("nccl_halo_declines", {"NTB_P2P": "0", "NTB_HALO_TILE_LIMIT": "1", "NTB_WORKER_GRIDS": "1x2x1",
                            "NTB_WORKER_EXPECT": "gather"}),
    # a process column with more ranks than a kernel argument can list (> 16 in production): the part list of the
    # stacked B panel lives in device memory - forced here on two process rows
    ("device_part_list", {"NTB_STACK_INLINE_PARTS": "1", "NTB_WORKER_GRIDS": "2x1x1"}),
])
def test_column_split_fallbacks(name, env):
    _run(2, env, 29640)
