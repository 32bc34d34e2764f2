"""CPU-only, world_size 2 over gloo: the host-side logic of the N>1 path (no GPU, no NCCL)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world_size_two_gloo():
    env = dict(os.environ)
    env.pop("RANK", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "mp_gloo_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "MP_GLOO_OK" in r.stdout
