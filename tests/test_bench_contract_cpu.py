"""The reference arm of bench.py (the CPU restatement timed on the host cores) runs without a GPU: check its JSON line
against the contract the driver parses, on the smallest configuration (c1) and on a reduced c4."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(args, env_extra=None):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + args,
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


@pytest.mark.parametrize("args,metric", [
    (["--config", "c1", "--steps", "1", "--warmup", "0"], "spgemm_useful_gflops_single_multiply"),
    (["--config", "c4", "--n", "4096", "--steps", "1", "--warmup", "0"], "spgemm_useful_gflops_per_sign_iteration"),
])
def test_reference_arm_line(args, metric):
    d = _line(args)
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == "GFLOP/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "grid" in d["config"]


def test_reference_arm_under_torchrun_only_rank_zero_prints():
    env = {k: v for k, v in os.environ.items()}
    env.update({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
