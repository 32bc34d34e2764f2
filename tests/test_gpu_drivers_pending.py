"""Drivers of the C ABI that had no parity test of their own when the GPU budget of round 1 ran out: HPCP,
PolarDecomposition, PowerBounds, McWeenyStep(S), EnergyDensityMatrix (cases in tests/pending_driver_worker.py, which
follow the reference's own tests at its tolerance 1e-4).

They have NOT run on hardware yet. Each case runs in a process of its own (an abort inside the library must not take
the session down) and is marked xfail(strict=False): a defect found by the first run shows up as XFAIL in the log
instead of stopping the suite, a pass as XPASS; round 2 removes the mark."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="added after the round-1 GPU budget was spent: first hardware run pending")]
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("case", ["hpcp_premade_density", "polar_decomposition", "power_bounds",
                                  "mcweeny_step_and_energy_density"])
def test_pending_driver(case):
    r = subprocess.run([sys.executable, os.path.join(HERE, "pending_driver_worker.py"), case], capture_output=True,
                       text=True, timeout=120)
    assert r.returncode == 0 and "PENDING_CASE_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]
