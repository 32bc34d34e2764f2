"""Drivers of the C ABI beyond the ones tests/test_gpu_solvers.py covers: HPCP, PolarDecomposition, PowerBounds,
McWeenyStep(S), EnergyDensityMatrix (cases in tests/driver_case_worker.py, which follow the reference's own tests at
its tolerance 1e-4). Each case runs in a process of its own: an abort inside the library must not take the session
down. First hardware run: round 2, call 1 (profiles/r02a_pending_first_run.log) - all green, marks removed."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("case", ["hpcp_premade_density", "polar_decomposition", "power_bounds",
                                  "mcweeny_step_and_energy_density"])
def test_driver_case(case):
    r = subprocess.run([sys.executable, os.path.join(HERE, "driver_case_worker.py"), case], capture_output=True,
                       text=True, timeout=120)
    assert r.returncode == 0 and "DRIVER_CASE_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]
