#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( timeout 300 python -m pytest tests/test_gpu_tile_space.py "tests/test_gpu_solvers.py::test_fused_steps_equal_the_reference_call_sequence" -m gpu -q --timeout 200 ) > gpurun_out/r2c9_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c9_pytest.log
grep -v "^  File\|site-packages" gpurun_out/r2c9_pytest.log | grep -v "^$" | tail -n 70
