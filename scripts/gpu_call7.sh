#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( timeout 600 python -m pytest tests/test_gpu_solvers.py -m gpu -q --timeout 600 -x -s -k "inverse_square_root" ) > gpurun_out/c7_isr.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_solvers.py -m gpu -q --timeout 600 -s -k "not inverse_square_root" ) > gpurun_out/c7_rest.log 2>&1
grep -v "^  File\|site-packages" gpurun_out/c7_isr.log | tail -n 25
tail -n 15 gpurun_out/c7_rest.log
