#!/bin/bash
# round 2, call 33 (1 GPU): whole GPU suite + smoke on the final build
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c33_smoke.log 2>&1; echo "smoke exit $?"; grep -v NCCL gpurun_out/r2c33_smoke.log | tail -n 2
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c33_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c33_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c33_pytest.log | grep -v "^$" | tail -n 12
