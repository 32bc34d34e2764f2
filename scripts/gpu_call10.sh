#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 600 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/c10_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c10_pytest.log
grep -v "^  File\|site-packages" gpurun_out/c10_pytest.log | tail -n 60
