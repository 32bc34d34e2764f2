#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python scripts/r2/debug_trs4.py > gpurun_out/r2c10_debug_trs4.log 2>&1
echo "exit $?"; tail -n 60 gpurun_out/r2c10_debug_trs4.log
