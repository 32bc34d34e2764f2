#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python scripts/e2e_timing.py > gpurun_out/c11_e2e_timing.txt 2>&1
cat gpurun_out/c11_e2e_timing.txt | tail -40
