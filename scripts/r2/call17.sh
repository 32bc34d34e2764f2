#!/bin/bash
# round 2, call 17 (1 GPU): atomic bins with RED + barrier-free sweep, c5 with the convergent shift, ncu of the numeric kernel
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
( time timeout 300 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_smatrix.py tests/test_gpu_helpers.py -m gpu -q --timeout 200 ) > gpurun_out/r2c17_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2c17_pytest.log
grep -v "^  File\|site-packages\|NCCL" gpurun_out/r2c17_pytest.log | grep -v "^$" | tail -n 12
timeout 100 python scripts/r2/probe_c5.py products 32768 2>&1 | grep -v NCCL > gpurun_out/r2c17_probe_products.log; echo "products exit $?"; cat gpurun_out/r2c17_probe_products.log
timeout 150 python scripts/r2/probe_c5.py inv 32768 0.0125 2>&1 | grep -v NCCL > gpurun_out/r2c17_probe_inv32768.log; echo "inv32768 exit $?"; tail -n 24 gpurun_out/r2c17_probe_inv32768.log
timeout 150 python scripts/r2/probe_c5.py exp 32768 0.0125 2>&1 | grep -v NCCL > gpurun_out/r2c17_probe_exp32768.log; echo "exp32768 exit $?"; tail -n 24 gpurun_out/r2c17_probe_exp32768.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile_numeric -s 12 -c 4 -f -o gpurun_out/r2c17_numeric \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check --no-peaks > gpurun_out/r2c17_ncu_numeric.out 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/r2c17_numeric.ncu-rep --page raw --csv > gpurun_out/r2c17_numeric_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c17_numeric.ncu-rep --page source --csv > gpurun_out/r2c17_numeric_source.csv 2>/dev/null
ls -la gpurun_out/r2c17_numeric*
