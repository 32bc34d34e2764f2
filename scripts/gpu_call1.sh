#!/bin/bash
# GPU call 1 of this session: parity suite, bench (both pipeline shapes), phase probe, ncu launch list + source-level capture, micro
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/c1_pytest.log 2>&1
echo "pytest exit: $?" >> gpurun_out/c1_pytest.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
NTB_NUMERIC_SHAPE=23 timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c1_bench_shape23.json 2> gpurun_out/c1_bench_shape23.err
NTB_TILE_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c1_phase.json 2> gpurun_out/c1_phase.err
scripts/micro/dmma_dfma_mix > gpurun_out/c1_micro_mix.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c1_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c1_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_numeric -s 14 -c 2 -f -o gpurun_out/c1_numeric \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c1_ncu_numeric.out 2>&1
ls -la gpurun_out > gpurun_out/c1_ls.txt
tail -5 gpurun_out/c1_pytest.log; cat gpurun_out/c1_bench.json; cat gpurun_out/c1_bench_shape23.json; cat gpurun_out/c1_micro_mix.txt
