#!/bin/bash
# Round-1 measurement recipe (run on the B200 box through gpurun from the repo root):
#   bench line (default command: 4 plain steps fused per launch + the one-launch-per-step leg), launch list of the same
#   command (ncu gpu__time_duration), one full capture of single-step launches (defender + attacker turn) and one of a
#   fused launch
set -e
mkdir -p gpurun_out
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cyg_ -c 80 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
python bench.py --fuse 1 --steps 9 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cyg_step_kernel -s 8 -c 2 -o gpurun_out/r01_step_kernel \
    python bench.py --fuse 1 --steps 9 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:cyg_step_kernel -s 4 -c 2 -o gpurun_out/r01_step_kernel_fused \
    python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_fused.log 2>&1
tail -1 gpurun_out/r01_bench.json | cut -c1-300
