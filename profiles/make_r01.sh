#!/bin/bash
# Round-1 measurement recipe (run on the B200 box through gpurun from the repo root):
#   bench line, launch list (ncu gpu__time_duration), one full capture of the step kernel (defender + attacker turn)
set -e
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cyg_ -c 60 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
python bench.py --steps 9 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cyg_step_kernel -s 8 -c 2 -o gpurun_out/r01_step_kernel \
    python bench.py --steps 9 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/r01_bench.json | cut -c1-300
