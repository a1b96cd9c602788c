#!/bin/bash
# ncu launch list of the bench command + one full capture of the tcgen05 conv kernel (B200_PROFILING.md)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 12 -c 3 -f -o gpurun_out/prof_conv_tc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
