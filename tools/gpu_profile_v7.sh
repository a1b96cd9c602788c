#!/bin/bash
# round-1 v7: conv on SM pairs (256-wide), halo staging (128-wide), final launch list + per-layer table
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv3x3_tc_kernel<\(int\)256" -s 6 -c 1 -f -o gpurun_out/prof_conv256_pair \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c256.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv3x3_tc_kernel<\(int\)128" -s 3 -c 1 -f -o gpurun_out/prof_conv128_halo \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c128.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_v7.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/bench_layers.py > gpurun_out/bench_layers_v3.txt 2>&1
