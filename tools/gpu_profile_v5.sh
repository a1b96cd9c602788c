#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tc_kernel<\(int\)256" -s 1 -c 2 -f -o gpurun_out/prof_dist_tc_v2 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_dist.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv1_1_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_conv1_tc_v2 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c1.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv3x3_tc_kernel<\(int\)64" -s 1 -c 1 -f -o gpurun_out/prof_conv64_tc_v2 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c64.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_v5.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
