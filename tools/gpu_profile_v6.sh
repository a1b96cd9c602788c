#!/bin/bash
# round-1 v6: distance kernel on SM pairs (tcgen05.mma.cta_group::2)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm2_top16_kernel -s 1 -c 2 -f -o gpurun_out/prof_dist_2sm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_dist2.log 2>&1
