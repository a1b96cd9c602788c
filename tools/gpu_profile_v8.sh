#!/bin/bash
# round-1 v8: software-pipelined NetVLAD kernel
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:netvlad_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_netvlad_tc_v3 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_nv3.log 2>&1
