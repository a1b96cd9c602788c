#!/bin/bash
# ncu full capture of the tcgen05 distance/top-16 kernel and the PCA GEMM (one launch each)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 4 -f -o gpurun_out/prof_gemm_tc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
