#!/bin/bash
# kernel-only durations of the fused NetVLAD path (ncu, no clock control)
mkdir -p gpurun_out
IBL_NV_DEBUG=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:netvlad --csv --log-file gpurun_out/nv_time.csv \
    python tools/nv_check.py > gpurun_out/nv_time.log 2>&1
grep "nv4\]" gpurun_out/nv_time.log | head -2
tail -12 gpurun_out/nv_time.csv | cut -d, -f5,12-
