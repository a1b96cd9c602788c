#!/bin/bash
# short 1-GPU call: launch list + three ncu --set full captures (bench without the strong / eager / cpu legs)
mkdir -p gpurun_out
timeout 300 tools/gpu_profile.sh launches r02_launches_s4
timeout 240 tools/gpu_profile.sh full gemm2_f16_top16 dist_f16 1 1
timeout 240 tools/gpu_profile.sh full conv1_fused conv1_fused 2 1
timeout 240 tools/gpu_profile.sh full netvlad_tc netvlad_1k 2 1
ls -la gpurun_out/*.ncu-rep | tail -4
