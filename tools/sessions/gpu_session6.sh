#!/bin/bash
# short session: distance variants after the chain-free insertion, retrieval tests, ncu of the dist kernel, launch list
mkdir -p gpurun_out
timeout 300 python tools/check_pair_kernels.py > gpurun_out/r02_pair_check.log 2>&1; prc=$?; tail -3 gpurun_out/r02_pair_check.log
if [ $prc -ne 0 ]; then echo "PAIR KERNEL CHECK FAILED -> stopping"; exit 1; fi
: > gpurun_out/r02_dist_variants.jsonl
timeout 120 python tools/bench_dist.py >> gpurun_out/r02_dist_variants.jsonl 2>gpurun_out/r02_dist_variants.err
IBL_DIST_BN=256 timeout 120 python tools/bench_dist.py >> gpurun_out/r02_dist_variants.jsonl 2>>gpurun_out/r02_dist_variants.err
IBL_DIST_SCREEN=3 timeout 120 python tools/bench_dist.py >> gpurun_out/r02_dist_variants.jsonl 2>>gpurun_out/r02_dist_variants.err
timeout 120 python tools/bench_dist.py 6800 31250 4096 10 >> gpurun_out/r02_dist_variants.jsonl 2>>gpurun_out/r02_dist_variants.err
IBL_DIST_BN=256 timeout 120 python tools/bench_dist.py 6800 31250 4096 10 >> gpurun_out/r02_dist_variants.jsonl 2>>gpurun_out/r02_dist_variants.err
timeout 180 python tools/bench_dist.py 6800 250000 4096 10 >> gpurun_out/r02_dist_variants.jsonl 2>>gpurun_out/r02_dist_variants.err
cat gpurun_out/r02_dist_variants.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -q -p no:cacheprovider -k "retrieval or topk or single_pass or DIST" > gpurun_out/r02_tests_s6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_tests_s6.log
timeout 240 tools/gpu_profile.sh full gemm2_f16_top16 dist_f16_v3 1 1
IBL_DIST_BN=256 timeout 240 tools/gpu_profile.sh full gemm2_f16_top16 dist_f16_v3_bn256 1 1
