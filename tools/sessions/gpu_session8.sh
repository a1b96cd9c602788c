#!/bin/bash
# session 8: pending-list distance epilogue -- variants, retrieval tests first (stop on failure), ncu, then the full suite + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -q -p no:cacheprovider -k "retrieval or topk or single_pass or DIST or guard or fallback" > gpurun_out/r02_tests_s8_retr.log 2>&1; rc=$?; echo "retrieval pytest rc=$rc"; tail -3 gpurun_out/r02_tests_s8_retr.log
if [ $rc -ne 0 ]; then echo "RETRIEVAL TESTS FAILED -> stopping"; exit 1; fi
J=gpurun_out/r02_dist_variants_s8.jsonl; E=gpurun_out/r02_dist_variants_s8.err; : > $J; : > $E
timeout 120 python tools/bench_dist.py >> $J 2>>$E
IBL_DIST_BN=256 timeout 120 python tools/bench_dist.py >> $J 2>>$E
timeout 120 python tools/bench_dist.py 6800 31250 4096 10 >> $J 2>>$E
IBL_DIST_BN=256 timeout 120 python tools/bench_dist.py 6800 31250 4096 10 >> $J 2>>$E
timeout 180 python tools/bench_dist.py 6800 250000 4096 10 >> $J 2>>$E
IBL_DIST_BN=256 timeout 180 python tools/bench_dist.py 6800 250000 4096 10 >> $J 2>>$E
cat $J
IBL_DIST_BN=256 timeout 240 tools/gpu_profile.sh full gemm2_f16_top16 dist_f16_v4_bn256 1 1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_tests_s8.log 2>&1; echo "full pytest rc=$?"; tail -3 gpurun_out/r02_tests_s8.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-strong > gpurun_out/r02_bench_s8.json 2>gpurun_out/r02_bench_s8.err; echo "bench rc=$?"; cat gpurun_out/r02_bench_s8.json
