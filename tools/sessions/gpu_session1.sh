#!/bin/bash
# round-2 session 1: full GPU test suite, default bench line, launch list, per-layer conv DRAM traffic
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02_tests_s1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests_s1.log
tail -5 gpurun_out/r02_tests_s1.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_s1.json 2> gpurun_out/r02_bench_s1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_s1.json
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'conv3x3_tc|conv1_1_tc' --csv --log-file gpurun_out/r02_conv_traffic.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-strong > gpurun_out/r02_conv_traffic.log 2>&1
timeout 600 tools/gpu_profile.sh launches r02_launches_s1
