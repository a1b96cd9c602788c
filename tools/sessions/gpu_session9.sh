#!/bin/bash
# session 9: uniform-register MMA issue (conv3x3 / fused conv1 / distance) -- parity first, then per-layer times, bench, launch list
mkdir -p gpurun_out
timeout 300 python tools/check_pair_kernels.py > gpurun_out/r02_pair_check9.log 2>&1; prc=$?; tail -3 gpurun_out/r02_pair_check9.log
frc=0
if [ $prc -ne 0 ] || [ $frc -ne 0 ]; then echo "KERNEL CHECK FAILED ($prc $frc) -> stopping"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_tests_s9.log 2>&1; echo "full pytest rc=$?"; tail -3 gpurun_out/r02_tests_s9.log
timeout 300 python tools/bench_layers.py > gpurun_out/r02_bench_layers_s9.txt 2>&1; tail -30 gpurun_out/r02_bench_layers_s9.txt
timeout 120 python tools/bench_dist.py > gpurun_out/r02_dist_variants_s9.jsonl 2>gpurun_out/r02_dist_variants_s9.err
timeout 180 python tools/bench_dist.py 6800 250000 4096 10 >> gpurun_out/r02_dist_variants_s9.jsonl 2>>gpurun_out/r02_dist_variants_s9.err
cat gpurun_out/r02_dist_variants_s9.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 --no-strong > gpurun_out/r02_bench_s9.json 2>gpurun_out/r02_bench_s9.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_s9.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['mma_issue_frac'], d['e2e']['value'], d['retrieval']['ms'], d['clocks'])"
timeout 300 tools/gpu_profile.sh launches r02_launches_s9
timeout 240 tools/gpu_profile.sh full conv1_fused conv1_fused_v2 1 1
