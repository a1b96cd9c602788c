#!/bin/bash
# session 10: elect.sync issue blocks (no per-instruction ELECT loops) -- parity, per-layer times, bench, launch list, ncu
mkdir -p gpurun_out
timeout 300 python tools/check_pair_kernels.py > gpurun_out/r02_pair_check10.log 2>&1; prc=$?; tail -3 gpurun_out/r02_pair_check10.log
frc=0
if [ $prc -ne 0 ] || [ $frc -ne 0 ]; then echo "KERNEL CHECK FAILED ($prc $frc) -> stopping"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_tests_s10.log 2>&1; echo "full pytest rc=$?"; tail -3 gpurun_out/r02_tests_s10.log
timeout 300 python tools/bench_layers.py > gpurun_out/r02_bench_layers_s10.txt 2>&1; tail -30 gpurun_out/r02_bench_layers_s10.txt
timeout 120 python tools/bench_dist.py > gpurun_out/r02_dist_variants_s10.jsonl 2>gpurun_out/r02_dist_variants_s10.err
timeout 180 python tools/bench_dist.py 6800 250000 4096 10 >> gpurun_out/r02_dist_variants_s10.jsonl 2>>gpurun_out/r02_dist_variants_s10.err
cat gpurun_out/r02_dist_variants_s10.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 --no-strong > gpurun_out/r02_bench_s10.json 2>gpurun_out/r02_bench_s10.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_s10.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['mma_issue_frac'], d['e2e']['value'], d['retrieval']['ms'], d['clocks'])"
timeout 300 tools/gpu_profile.sh launches r02_launches_s10
timeout 240 tools/gpu_profile.sh full conv1_fused conv1_fused_v3 1 1
timeout 240 tools/gpu_profile.sh full netvlad_tc netvlad_v4_elect 1 1
timeout 240 tools/gpu_profile.sh full "conv3x3_tc_kernel<128" conv128_halo_v2_elect 2 1
