#!/bin/bash
# short session: pair check, distance-call timing per variant, fused-conv1 A/B bench, retrieval tests, ncu of the dist kernel
mkdir -p gpurun_out
timeout 300 python tools/check_pair_kernels.py > gpurun_out/r02_pair_check.log 2>&1; prc=$?; tail -4 gpurun_out/r02_pair_check.log
if [ $prc -ne 0 ]; then echo "PAIR KERNEL CHECK FAILED -> stopping"; exit 1; fi
: > gpurun_out/r02_dist_variants.jsonl
timeout 120 python tools/bench_dist.py >> gpurun_out/r02_dist_variants.jsonl 2>gpurun_out/r02_dist_variants.err
IBL_DIST_BN=256 timeout 120 python tools/bench_dist.py >> gpurun_out/r02_dist_variants.jsonl 2>>gpurun_out/r02_dist_variants.err
IBL_DIST_SCREEN=3 timeout 120 python tools/bench_dist.py >> gpurun_out/r02_dist_variants.jsonl 2>>gpurun_out/r02_dist_variants.err
timeout 120 python tools/bench_dist.py 6800 31250 4096 10 >> gpurun_out/r02_dist_variants.jsonl 2>>gpurun_out/r02_dist_variants.err
cat gpurun_out/r02_dist_variants.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "retrieval or topk or single_pass or netvlad or pipelin or tokyo" > gpurun_out/r02_tests_s5.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_tests_s5.log
for f in 1 0; do
  IBL_CONV1_FUSED=$f timeout 300 python bench.py --steps 30 --warmup 5 --no-strong --no-eager --no-cpu-baseline > gpurun_out/r02_bench_fused$f.json 2>gpurun_out/r02_bench_fused$f.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_fused$f.json'))
    print('fused=$f', {k:d[k] for k in ('value','ms_per_step')}, 'retrieval ms', d['retrieval']['ms'], 'e2e', d['e2e']['value'], d['e2e'].get('blocking_call_value'), 'roofline', round(d['roofline']['frac'],4), round(d['roofline']['ms_per_launch_group'],3), d['clocks'])
except Exception as e: print('bench parse failed', e)
PY
done
timeout 240 tools/gpu_profile.sh full gemm2_f16_top16 dist_f16_v2 1 1
timeout 200 tools/gpu_profile.sh launches r02_launches_s5
