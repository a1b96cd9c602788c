#!/bin/bash
# round-2 session 3: fused conv1 check (guarded by a short timeout), full GPU suite, bench, ncu captures
mkdir -p gpurun_out
timeout 120 python tools/check_fused_conv1.py gpurun_out/f1.pt --time > gpurun_out/r02_fused_on.log 2>&1; echo "fused-on rc=$?"
IBL_CONV1_FUSED=0 timeout 120 python tools/check_fused_conv1.py gpurun_out/f0.pt --time > gpurun_out/r02_fused_off.log 2>&1; echo "fused-off rc=$?"
cat gpurun_out/r02_fused_on.log gpurun_out/r02_fused_off.log | tail -14
python - <<'PY'
import torch
try:
    a, b = torch.load('gpurun_out/f1.pt'), torch.load('gpurun_out/f0.pt')
    for x, y in zip(a, b):
        print('fused vs separate rel-L2', float((x - y).norm() / y.norm()), 'equal', bool(torch.equal(x, y)))
except Exception as e:
    print('compare failed', e)
PY
rm -f gpurun_out/f1.pt gpurun_out/f0.pt
if grep -q "layer 13" gpurun_out/r02_fused_on.log; then export FUSED_OK=1; else export IBL_CONV1_FUSED=0; echo "FUSED KERNEL FAILED -> running the rest with IBL_CONV1_FUSED=0"; fi
timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/r02_tests_s3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests_s3.log
tail -15 gpurun_out/r02_tests_s3.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-strong > gpurun_out/r02_bench_s3.json 2> gpurun_out/r02_bench_s3.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_s3.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'retrieval ms', d['retrieval']['ms'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['ms_per_launch_group'])
except Exception as e: print('bench parse failed', e)
PY
timeout 600 tools/gpu_profile.sh launches r02_launches_s3
timeout 600 tools/gpu_profile.sh full gemm2_f16_top16 dist_f16 1 1
timeout 600 tools/gpu_profile.sh full conv1_fused conv1_fused 2 1
timeout 600 tools/gpu_profile.sh full netvlad_tc netvlad_1k 2 1
ls -la gpurun_out/*.ncu-rep | tail -5
