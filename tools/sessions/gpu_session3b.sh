#!/bin/bash
# guarded re-run after a lost box: fused-conv1 check first (short timeouts, small then full size), then the suite, then bench.
# No ncu in this call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.used,memory.total --format=csv > gpurun_out/r02_gpu3b.txt; free -g >> gpurun_out/r02_gpu3b.txt
# the SM-pair kernels changed their barrier protocol: check them first under a short timeout, stop the session if they fail
timeout 300 python tools/check_pair_kernels.py > gpurun_out/r02_pair_check.log 2>&1; prc=$?
tail -5 gpurun_out/r02_pair_check.log
if [ $prc -ne 0 ]; then echo "PAIR KERNEL CHECK FAILED (rc=$prc) -> stopping"; exit 1; fi
IBL_DIST_SCREEN=3 timeout 90 python tools/check_pair_kernels.py > gpurun_out/r02_pair_check3.log 2>&1; prc=$?
tail -2 gpurun_out/r02_pair_check3.log
if [ $prc -ne 0 ]; then echo "PAIR KERNEL CHECK (bf16x3 distance) FAILED (rc=$prc) -> stopping"; exit 1; fi
timeout 100 python tools/check_fused_conv1.py gpurun_out/f1.pt --time > gpurun_out/r02_fused_on.log 2>&1; echo "fused-on rc=$?"
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
IBL_CONV1_FUSED=0 timeout 100 python tools/check_fused_conv1.py gpurun_out/f0.pt --time > gpurun_out/r02_fused_off.log 2>&1; echo "fused-off rc=$?"
tail -6 gpurun_out/r02_fused_on.log; tail -6 gpurun_out/r02_fused_off.log
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
if grep -q "layer 13" gpurun_out/r02_fused_on.log; then echo FUSED_OK; else export IBL_CONV1_FUSED=0; echo "FUSED KERNEL FAILED -> running the rest with IBL_CONV1_FUSED=0"; fi
timeout 1800 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/r02_tests_s3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests_s3.log
tail -15 gpurun_out/r02_tests_s3.log
free -g | head -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-strong > gpurun_out/r02_bench_s3.json 2> gpurun_out/r02_bench_s3.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_s3.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'retrieval ms', d['retrieval']['ms'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['ms_per_launch_group'])
except Exception as e: print('bench parse failed', e)
PY
