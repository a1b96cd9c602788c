#!/bin/bash
# 8-GPU call (charged 8x): bench.py at N = 8 (weak-scaling line + BASELINE configs[3] strong_250k leg), configs[4]
# (SFRS step under 8 x DDP), the world>1 API tests.  Keep it short.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r02_8gpu_devices.txt
timeout 420 $TR --master-port 29811 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu.json 2>gpurun_out/r02_bench_8gpu.err
echo "bench8 rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_8gpu.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, 'e2e', d.get('e2e',{}).get('value'), 'strong', d.get('strong_250k'))
except Exception as e:
    print('bench8 parse failed', e)
PY
timeout 300 $TR --master-port 29812 examples/sfrs_step_synthetic.py --launcher pytorch --tuple-size 4 --neg-num 10 --diff-num 10 \
    --height 480 --width 640 --steps 2 > gpurun_out/r02_sfrs_step_8gpu.log 2>&1
echo "sfrs rc=$?"; grep SFRS_STEP gpurun_out/r02_sfrs_step_8gpu.log | tail -1
timeout 600 python -m pytest tests/test_gpu_e2e_api.py tests/test_gpu_dropin_reference_script.py tests/test_gpu_train.py -q -p no:cacheprovider > gpurun_out/r02_tests_8gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_tests_8gpu.log
