#!/bin/bash
# 8-GPU call (charged 8x): BASELINE configs[3] at N = 8 and configs[4] (SFRS step under 8 x DDP); keep it short.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r02_8gpu_devices.txt
timeout 300 $TR --master-port 29811 tools/bench_gallery.py --n-db 250000 --n-q 6800 --no-exact > gpurun_out/r02_gallery250k_8gpu.log 2>&1
echo "gallery rc=$?"; grep '^{' gpurun_out/r02_gallery250k_8gpu.log | tail -1 > gpurun_out/r02_gallery250k_8gpu.json; cat gpurun_out/r02_gallery250k_8gpu.json | cut -c1-600
timeout 300 $TR --master-port 29812 examples/sfrs_step_synthetic.py --launcher pytorch --tuple-size 4 --neg-num 10 --diff-num 10 \
    --height 480 --width 640 --steps 2 > gpurun_out/r02_sfrs_step_8gpu.log 2>&1
echo "sfrs rc=$?"; grep SFRS_STEP gpurun_out/r02_sfrs_step_8gpu.log | tail -1
timeout 600 python -m pytest tests/test_gpu_e2e_api.py tests/test_gpu_dropin_reference_script.py -q -p no:cacheprovider > gpurun_out/r02_tests_8gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_tests_8gpu.log
