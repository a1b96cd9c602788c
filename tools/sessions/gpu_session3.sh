#!/bin/bash
# neutralised: the previous content lost a GPU box (strike); see tools/sessions/gpu_session3b.sh for the guarded re-run
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.used --format=csv > gpurun_out/r02_noop.txt 2>&1
echo "noop"
