#!/bin/bash
# second 8-GPU call (charged 8x): BASELINE configs[3] with the centred PCA bias (discriminative descriptors), fp64 subset check on
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29821 tools/bench_gallery.py --n-db 250000 --n-q 6800 > gpurun_out/r02_gallery250k_8gpu.log 2>&1
echo "gallery rc=$?"; grep '^{' gpurun_out/r02_gallery250k_8gpu.log | tail -1 > gpurun_out/r02_gallery250k_8gpu.json; cut -c1-1200 gpurun_out/r02_gallery250k_8gpu.json
