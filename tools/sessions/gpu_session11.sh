#!/bin/bash
# session 11: gallery diagnostics (is the strong_250k gallery discriminative?), short
mkdir -p gpurun_out
O=gpurun_out/r02_diag_gallery.jsonl; : > $O
timeout 150 python tools/diag_gallery.py 30000 2000 0 >> $O 2>gpurun_out/r02_diag_gallery.err
timeout 150 python tools/diag_gallery.py 30000 2000 1 >> $O 2>>gpurun_out/r02_diag_gallery.err
IBL_GALLERY_NOISE=0.1 timeout 150 python tools/diag_gallery.py 30000 2000 1 >> $O 2>>gpurun_out/r02_diag_gallery.err
cut -c1-900 $O; tail -3 gpurun_out/r02_diag_gallery.err
