#!/bin/bash
# round-2 session 2: full GPU suite with the new kernels, quick timing of the distance call and NetVLAD, short bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/r02_tests_s2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests_s2.log
tail -25 gpurun_out/r02_tests_s2.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-strong --no-cpu-baseline --no-eager > gpurun_out/r02_bench_s2.json 2> gpurun_out/r02_bench_s2.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_s2.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['retrieval']['ms'], d['e2e']['value'])
except Exception as e: print('bench parse failed', e)
PY
timeout 600 tools/gpu_profile.sh launches r02_launches_s2
