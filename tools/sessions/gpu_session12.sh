#!/bin/bash
# session 12 (final 1-GPU pass of round 2): full GPU suite, default bench line (with strong_250k at N = 1), launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_tests_s12.log 2>&1; echo "full pytest rc=$?"; tail -3 gpurun_out/r02_tests_s12.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_s12.json 2>gpurun_out/r02_bench_s12.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_s12.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['mma_issue_frac'], d['e2e']['value'], d['retrieval']['ms'], d['clocks']); print(d.get('strong_250k'))"
timeout 300 tools/gpu_profile.sh launches r02_launches_s12
