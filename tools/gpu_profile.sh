#!/bin/bash
# One profiling entry point for the GPU box (B200_PROFILING.md).  Run under gpurun from the repo root.
#   tools/gpu_profile.sh launches [tag]                 ncu launch list (gpu__time_duration) of `bench.py --steps 2`
#   tools/gpu_profile.sh full <kernel-regex> <tag> [skip] [count] [-- cmd...]
#                                                       ncu --set full capture of matching kernels; default cmd = bench.py
# Outputs land in gpurun_out/ (merged back by gpurun); summarise with tools/ncu_summary.py into profiles/.
set -u
mkdir -p gpurun_out
mode=${1:-launches}; shift || true
case "$mode" in
  launches)
    tag=${1:-launches}
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager --no-strong > gpurun_out/${tag}_bench.log 2>&1
    ;;
  full)
    regex=$1; tag=$2; skip=${3:-0}; count=${4:-2}
    shift 4 2>/dev/null || shift $#
    if [ "${1:-}" = "--" ]; then shift; fi
    if [ $# -eq 0 ]; then set -- python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-strong; fi
    ncu --set full --clock-control none --import-source on -k regex:${regex} -s ${skip} -c ${count} -f \
        -o gpurun_out/prof_${tag} "$@" > gpurun_out/ncu_${tag}.log 2>&1
    ;;
  *) echo "unknown mode $mode"; exit 2;;
esac
