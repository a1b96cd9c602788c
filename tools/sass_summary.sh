#!/bin/bash
# Opcode histogram of the product library: the tracked evidence that the hot path is tcgen05 / TMEM / TMA code.
# R2UR.BROADCAST / BRA.U.ANY count the per-instruction issue loops ptxas builds around tcgen05 / TMA instructions that
# are not behind elect.sync ("The issue path"): 0 in the conv, distance, NetVLAD, dense-GEMM MMA and wgrad kernels.
#   tools/sass_summary.sh > profiles/rNN_sass_summary.txt
LIB=${1:-openibl_b200/lib/libiblb200.so}
echo "# cuobjdump -sass $LIB  (sm_100a)  -- Blackwell-specific opcodes per kernel and in total"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3 }
  { for (i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|UTCQMMA|UTCOMMA|UTMALDG|UTMASTG|UTMAPF|LDTM|STTM|UTCBAR|UTCCP|STAS|SYNCS|UBLKCP|UTCATOMSWS|REDAS|R2UR.BROADCAST|BRA.U.ANY|ELECT|REDUX)/) { op=$i; sub(/;$/,"",op); tot[op]++; per[fn" "op]++ } }
  END {
    print "## total"; for (o in tot) printf "%8d  %s\n", tot[o], o | "sort -k2"; close("sort -k2");
    print "## per kernel"; for (k in per) printf "%8d  %s\n", per[k], k | "sort -k2"; close("sort -k2");
  }'
