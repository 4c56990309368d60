#!/bin/bash
# compute-sanitizer memcheck + racecheck over every tcgen05 kernel instantiation on small shapes (GPU box).
# Usage: bash tools/sanitize.sh <outdir>      (logs are summarised into profiles/)
O=${1:-gpurun_out/sanitize}; mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
MEM="c1x1_k32_n16 c1x1_k32_n32 c1x1_k32_n64 c1x1_k32_n128 c1x1_k256_n256 c3x3_64_64 c3x3_16_32_s2 c3x3_128_128_s2 c3x3_512_512 root_448_128 dcn_64_64 dcn_512_256 dcn_128_128 heads_mot"
RACE="c1x1_k32_n32 c1x1_k32_n128 c3x3_64_64 c3x3_512_512 root_448_128 dcn_64_64 dcn_128_128 heads_mot"
for c in $MEM; do
  timeout 300 $CS --tool memcheck --report-api-errors no --error-exitcode 77 --print-limit 20 python tools/tc_diag.py --one $c > $O/memcheck_$c.log 2>&1; echo "memcheck $c rc=$?" >> $O/summary.txt
done
for c in $RACE; do
  timeout 600 $CS --tool racecheck --racecheck-report all --error-exitcode 77 --print-limit 20 python tools/tc_diag.py --one $c > $O/racecheck_$c.log 2>&1; echo "racecheck $c rc=$?" >> $O/summary.txt
done
timeout 900 $CS --tool memcheck --report-api-errors no --error-exitcode 77 --print-limit 20 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "track_similarity_op or afe_forward or uint8 or dcn_backward or pipeline_track" > $O/memcheck_round2_ops.log 2>&1; echo "memcheck round2 ops rc=$?" >> $O/summary.txt
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" $O/*.log | sort | uniq -c >> $O/summary.txt
cat $O/summary.txt
