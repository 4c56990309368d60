#!/bin/bash
# compute-sanitizer over what changed after the first round-2 sweep (tools/sanitize.sh): the merged issuer waits (every TS kernel and
# the heads kernel), the 16-channel halo-patch producer, the halo-patch DCN gather kernel.   Usage: bash tools/sanitize_final.sh <outdir>
O=${1:-gpurun_out/sanitize_final}; mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
for c in c3x3_64_64 dcn_64_64 dcn_128_128 heads_mot c1x1_k32_n32; do
  timeout 600 $CS --tool racecheck --racecheck-report all --error-exitcode 77 --print-limit 20 python tools/tc_diag.py --one $c > $O/racecheck_$c.log 2>&1; echo "racecheck $c rc=$?" >> $O/summary.txt
  timeout 300 $CS --tool memcheck --report-api-errors no --error-exitcode 77 --print-limit 20 python tools/tc_diag.py --one $c > $O/memcheck_$c.log 2>&1; echo "memcheck $c rc=$?" >> $O/summary.txt
done
K16="conv_bn_act and (16-16-3-1-hw12-tc3x or 16-16-3-1-hw13-tc3x or 16-32-3-1-hw14-tc3x)"
timeout 600 $CS --tool racecheck --racecheck-report all --error-exitcode 77 --print-limit 20 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "$K16" > $O/racecheck_patch16.log 2>&1; echo "racecheck patch16 rc=$?" >> $O/summary.txt
timeout 600 $CS --tool memcheck --report-api-errors no --error-exitcode 77 --print-limit 20 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "$K16" > $O/memcheck_patch16.log 2>&1; echo "memcheck patch16 rc=$?" >> $O/summary.txt
KD="halo_patch and (64-64-24-48 or 64-64-40-46 or 128-128-32-32 or 256-256-16-32)"
timeout 900 $CS --tool racecheck --racecheck-report all --error-exitcode 77 --print-limit 20 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "$KD" > $O/racecheck_dcn_patch.log 2>&1; echo "racecheck dcn_patch rc=$?" >> $O/summary.txt
timeout 900 $CS --tool memcheck --report-api-errors no --error-exitcode 77 --print-limit 20 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "$KD" > $O/memcheck_dcn_patch.log 2>&1; echo "memcheck dcn_patch rc=$?" >> $O/summary.txt
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" $O/*.log | sort | uniq -c >> $O/summary.txt
cat $O/summary.txt
