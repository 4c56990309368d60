#!/bin/bash
# GPU call 1 (round 2): correctness of the DCN table rewrite / BN=128 policy, protocol traces, DCN knob sweep, quick bench.
mkdir -p gpurun_out/r2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 300 python tools/tc_diag.py dcn_64_64 dcn_512_256 dcn_128_128 big_dcn_64_64 c3x3_64_64 > $O/diag1.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops1.log 2>&1
for c in conv64 conv128 dcn64 dcn128; do timeout 120 python tools/tc_trace.py $c > $O/trace_$c.log 2>&1; done
for sb in 2 3 4 8; do DEFT_DCN_SB=$sb timeout 200 python tools/tc_perf.py > $O/perf_sb$sb.log 2>&1; done
DEFT_DCN_MAXBN=64 timeout 200 python tools/tc_perf.py > $O/perf_bn64.log 2>&1
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_a.json 2> $O/bench_a.err
DEFT_DCN_MAXBN=64 DEFT_DCN_SB=8 timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_r1like.json 2> $O/bench_r1like.err
tail -3 $O/pytest_ops1.log; cat $O/diag1.log; grep -h "dcn" $O/perf_*.log; cat $O/bench_a.json | head -c 600
