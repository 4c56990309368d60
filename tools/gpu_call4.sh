#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
timeout 300 python tools/tc_bias.py > $O/tc_bias.log 2>&1
timeout 200 python tools/exp_f16ts.py > $O/exp_f16ts.log 2>&1
timeout 600 python tools/stress_affinity.py 400 > $O/stress_aff.log 2>&1
timeout 600 python bench.py --config c4 --steps 150 --warmup 10 --no-cpu-baseline > $O/bench_c4_a.json 2> $O/bench_c4_a.err
timeout 600 python bench.py --config c4 --steps 150 --warmup 10 --no-cpu-baseline > $O/bench_c4_b.json 2> $O/bench_c4_b.err
cat $O/tc_bias.log; cat $O/exp_f16ts.log; tail -3 $O/stress_aff.log; head -c 400 $O/bench_c4_a.json; head -c 400 $O/bench_c4_b.json; grep -h "roles\|stall" $O/*.json $O/*.err $O/*.log | head -20
