#!/bin/bash
# split-K target sweep on the final build (same box): CTAs a split-K layer should reach
O=gpurun_out/r2x; mkdir -p $O
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline"
for t in 74 1 24 37 50 60 74; do DEFT_TC_SPLIT_TARGET=$t timeout 300 $B > $O/bench_split${t}_$RANDOM.json 2> $O/err_$t.log; done
DEFT_TC_SPLIT_TARGET=50 timeout 300 $B --depth 1 > $O/bench_d1_split50.json 2> $O/err_d1_50.log
timeout 300 $B --depth 1 > $O/bench_d1_split74.json 2> $O/err_d1_74.log
echo DONE
