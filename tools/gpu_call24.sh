#!/bin/bash
# tensor peaks + a last knob sweep on the final build (same box)
O=gpurun_out/r2w; mkdir -p $O
timeout 300 python tools/measure_tensor_peaks.py > $O/tensor_peaks.json 2> $O/tensor_peaks.err
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline"
timeout 300 $B > $O/bench_default.json 2> $O/bench_default.err
DEFT_PDL=1 timeout 300 $B > $O/bench_pdl1.json 2> $O/bench_pdl1.err
timeout 300 $B --depth 5 --dense-streams 4 > $O/bench_d5s4.json 2> $O/bench_d5s4.err
timeout 300 $B --depth 6 --dense-streams 5 > $O/bench_d6s5.json 2> $O/bench_d6s5.err
DEFT_TC_WIDE=0 timeout 300 $B > $O/bench_wide0.json 2> $O/bench_wide0.err
DEFT_TC_SPLIT_TARGET=50 timeout 300 $B > $O/bench_split50.json 2> $O/bench_split50.err
timeout 300 $B > $O/bench_default_b.json 2> $O/bench_default_b.err
echo DONE
