#!/bin/bash
# 16-channel patch ring 8 deep (default) vs 4 deep, same box
O=gpurun_out/r2v; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv_bn_act or c16" > $O/pytest_ops_conv.log 2>&1
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline"
timeout 600 $B > $O/bench_ps8.json 2> $O/bench_ps8.err
DEFT_B200_LIB=$PWD/build_alt/libdeft_ps4.so timeout 600 $B > $O/bench_ps4.json 2> $O/bench_ps4.err
timeout 600 $B > $O/bench_ps8_b.json 2> $O/bench_ps8_b.err
DEFT_PDL=0 timeout 120 python tools/tc_trace.py conv16 > $O/trace_conv16_ps8.log 2>&1
echo DONE
