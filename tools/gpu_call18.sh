#!/bin/bash
# A/B on one box: default build, no progress board (DEFT_TS_PROGRESS=0), accumulate chains at N <= 32 (DEFT_TS_CHAINS=1)
O=gpurun_out/r2r; mkdir -p $O
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline"
timeout 600 $B > $O/bench_default_1.json 2> $O/bench_default_1.err
DEFT_B200_LIB=$PWD/build_alt/libdeft_noboard.so timeout 600 $B > $O/bench_noboard.json 2> $O/bench_noboard.err
DEFT_B200_LIB=$PWD/build_alt/libdeft_chains.so timeout 600 $B > $O/bench_chains.json 2> $O/bench_chains.err
timeout 600 $B > $O/bench_default_2.json 2> $O/bench_default_2.err
DEFT_B200_LIB=$PWD/build_alt/libdeft_noboard.so timeout 600 $B > $O/bench_noboard_2.json 2> $O/bench_noboard_2.err
DEFT_B200_LIB=$PWD/build_alt/libdeft_chains.so timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops_chains.log 2>&1
echo DONE
