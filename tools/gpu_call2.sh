#!/bin/bash
# GPU call 2 (round 2): engine variants (single issuer N=128, pair-MMA N=32), new rows a9/f1/f2, full-size parity, bench, ncu.
O=gpurun_out/r2b; mkdir -p $O
timeout 400 python tools/tc_diag.py c1x1_k32_n32 c1x1_k32_n128 c1x1_k256_n256 c3x3_64_64 c3x3_512_512 root_448_128 dcn_64_64 dcn_512_256 dcn_128_128 big_dcn_64_64 big_c3x3_64_64 c3x3_256_256_l4 > $O/diag.log 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops.log 2>&1
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -s > $O/pytest_r2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_frame.py -x -q -m gpu > $O/pytest_frame.log 2>&1
for c in conv32 conv128 dcn64; do timeout 120 python tools/tc_trace.py $c > $O/trace_$c.log 2>&1; done
timeout 200 python tools/tc_perf.py > $O/perf_new.log 2>&1
DEFT_B200_LIB=$PWD/deft_b200/libdeft_b200_alt.so timeout 200 python tools/tc_perf.py > $O/perf_alt.log 2>&1
timeout 600 python bench.py --steps 100 --warmup 10 > $O/bench.json 2> $O/bench.err
DEFT_B200_LIB=$PWD/deft_b200/libdeft_b200_alt.so timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_alt.json 2> $O/bench_alt.err
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --fp32-input --no-tracks > $O/bench_r1cfg.json 2> $O/bench_r1cfg.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_ts_kernel --launch-skip 6 --launch-count 2 -o $O/ncu_dcn64 -f python tools/one_layer.py dcn64 4 > $O/ncu_dcn64.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_ts_kernel --launch-skip 3 --launch-count 1 -o $O/ncu_conv128 -f python tools/one_layer.py conv128 4 > $O/ncu_conv128.log 2>&1
tail -3 $O/pytest_ops.log $O/pytest_r2.log $O/pytest_frame.log; cat $O/diag.log | cut -c1-150; head -c 1500 $O/bench.json; tail -5 $O/bench.err
