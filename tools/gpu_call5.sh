#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
timeout 300 python tools/tc_bias.py > $O/tc_bias.log 2>&1
timeout 400 python tools/tc_diag.py c1x1_k32_n16 c1x1_k32_n32 c1x1_k32_n64 c1x1_k32_n128 c3x3_64_64 c3x3_16_32_s2 c3x3_512_512 root_448_128 dcn_64_64 dcn_512_256 dcn_128_128 > $O/diag.log 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -s > $O/pytest_r2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_frame.py -x -q -m gpu > $O/pytest_frame.log 2>&1
timeout 900 python bench.py --steps 100 --warmup 10 > $O/bench_c2.json 2> $O/bench_c2.err
DEFT_TC_RZ_COMP=0 timeout 900 python bench.py --steps 50 --warmup 10 > $O/bench_c2_norz.json 2> $O/bench_c2_norz.err
timeout 600 python bench.py --config c4 --steps 150 --warmup 10 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
timeout 600 python bench.py --config c4 --streams 3 --steps 60 --warmup 10 --no-cpu-baseline > $O/bench_c4_3streams.json 2> $O/bench_c4_3streams.err
timeout 600 python bench.py --config c3 --steps 150 --warmup 10 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
# ---- 3xFP16 mode (opt-in): numerics vs the SIMT engine, per-layer time, frame rate + parity
DEFT_TC_F16=1 timeout 400 python tools/tc_diag.py c1x1_k256_n256 c3x3_64_64 c3x3_128_128_s2 c3x3_512_512 root_448_128 dcn_64_64 dcn_512_256 dcn_128_128 big_c3x3_64_64 big_dcn_64_64 c3x3_256_256_l4 > $O/diag_f16.log 2>&1
DEFT_TC_F16=1 timeout 200 python tools/tc_perf.py > $O/perf_f16.log 2>&1
timeout 200 python tools/tc_perf.py > $O/perf_tf32.log 2>&1
DEFT_TC_F16=1 timeout 900 python bench.py --steps 100 --warmup 10 > $O/bench_c2_f16.json 2> $O/bench_c2_f16.err
DEFT_TC_F16=1 timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops_f16.log 2>&1
echo DONE
