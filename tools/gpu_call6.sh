#!/bin/bash
# default build = 3xFP16 where eligible (DEFT_TC_F16 default 1), split b_full barriers, global issuer ownership
O=gpurun_out/r2f; mkdir -p $O
timeout 300 python tools/tc_diag.py c1x1_k32_n32 c1x1_k256_n256 c3x3_64_64 c3x3_512_512 root_448_128 dcn_64_64 dcn_512_256 dcn_128_128 heads_mot heads_nusc big_heads_mot > $O/diag.log 2>&1
timeout 600 python bench.py --config c4 --streams 3 --steps 60 --warmup 10 --no-cpu-baseline > $O/bench_c4_3streams.json 2> $O/bench_c4_3streams.err
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -s > $O/pytest_r2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_frame.py -x -q -m gpu > $O/pytest_frame.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
DEFT_TC_F16=0 timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_c2_tf32.json 2> $O/bench_c2_tf32.err
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --assign > $O/bench_c2_assign.json 2> $O/bench_c2_assign.err
timeout 600 python bench.py --config c4 --steps 150 --warmup 10 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
timeout 600 python bench.py --config c3 --steps 150 --warmup 10 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 200 python tools/tc_perf.py > $O/perf.log 2>&1
echo DONE
