#!/bin/bash
# final verification of the round-2 build: every GPU test, smoke, the three configs, launch list, sanitizer on the new kernels
O=gpurun_out/r2t; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_round2.py -x -q -m gpu > $O/pytest_r2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_frame.py -x -q -m gpu > $O/pytest_frame.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 600 python bench.py --config c3 --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 python bench.py --config c4 --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --depth 1 > $O/bench_d1.json 2> $O/bench_d1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_r2_final.csv python bench.py --steps 2 --warmup 1 --no-graph --depth 1 --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --report-api-errors no python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv_bn_act and 16-16" > $O/sanitizer_patch16_memcheck.log 2>&1
echo DONE
