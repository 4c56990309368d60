#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python tools/parity_c2.py > $O/parity_c2.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_round2.py -q -m gpu -s > $O/pytest_r2.log 2>&1
timeout 900 python bench.py --steps 100 --warmup 10 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 400 python bench.py --steps 100 --warmup 10 --assign --no-cpu-baseline > $O/bench_c2_assign.json 2> $O/bench_c2_assign.err
timeout 400 python bench.py --config c3 --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 python bench.py --config c4 --steps 60 --warmup 10 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
timeout 1500 bash tools/sanitize.sh $O/sanitize > $O/sanitize.out 2>&1
grep -E "passed|failed|error" $O/pytest_r2.log | tail -5; cat $O/parity_c2.log | tail -20; head -c 700 $O/bench_c2.json; tail -3 $O/bench_c2.err; cat $O/sanitize/summary.txt | tail -30
