#!/bin/bash
# re-verification after the split-K target change (default 50): every GPU test, smoke, the three configs
O=gpurun_out/r2y; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_round2.py -x -q -m gpu > $O/pytest_r2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_frame.py -x -q -m gpu > $O/pytest_frame.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 600 python bench.py --config c3 --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 python bench.py --config c4 --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --depth 1 > $O/bench_d1.json 2> $O/bench_d1.err
echo DONE
