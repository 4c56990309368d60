#!/bin/bash
O=gpurun_out/r2i; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -s > $O/pytest_r2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_frame.py -x -q -m gpu > $O/pytest_frame.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
for v in "d4s2:--depth 4 --dense-streams 2" "d3:--depth 3" "d4s4:--depth 4 --dense-streams 4"; do n=${v%%:*}; a=${v#*:}; timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline $a > $O/bench_$n.json 2> $O/bench_$n.err; done
DEFT_TC_SPLIT_TARGET=148 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_split148.json 2> $O/bench_split148.err
DEFT_TC_SPLIT_TARGET=37 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_split37.json 2> $O/bench_split37.err
DEFT_DCN_SB=2 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_dcnsb2.json 2> $O/bench_dcnsb2.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --depth 1 > $O/bench_d1.json 2> $O/bench_d1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
echo DONE
