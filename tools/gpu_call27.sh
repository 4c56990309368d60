#!/bin/bash
# split-K target by kind (convs 50, DCN gather 74): parity + the benchmark line
O=gpurun_out/r2z; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 600 python -m pytest tests/test_gpu_frame.py -x -q -m gpu -k "depths or fallback or graph" > $O/pytest_frame_subset.log 2>&1
echo DONE
