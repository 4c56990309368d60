#!/bin/bash
# accumulate chains at N <= 32 (TsCfg::CH): parity, per-layer time, frame rate
O=gpurun_out/r2p; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops.log 2>&1
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "not full_size and not second_gpu" > $O/pytest_r2.log 2>&1
timeout 200 python tools/tc_perf.py > $O/perf.log 2>&1
DEFT_PDL=0 timeout 120 python tools/tc_trace.py conv32 > $O/trace_conv32.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
echo DONE
