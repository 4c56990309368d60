#!/bin/bash
# issuer waits on the A stage and the weight stage together (mbar_wait2): parity, traces, frame rate
O=gpurun_out/r2q; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops.log 2>&1
DEFT_PDL=0 timeout 120 python tools/tc_trace.py conv32 > $O/trace_conv32.log 2>&1
DEFT_PDL=0 timeout 120 python tools/tc_trace.py conv64 > $O/trace_conv64.log 2>&1
timeout 200 python tools/tc_perf.py > $O/perf.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 900 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_c2_b.json 2> $O/bench_c2_b.err
echo DONE
