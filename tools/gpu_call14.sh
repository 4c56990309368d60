#!/bin/bash
# DCN gather with corner tables built one tile ahead by the epilogue warps (no tile barrier): parity, trace, per-layer time, frame rate
O=gpurun_out/r2n; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "halo_patch or odd_kblock" > $O/pytest_patch.log 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "dcn" > $O/pytest_ops_dcn.log 2>&1
DEFT_PDL=0 timeout 120 python tools/tc_trace.py dcn64 > $O/trace_dcn64.log 2>&1
DEFT_PDL=0 timeout 120 python tools/tc_trace.py dcn128 > $O/trace_dcn128.log 2>&1
timeout 200 python tools/tc_perf.py > $O/perf.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 600 python -m pytest tests/test_gpu_frame.py -x -q -m gpu > $O/pytest_frame.log 2>&1
echo DONE
