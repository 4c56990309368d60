#!/bin/bash
# DCN gather: row-per-thread table build (one ahead, by the epilogue warps) and 16x8 patch tiles (DEFT_DCN_TILE2D)
O=gpurun_out/r2o; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "halo_patch or odd_kblock" > $O/pytest_patch.log 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "dcn" > $O/pytest_ops_dcn.log 2>&1
DEFT_DCN_TILE2D=0 timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "halo_patch" > $O/pytest_patch_linear.log 2>&1
DEFT_PDL=0 timeout 120 python tools/tc_trace.py dcn64 > $O/trace_dcn64_tile2d.log 2>&1
DEFT_DCN_TILE2D=0 DEFT_PDL=0 timeout 120 python tools/tc_trace.py dcn64 > $O/trace_dcn64_linear.log 2>&1
timeout 200 python tools/tc_perf.py > $O/perf_tile2d.log 2>&1
DEFT_DCN_TILE2D=0 timeout 200 python tools/tc_perf.py > $O/perf_linear.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_c2_tile2d.json 2> $O/bench_c2_tile2d.err
DEFT_DCN_TILE2D=0 timeout 900 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_c2_linear.json 2> $O/bench_c2_linear.err
echo DONE
