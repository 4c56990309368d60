#!/bin/bash
# DCN gather out of a halo patch (TsDcnPatchAP): parity, per-layer time against the global-gather kernel, frame rate
O=gpurun_out/r2k; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "halo_patch" > $O/pytest_patch.log 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "dcn" > $O/pytest_ops_dcn.log 2>&1
DEFT_DCN_PATCH=0 timeout 200 python tools/tc_perf.py > $O/perf_gather.log 2>&1
timeout 200 python tools/tc_perf.py > $O/perf_patch.log 2>&1
DEFT_DCN_MAXBN=64 timeout 200 python tools/tc_perf.py > $O/perf_patch_bn64.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
DEFT_DCN_PATCH=0 timeout 900 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_c2_gather.json 2> $O/bench_c2_gather.err
echo DONE
