#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
DEFT_DEBUG_ENGINE=1 timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "halo_patch" > $O/pytest_patch.log 2>&1
DEFT_DEBUG_ENGINE=1 timeout 200 python tools/tc_perf.py > $O/perf_patch.log 2>&1
echo DONE
