#!/bin/bash
# 16-channel halo-patch producer (TsPatch16AP) for DLA level0: parity, frame parity, frame rate A/B on one box
O=gpurun_out/r2s; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops.log 2>&1
timeout 300 python tools/tc_diag.py c3x3_16_16 > $O/diag.log 2>&1
B="python bench.py --steps 200 --warmup 20"
timeout 900 $B > $O/bench_c2.json 2> $O/bench_c2.err
DEFT_TC_PATCH16=0 timeout 600 $B --no-cpu-baseline > $O/bench_c2_nopatch16.json 2> $O/bench_c2_nopatch16.err
timeout 600 $B --no-cpu-baseline > $O/bench_c2_b.json 2> $O/bench_c2_b.err
echo DONE
