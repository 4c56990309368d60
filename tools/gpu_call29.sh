#!/bin/bash
O=gpurun_out/r2zz; mkdir -p $O
timeout 170 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "full_size and c3" > $O/pytest_full_c3.log 2>&1
echo DONE
