#!/bin/bash
# final combination (conv split target 50, DCN 74): the full-size C2 frame through FramePipeline(depth=4) against the oracle
O=gpurun_out/r2zz; mkdir -p $O
timeout 280 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "full_size and c2" > $O/pytest_full_c2.log 2>&1
echo DONE
