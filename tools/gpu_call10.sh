#!/bin/bash
# 2-GPU call: stream-parallel scaling (C2, one stream per GPU), the nuScenes split (6 cameras over 2 GPUs), multi-device launch test
O=gpurun_out/r2j; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "second_gpu" > $O/pytest_2gpu.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_c2_2gpu.json 2> $O/bench_c2_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c4 --streams 6 --steps 60 --warmup 10 --no-cpu-baseline > $O/bench_c4_6cams_2gpu.json 2> $O/bench_c4_6cams_2gpu.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_c2_1gpu_same_box.json 2> $O/bench_c2_1gpu_same_box.err
echo DONE
