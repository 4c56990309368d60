#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 300 python tools/tc_diag.py c3x3_16_16 c3x3_16_32_s2 c1x1_k32_n16 c3x3_64_64 dcn_64_64 root_448_128 c3x3_512_512 > $O/diag.log 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/pytest_ops.log 2>&1
timeout 200 python tools/tc_perf.py > $O/perf.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_r2_final.csv python bench.py --steps 2 --warmup 1 --no-graph --depth 1 --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k regex:TsDcnAP -c 16 -o $O/prof_dcn_r2 -f python bench.py --steps 1 --warmup 3 --no-graph --depth 1 --no-cpu-baseline > $O/ncu_dcn.log 2>&1
timeout 120 python tools/ncu_traffic.py $O/prof_dcn_r2.ncu-rep $O/r2_dcn_ncu_traffic.json > $O/ncu_traffic.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_ts_kernel --launch-skip 6 --launch-count 2 -o $O/ncu_dcn256 -f python tools/one_layer.py dcn256 4 > $O/ncu_dcn256.log 2>&1
ncu -i $O/ncu_dcn256.ncu-rep --page details --csv > $O/ncu_dcn256.details.csv 2>/dev/null
echo DONE
