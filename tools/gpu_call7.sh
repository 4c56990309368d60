#!/bin/bash
# final captures of round 2 (default build): launch list, DCN ncu traffic, ncu detail pages, sanitizer sweep
O=gpurun_out/r2g; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2_final.csv python bench.py --steps 2 --warmup 1 --no-graph --depth 1 --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:TsDcnAP -c 16 -o $O/prof_dcn_r2 -f python bench.py --steps 2 --warmup 3 --no-graph --depth 1 --no-cpu-baseline > $O/ncu_dcn.log 2>&1
timeout 120 python tools/ncu_traffic.py $O/prof_dcn_r2.ncu-rep $O/r2_dcn_ncu_traffic.json > $O/ncu_traffic.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_ts_kernel --launch-skip 6 --launch-count 2 -o $O/ncu_dcn64 -f python tools/one_layer.py dcn64 4 > $O/ncu_dcn64.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_ts_kernel --launch-skip 9 --launch-count 3 -o $O/ncu_dcn256 -f python tools/one_layer.py dcn256 4 > $O/ncu_dcn256.log 2>&1
for r in ncu_dcn64 ncu_dcn256; do ncu -i $O/$r.ncu-rep --page details --csv > $O/$r.details.csv 2>/dev/null; done
timeout 1500 bash tools/sanitize.sh $O/sanitize > $O/sanitize.out 2>&1
echo DONE
