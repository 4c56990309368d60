#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
DEFT_PDL=0 timeout 120 python tools/tc_trace.py conv16 > $O/trace_conv16_patch.log 2>&1
DEFT_TC_PATCH16=0 DEFT_PDL=0 timeout 120 python tools/tc_trace.py conv16 > $O/trace_conv16_gather.log 2>&1
for b in 0 8 6 16 14 30; do DEFT_TC_DBG=$b timeout 100 python tools/tc_perf.py > $O/perf_dbg$b.log 2>&1; done
echo DONE
