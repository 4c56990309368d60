#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
timeout 120 python tools/tc_trace.py dcn64 > $O/trace_dcn64_patch.log 2>&1
timeout 120 python tools/tc_trace.py dcn128 > $O/trace_dcn128_patch.log 2>&1
DEFT_PDL=0 timeout 120 python tools/tc_trace.py dcn64 > $O/trace_dcn64_patch_nopdl.log 2>&1
DEFT_DCN_PATCH=0 DEFT_PDL=0 timeout 120 python tools/tc_trace.py dcn64 > $O/trace_dcn64_gather_nopdl.log 2>&1
echo DONE
