#!/bin/bash
# round 2, call 10: deep levels — CUDA-event times of the weight-gradient GEMMs, then a source-level ncu capture of tn_gemm
O=gpurun_out/r2c10
mkdir -p $O
(timeout 300 python tools/profile_deep.py --time 2>&1 | tail -30) > $O/time_deep.log
cat $O/time_deep.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'tn_gemm_kernel' -c 8 -o $O/tn_gemm python tools/profile_deep.py > $O/ncu.log 2>&1
tail -3 $O/ncu.log
