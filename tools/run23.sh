#!/bin/bash
mkdir -p gpurun_out
PROF_ONLY=attn_time,attn_cls timeout 300 python tools/prof_kernels.py 2>&1 | tail -4
PROF_NO_TIMING=1 PROF_ONLY=gemm_qkv_fwd,gemm_fc1_gelu timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/kern python tools/prof_kernels.py > gpurun_out/ncu_kern.log 2>&1
echo "== ncu: exit $?"
ncu -i /tmp/kern.ncu-rep --page source --csv > gpurun_out/gemm_source.csv 2>/dev/null
ncu -i /tmp/kern.ncu-rep --page details --csv > gpurun_out/gemm_details.csv 2>/dev/null
gzip -f gpurun_out/gemm_source.csv
