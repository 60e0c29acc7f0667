#!/bin/bash
mkdir -p gpurun_out
EGV_ATTN_TINY_STRICT=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k attention > gpurun_out/attn_tests.log 2>&1
echo "== attention tests: exit $? : $(tail -n 1 gpurun_out/attn_tests.log)"; grep -E "^E|FAILED|egv:" gpurun_out/attn_tests.log | head -30
echo default; PROF_ONLY=attn_space,attn_time timeout 300 python tools/prof_kernels.py 2>&1 | tail -4
echo "wide=7 (all modes)"; EGV_ATTN_GROUP_WIDE=7 PROF_ONLY=attn_space timeout 300 python tools/prof_kernels.py 2>&1 | tail -2
echo "wide=0"; EGV_ATTN_GROUP_WIDE=0 PROF_ONLY=attn_space timeout 300 python tools/prof_kernels.py 2>&1 | tail -2
echo "tiny off"; EGV_ATTN_TINY=0 PROF_ONLY=attn_time timeout 300 python tools/prof_kernels.py 2>&1 | tail -2
