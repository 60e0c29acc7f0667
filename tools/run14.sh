#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k attention > gpurun_out/attn_tests.log 2>&1
echo "== attention tests (wide): exit $? : $(tail -n 1 gpurun_out/attn_tests.log)"; grep -E "^E|FAILED|egv:" gpurun_out/attn_tests.log | head -30
EGV_ATTN_GROUP_WIDE=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k attention > gpurun_out/attn_tests0.log 2>&1
echo "== attention tests (7 warps): exit $? : $(tail -n 1 gpurun_out/attn_tests0.log)"; grep -E "^E|FAILED|egv:" gpurun_out/attn_tests0.log | head -30
PROF_ONLY=attn_space timeout 300 python tools/prof_kernels.py 2>&1 | tail -2
EGV_ATTN_GROUP_WIDE=0 PROF_ONLY=attn_space timeout 300 python tools/prof_kernels.py 2>&1 | tail -2
