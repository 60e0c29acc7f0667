#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "gemm" > gpurun_out/tests.log 2>&1
echo "== tests: exit $? : $(tail -n 1 gpurun_out/tests.log)"; grep -E "^E|FAILED|egv:" gpurun_out/tests.log | head -20
PROF_ONLY=gemm_fc1,gemm_dgrad_fc2 timeout 300 python tools/prof_kernels.py 2>&1 | tail -4
