#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "cluster_pairs" -x > gpurun_out/pair_tests.log 2>&1
echo "== pair tests: exit $? : $(tail -n 1 gpurun_out/pair_tests.log)"; grep -E "^E|FAILED|egv:" gpurun_out/pair_tests.log | head -12
for c in 1; do echo CLUSTER=$c; EGV_GEMM_CLUSTER=$c GEMM_LAYOUTS=NT,TN timeout 240 python tools/gemm_bench.py 25096x2304x768 25096x3072x768 25096x768x3072 2>&1 | grep -E "TN|bf16 out|mainloop|residual|gelu" | grep -v "bias + bf16"; done | tee gpurun_out/pair_bench.txt
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "two_streams" > gpurun_out/ts.log 2>&1
echo "== two-stream test: exit $? : $(tail -n 1 gpurun_out/ts.log)"; grep -E "^E|FAILED" gpurun_out/ts.log | head
