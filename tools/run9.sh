mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k gemm > gpurun_out/gpu_tests.log 2>&1
echo "== gemm tests: exit $? : $(tail -n 1 gpurun_out/gpu_tests.log)"; grep -E "^E|FAILED" gpurun_out/gpu_tests.log | head -8
for c in 1 0; do echo CLUSTER=$c; EGV_GEMM_CLUSTER=$c GEMM_LAYOUTS=NT,NN,TN timeout 300 python tools/gemm_bench.py 25096x2304x768 25096x768x768 25096x768x3072 2>&1 | grep -E "TN|bias \+ bf16|mainloop|residual"; done
