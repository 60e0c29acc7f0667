#!/bin/bash
# GPU call: planner A/B on the weight-gradient GEMMs, per-kernel timing table, ncu --set full of the hot kernels, bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k gemm > gpurun_out/gemm_tests.log 2>&1
echo "== gemm tests (plan 1): exit $? : $(tail -n 1 gpurun_out/gemm_tests.log)"; grep -E "^E|FAILED" gpurun_out/gemm_tests.log | head -8
EGV_GEMM_PLAN=2 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k gemm > gpurun_out/gemm_tests2.log 2>&1
echo "== gemm tests (plan 2): exit $? : $(tail -n 1 gpurun_out/gemm_tests2.log)"; grep -E "^E|FAILED" gpurun_out/gemm_tests2.log | head -8
for pm in 0 1 2; do echo PLAN=$pm; EGV_GEMM_PLAN=$pm GEMM_LAYOUTS=TN timeout 300 python tools/gemm_bench.py 25096x2304x768 25096x768x768 25096x3072x768 25096x768x3072 25096x1536x768 2>&1 | grep TN; done | tee gpurun_out/tn_plan.txt
timeout 600 python tools/prof_kernels.py > gpurun_out/kernel_times.txt 2>&1; cat gpurun_out/kernel_times.txt
PROF_NO_TIMING=1 PROF_ONLY=attn_space,attn_time,attn_cls_bwd,ln_bwd,gemm_wgrad_fc2,gemm_wgrad_proj,gemm_fc1_gelu,gemm_proj_fwd_res,gemm_dgrad_fc2,attn_i2t timeout 900 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/kern python tools/prof_kernels.py > gpurun_out/ncu_kern.log 2>&1
echo "== ncu: exit $?"; ls -la gpurun_out/*.ncu-rep
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench: exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
