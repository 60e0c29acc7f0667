#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_kernels.py > gpurun_out/kernel_times.txt 2>&1; cat gpurun_out/kernel_times.txt | tail -27
EGV_ATTN_TINY=3 PROF_ONLY=attn_time timeout 300 python tools/prof_kernels.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "gemm or attention" > gpurun_out/gpu_tests.log 2>&1
echo "== gpu tests: exit $? : $(tail -n 1 gpurun_out/gpu_tests.log)"; grep -E "^E|FAILED" gpurun_out/gpu_tests.log | head -12
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench N=1: exit $?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline']['achieved'], d['config']['gemm_share_of_kernel_time'], d['config']['last_loss'])"
