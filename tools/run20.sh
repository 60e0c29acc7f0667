#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "== gpu tests: exit $? : $(tail -n 1 gpurun_out/gpu_tests.log)"; grep -E "^E|FAILED" gpurun_out/gpu_tests.log | head -12
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench N=1: exit $?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline']['achieved'], d['config']['gemm_share_of_kernel_time'], d['config']['last_loss'])"
PROF_ONLY=gemm_fc1_gelu,gemm_dgrad_fc2 timeout 300 python tools/prof_kernels.py 2>&1 | tail -2
EGV_ATTN_TINY=3 PROF_NO_TIMING=1 PROF_ONLY=attn_time timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/kern python tools/prof_kernels.py > gpurun_out/ncu_kern.log 2>&1
echo "== ncu: exit $?"
ncu -i /tmp/kern.ncu-rep --page raw --csv > gpurun_out/tiny_raw.csv 2>/dev/null
ncu -i /tmp/kern.ncu-rep --page source --csv > gpurun_out/tiny_source.csv 2>/dev/null
ncu -i /tmp/kern.ncu-rep --page details --csv > gpurun_out/tiny_details.csv 2>/dev/null
gzip -f gpurun_out/tiny_source.csv
