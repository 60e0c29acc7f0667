#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "== gpu tests: exit $? : $(tail -n 1 gpurun_out/gpu_tests.log)"; grep -E "^E|FAILED" gpurun_out/gpu_tests.log | head -12
PROF_ONLY=ln_,attn_space timeout 300 python tools/prof_kernels.py 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench: exit $?"; tail -3 gpurun_out/bench.err; cut -c1-330 gpurun_out/bench.json
PROF_NO_TIMING=1 PROF_ONLY=attn_space,ln_bwd timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/kern python tools/prof_kernels.py > gpurun_out/ncu_kern.log 2>&1
echo "== ncu: exit $?"
ncu -i /tmp/kern.ncu-rep --page raw --csv > gpurun_out/kern_raw.csv 2>/dev/null
ncu -i /tmp/kern.ncu-rep --page source --csv > gpurun_out/kern_source.csv 2>/dev/null
ncu -i /tmp/kern.ncu-rep --page details --csv > gpurun_out/kern_details.csv 2>/dev/null
gzip -f gpurun_out/kern_source.csv
