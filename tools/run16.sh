#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "== gpu tests: exit $? : $(tail -n 1 gpurun_out/gpu_tests.log)"; grep -E "^E|FAILED" gpurun_out/gpu_tests.log | head -12
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench two streams: exit $?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline']['achieved'], d['config']['gemm_share_of_kernel_time'], d['config']['last_loss'])"
EGV_TEXT_STREAM=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench1.json 2> gpurun_out/bench1.err
echo "== bench one stream: exit $?"; tail -3 gpurun_out/bench1.err; python -c "
import json; d=json.load(open('gpurun_out/bench1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline']['achieved'], d['config']['gemm_share_of_kernel_time'], d['config']['last_loss'])"
