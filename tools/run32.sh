#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "gemm or colsum or elementwise" > gpurun_out/tests.log 2>&1
echo "== tests: exit $? : $(tail -n 1 gpurun_out/tests.log)"; grep -E "^E|FAILED|egv:" gpurun_out/tests.log | head -20
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/model_tests.log 2>&1
echo "== model tests: exit $? : $(tail -n 1 gpurun_out/model_tests.log)"; grep -E "^E|FAILED" gpurun_out/model_tests.log | head -12
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench: exit $?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline']['achieved'], d['roofline']['frac'], d['config']['gemm_share_of_kernel_time'], d['config']['last_loss'])"
