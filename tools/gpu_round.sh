#!/bin/bash
# Full GPU check of the current tree: kernel tests, model parity tests, smoke, a short bench, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "== kernels: exit $? : $(tail -n 1 gpurun_out/kernels.log)"
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/model.log 2>&1
echo "== model: exit $? : $(tail -n 1 gpurun_out/model.log)"
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "== smoke: exit $? : $(tail -n 1 gpurun_out/smoke.log)"
timeout 1200 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench: exit $? : $(tail -c 1500 gpurun_out/bench.json)"
tail -n 5 gpurun_out/bench.err
