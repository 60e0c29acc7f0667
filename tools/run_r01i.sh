#!/bin/bash
# GPU tests + the fine-tuning (Dual) step bench for both datasets.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "== gpu tests: exit $? : $(tail -n 1 gpurun_out/gpu_tests.log)"; grep -E "^E|FAILED" gpurun_out/gpu_tests.log | head -12
timeout 300 python tools/bench_dual.py --dataset epic > gpurun_out/bench_dual_epic.json 2> gpurun_out/bench_dual_epic.err
echo "== dual epic: exit $?"; tail -3 gpurun_out/bench_dual_epic.err; cat gpurun_out/bench_dual_epic.json
timeout 300 python tools/bench_dual.py --dataset charades > gpurun_out/bench_dual_charades.json 2> gpurun_out/bench_dual_charades.err
echo "== dual charades: exit $?"; tail -3 gpurun_out/bench_dual_charades.err; cat gpurun_out/bench_dual_charades.json
