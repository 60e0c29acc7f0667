#!/bin/bash
# Round-1 late check: full GPU tests (incl. the uint8 input path and the fine-tuning Dual path), smoke, default bench,
# the uint8-input bench, per-kernel times.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "== gpu tests: exit $? : $(tail -n 1 gpurun_out/gpu_tests.log)"; grep -E "^E|FAILED" gpurun_out/gpu_tests.log | head -12
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
echo "== smoke: exit $? : $(tail -n 1 gpurun_out/smoke.log)"
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench: exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --input-dtype u8 --no-cpu-baseline > gpurun_out/bench_u8.json 2> gpurun_out/bench_u8.err
echo "== bench u8: exit $?"; tail -3 gpurun_out/bench_u8.err; cut -c1-400 gpurun_out/bench_u8.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_u8.json
