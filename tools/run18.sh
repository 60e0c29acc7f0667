#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k attention > gpurun_out/attn_tests.log 2>&1
echo "== attention tests: exit $? : $(tail -n 1 gpurun_out/attn_tests.log)"; grep -E "^E|FAILED|egv:" gpurun_out/attn_tests.log | head -30
PROF_ONLY=attn_cls timeout 300 python tools/prof_kernels.py 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench: exit $?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline']['achieved'], d['config']['gemm_share_of_kernel_time'], d['config']['last_loss'])"
EGV_TEXT_STREAM=0 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
echo "== ncu: exit $?"; python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.md; head -40 gpurun_out/launches_summary.md
