#!/bin/bash
# Round-2 evidence (on the GPU box): GPU tests, smoke, bench (with the CPU baseline), ncu launch list of one step
# (time + DRAM bytes), ncu --set full of the hot kernels exported to CSV, per-kernel event times, cross-attention launches.
# usage: bash tools/run_r02.sh   -> files under gpurun_out/ (copied to profiles/r02_* by hand)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/ -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "== gpu tests: exit $? : $(tail -n 1 gpurun_out/gpu_tests.log)"; grep -E "^E|FAILED" gpurun_out/gpu_tests.log | head -12
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
echo "== smoke: exit $? : $(tail -n 1 gpurun_out/smoke.log)"
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench: exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
EGV_TEXT_STREAM=0 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
echo "== ncu launch list: exit $?"; python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.md; head -12 gpurun_out/launches_summary.md; tail -1 gpurun_out/launches_summary.md
python tools/gemm_traffic.py gpurun_out/launches.csv gpurun_out/gemm_traffic.json
gzip -f gpurun_out/launches.csv
PROF_NO_TIMING=1 PROF_ONLY=gemm_qkv_fwd,gemm_fc1_gelu_dg,gemm_dgrad_fc2_mulaux,gemm_wgrad_fc2,gemm_dgrad_fc1,gemm_proj_fwd_res,attn_space,attn_time,attn_cls,ln_bwd,ln_fwd timeout 300 ncu --set full --clock-control none --profile-from-start off -o /tmp/kern python tools/prof_kernels.py > gpurun_out/ncu_kern.log 2>&1
echo "== ncu full: exit $?"
ncu -i /tmp/kern.ncu-rep --page raw --csv > gpurun_out/final_raw.csv 2>/dev/null
ncu -i /tmp/kern.ncu-rep --page source --csv > gpurun_out/final_source.csv 2>/dev/null
gzip -f gpurun_out/final_source.csv
timeout 200 ncu --set full --clock-control none -k regex:bgemm -c 12 -o /tmp/xa python tools/xattn_bench.py --iters 1 > gpurun_out/ncu_xattn.log 2>&1
echo "== ncu xattn: exit $?"
ncu -i /tmp/xa.ncu-rep --page raw --csv > gpurun_out/xattn_raw.csv 2>/dev/null
timeout 300 python tools/prof_kernels.py > gpurun_out/kernel_times.txt 2>&1
timeout 200 python tools/xattn_bench.py > gpurun_out/xattn_times.txt 2>&1
timeout 200 python tools/attn_bench.py > gpurun_out/attn_times.txt 2>&1
echo "== done"
