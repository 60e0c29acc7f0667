mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "== kernels: exit $? : $(tail -n 1 gpurun_out/kernels.log)"
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/model.log 2>&1
echo "== model: exit $? : $(tail -n 1 gpurun_out/model.log)"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench: exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']['value']); print(d['roofline']['achieved'], d['config']['gemm_share_of_kernel_time'])"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
echo "== ncu: exit $?"; python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.md; head -34 gpurun_out/launches_summary.md
