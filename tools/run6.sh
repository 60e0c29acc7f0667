mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/model.log 2>&1
echo "== model: exit $? : $(tail -n 1 gpurun_out/model.log)"
for mode in "" "--no-graph"; do
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline $mode > gpurun_out/bench$mode.json 2> gpurun_out/bench$mode.err
echo "== bench $mode: exit $?"; tail -3 gpurun_out/bench$mode.err; python -c "
import json; d=json.load(open('gpurun_out/bench$mode.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline']['achieved'], d['config']['gemm_share_of_kernel_time'], d['config']['last_loss'])"
done
