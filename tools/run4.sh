mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "gemm" > gpurun_out/kernels.log 2>&1
echo "== kernels: exit $? : $(tail -n 1 gpurun_out/kernels.log)"
GEMM_LAYOUTS=NT,TN python tools/gemm_bench.py 25096x2304x768 25096x768x768 25096x768x3072 > gpurun_out/gemm_bench.txt 2>&1; cat gpurun_out/gemm_bench.txt
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/model.log 2>&1
echo "== model: exit $? : $(tail -n 1 gpurun_out/model.log)"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench: exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']['value']); print(d['roofline']['achieved'], d['config']['gemm_share_of_kernel_time'])"
