#!/bin/bash
# 2-GPU: P2P gather + step parity check, then the 2-GPU bench (two streams + graph + NCCL all-reduce)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/dist_check.log 2>&1
echo "== dist_check: exit $?"; grep -v "^W\|^\[W\|NCCL version" gpurun_out/dist_check.log | tail -8
BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "== bench N=2: exit $?"; grep -E "^\[bench|egv:|Error" gpurun_out/bench_n2.err | tail -6; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', d['e2e']['value']); print(d['config']['embedding_gather'], d['config']['last_loss'])"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench N=1: exit $?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print(d['roofline']['achieved'], d['config']['gemm_share_of_kernel_time'], d['config']['last_loss'])"
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "attention or adamw" > gpurun_out/tests.log 2>&1
echo "== tests: exit $? : $(tail -n 1 gpurun_out/tests.log)"; grep -E "^E|FAILED" gpurun_out/tests.log | head
PROF_ONLY=attn_cls timeout 300 python tools/prof_kernels.py 2>&1 | tail -2
