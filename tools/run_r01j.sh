#!/bin/bash
# 2-GPU sanity of the final tree: multi-rank parity check + the N=2 bench line.
N=2
mkdir -p gpurun_out
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/dist_check.log 2>&1
echo "== dist_check N=$N: exit $?"; grep -E "ok|Error|error|mismatch|assert" gpurun_out/dist_check.log | tail -6
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "== bench N=$N: exit $?"; grep -E "egv:|Error" gpurun_out/bench_n$N.err | tail -6; python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', d['e2e']['value']); print(d['config']['embedding_gather'], d['config']['last_loss'], d['clocks'])"
