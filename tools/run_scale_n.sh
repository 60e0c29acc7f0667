#!/bin/bash
# bench.py on N GPUs of one box (torchrun), then the same box's 1-GPU line.  usage: bash tools/run_scale_n.sh 8
N=${1:-8}
mkdir -p gpurun_out
BENCH_VERBOSE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "== bench N=$N: exit $?"; grep -E "egv:|Error" gpurun_out/bench_n$N.err | tail -4
python -c "
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', d['e2e']['value'], d['config']['embedding_gather'], d['config']['last_loss'], d['clocks'])"
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_same_box.json 2> gpurun_out/bench_n1_same_box.err
echo "== bench N=1 (same box): exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n1_same_box.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}); print('e2e', d['e2e']['value'])"
