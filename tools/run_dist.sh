mkdir -p gpurun_out
N=${1:-2}
for mode in ""; do
BENCH_VERBOSE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 $mode > gpurun_out/bench_n$N$mode.json 2> gpurun_out/bench_n$N$mode.err
echo "== bench N=$N $mode: exit $?"; grep -E "^\[bench|egv:|Error" gpurun_out/bench_n$N$mode.err | tail -12; tail -c 900 gpurun_out/bench_n$N$mode.json
done
