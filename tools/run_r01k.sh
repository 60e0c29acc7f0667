#!/bin/bash
mkdir -p gpurun_out
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/dist_check.log 2>&1
echo "== dist_check N=2: exit $?"; grep -E "ok|AssertionError|mismatch" gpurun_out/dist_check.log | tail -6
