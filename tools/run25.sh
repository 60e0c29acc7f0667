#!/bin/bash
mkdir -p gpurun_out
EGV_GEMM_CLUSTER=1 GEMM_ONLY="bf16 out" GEMM_LAYOUTS=NT timeout 300 ncu --set full --clock-control none -k regex:gemm_tc -s 2 -c 2 -o /tmp/pair python tools/gemm_bench.py 25096x2304x768 > gpurun_out/ncu_pair.log 2>&1
echo "== ncu: exit $?"; tail -3 gpurun_out/ncu_pair.log
ncu -i /tmp/pair.ncu-rep --page details --csv > gpurun_out/pair_details.csv 2>/dev/null
ncu -i /tmp/pair.ncu-rep --page source --csv > gpurun_out/pair_source.csv 2>/dev/null
gzip -f gpurun_out/pair_source.csv
python - <<'PY'
import torch
from egovlpv2_b200 import lib as L
print(torch.cuda.get_device_properties(0))
PY
