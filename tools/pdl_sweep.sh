#!/bin/bash
# Step time per programmatic-dependent-launch class mask (EGV_PDL, see csrc/host_common.h).  usage: bash tools/pdl_sweep.sh "0 1 4 8 16 31"
mkdir -p gpurun_out
for m in $1; do
  EGV_PDL=$m python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl_m$m.json 2> gpurun_out/bench_pdl_m$m.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_pdl_m$m.json").read().strip().splitlines()[-1])
    print("EGV_PDL=$m", round(d["ms_per_step"], 3), "ms", round(d["value"], 2), "clips/s", "loss", d["config"]["last_loss"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("EGV_PDL=$m", "ERR", e)
PY
done
