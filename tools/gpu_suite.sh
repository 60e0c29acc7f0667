#!/bin/bash
# Run the GPU kernel tests group by group (a CUDA fault in one group must not hide the others).
# usage (on the GPU box): bash tools/gpu_suite.sh [extra pytest args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in gemm_plain gemm_epilogue "gemm_inplace or gemm_split or gemm_simt or gemm_linearity or gemm_argument" layernorm attention "elementwise or patch_embed" "xent or egonce or adamw"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" --tb=short -p no:cacheprovider "$@" > "gpurun_out/kernels_${name}.log" 2>&1
  echo "== $grp: exit $? : $(tail -n 1 gpurun_out/kernels_${name}.log)"
done
