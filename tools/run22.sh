#!/bin/bash
mkdir -p gpurun_out
GEMM_LAYOUTS=NT,NN,REF timeout 600 python tools/gemm_bench.py 25096x2304x768 25096x768x768 25096x3072x768 25096x768x3072 2>&1 | tee gpurun_out/gemm_bench.txt
PROF_ONLY=attn_time timeout 300 python tools/prof_kernels.py 2>&1 | tail -2
