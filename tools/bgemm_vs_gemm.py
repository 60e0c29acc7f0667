#!/usr/bin/env python
"""bgemm (csrc/xgemm.cu) against the plain tcgen05 GEMM (csrc/gemm.cu) on the same flat problem, and per-clip batching
overhead: CUDA-event medians, L2 flushed."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egovlpv2_b200 import lib as L  # noqa: E402
from egovlpv2_b200.lib import BV, GEMM_NN, GEMM_NT, GEMM_TN  # noqa: E402

DEV = "cuda"
K = L.kernels()
flush = torch.empty(160 * 2 ** 20, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=7):
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def rnd(*shape, dtype=torch.bfloat16):
    return torch.randn(*shape, device=DEV).to(dtype)


B, N, C, HS = 8, 3137, 768, 384
M = B * N
A, W = rnd(M, C), rnd(HS, C)
Wb = rnd(B, HS, C)
o16 = torch.empty(M, HS, dtype=torch.bfloat16, device=DEV)
fl = 2.0 * M * HS * C
for name, fn in [
    ("gemm   NT M=25096 N=384 K=768 -> bf16", lambda: K.gemm(GEMM_NT, A, W, out_bf16=o16)),
    ("bgemm  NT same, nb=(1,1)", lambda: K.bgemm(GEMM_NT, M, HS, C, BV(A, C), BV(W, C), out_bf16=BV(o16, HS))),
    ("bgemm  NT per clip nb=(8,1)", lambda: K.bgemm(GEMM_NT, N, HS, C, BV(A, C, N * C), BV(Wb, C, HS * C), nb=(B, 1), out_bf16=BV(o16, HS, N * HS))),
]:
    us = timeit(fn)
    print("%-50s %7.1f us %7.1f TF/s" % (name, us, fl / us / 1e6))
P, U = rnd(M, HS), rnd(HS, C)
Ub = rnd(B, HS, C)
o32 = torch.empty(M, C, device=DEV)
res = torch.randn(M, C, device=DEV)
fl = 2.0 * M * HS * C
for name, fn in [
    ("gemm   NN M=25096 N=768 K=384 +res -> f32", lambda: K.gemm(GEMM_NN, P, U, residual=res, out_f32=o32)),
    ("bgemm  NN same nb=(1,1)", lambda: K.bgemm(GEMM_NN, M, C, HS, BV(P, HS), BV(U, C), residual=BV(res, C), out_f32=BV(o32, C))),
    ("bgemm  NN per clip", lambda: K.bgemm(GEMM_NN, N, C, HS, BV(P, HS, N * HS), BV(Ub, C, HS * C), nb=(B, 1), residual=BV(res, C, N * C), out_f32=BV(o32, C, N * C))),
    ("gemm   NN no residual -> bf16 [M,768]", lambda: K.gemm(GEMM_NN, P, U, out_bf16=A)),
    ("bgemm  NN per clip -> bf16", lambda: K.bgemm(GEMM_NN, N, C, HS, BV(P, HS, N * HS), BV(Ub, C, HS * C), nb=(B, 1), out_bf16=BV(A, C, N * C))),
]:
    us = timeit(fn)
    print("%-50s %7.1f us %7.1f TF/s" % (name, us, fl / us / 1e6))
dU = torch.zeros(HS, C, device=DEV)
dUb = torch.zeros(B, HS, C, device=DEV)
for name, fn in [
    ("gemm   TN M=384 N=768 K=25096", lambda: K.gemm(GEMM_TN, P, A, out_f32=dU)),
    ("bgemm  TN per clip K=3137 accumulate", lambda: K.bgemm(GEMM_TN, HS, C, N, BV(P, HS, N * HS), BV(A, C, N * C), nb=(B, 1), out_f32=BV(dUb, C, HS * C), accumulate=True)),
]:
    us = timeit(fn)
    print("%-50s %7.1f us %7.1f TF/s" % (name, us, fl / us / 1e6))
# launch floor: a one-tile problem
a1, b1 = rnd(128, 64), rnd(64, 64)
o1 = torch.empty(128, 64, dtype=torch.bfloat16, device=DEV)
print("one-tile bgemm %.1f us; one-tile gemm (>= 2^18 MACs) %.1f us; cast 256x768 %.1f us" % (
    timeit(lambda: K.bgemm(GEMM_NT, 128, 64, 64, BV(a1, 64), BV(b1, 64), out_bf16=BV(o1, 64))),
    timeit(lambda: K.gemm(GEMM_NT, a1, b1, out_bf16=o1)),
    timeit(lambda: K.cast(res[:256], A[:256]))))
