"""GEMM microbenchmark: TFLOP/s of egv_gemm_bf16 per shape / layout / epilogue variant (CUDA events, L2 flushed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlpv2_b200 import lib as L  # noqa: E402

K = L.Kernels()
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def operands(layout, M, N, Kd):
    if layout == L.GEMM_NT:
        return torch.randn(M, Kd, device=dev).bfloat16(), torch.randn(N, Kd, device=dev).bfloat16()
    if layout == L.GEMM_NN:
        return torch.randn(M, Kd, device=dev).bfloat16(), torch.randn(Kd, N, device=dev).bfloat16()
    return torch.randn(Kd, M, device=dev).bfloat16(), torch.randn(Kd, N, device=dev).bfloat16()


shapes = [(25096, 2304, 768), (25096, 768, 768), (25096, 3072, 768), (25096, 768, 3072), (25096, 1536, 768)]
ONLY = os.environ.get("GEMM_ONLY")       # substring of the variant name
LAYOUTS = os.environ.get("GEMM_LAYOUTS", "NT,NN,TN,REF").split(",")
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
print("%-22s %-4s %-26s %9s %9s" % ("shape MxNxK", "lay", "epilogue", "us", "TFLOP/s"))
for (M, N, Kd) in shapes:
    for layout, lname in ((L.GEMM_NT, "NT"), (L.GEMM_NN, "NN")):
        if lname not in LAYOUTS:
            continue
        A, B = operands(layout, M, N, Kd)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        o32 = torch.empty(M, N, device=dev)
        o16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        opre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        variants = [
            ("mainloop only (no stores)", dict(out_f32=o32, act=99)),
            ("bf16 out", dict(out_bf16=o16)),
            ("bias + bf16 out", dict(bias=bias, out_bf16=o16)),
            ("bias + gelu + pre + bf16", dict(bias=bias, act=L.ACT_GELU, out_bf16=o16, out_pre=opre)),
            ("bias + residual + f32 out", dict(bias=bias, residual=res, out_f32=o32)),
            ("gelu_bwd(aux) + bf16 out", dict(aux=opre, act=L.ACT_GELU_BWD, out_bf16=o16)),
        ]
        for name, kw in variants:
            if ONLY and ONLY not in name:
                continue
            t = timeit(lambda: K.gemm(layout, A, B, **kw))
            print("%-22s %-4s %-26s %9.1f %9.1f" % ("%dx%dx%d" % (M, N, Kd), lname, name, t * 1e6, 2.0 * M * N * Kd / t / 1e12))
    if "TN" not in LAYOUTS:
        continue
    # weight-gradient shape: out [N, Kd] = dy^T x with reduction over M
    A, B = torch.randn(M, N, device=dev).bfloat16(), torch.randn(M, Kd, device=dev).bfloat16()
    out = torch.empty(N, Kd, device=dev)
    t = timeit(lambda: K.gemm(L.GEMM_TN, A, B, out_f32=out))
    print("%-22s %-4s %-26s %9.1f %9.1f" % ("%dx%dx%d" % (N, Kd, M), "TN", "f32 out (auto split-K)", t * 1e6, 2.0 * M * N * Kd / t / 1e12))
# reference point: cuBLAS through torch
for (M, N, Kd) in (shapes if "REF" in LAYOUTS else []):
    A, B = torch.randn(M, Kd, device=dev).bfloat16(), torch.randn(N, Kd, device=dev).bfloat16()
    t = timeit(lambda: torch.matmul(A, B.t()))
    print("%-22s %-4s %-26s %9.1f %9.1f" % ("%dx%dx%d" % (M, N, Kd), "NT", "torch.matmul (cuBLAS) ref", t * 1e6, 2.0 * M * N * Kd / t / 1e12))
