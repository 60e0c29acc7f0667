"""Attention microbenchmark at the cfg-3 shapes (B=8, T=16, Nf=196, H=12, S=32): fwd and bwd per attention type."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlpv2_b200 import lib as L  # noqa: E402

K = L.Kernels()
dev = "cuda"
B, T, Nf, H, S = 8, 16, 196, 12, 32
C = H * 64
N = 1 + T * Nf
sc = 64 ** -0.5
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
    return ts[len(ts) // 2] * 1e3


qkv = torch.randn(B, N, 3 * C, device=dev).bfloat16()
q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
specs = {
    "time": L.AttnSpec(H=H, G=Nf, Lq=T, Lk=T, q_row0=1, q_gstride=1, q_istride=Nf, k_row0=1, k_gstride=1, k_istride=Nf,
                       has_cls_key=True, cls_row=0, scale=sc),
    "space": L.AttnSpec(H=H, G=T, Lq=Nf, Lk=Nf, q_row0=1, q_gstride=Nf, q_istride=1, k_row0=1, k_gstride=Nf, k_istride=1,
                        has_cls_key=True, cls_row=0, scale=sc),
    "cls": L.AttnSpec(H=H, G=1, Lq=1, Lk=N - 1, q_row0=0, k_row0=1, has_cls_key=True, cls_row=0, scale=sc),
}
print("%-8s %10s %10s" % ("type", "fwd us", "bwd us"))
for name, spec in specs.items():
    o = torch.zeros(B, N, C, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B * H * spec.G * spec.Lq, device=dev)
    d_o = torch.randn(B, N, C, device=dev).bfloat16()
    dqkv = torch.zeros_like(qkv)
    cls = torch.zeros(B * H * 128, device=dev)
    tf = timeit(lambda: K.attention_fwd(spec, q, k, v, o, lse))
    tb = timeit(lambda: K.attention_bwd(spec, q, k, v, o, lse, d_o, dqkv[:, :, :C], dqkv[:, :, C:2 * C], dqkv[:, :, 2 * C:],
                                        torch.empty_like(lse), dkv_cls=cls, dkv_accumulate=(name == "cls")))
    print("%-8s %10.1f %10.1f" % (name, tf, tb))
# cross attention
xq = torch.randn(B, N, C, device=dev).bfloat16()
tkv = torch.randn(B, S, 2 * C, device=dev).bfloat16()
kb = torch.zeros(B, S, device=dev)
for name, (qq, kk, vv, spec, bias) in {
    "i2t": (xq, tkv[:, :, :C], tkv[:, :, C:], L.AttnSpec(H=H, G=1, Lq=N, Lk=S, scale=sc), kb),
    "t2i": (tkv[:, :, :C].contiguous(), qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], L.AttnSpec(H=H, G=1, Lq=S, Lk=N, scale=sc), None),
    "text": (tkv[:, :, :C].contiguous(), tkv[:, :, :C], tkv[:, :, C:], L.AttnSpec(H=H, G=1, Lq=S, Lk=S, scale=sc), kb),
}.items():
    Bq, Lq = qq.shape[0], qq.shape[1]
    o = torch.zeros(Bq, Lq, C, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(Bq * H * Lq, device=dev)
    d_o = torch.randn(Bq, Lq, C, device=dev).bfloat16()
    dq = torch.zeros_like(qq)
    dkvb = torch.zeros(kk.shape[0], kk.shape[1], 2 * C, device=dev, dtype=torch.bfloat16)
    tf = timeit(lambda: K.attention_fwd(spec, qq, kk, vv, o, lse, key_bias=bias))
    tb = timeit(lambda: K.attention_bwd(spec, qq, kk, vv, o, lse, d_o, dq, dkvb[:, :, :C], dkvb[:, :, C:], torch.empty_like(lse),
                                        key_bias=bias))
    print("%-8s %10.1f %10.1f" % (name, tf, tb))
