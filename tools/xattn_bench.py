#!/usr/bin/env python
"""Per-launch timing of the re-associated cross-attention (egovlpv2_b200/xattn_reassoc.py) at the BASELINE cfg-3 shapes
(B = 8 clips, N = 3137 tokens, C = 768, 12 heads, S = 32): CUDA events around every kernel call, L2 flushed between
calls, median of `--iters` runs.  Prints one line per launch: method, shape signature, microseconds, TFLOP/s."""
import argparse
import collections
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egovlpv2_b200 import lib as L  # noqa: E402
from egovlpv2_b200 import xattn_reassoc as XR  # noqa: E402

DEV = "cuda"


class Timed:
    """wraps a Kernels object: every method call is bracketed by CUDA events (after an L2 flush)"""

    def __init__(self, K, flush=True):
        self.K, self.rec, self.flush = K, [], flush
        self.buf = torch.empty(160 * 2 ** 20, dtype=torch.uint8, device=DEV) if flush else None

    def __getattr__(self, name):
        fn = getattr(self.K, name)

        def call(*a, **k):
            if self.flush:
                self.buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            sig = name
            fl = 0.0
            if name == "bgemm":
                layout, M, N, Kd = a[0], a[1], a[2], a[3]
                nb = k.get("nb", (1, 1))
                fl = 2.0 * M * N * Kd * nb[0] * nb[1]
                sig = "bgemm %s M=%d N=%d K=%d nb=%s epi=%d%s" % ("NT NN TN".split()[layout], M, N, Kd, nb, k.get("epilogue", 0),
                                                                 " +res" if k.get("residual") is not None else "")
            self.rec.append((sig, fl, e0, e1))
            return r
        return call


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no-flush", action="store_true")
    a = ap.parse_args()
    B, N, C, H, S = a.batch, 3137, 768, 12, 32
    g = torch.Generator().manual_seed(0)

    def rnd(*shape, dtype=torch.bfloat16, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)
    lnc, kv = rnd(B * N, C), rnd(B * S, 2 * C)
    wq, wp, wk, wv = (rnd(C, C, scale=0.03) for _ in range(4))
    bq, bp, bv = (rnd(C, dtype=torch.float32, scale=0.1) for _ in range(3))
    xa, dout = rnd(B * N, C, dtype=torch.float32), rnd(B * N, C)
    mask = torch.zeros(B, S, device=DEV)
    alpha = torch.tensor([0.5], device=DEV)
    q, x, dox = rnd(B * S, C), rnd(B * N, C), rnd(B * S, C)
    K = Timed(L.kernels(), flush=not a.no_flush)
    times = collections.OrderedDict()
    for it in range(a.iters + 1):
        K.rec = []
        out = torch.empty(B * N, C, device=DEV)
        s = XR.i2t_fwd(K, lnc, kv, mask, wq, bq, wp, bp, alpha, xa, out, B, N, H)
        n_i2t_fwd = len(K.rec)
        da, dwq, dbq, dwp = (torch.zeros(n, device=DEV) for n in ((1,), (C, C), (C,), (C, C)))
        XR.i2t_bwd(K, s, dout, dout.float().sum(0), wq, bq, wp, bp, alpha, da, dwq, dbq, dwp)
        n_i2t = len(K.rec)
        ox = torch.empty(B * S, C, dtype=torch.bfloat16, device=DEV)
        s2 = XR.t2i_fwd(K, q, x, wk, wv, bv, ox, B, N, H)
        n_t2i_fwd = len(K.rec)
        dwk, dwv, dbv = (torch.zeros(n, device=DEV) for n in ((C, C), (C, C), (C,)))
        dx = torch.empty(B * N, C, device=DEV)
        XR.t2i_bwd(K, s2, dox, wk, wv, bv, dwk, dwv, dbv, dx)
        torch.cuda.synchronize()
        if it == 0:
            continue   # warm-up (tensor-map encodes, module loads)
        for i, (sig, fl, e0, e1) in enumerate(K.rec):
            times.setdefault((i, sig, fl), []).append(e0.elapsed_time(e1) * 1e3)
    bounds = [(0, "i2t fwd"), (n_i2t_fwd, "i2t bwd"), (n_i2t, "t2i fwd"), (n_t2i_fwd, "t2i bwd")]
    tot = 0.0
    section = collections.OrderedDict()
    for (i, sig, fl), ts in times.items():
        for b0, name in bounds:
            if i == b0:
                print("---- %s" % name)
        us = statistics.median(ts)
        cur = [name for b0, name in bounds if i >= b0][-1]
        section[cur] = section.get(cur, 0.0) + us
        print("%-62s %8.1f us %8.1f TF/s" % (sig, us, fl / us / 1e6 if fl else 0.0))
    print("---- totals (us):", {k: round(v, 1) for k, v in section.items()})


if __name__ == "__main__":
    main()
