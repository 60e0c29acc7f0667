"""Gradient all-reduce of the flat arena, overlapped with the backward (reference: DistributedDataParallel's bucketed
reducer, base_trainer.py:267-269).

The three passes of a step share their parameters, and autograd replays them last-created-first: ITM pass, MLM pass, then
the EgoNCE pass.  A parameter's gradient is therefore FINAL once the EgoNCE pass's backward has gone through its module --
for the parameters only the fused passes use (cross-attention layers, MLM / ITM heads, FrozenInTime.cls_token / norm) as
soon as the EgoNCE pass's backward starts.  The towers register a hook on the input of every block / layer of the EgoNCE
pass (`watch`); when its gradient arrives, the ranges of that block's parameters in the flat gradient buffer are
all-reduced on a communication stream while the backward carries on.  `finish()` reduces what is left (embeddings, patch
embedding, ...) and joins the streams.  The arena is flat, so a bucket is a list of slices: no copies, no flattening.

Works without CUDA too (gloo, synchronous): the 2-rank CPU tests pin that the bucketed result equals one all-reduce of the
whole buffer bit for bit.

EGV_ALLREDUCE_BF16=1 (opt-in, off by default): every bucket travels as bf16 -- half the bytes over NVLink -- and is written
back into the fp32 gradient buffer (AdamW keeps accumulating in fp32).  This changes the reduction's rounding (each rank's
partial gradient is rounded to 8 mantissa bits before the sum; the reference's DDP reduces fp32), so it is not the default;
the 2-rank CPU test bounds the difference.  Not measured at N = 8."""
import os

import torch
import torch.distributed as dist

_active = None
FUSED_ONLY = ("i2t", "t2i", "cross_modal", "mlm_score", "itm_score")


def active():
    return _active


class OverlappedGradReducer:
    def __init__(self, model, arena, enabled=True):
        self.arena, self.enabled = arena, enabled
        self.cuda = arena.grad.is_cuda
        self.comm = torch.cuda.Stream(device=arena.grad.device) if self.cuda else None
        self.done = []
        self.calls = 0
        self.bf16 = os.environ.get("EGV_ALLREDUCE_BF16", "0") == "1"
        names = {id(p): n for n, p in model.named_parameters()}
        fused = [p for p in arena.params if any(s in names.get(id(p), "") for s in FUSED_ONLY)
                 or names.get(id(p), "") in ("cls_token", "norm.weight", "norm.bias")]
        self.fused_ranges = self.ranges_of(fused)
        self._fused_sent = False

    # ------------------------------------------------------------------ ranges
    def ranges_of(self, params):
        a = self.arena
        spans = sorted((a.offsets[id(p)][0], (a.offsets[id(p)][0] + a.offsets[id(p)][1] + a.ALIGN - 1) // a.ALIGN * a.ALIGN)
                       for p in params if id(p) in a.offsets)
        out = []
        for lo, hi in spans:
            if out and lo <= out[-1][1]:
                out[-1][1] = max(out[-1][1], hi)
            else:
                out.append([lo, hi])
        return [(lo, hi) for lo, hi in out]

    @staticmethod
    def _subtract(ranges, done):
        out = []
        for lo, hi in ranges:
            cur = lo
            for dlo, dhi in sorted(done):
                if dhi <= cur or dlo >= hi:
                    continue
                if dlo > cur:
                    out.append((cur, dlo))
                cur = max(cur, dhi)
            if cur < hi:
                out.append((cur, hi))
        return out

    # ------------------------------------------------------------------ step protocol
    def begin_step(self):
        global _active
        self.done, self._fused_sent, self.calls = [], False, 0
        _active = self if self.enabled else None

    def watch(self, tensor, module):
        """`tensor` is the input of `module` in the EgoNCE pass: when its gradient arrives, module's parameter gradients
        are final"""
        if tensor.requires_grad:
            ranges = self.ranges_of(list(module.parameters()))
            tensor.register_hook(lambda g, r=ranges: self._on_final(r))

    @staticmethod
    def _union(ranges):
        out = []
        for lo, hi in sorted(ranges):
            if out and lo <= out[-1][1]:
                out[-1][1] = max(out[-1][1], hi)
            else:
                out.append([lo, hi])
        return [(lo, hi) for lo, hi in out]

    def _on_final(self, ranges):
        if not self._fused_sent:      # the fused passes ran their whole backward before the EgoNCE pass's first node
            ranges = ranges + self.fused_ranges
            self._fused_sent = True
        ranges = self._subtract(self._union(ranges), self.done)   # a fused block's own cross-attention parameters are in both lists
        self._reduce(ranges)
        self.done = self._union(self.done + ranges)
        return None

    def _reduce(self, ranges):
        if not ranges:
            return
        g = self.arena.grad
        if self.cuda:
            from . import streams
            self.comm.wait_stream(torch.cuda.current_stream())
            if streams.enabled():
                self.comm.wait_stream(streams._side_stream())
            with torch.cuda.stream(self.comm):
                for lo, hi in ranges:
                    self._all_reduce(g[lo:hi])
        else:
            for lo, hi in ranges:
                self._all_reduce(g[lo:hi])
        self.calls += len(ranges)

    def _all_reduce(self, view):
        if self.bf16:
            t = view.to(torch.bfloat16)
            dist.all_reduce(t)
            view.copy_(t)
        else:
            dist.all_reduce(view)

    def finish(self):
        """after loss.backward() (and streams.join()): reduce the rest, make the main stream wait for the communication"""
        global _active
        _active = None
        rest = self._subtract([(0, self.arena.numel)], self.done)
        self._reduce(rest)
        self.done = self._union(self.done + rest)
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.comm)
