"""bf16 operand copies of the fp32 master parameters, and the flat parameter arena.

The reference keeps fp32 nn.Parameters and lets autocast cast them on every use (base_trainer.py:334,
trainer_egoclip.py:143).  Here every GEMM reads a persistent bf16 copy:

* `WeightCache` -- per-parameter bf16 copies refreshed by the cast kernel whenever the parameter's autograd
  version counter (bumped by any in-place optimiser update) or storage changes.  Works with any optimiser.
* `ParamArena` -- all parameters of a model re-homed as views of ONE flat fp32 buffer, with a flat bf16 shadow
  and a flat fp32 gradient buffer.  The fused AdamW kernel then updates master weights and refreshes the bf16
  operands in a single launch per hyper-parameter group, and q/k/v (and cross-attention k/v) weights are laid out
  adjacently so their concatenations are plain views.
"""
import weakref

import torch

from . import functional as _F
from . import lib as _lib


class WeightCache:
    """Freshness: a copy is reused while the parameter object is the SAME (weak reference: a recycled id() of a freed
    model never matches), its autograd version counter is unchanged and its storage has not moved.  An update that goes
    through `p.data` or a raw pointer (older HF AdamW, apex / DeepSpeed fused optimisers, EMA via p.data.copy_) does not bump
    the version counter: call `invalidate()` after such a step, or `attach(optimizer)` once -- it registers a post-step
    hook that does."""

    def __init__(self):
        self._single = {}
        self._cat = {}
        self._arenas = []

    @property
    def arena(self):
        """the most recently created live parameter arena (FusedAdamW), or None"""
        self._arenas = [r for r in self._arenas if r() is not None]
        return self._arenas[-1]() if self._arenas else None

    @arena.setter
    def arena(self, a):   # tests reset with `cache().arena = None`
        self._arenas = [] if a is None else [weakref.ref(a)]

    def add_arena(self, a):
        """several models / optimisers may be alive at once: every arena serves its own parameters"""
        self._arenas = [r for r in self._arenas if r() is not None] + [weakref.ref(a)]
        self.clear()

    def _arena_of(self, p):
        for r in reversed(self._arenas):
            a = r()
            if a is not None and a.owns(p):
                return a
        return None

    def clear(self):
        self._single.clear()
        self._cat.clear()

    def invalidate(self):
        """forget every cached copy (next use re-casts) and refresh the arenas' bf16 shadows from their fp32 masters"""
        self.clear()
        for r in self._arenas:
            a = r()
            if a is not None:
                a.refresh_shadow()

    def attach(self, optimizer):
        """keep the copies fresh under an external torch optimiser, whatever way it writes the parameters"""
        return optimizer.register_step_post_hook(lambda *_: self.invalidate())

    @staticmethod
    def _stamp(p):
        return (p._version, p.data_ptr(), p.device)

    def bf16(self, p, shape2d=None):
        """bf16 copy of parameter `p` (optionally viewed as 2-D `shape2d`)."""
        a = self._arena_of(p)
        if a is not None:
            a.touch(p)
            v = a.bf16_view(p)
            return v.view(shape2d) if shape2d is not None else v
        key = id(p)
        ent = self._single.get(key)
        st = self._stamp(p)
        if ent is not None and ent[2]() is not p:
            ent = None        # id() recycled by another tensor
        if ent is None or ent[0] != st:
            buf = ent[1] if ent is not None and ent[1].numel() == p.numel() and ent[1].device == p.device else \
                torch.empty(p.shape, dtype=_F.BF16, device=p.device)
            _lib.kernels().cast(p.detach().contiguous(), buf)
            ent = (st, buf, weakref.ref(p))
            self._single[key] = ent
        return ent[1].view(shape2d) if shape2d is not None else ent[1]

    def _cat_arena(self, params, bf16):
        a = self._arena_of(params[0])
        if a is None:
            return None
        v = a.cat_view(params, bf16=bf16)
        if v is not None:
            for p in params:
                a.touch(p)
        return v

    @staticmethod
    def _same(ent, params):
        return ent is not None and all(r() is p for r, p in zip(ent[2], params))

    def cat_bf16(self, params):
        """bf16 copy of torch.cat(params, 0) (2-D weights with equal inner size)."""
        v = self._cat_arena(params, True)
        if v is not None:
            return v
        key = tuple(id(p) for p in params)
        st = tuple(self._stamp(p) for p in params)
        ent = self._cat.get(key)
        if not self._same(ent, params):
            ent = None
        if ent is None or ent[0] != st:
            rows = sum(p.shape[0] for p in params)
            buf = ent[1] if ent is not None else torch.empty((rows,) + tuple(params[0].shape[1:]), dtype=_F.BF16,
                                                             device=params[0].device)
            r = 0
            for p in params:
                _lib.kernels().cast(p.detach().contiguous(), buf[r:r + p.shape[0]])
                r += p.shape[0]
            ent = (st, buf, [weakref.ref(p) for p in params])
            self._cat[key] = ent
        return ent[1]

    def cat_f32(self, params):
        """fp32 torch.cat(params, 0) of small vectors (biases); a view when the arena laid them out adjacently."""
        v = self._cat_arena(params, False)
        if v is not None:
            return v
        key = ("f32",) + tuple(id(p) for p in params)
        st = tuple(self._stamp(p) for p in params)
        ent = self._cat.get(key)
        if not self._same(ent, params):
            ent = None
        if ent is None or ent[0] != st:
            ent = (st, torch.cat([p.detach() for p in params], 0), [weakref.ref(p) for p in params])
            self._cat[key] = ent
        return ent[1]


_CACHE = WeightCache()


def cache():
    return _CACHE


class ParamArena:
    """Flat fp32 master / bf16 operand / fp32 gradient buffers for all parameters of `model`.

    `groups`: list of (name, [parameters]) -- one contiguous range per optimiser hyper-parameter group, in the
    order given.  Inside a group, `adjacent` lists of parameters (q/k/v weights ...) are placed back to back."""

    ALIGN = 64  # elements: keeps every parameter 128-byte (bf16) / 256-byte (fp32) aligned

    def __init__(self, groups, adjacent=()):
        device = groups[0][1][0].device
        adj_of = {}
        for run in adjacent:
            for p in run:
                adj_of[id(p)] = run
        self.offsets = {}
        self.ranges = []
        off = 0
        placed = set()
        for name, params in groups:
            start = off
            for p in params:
                if id(p) in placed:
                    continue
                run = adj_of.get(id(p), [p])
                for q in run:
                    assert id(q) not in placed, "parameter listed in two adjacency runs / groups"
                    self.offsets[id(q)] = (off, q.numel(), tuple(q.shape))
                    placed.add(id(q))
                    off += q.numel()
                off = (off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            self.ranges.append((name, start, off))
        self.numel = off
        self.master = torch.zeros(off, dtype=torch.float32, device=device)
        self.shadow = torch.zeros(off, dtype=_F.BF16, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device)
        self.params = [p for _, ps in groups for p in ps]
        self._refs = {id(p): weakref.ref(p) for p in self.params}
        self.touched = set()      # ids of the parameters used by a forward pass since the last zero_grad()
        for p in self.params:
            o, n, shp = self.offsets[id(p)]
            self.master[o:o + n].copy_(p.detach().reshape(-1))
            p.data = self.master[o:o + n].view(shp)
        self.refresh_shadow()

    def owns(self, p):
        r = self._refs.get(id(p))
        return r is not None and r() is p

    def touch(self, p):
        self.touched.add(id(p))

    def refresh_shadow(self):
        """bf16 operand shadow <- fp32 masters (after anything but the fused AdamW kernel wrote the parameters:
        load_state_dict, re-initialisation, an external optimiser)"""
        _lib.kernels().cast(self.master, self.shadow)

    def bf16_view(self, p):
        ent = self.offsets.get(id(p))
        if ent is None:
            return None
        o, n, shp = ent
        return self.shadow[o:o + n].view(shp)

    def grad_view(self, p):
        o, n, shp = self.offsets[id(p)]
        return self.grad[o:o + n].view(shp)

    def bind_grads(self):
        """Point every p.grad at its slice of the flat gradient buffer (autograd then accumulates in place)."""
        for p in self.params:
            p.grad = self.grad_view(p)

    def grad_cat_view(self, params):
        """gradient-buffer view of torch.cat(params, 0) when the arena laid them out adjacently, else None"""
        return self.cat_view(params, bf16=False, buf=self.grad)

    def cat_view(self, params, bf16, buf=None):
        ents = [self.offsets.get(id(p)) for p in params]
        if any(e is None for e in ents):
            return None
        off = ents[0][0]
        for o, n, _ in ents:
            if o != off:
                return None
            off += n
        if buf is None:
            buf = self.shadow if bf16 else self.master
        rows = sum(e[2][0] for e in ents)
        return buf[ents[0][0]:off].view((rows,) + ents[0][2][1:])
