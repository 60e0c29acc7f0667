"""Fused AdamW over the flat parameter arena, with the reference's six name-based hyper-parameter groups
(EgoVLPv2/set_optim_schedule.py:16-129: decay / no-decay x backbone / heads / cross-modal) and its HF cosine schedule
with warm-up.  One kernel launch per group updates the fp32 master weights, both moments and the bf16 operand copies."""
import math

from . import lib as _lib
from .weights import ParamArena, cache

NO_DECAY = ["bias", "LayerNorm.bias", "LayerNorm.weight", "norm.bias", "norm.weight", "norm1.bias", "norm1.weight",
            "norm2.bias", "norm2.weight"]
HEAD_NAMES = ["mlm_score", "itm_score", "txt_proj", "vid_proj"]
CROSS_MODAL_NAMES = ["cross_modal", "i2t", "t2i"]


def param_groups(model, lr, weight_decay, lr_mult_head=1.0, lr_mult_cross_modal=1.0):
    """Same membership rules (substring matches on parameter names) as set_optim_schedule.py:38-106."""
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    groups = []
    for kind, mult in (("backbone", 1.0), ("head", lr_mult_head), ("cross", lr_mult_cross_modal)):
        for decay in (True, False):
            ps = []
            for n, p in named:
                nd = any(s in n for s in NO_DECAY)
                hd = any(s in n for s in HEAD_NAMES)
                cm = any(s in n for s in CROSS_MODAL_NAMES)
                k = "cross" if cm else ("head" if hd else "backbone")
                if k == kind and (not nd) == decay:
                    ps.append(p)
            groups.append(dict(name="%s_%s" % (kind, "decay" if decay else "nodecay"), params=ps,
                               weight_decay=weight_decay if decay else 0.0, lr=lr * mult))
    return groups


def adjacency_runs(model):
    """q/k/v (and cross-attention k/v) weights and biases of every RoBERTa layer, to be laid out back to back."""
    runs = []
    tm = getattr(model, "text_model", None)
    if tm is None:
        return runs
    for layer in tm.encoder.layer:
        sa = layer.attention.self
        runs.append([sa.query.weight, sa.key.weight, sa.value.weight])
        runs.append([sa.query.bias, sa.key.bias, sa.value.bias])
        if hasattr(layer, "crossattention_t2i"):
            ca = layer.crossattention_t2i.self
            runs.append([ca.key.weight, ca.value.weight])
            runs.append([ca.key.bias, ca.value.bias])
    return runs


class FusedAdamW:
    """transformers.AdamW semantics (betas (0.9, 0.98), eps 1e-8, correct_bias=True; set_optim_schedule.py:108)."""

    def __init__(self, model, lr, weight_decay, lr_mult_head=1.0, lr_mult_cross_modal=1.0, betas=(0.9, 0.98), eps=1e-8,
                 max_steps=None, warmup_steps=0):
        import torch
        self.groups = [g for g in param_groups(model, lr, weight_decay, lr_mult_head, lr_mult_cross_modal) if g["params"]]
        # a run of adjacent parameters must live inside one hyper-parameter group (q/k/v weights always do)
        self.arena = ParamArena([(g["name"], g["params"]) for g in self.groups], adjacent=adjacency_runs(model))
        cache().arena = self.arena
        cache().clear()
        self.m = torch.zeros_like(self.arena.master)
        self.v = torch.zeros_like(self.arena.master)
        self.betas, self.eps = betas, eps
        self.step_count = 0
        # {lr multiplier, 1-beta1^t, 1-beta2^t} on the device: lets a captured CUDA graph serve every step
        self.hyper_dev = torch.ones(3, dtype=torch.float32, device=self.arena.master.device)
        self.hyper_host = torch.ones(3, dtype=torch.float32)
        if self.arena.master.is_cuda:
            self.hyper_host = self.hyper_host.pin_memory()
        self.max_steps, self.warmup_steps = max_steps, warmup_steps
        self.arena.bind_grads()

    def lr_scale(self):
        """get_cosine_schedule_with_warmup (set_optim_schedule.py:115-119)."""
        if not self.max_steps:
            return 1.0
        s = self.step_count
        if s < self.warmup_steps:
            return s / max(1, self.warmup_steps)
        prog = (s - self.warmup_steps) / max(1, self.max_steps - self.warmup_steps)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * min(1.0, prog))))

    def zero_grad(self):
        _lib.kernels().zero(self.arena.grad)
        self.arena.bind_grads()

    def advance(self):
        """Host part of a step: bump the counter and upload the step-dependent scalars (async H2D, outside any graph)."""
        scale = self.lr_scale()
        self.step_count += 1
        self.hyper_host[0] = scale
        self.hyper_host[1] = 1.0 - self.betas[0] ** self.step_count
        self.hyper_host[2] = 1.0 - self.betas[1] ** self.step_count
        self.hyper_dev.copy_(self.hyper_host, non_blocking=True)

    def launch(self, grad_scale=1.0):
        """Device part of a step (capturable): one fused AdamW launch per hyper-parameter group."""
        K = _lib.kernels()
        a = self.arena
        for g, (name, lo, hi) in zip(self.groups, a.ranges):
            K.adamw(a.master[lo:hi], a.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], a.shadow[lo:hi], g["lr"],
                    self.betas[0], self.betas[1], self.eps, g["weight_decay"], 1, grad_scale=grad_scale,
                    hyper_dev=self.hyper_dev)

    def step(self, grad_scale=1.0):
        self.advance()
        self.launch(grad_scale)
