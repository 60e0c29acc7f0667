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
        cache().add_arena(self.arena)
        self.m = torch.zeros_like(self.arena.master)
        self.v = torch.zeros_like(self.arena.master)
        self.betas, self.eps = betas, eps
        self.step_count = 0      # host mirror of step_dev (exact in eager use; PretrainStep keeps it in step with graph replays)
        dev = self.arena.master.device
        # the step counter and {lr multiplier, 1-beta1^t, 1-beta2^t} live on the device and are advanced by a kernel
        # (egv_adamw_schedule): a captured CUDA graph serves every step and no host buffer can be overwritten under a
        # replay that is still in flight
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.hyper_dev = torch.ones(3, dtype=torch.float32, device=dev)
        self.max_steps, self.warmup_steps = max_steps, warmup_steps
        self.arena.bind_grads()
        # a checkpoint loaded AFTER the optimiser exists (the reference's _resume_checkpoint order) writes the fp32 masters in
        # place: refresh the bf16 operand shadow right behind it
        self._hook = model.register_load_state_dict_post_hook(lambda module, incompatible: self.arena.refresh_shadow())

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
        self.arena.touched.clear()

    def advance(self):
        """Start of an optimiser step (capturable: one tiny kernel): device step counter += 1, schedule scalars refreshed."""
        self.step_count += 1
        _lib.kernels().adamw_schedule(self.step_dev, self.hyper_dev, self.warmup_steps, self.max_steps, self.betas[0], self.betas[1])

    def active_ranges(self):
        """Per hyper-parameter group, the maximal runs of parameters that took part in a forward pass since the last
        zero_grad() (weights.ParamArena.touch).  The reference's optimiser skips parameters whose .grad is None
        (transformers.AdamW): e.g. the fusion layers and the MLM / ITM heads under task_names='Dual' get neither an Adam
        update nor weight decay.  When every parameter was used this is one range per group."""
        a = self.arena
        if not a.touched or len(a.touched) == len(a.params):
            return [(g, lo, hi) for g, (name, lo, hi) in zip(self.groups, a.ranges)]
        out = []
        for g in self.groups:
            run = None
            for o, n, pid in sorted((a.offsets[id(p)][0], a.offsets[id(p)][1], id(p)) for p in g["params"]):
                end = (o + n + a.ALIGN - 1) // a.ALIGN * a.ALIGN
                if pid in a.touched:
                    if run is not None and o <= run[1]:
                        run = (run[0], max(run[1], end))
                    else:
                        if run is not None:
                            out.append((g, run[0], run[1]))
                        run = (o, end)
                elif run is not None:
                    out.append((g, run[0], run[1]))
                    run = None
            if run is not None:
                out.append((g, run[0], run[1]))
        return out

    def launch(self, grad_scale=1.0):
        """Device part of a step (capturable): one fused AdamW launch per hyper-parameter group (per run of used parameters)."""
        K = _lib.kernels()
        a = self.arena
        for g, lo, hi in self.active_ranges():
            K.adamw(a.master[lo:hi], a.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], a.shadow[lo:hi], g["lr"],
                    self.betas[0], self.betas[1], self.eps, g["weight_decay"], 1, grad_scale=grad_scale,
                    hyper_dev=self.hyper_dev)

    # ------------------------------------------------------------------ checkpointing (base_trainer.py:_save / _resume_checkpoint)
    def state_dict(self):
        return {"m": self.m.clone(), "v": self.v.clone(), "step": int(self.step_dev.item()), "numel": self.arena.numel,
                "layout": [(name, lo, hi) for name, lo, hi in self.arena.ranges]}

    def load_state_dict(self, sd):
        if sd["numel"] != self.arena.numel or [tuple(x) for x in sd["layout"]] != [tuple(x) for x in self.arena.ranges]:
            raise ValueError("optimizer state was saved for a different parameter layout")
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.step_dev.fill_(int(sd["step"]))
        self.step_count = int(sd["step"])

    def step(self, grad_scale=1.0):
        self.advance()
        self.launch(grad_scale)
