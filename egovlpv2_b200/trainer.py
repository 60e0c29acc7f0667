"""One EgoNCE+MLM+ITM pre-training step (trainer_egoclip.py:106-150 + base_trainer.py:324-393) on the B200 kernels:
H2D of the batch, FrozenInTime.forward (three passes), backward, gradient all-reduce (N > 1), fused AdamW.

This is the host driver used by bench.py and __graft_entry__.smoke(); the reference's own trainer can drive the same
model unchanged."""
import os
import types

import torch
import torch.distributed as dist

from . import lib as _lib
from . import streams
from .comm import NcclAllGather, P2PAllGather
from .model.loss import EgoNCE
from .model.model import DEFAULT_CONFIG, FrozenInTime
from .optim import FusedAdamW


def model_config(C=768, heads=12, depth=12, n_fuse=6, vocab=50265):
    return dict(DEFAULT_CONFIG, input_image_embed_size=C, input_text_embed_size=C, hidden_size=C, num_heads=heads,
                num_layers=depth, num_fuse_block=n_fuse, vocab_size=vocab)


def build_model(T=16, img=224, C=768, heads=12, depth=12, n_fuse=6, vocab=50265, proj=4096, tasks="EgoNCE_ITM_MLM", patch=16):
    return FrozenInTime(
        video_params=dict(model="SpaceTimeTransformer", arch_config="base_patch16_224", num_frames=T, pretrained=True,
                          time_init="zeros", img_size=img, patch_size=patch, embed_dim=C, depth=depth, num_heads=heads),
        text_params=dict(model="roberta-base", pretrained=True, input="text", allow_random_init=True,   # synthetic benchmarks / tests
                         config=dict(hidden_size=C, num_hidden_layers=depth, num_attention_heads=heads,
                                     intermediate_size=4 * C, vocab_size=vocab)),
        projection_dim=proj, config=model_config(C, heads, depth, n_fuse, vocab), task_names=tasks, embed_dim=C)


def randomize_gates(model, seed=0):
    """The reference zero-initialises the fusion gates and the time attention, which would let a
    benchmark skip real work numerically; give them non-trivial values (same recipe as the parity oracle)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("alpha_i2t") or n.endswith("alpha_t2i"):
                p.fill_(0.5)
            elif ".timeattn." in n:
                p.copy_((0.02 * torch.randn(p.shape, generator=g)).to(p.device))


class PretrainStep:
    def __init__(self, model, device, lr=3e-5, weight_decay=0.01, lr_mult_head=1.0, lr_mult_cross_modal=4.0,
                 tasks="EgoNCE_MLM_ITM", gather="auto", max_steps=None, warmup_steps=0):
        self.device = device
        self.model = model.to(device)
        self.tasks = tasks
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.opt = FusedAdamW(self.model, lr, weight_decay, lr_mult_head, lr_mult_cross_modal, max_steps=max_steps,
                              warmup_steps=warmup_steps)
        self.loss_fn = EgoNCE()
        self.args = types.SimpleNamespace(world_size=self.world, rank=self.rank)
        self.gather_kind = "none"
        if self.world > 1:
            if gather in ("auto", "p2p"):
                try:
                    self.allgather = P2PAllGather(device)
                    self.gather_kind = "p2p"
                except Exception:
                    if gather == "p2p":
                        raise
                    self.allgather = NcclAllGather()
                    self.gather_kind = "nccl"
            else:
                self.allgather = NcclAllGather()
                self.gather_kind = "nccl"
        else:
            self.allgather = lambda t, n=None, a=None: t
        self.cfg = {"loss": {"type": "EgoNCE"}}
        # gradient all-reduce in buckets on a communication stream, overlapped with the backward (reduce.py);
        # EGV_OVERLAP_ALLREDUCE=0: one all-reduce of the whole flat buffer after the backward (round 1)
        import os
        from .reduce import OverlappedGradReducer
        self.reducer = OverlappedGradReducer(self.model, self.opt.arena,
                                             enabled=os.environ.get("EGV_OVERLAP_ALLREDUCE", "1") != "0") if self.world > 1 else None
        self.two_streams = streams.enable(True)   # text tower on a side stream (env EGV_TEXT_STREAM=0 turns it off)

    def to_device(self, host_batch):
        """H2D of one step's inputs from pinned host memory (non-blocking on the current stream)."""
        return {k: v.to(self.device, non_blocking=True) for k, v in host_batch.items()}

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def capture(self, example_batch, warmup=3):
        """Capture forward + backward + all-reduce + AdamW of one step into a CUDA graph with static input buffers.
        Afterwards `step_graph(batch)` copies the batch into the static buffers and replays the graph: one host call
        instead of ~3000 kernel launches and the Python autograd bookkeeping."""
        self.static = {k: v.clone() for k, v in example_batch.items()}
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._device_step(self.static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        # the warm-up's cached blocks belong to the side stream's pool; the capture allocates a private pool of the same
        # size next to them (cfg 5 at B = 2: 57 + 119 GB): give them back first
        torch.cuda.empty_cache()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.kernels().launch_count()
        # EGV_MAIN_PRIORITY=1: capture on a HIGH-priority stream (the text tower's side stream and the communication
        # stream keep the default, lower priority).  Measured on B200: no difference (87.1 / 87.5 vs 87.6 ms) -- off.
        cap_stream = None
        if os.environ.get("EGV_MAIN_PRIORITY", "0") == "1":
            cap_stream = torch.cuda.Stream(device=self.device, priority=-1)
        with (torch.cuda.graph(self.graph, stream=cap_stream) if cap_stream is not None else torch.cuda.graph(self.graph)):
            self.static_loss, self.static_loss_dict = self._device_step(self.static)
        self.launches_per_step = _lib.kernels().launch_count() - n0   # kernels of this library inside one replay
        return self.graph

    def step_graph(self, batch):
        """`batch`: device or pinned-host tensors (copied into the static input buffers), then one graph replay."""
        for k, v in batch.items():
            if v is not self.static[k]:
                self.static[k].copy_(v, non_blocking=True)
        self.opt.step_count += 1     # host mirror; the device counter advances inside the graph
        self.graph.replay()
        return self.static_loss, self.static_loss_dict

    # ------------------------------------------------------------------ input prefetch for the captured step
    def prefetch(self, host_batch):
        """Start the H2D copy of a (pinned) host batch into a staging buffer on a copy stream; it overlaps whatever the
        compute stream is doing.  `step_graph_prefetched()` then moves it into the graph's static inputs (a device-side
        copy) and replays.  The reference gets the same overlap from its DataLoader workers + non_blocking .cuda()."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staging = {k: torch.empty_like(v) for k, v in self.static.items()}
            self._h2d_done = torch.cuda.Event()
            self._staging_free = torch.cuda.Event()
            self._staging_free.record()
        cs = self._copy_stream
        cs.wait_event(self._staging_free)          # the previous step's device-side copy has drained the staging buffer
        with torch.cuda.stream(cs):
            for k, v in host_batch.items():
                self._staging[k].copy_(v, non_blocking=True)
            self._h2d_done.record(cs)

    def step_graph_prefetched(self):
        """Consume the batch started with `prefetch()`: staging -> static inputs, then one graph replay."""
        main = torch.cuda.current_stream()
        main.wait_event(self._h2d_done)
        for k, v in self._staging.items():
            self.static[k].copy_(v, non_blocking=True)
        self._staging_free.record(main)
        self.opt.step_count += 1
        self.graph.replay()
        return self.static_loss, self.static_loss_dict

    def step(self, dev_batch):
        """forward + backward + (all-reduce) + AdamW, eager launches; returns the detached total loss (device scalar)."""
        return self._device_step(dev_batch)

    def _device_step(self, dev_batch):
        d = dev_batch
        self.opt.advance()      # device-side step counter / schedule scalars: a kernel, captured with the step
        data = {"video": d["video"], "text": {"input_ids": d["input_ids"], "attention_mask": d["attention_mask"]},
                "text_mlm_ids": d["text_mlm_ids"], "text_mlm_labels": d["text_mlm_labels"]}
        self.opt.zero_grad()
        if self.reducer is not None:
            self.reducer.begin_step()
        loss, loss_dict, _ = self.model(data, d["noun_vec"], d["verb_vec"], self.allgather, self.world, self.args, self.cfg,
                                        self.loss_fn, self.rank, task_names=self.tasks)
        loss.backward()
        # the backward of the text tower ran on the side stream and wrote into the gradient arena directly
        streams.join()
        if self.world > 1:
            self.reducer.finish()                         # DDP semantics: average over ranks (base_trainer.py:269)
            self.opt.launch(grad_scale=1.0 / self.world)
        else:
            self.opt.launch()
        return loss.detach(), {k: v.detach() for k, v in loss_dict.items()}


def build_dual_model(T=16, img=224, C=768, heads=12, depth=12, n_fuse=6, vocab=50265):
    """model_epic_charades.FrozenInTime as configs/ft/epic.json / charades.json build it (projection 'minimal')."""
    from .model.model_epic_charades import FrozenInTime as DualFrozenInTime
    return DualFrozenInTime(
        video_params=dict(model="SpaceTimeTransformer", arch_config="base_patch16_224", num_frames=T, pretrained=True,
                          time_init="zeros", drop_path_rate=0.0, img_size=img, embed_dim=C, depth=depth, num_heads=heads),
        text_params=dict(model="roberta-base", pretrained=True, input="text", allow_random_init=True,
                         config=dict(hidden_size=C, num_hidden_layers=depth, num_attention_heads=heads,
                                     intermediate_size=4 * C, vocab_size=vocab)),
        projection="minimal", config=model_config(C, heads, depth, n_fuse, vocab), task_names="EgoNCE_ITM_MLM", embed_dim=C)


class FinetuneStep(PretrainStep):
    """One fine-tuning step of the dual-encoder mains (trainer_epic.py:128-150 / trainer_charades.py): H2D, forward
    (task_names='Dual'), backward, gradient all-reduce, fused AdamW.  Same graph capture / prefetch machinery as the
    pre-training step; the batch additionally carries `relation` [B] for dataset_name='epic'."""

    def __init__(self, model, device, loss_fn, dataset_name="epic", **kw):
        super().__init__(model, device, tasks="Dual", **kw)
        self.loss_fn, self.dataset_name = loss_fn, dataset_name

    def _device_step(self, dev_batch):
        d = dev_batch
        data = {"video": d["video"], "text": {"input_ids": d["input_ids"], "attention_mask": d["attention_mask"]}}
        if self.dataset_name == "epic":
            data["relation"] = d["relation"]
        self.opt.zero_grad()
        loss, loss_dict, _ = self.model(data, self.allgather, self.world, self.args, self.cfg, self.loss_fn, self.rank,
                                        task_names="Dual", dataset_name=self.dataset_name)
        loss.backward()
        streams.join()
        if self.world > 1:
            dist.all_reduce(self.opt.arena.grad)
            self.opt.launch(grad_scale=1.0 / self.world)
        else:
            self.opt.launch()
        return loss.detach(), {k: v.detach() for k, v in loss_dict.items()}


def step_flops(B, T, img=224, patch=16, S=32, C=768, depth=12, n_fuse=6, P=4096, V=50265):
    """Algorithmic FLOPs of one step (fwd + bwd = 3x fwd, no recompute) -- SURVEY.md section 8(d)."""
    Nf = (img // patch) ** 2
    N = T * Nf + 1
    vblk = 32 * N * C * C + 4 * (N - 1) * C * ((T + 1) + (Nf + 1)) + 8 * N * C
    xattn = 4 * N * C * C + 4 * S * C * C + 4 * N * S * C
    tblk = 24 * S * C * C + 4 * S * S * C
    patch_f = 2 * T * Nf * (3 * patch * patch) * C
    proj = 2 * (C * P + 2 * P * P)
    mlm_head = 2 * S * C * C + S * (2 * C * C + 2 * C * V)
    itm_head = 8 * C * C
    pass_nce = patch_f + depth * vblk + depth * tblk + 2 * proj
    pass_fused = patch_f + depth * vblk + n_fuse * xattn + depth * tblk + n_fuse * xattn
    fwd = pass_nce + (pass_fused + mlm_head) + (pass_fused + itm_head)
    return 3.0 * B * fwd, dict(vblk=vblk, xattn=xattn, fwd_per_sample=fwd)
