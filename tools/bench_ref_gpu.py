#!/usr/bin/env python
"""MEASUREMENT INFRASTRUCTURE ONLY -- the reference's own 1-GPU PyTorch eager training step (the ">= 6x" denominator of
BASELINE.json's north star; BASELINE.md B1/B2, SURVEY.md 8(d) "Reference GPU baseline").

    python tools/install_ref.py                       # build container: stage the unmodified reference under baseline/_ref
    gpurun -- python tools/bench_ref_gpu.py --cfg 3   # GPU box

Runs the UNMODIFIED reference model code (EgoVLPv2/model/*.py through oracle/ref_shim.py, which only stubs absent
third-party names: timm DropPath / trunc_normal_, removed transformers helpers, the hard-coded ViT checkpoint path) the
way trainer_egoclip.py:139-149 + base_trainer.py:267-269,334 drive it: yaml defaults (use_checkpoint: True),
torch.cuda.amp.autocast() fp16 + GradScaler, DistributedDataParallel(static_graph=True), AdamW(betas=(0.9, 0.98), eps=1e-8)
over the six name-based groups of set_optim_schedule.py:20-106, model.train() (text dropout active), the same synthetic
batch as bench.py.  None of this repository's kernels, models or engine is on this path.  CUDA-event timing, >= 3 warm-up
steps, nvidia-smi clocks sampled during the timed region.  One JSON line per configuration."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFGS = {
    2: dict(frames=4, tasks="EgoNCE", desc="cfg 2: TimeSformer-B/16 + RoBERTa-base, 4 frames 224^2, seq=32, bs=8, fusion OFF (dual-encoder, EgoNCE only)"),
    3: dict(frames=16, tasks="EgoNCE_MLM_ITM", desc="cfg 3: TimeSformer-B/16 + RoBERTa-base, 16 frames 224^2, seq=32, bs=8, fusion ON (top-6), EgoNCE+MLM+ITM"),
}


def groups(model, lr=3e-5, wd=0.01, mult_head=1.0, mult_cross=4.0):
    """the reference's parameter grouping rule (set_optim_schedule.py:20-106), restated: substring matches on names"""
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight", "norm.bias", "norm.weight", "norm1.bias", "norm1.weight",
                "norm2.bias", "norm2.weight"]
    head, cross = ["mlm_score", "itm_score", "txt_proj", "vid_proj"], ["cross_modal", "i2t", "t2i"]
    out = []
    for is_head, is_cross, mult in ((False, False, 1.0), (True, False, mult_head), (False, True, mult_cross)):
        for nd in (False, True):
            ps = [p for n, p in model.named_parameters()
                  if any(x in n for x in no_decay) == nd and any(x in n for x in head) == is_head
                  and any(x in n for x in cross) == is_cross and not (is_head and is_cross)]
            out.append({"params": ps, "weight_decay": 0.0 if nd else wd, "lr": lr * mult})
    return out


class Clocks:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        self.p.wait(timeout=5)
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons)}


def run(cfg_id, a, ref):
    from oracle import egovlp_oracle as O
    c = CFGS[cfg_id]
    dev = torch.device("cuda", 0)
    mm, vt, rb = ref.mm, ref.vt, ref.rb
    for mod_cfg in (mm.config, vt.config_yaml, rb.config_yaml):
        mod_cfg["use_checkpoint"] = not a.no_checkpoint
    torch.manual_seed(0)
    torch.load = mm._shim_fake_load
    try:
        model = mm.FrozenInTime(
            video_params=dict(model="SpaceTimeTransformer", arch_config="base_patch16_224", num_frames=c["frames"],
                              pretrained=True, time_init="zeros"),
            text_params=dict(model="roberta-base", pretrained=True, input="text"),
            projection_dim=4096, config=dict(mm.config), task_names="EgoNCE_ITM_MLM")
    finally:
        torch.load = mm._shim_real_load
    with torch.no_grad():   # same non-trivial gates / time attention as bench.py's own arm
        g = torch.Generator().manual_seed(0)
        for n, p in model.named_parameters():
            if n.endswith("alpha_i2t") or n.endswith("alpha_t2i"):
                p.fill_(0.5)
            elif ".timeattn." in n:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    model = model.to(dev).train()
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[0], static_graph=True, find_unused_parameters=False)
    opt = torch.optim.AdamW(groups(model), lr=3e-5, eps=1e-8, betas=(0.9, 0.98))
    scaler = torch.amp.GradScaler("cuda")
    host = O.synthetic_batch(a.batch, c["frames"], 224, a.seq, seed=1234)
    d = {k: v.to(dev) for k, v in host.items()}
    args = types.SimpleNamespace(world_size=1, rank=0)
    loss_fn = ref.loss.EgoNCE()
    config = {"loss": {"type": "EgoNCE"}}

    def step():
        data = {"video": d["video"], "text": {"input_ids": d["input_ids"], "attention_mask": d["attention_mask"]},
                "text_mlm_ids": d["text_mlm_ids"], "text_mlm_labels": d["text_mlm_labels"]}
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            loss, loss_dict, ret = ddp(data, d["noun_vec"], d["verb_vec"], lambda t, n, ar: t, 1, args, config, loss_fn, 0,
                                       task_names=c["tasks"])
        scaler.scale(loss).backward()
        if isinstance(ret, dict):
            ret.clear()   # infer()'s mutable default `ret={}` would otherwise carry last step's graph into the next
        scaler.step(opt)
        scaler.update()
        return loss

    for _ in range(max(3, a.warmup)):
        step()
    torch.cuda.synchronize()
    clocks = Clocks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    line = {"impl": "reference_gpu_eager", "metric": "pretrain_clips_per_sec", "value": a.batch / (ms * 1e-3), "unit": "clips/s",
            "n_gpus": 1, "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": ms, "dtype": "fp16 autocast + GradScaler",
            "config": {"workload": c["desc"], "use_checkpoint": not a.no_checkpoint, "ddp_static_graph": True, "train_mode": True,
                       "optimizer": "torch.optim.AdamW, 6 name-based groups", "torch": torch.__version__},
            "last_loss": float(loss), "clocks": clocks.stop(),
            "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    print(json.dumps(line), flush=True)
    del ddp, model, opt
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, nargs="+", default=[3, 2])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-checkpoint", action="store_true", help="use_checkpoint: False (the yaml default is True)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from oracle import ref_shim
    root = ref_shim.REF_ROOT if os.path.isdir(ref_shim.REF_ROOT) else os.path.join(ROOT, "baseline", "_ref", "EgoVLPv2")
    if not os.path.isdir(root):
        print(json.dumps({"impl": "reference_gpu_eager", "unavailable": "run tools/install_ref.py in the build container first"}))
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29591")
    ref = ref_shim.load(use_checkpoint=not a.no_checkpoint, root=root)
    for cid in a.cfg:
        try:
            line = run(cid, a, ref)
        except Exception as e:   # keep the other configurations' numbers
            line = {"impl": "reference_gpu_eager", "cfg": cid, "error": repr(e)[:300]}
            print(json.dumps(line), flush=True)
        if a.out:
            with open(a.out, "a") as f:
                f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
