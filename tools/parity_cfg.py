#!/usr/bin/env python
"""TEST / MEASUREMENT INFRASTRUCTURE -- end-to-end parity of one training step at a BENCHMARKED configuration
(BASELINE.json cfg 2 / cfg 3: depth 12, C = 768, 224^2) against the fp32 oracle run on the same device, next to the
error of the reference-style fp16-autocast evaluation of the same oracle ("not worse than 2x the error of
the reference's own fp16-autocast run against the same fp32 oracle").

    python tools/parity_cfg.py --cfg 3 [--batch 2]      -> one JSON report on stdout

`report()` is what tests/test_model_gpu.py::test_benchmarked_config_step_vs_oracle asserts on."""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFGS = {2: dict(T=4, tasks="EgoNCE"), 3: dict(T=16, tasks="EgoNCE_MLM_ITM")}
# gradients sampled for the report: every module family on the path, early / middle / late layers
GRAD_KEYS = [
    "video_model.patch_embed.proj.weight", "video_model.pos_embed", "video_model.temporal_embed", "video_model.cls_token",
    "video_model.blocks.0.timeattn.qkv.weight", "video_model.blocks.0.attn.proj.weight", "video_model.blocks.3.mlp.fc1.weight",
    "video_model.blocks.5.norm1.weight", "video_model.blocks.6.attn.qkv_i2t.weight", "video_model.blocks.6.attn.alpha_i2t",
    "video_model.blocks.8.attn.qkv_text_i2t.weight", "video_model.blocks.9.attn.proj_i2t.weight",
    "video_model.blocks.11.mlp.fc2.weight", "video_model.blocks.11.attn.qkv.weight", "video_model.norm.weight",
    "text_model.embeddings.word_embeddings.weight", "text_model.encoder.layer.0.attention.self.query.weight",
    "text_model.encoder.layer.4.intermediate.dense.weight", "text_model.encoder.layer.7.crossattention_t2i.self.key.weight",
    "text_model.encoder.layer.7.alpha_t2i", "text_model.encoder.layer.9.crossattention_t2i.output.dense.weight",
    "text_model.encoder.layer.11.output.dense.weight", "txt_proj.0.weight", "vid_proj.4.weight",
    "cross_modal_text_transform.weight", "mlm_score.decoder.weight", "itm_score.fc.weight", "norm.weight",
]


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _oracle(O, data, sd, plan, tasks, autocast):
    # fp16 backward needs loss scaling; GradScaler starts at 2^16 and halves on every overflow -- same rule here
    scale = 65536.0 if autocast else 1.0
    while True:
        sdg = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
            out = O.pretrain_step(data, sdg, 12, 12, 6, plan, tasks=tasks)
        (out["loss_total"].float() * scale).backward()
        grads = {k: (sdg[k].grad / scale) for k in GRAD_KEYS if k in sdg and sdg[k].grad is not None}
        if all(bool(torch.isfinite(g).all()) for g in grads.values()) or scale <= 1.0:
            break
        scale /= 2.0
        del sdg, out, grads
    keep = {k: out[k].detach().float() for k in ("EgoNCE", "loss_mlm", "loss_itm", "loss_total", "sim_v2t", "text_embeds",
                                                  "video_embeds", "cross_attn_itm_logits") if k in out}
    keep["_loss_scale"] = scale
    if "cross_attn_mlm_logits" in out:
        keep["mlm_logits_slice"] = out["cross_attn_mlm_logits"].detach().float()[:, :, ::97].contiguous()
    return keep, grads


def report(cfg_id=3, B=2, S=32, seed=0):
    from egovlpv2_b200.model.loss import EgoNCE
    from egovlpv2_b200.trainer import build_model
    from oracle import egovlp_oracle as O
    c = CFGS[cfg_id]
    dev = torch.device("cuda", 0)
    shapes = O.key_shapes(T=c["T"])
    sd = O.seeded_state(shapes, seed, device=dev)
    data = O.synthetic_batch(B, c["T"], 224, S, seed=1234, device=dev)
    plan = {k: v.to(dev) for k, v in O.synthetic_itm_plan(B, seed=4321).items()}
    ref, gref = _oracle(O, data, sd, plan, c["tasks"], autocast=False)
    h16, g16 = _oracle(O, data, sd, plan, c["tasks"], autocast=True)
    torch.cuda.empty_cache()

    model = build_model(T=c["T"])
    model.load_state_dict(sd, strict=False)
    model.eval().to(dev)
    model.itm_plan = plan
    batch = {"video": data["video"], "text": {"input_ids": data["input_ids"], "attention_mask": data["attention_mask"]},
             "text_mlm_ids": data["text_mlm_ids"], "text_mlm_labels": data["text_mlm_labels"]}
    args = types.SimpleNamespace(world_size=1, rank=0)
    loss, ld, ret = model(batch, data["noun_vec"], data["verb_vec"], lambda t, n, a: t, 1, args, {"loss": {"type": "EgoNCE"}},
                          EgoNCE(), 0, task_names=c["tasks"])
    loss.backward()
    torch.cuda.synchronize()
    mine = {k: ld[k].detach().float() for k in ("EgoNCE", "loss_mlm", "loss_itm", "loss_total") if k in ld}
    mine.update(sim_v2t=ret["sim_v2t"].float(), text_embeds=ret["text_embeds"].float(), video_embeds=ret["video_embeds"].float())
    if "cross_attn_itm_logits" in ret:
        mine["cross_attn_itm_logits"] = ret["cross_attn_itm_logits"].float()
        mine["mlm_logits_slice"] = ret["cross_attn_mlm_logits"].float()[:, :, ::97]
    params = dict(model.named_parameters())
    out = {"cfg": cfg_id, "B": B, "T": c["T"], "tasks": c["tasks"], "fp16_loss_scale": h16["_loss_scale"], "loss": {}, "tensors": {},
           "grads": {}}
    for k in ("EgoNCE", "loss_mlm", "loss_itm", "loss_total"):
        if k in ref:
            r = float(ref[k])
            out["loss"][k] = {"oracle": r, "ours": float(mine[k]), "fp16_autocast": float(h16[k]),
                              "rel_ours": abs(float(mine[k]) - r) / max(abs(r), 1e-9),
                              "rel_fp16": abs(float(h16[k]) - r) / max(abs(r), 1e-9)}
    out["tensors"]["sim_v2t_abs"] = {"ours": (mine["sim_v2t"] - ref["sim_v2t"]).abs().max().item(),
                                     "fp16": (h16["sim_v2t"] - ref["sim_v2t"]).abs().max().item()}
    for k in ("text_embeds", "video_embeds", "cross_attn_itm_logits", "mlm_logits_slice"):
        if k in ref:
            out["tensors"][k] = {"ours": rel(mine[k], ref[k]), "fp16": rel(h16[k], ref[k])}
    for k in GRAD_KEYS:
        if k in gref and params[k].grad is not None:
            out["grads"][k] = {"ours": rel(params[k].grad, gref[k]), "fp16": rel(g16[k], gref[k]),
                               "finite": bool(torch.isfinite(params[k].grad).all())}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2)
    a = ap.parse_args()
    print(json.dumps(report(a.cfg, a.batch), indent=1))
