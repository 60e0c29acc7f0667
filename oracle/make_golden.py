"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt by running the UNMODIFIED
reference (/root/reference, via oracle/ref_shim.py) on seeded inputs.

Run in the build container only:   python -m oracle.make_golden
The fixtures hold inputs (or the seeds that regenerate them), the reference's
outputs, and never weights: weights are regenerated from (schema, seed) by
`egovlp_oracle.seeded_state`, loaded into the reference with load_state_dict and
handed to the oracle as the same dict.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import egovlp_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY = dict(C=128, heads=2, depth=8, n_fuse=2, T=2, img=64, patch=16, S=8, B=4, proj=256, vocab=50265)


def _yaml_cfg(c):
    return dict(input_image_embed_size=c["C"], vocab_size=c["vocab"], mlm_prob=0.15,
                input_text_embed_size=c["C"], hidden_size=c["C"], num_heads=c["heads"],
                num_layers=c["depth"], mlp_ratio=4, drop_rate=0.1, num_fuse_block=c["n_fuse"],
                use_checkpoint=False)


def build_tiny_reference(ref, c, seed=0):
    text_cfg = dict(hidden_size=c["C"], num_hidden_layers=c["depth"], num_attention_heads=c["heads"],
                    intermediate_size=4 * c["C"])
    with ref_shim.tiny_wrapper(ref, img_size=c["img"], embed_dim=c["C"], depth=c["depth"],
                               num_heads=c["heads"], text_cfg=text_cfg, dim_cross=c["C"]):
        # text layers >= 12-NUM_FUSE_BLOCK get cross-attention (roberta.py:438); video blocks >= 6 (:302)
        ref.rb.NUM_FUSE_BLOCK = 12 - (c["depth"] - c["n_fuse"])
        model = ref.mm.FrozenInTime(
            video_params=dict(model="SpaceTimeTransformer", arch_config="base_patch16_224",
                              num_frames=c["T"], pretrained=True, time_init="zeros"),
            text_params=dict(model="roberta-base", pretrained=True, input="text"),
            projection_dim=c["proj"], config=_yaml_cfg(c), task_names="EgoNCE_ITM_MLM", embed_dim=c["C"])
    assert c["depth"] - c["n_fuse"] == 6, "reference hard-codes video fusion from block 6"
    shapes = O.key_shapes(C=c["C"], heads=c["heads"], depth=c["depth"], n_fuse=c["n_fuse"], T=c["T"],
                          img=c["img"], patch=c["patch"], vocab=c["vocab"], proj=c["proj"])
    ref_sd = model.state_dict()
    mine = set(shapes)
    theirs = {k for k in ref_sd if not k.endswith("position_ids")}
    assert mine == theirs, (sorted(mine - theirs)[:5], sorted(theirs - mine)[:5])
    for k in mine:
        assert tuple(ref_sd[k].shape) == tuple(shapes[k]), (k, ref_sd[k].shape, shapes[k])
    sd = O.seeded_state(shapes, seed)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.endswith("position_ids") for m in missing), (missing, unexpected)
    model.eval()
    return model, shapes, sd


class _Args:
    world_size = 1
    rank = 0


def _allgather(t, n_gpu, args):
    return t


def reference_step(ref, model, data, plan, grad=False):
    """Drive FrozenInTime.forward (model.py:370-487) with the host RNG of the ITM pass
    (:438 randperm, :459 np.random.rand, :460/465 multinomial) replaced by `plan`."""
    labels = plan["labels"]
    base = torch.cat([torch.ones(len(labels) // 2), torch.zeros(len(labels) - len(labels) // 2)])
    # find a permutation p with base[p] == labels
    ones = [i for i in range(len(base)) if base[i] == 1]
    zeros = [i for i in range(len(base)) if base[i] == 0]
    perm = torch.tensor([ones.pop() if l == 1 else zeros.pop() for l in labels.tolist()])
    neg_rows = [i for i in range(len(labels)) if labels[i] == 0]
    rand_seq = iter([0.9 if bool(plan["swap_video"][i]) else 0.1 for i in neg_rows])
    multi_seq = iter([int(plan["neg_idx"][i]) for i in neg_rows])
    real = (torch.randperm, np.random.rand, torch.multinomial)
    torch.randperm = lambda n, **k: perm
    np.random.rand = lambda *a: next(rand_seq)
    torch.multinomial = lambda w, n, **k: torch.tensor([next(multi_seq)])
    try:
        d = dict(video=data["video"].clone(),
                 text=dict(input_ids=data["input_ids"].clone(), attention_mask=data["attention_mask"].clone()),
                 text_mlm_ids=data["text_mlm_ids"].clone(), text_mlm_labels=data["text_mlm_labels"].clone())
        with torch.set_grad_enabled(grad):
            loss, loss_dict, ret = model(d, data["noun_vec"], data["verb_vec"], _allgather, 1, _Args(),
                                         {"loss": {"type": "EgoNCE"}}, ref.loss.EgoNCE(), 0,
                                         task_names="EgoNCE_MLM_ITM")
    finally:
        torch.randperm, np.random.rand, torch.multinomial = real
    return loss, loss_dict, ret


def _cmp(name, a, b, tol=2e-4):
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    print(f"  {name:28s} max|d|={err:.3e}  max|ref|={ref:.3e}")
    assert err <= tol * max(1.0, ref), name
    return err


def golden_tiny_step(ref):
    c = TINY
    model, shapes, sd = build_tiny_reference(ref, c, seed=0)
    data = O.synthetic_batch(c["B"], c["T"], c["img"], c["S"], seed=1234)
    plan = O.synthetic_itm_plan(c["B"], seed=4321)
    loss, ld, ret = reference_step(ref, model, data, plan)
    mine = O.pretrain_step(data, sd, c["heads"], c["depth"], c["n_fuse"], plan)
    print("tiny step (reference vs oracle):")
    _cmp("loss_total", mine["loss_total"], loss)
    _cmp("EgoNCE", mine["EgoNCE"], ld["EgoNCE"])
    _cmp("loss_mlm", mine["loss_mlm"], ld["loss_mlm"])
    _cmp("loss_itm", mine["loss_itm"], ld["loss_itm"])
    _cmp("sim_v2t", mine["sim_v2t"], ret["sim_v2t"])
    _cmp("itm_logits", mine["cross_attn_itm_logits"], ret["cross_attn_itm_logits"])
    _cmp("mlm_logits", mine["cross_attn_mlm_logits"], ret["cross_attn_mlm_logits"])
    _cmp("text_embeds", mine["text_embeds"], ret["text_embeds"])
    _cmp("video_embeds", mine["video_embeds"], ret["video_embeds"])
    mlm = ret["cross_attn_mlm_logits"]
    fx = dict(cfg=c, weight_seed=0, data_seed=1234, plan_seed=4321,
              loss_total=loss, EgoNCE=ld["EgoNCE"], loss_mlm=ld["loss_mlm"], loss_itm=ld["loss_itm"],
              sim_v2t=ret["sim_v2t"], itm_logits=ret["cross_attn_itm_logits"],
              text_embeds=ret["text_embeds"], video_embeds=ret["video_embeds"],
              mlm_logits_slice=mlm[:, :, ::997].clone(), mlm_logits_sum=mlm.double().sum(-1).float(),
              mlm_logits_lse=torch.logsumexp(mlm, -1))
    torch.save(fx, os.path.join(OUT, "tiny_step.pt"))

    # gradients of the reference (fp32, eval mode) for a handful of parameters
    model.zero_grad()
    labels = plan["labels"]
    real = (torch.randperm, np.random.rand, torch.multinomial)
    base_ones = [i for i in range(len(labels) // 2)]
    base_zeros = [i for i in range(len(labels) // 2, len(labels))]
    perm = torch.tensor([base_ones.pop() if l == 1 else base_zeros.pop() for l in labels.tolist()])
    neg_rows = [i for i in range(len(labels)) if labels[i] == 0]
    rand_seq = iter([0.9 if bool(plan["swap_video"][i]) else 0.1 for i in neg_rows])
    multi_seq = iter([int(plan["neg_idx"][i]) for i in neg_rows])
    torch.randperm = lambda n, **k: perm
    np.random.rand = lambda *a: next(rand_seq)
    torch.multinomial = lambda w, n, **k: torch.tensor([next(multi_seq)])
    try:
        d = dict(video=data["video"].clone(),
                 text=dict(input_ids=data["input_ids"].clone(), attention_mask=data["attention_mask"].clone()),
                 text_mlm_ids=data["text_mlm_ids"].clone(), text_mlm_labels=data["text_mlm_labels"].clone())
        loss, _, _ = model(d, data["noun_vec"], data["verb_vec"], _allgather, 1, _Args(),
                           {"loss": {"type": "EgoNCE"}}, ref.loss.EgoNCE(), 0, task_names="EgoNCE_MLM_ITM")
        loss.backward()
    finally:
        torch.randperm, np.random.rand, torch.multinomial = real
    grads = {}
    named = dict(model.named_parameters())
    for k in GRAD_KEYS:
        grads[k] = named[k].grad.clone()
    # oracle gradients through autograd on the same dict
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.pretrain_step(data, sdg, c["heads"], c["depth"], c["n_fuse"], plan)["loss_total"].backward()
    print("tiny step gradients (reference vs oracle):")
    for k in GRAD_KEYS:
        _cmp(k[-36:], sdg[k].grad, grads[k], tol=5e-4)
    torch.save(dict(cfg=c, weight_seed=0, data_seed=1234, plan_seed=4321,
                    grads={k: (g if g.numel() <= 70000 else g.flatten()[::37].clone()) for k, g in grads.items()}),
               os.path.join(OUT, "tiny_step_grads.pt"))


GRAD_KEYS = [
    "video_model.blocks.0.timeattn.qkv.weight", "video_model.blocks.0.norm3.weight",
    "video_model.blocks.7.attn.alpha_i2t", "video_model.blocks.7.attn.qkv_text_i2t.weight",
    "video_model.blocks.6.attn.norm_i2t_i.bias", "video_model.blocks.3.mlp.fc1.bias",
    "video_model.patch_embed.proj.weight", "video_model.pos_embed", "video_model.temporal_embed",
    "video_model.cls_token", "cls_token", "norm.weight",
    "text_model.encoder.layer.7.alpha_t2i", "text_model.encoder.layer.6.crossattention_t2i.self.key.weight",
    "text_model.encoder.layer.0.attention.self.query.weight", "text_model.encoder.layer.5.output.LayerNorm.weight",
    "text_model.embeddings.position_embeddings.weight", "txt_proj.0.weight", "vid_proj.4.bias",
    "mlm_score.transform.LayerNorm.weight", "mlm_score.bias", "itm_score.fc.weight",
    "cross_modal_video_pooler.dense.weight",
]


def golden_blocks_fullwidth(ref):
    """Block-level goldens at the real width (C=768, h=12, d=64) on short sequences:
    one fused SpaceTimeBlock and one fused RobertaLayer (reference classes built directly)."""
    torch.manual_seed(0)
    C, h, T, Nf, S, B = 768, 12, 3, 9, 16, 2
    shapes_all = O.key_shapes(C=C, heads=h, depth=7, n_fuse=1, T=T, img=48, patch=16, vocab=64, proj=64)
    sd = O.seeded_state(shapes_all, seed=7)
    vp, tp = "video_model.blocks.6.", "text_model.encoder.layer.6."
    blk = ref.vt.SpaceTimeBlock(dim=C, num_heads=h, qkv_bias=True, time_init="zeros", dim_text=C)
    blk.load_state_dict({k[len(vp):]: v for k, v in sd.items() if k.startswith(vp)}, strict=True)
    ref.rb.NUM_FUSE_BLOCK = 6
    cfg = ref.RobertaConfig(vocab_size=64, hidden_size=C, num_hidden_layers=7, num_attention_heads=h,
                            intermediate_size=4 * C, layer_norm_eps=1e-5)
    lay = ref.rb.RobertaLayer(cfg, layer_index=6)
    lay.load_state_dict({k[len(tp):]: v for k, v in sd.items() if k.startswith(tp)}, strict=True)
    blk.eval(), lay.eval()
    g = torch.Generator().manual_seed(99)
    x = torch.randn(B, 1 + T * Nf, C, generator=g)
    y = torch.randn(B, S, C, generator=g)
    am = torch.ones(B, S, dtype=torch.int64)
    am[0, 11:] = 0
    am[1, 5:] = 0
    m = O.extended_mask(am)
    with torch.no_grad():
        out_plain = blk(x, 'b (f n) d', '(b f) n d', 'b (f n) d', '(b n) f d', time_n=Nf, space_f=T)
        out_fused = blk(x, 'b (f n) d', '(b f) n d', 'b (f n) d', '(b n) f d', time_n=Nf, space_f=T, y=y, y_mask=m)
        t_plain = lay(y, m)[0]
        t_fused = lay(y, m, None, x, None, None, False, True)[0]
    print("full-width blocks (reference vs oracle):")
    _cmp("video block plain", O.space_time_block(x, sd, vp, h, T, Nf), out_plain)
    _cmp("video block fused", O.space_time_block(x, sd, vp, h, T, Nf, y=y, y_mask=m), out_fused)
    _cmp("text layer plain", O.roberta_layer(y, m, sd, tp, h), t_plain)
    _cmp("text layer fused", O.roberta_layer(y, m, sd, tp, h, video=x), t_fused)
    torch.save(dict(C=C, heads=h, T=T, Nf=Nf, S=S, B=B, weight_seed=7, x=x, y=y, attention_mask=am,
                    video_plain=out_plain, video_fused=out_fused, text_plain=t_plain, text_fused=t_fused),
               os.path.join(OUT, "blocks_fullwidth.pt"))


def golden_egonce(ref):
    g = torch.Generator().manual_seed(5)
    G = 12
    t = torch.randn(G, 64, generator=g)
    v = torch.randn(G, 64, generator=g) + 0.5 * t
    noun = (torch.rand(G, 30, generator=g) < 0.15).float()
    verb = (torch.rand(G, 12, generator=g) < 0.25).float()
    noun[:, 0] = 1
    verb[:, 0] = 1
    verb[3] = 0     # an all-zero row exercises the eps clamp of sim_matrix
    x = ref.mm.sim_matrix(t, v)
    sv, sn = ref.mm.sim_matrix(verb, verb), ref.mm.sim_matrix(noun, noun)
    loss, mask, temp = ref.loss.EgoNCE()(x, sv, sn)
    lo, mo = O.egonce(O.sim_matrix(t, v), O.sim_matrix(verb, verb), O.sim_matrix(noun, noun))
    print("EgoNCE (reference vs oracle):")
    _cmp("loss", lo, loss)
    assert torch.equal(mo, mask)
    torch.save(dict(t=t, v=v, noun=noun, verb=verb, sim=x, loss=loss, mask=mask, temperature=temp),
               os.path.join(OUT, "egonce.pt"))


def _load_ref_epic_charades(ref):
    """import the reference's model/model_epic_charades.py next to the already-shimmed model.* modules"""
    import importlib
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_ROOT)
    try:
        mec = importlib.import_module("model.model_epic_charades")
    finally:
        os.chdir(cwd)
    mec.config["use_checkpoint"] = False
    return mec


def dual_key_shapes(c):
    return O.dual_key_shapes(C=c["C"], heads=c["heads"], depth=c["depth"], n_fuse=c["n_fuse"], T=c["T"], img=c["img"],
                             patch=c["patch"], vocab=c["vocab"], proj=c["proj"])


def golden_dual(ref):
    """Fine-tuning dual-encoder step: model_epic_charades.FrozenInTime.forward with
    NormSoftmaxLoss (charades) and AdaptiveMaxMarginRankingLoss (epic), losses + similarities + gradients;
    plus the three loss classes on a fixed similarity matrix."""
    mec = _load_ref_epic_charades(ref)
    c = dict(TINY, T=4)
    text_cfg = dict(hidden_size=c["C"], num_hidden_layers=c["depth"], num_attention_heads=c["heads"],
                    intermediate_size=4 * c["C"])
    saved = mec.SpaceTimeTransformer
    import functools
    with ref_shim.tiny_wrapper(ref, img_size=c["img"], embed_dim=c["C"], depth=c["depth"], num_heads=c["heads"],
                               text_cfg=text_cfg, dim_cross=c["C"]):
        ref.rb.NUM_FUSE_BLOCK = 12 - (c["depth"] - c["n_fuse"])
        mec.SpaceTimeTransformer = functools.partial(ref.vt.SpaceTimeTransformer, img_size=c["img"], embed_dim=c["C"],
                                                     depth=c["depth"], num_heads=c["heads"])
        try:
            model = mec.FrozenInTime(
                video_params=dict(model="SpaceTimeTransformer", arch_config="base_patch16_224", num_frames=c["T"],
                                  pretrained=True, time_init="zeros", drop_path_rate=0.0),
                text_params=dict(model="roberta-base", pretrained=True, input="text"),
                projection="minimal", config=_yaml_cfg(c), task_names="EgoNCE_ITM_MLM", embed_dim=c["C"])
        finally:
            mec.SpaceTimeTransformer = saved
    shapes = dual_key_shapes(c)
    ref_sd = model.state_dict()
    theirs = {k for k in ref_sd if not k.endswith("position_ids")}
    assert set(shapes) == theirs, (sorted(set(shapes) - theirs)[:5], sorted(theirs - set(shapes))[:5])
    for k in shapes:
        assert tuple(ref_sd[k].shape) == tuple(shapes[k]), (k, ref_sd[k].shape, shapes[k])
    sd = O.seeded_state(shapes, 7)
    model.load_state_dict(sd, strict=False)
    model.eval()
    data = O.synthetic_batch(c["B"], c["T"], c["img"], c["S"], seed=77)
    g = torch.Generator().manual_seed(78)
    relation = torch.rand(c["B"], generator=g)
    out = dict(cfg=c, weight_seed=7, data_seed=77, relation=relation)
    for ds, loss_mod in (("charades", ref.loss.NormSoftmaxLoss()), ("epic", ref.loss.AdaptiveMaxMarginRankingLoss(margin=0.2))):
        model.zero_grad()
        d = dict(video=data["video"].clone(), relation=relation.clone(),
                 text=dict(input_ids=data["input_ids"].clone(), attention_mask=data["attention_mask"].clone()))
        loss, ld, ret = model(d, _allgather, 1, _Args(), {}, loss_mod, 0, task_names="Dual", dataset_name=ds)
        loss.backward()
        grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None and p.grad.abs().max() > 0}
        pick = ["txt_proj.1.weight", "txt_proj.1.bias", "vid_proj.0.weight", "vid_proj.0.bias", "video_model.cls_token",
                "video_model.blocks.0.attn.qkv.weight", "video_model.blocks.7.mlp.fc2.weight",
                "text_model.encoder.layer.0.attention.self.query.weight", "text_model.encoder.layer.7.output.dense.weight",
                "video_model.norm.weight"]
        out[ds] = dict(loss=loss.detach(), sim_v2t=ret["sim_v2t"].detach(), text_embeds=ret["text_embeds"].detach(),
                       video_embeds=ret["video_embeds"].detach(), grads={k: grads[k] for k in pick},
                       n_grads=len(grads))
        print(f"dual[{ds}]: loss {float(loss):.6f}, {len(grads)} tensors with gradient")
    # the loss classes on a fixed matrix (incl. fix_norm=False)
    x = torch.tanh(torch.randn(6, 6, generator=g))
    w = torch.rand(6, generator=g)
    out["loss_cases"] = dict(
        x=x, w=w,
        norm_softmax=ref.loss.NormSoftmaxLoss(0.07)(x)[0],
        max_margin=ref.loss.MaxMarginRankingLoss(0.2)(x),
        max_margin_nofix=ref.loss.MaxMarginRankingLoss(0.2, fix_norm=False)(x),
        adaptive=ref.loss.AdaptiveMaxMarginRankingLoss(0.4)(x, w),
        adaptive_nofix=ref.loss.AdaptiveMaxMarginRankingLoss(0.4, fix_norm=False)(x, w))
    torch.save(out, os.path.join(OUT, "dual_step.pt"))


def golden_optim(ref):
    """The reference's optimiser grouping and LR schedule (set_optim_schedule.py:16-129) applied to the reference's own
    tiny model: parameter name -> (weight_decay, lr) and the cosine-with-warm-up multipliers of the first steps."""
    import importlib
    import transformers.optimization as topt
    if not hasattr(topt, "AdamW"):          # removed from transformers 5.x; same constructor signature for what is read here
        topt.AdamW = lambda groups, lr, eps, betas: torch.optim.AdamW(groups, lr=lr, eps=eps, betas=betas)
    sys.path.insert(0, ref_shim.REF_ROOT)
    sos = importlib.import_module("set_optim_schedule")
    model, shapes, sd = build_tiny_reference(ref, TINY, seed=0)
    cfg = {"optimizer": {"type": "AdamW", "args": {"lr": 3e-5, "weight_decay": 0.01, "lr_mult_head": 2.0,
                                                   "lr_mult_cross_modal": 4.0}}}
    opt, sched = sos.set_schedule(model, cfg, {"end_lr": 1e-7, "decay_power": "cosine"}, max_steps=50, warmup_steps=5)
    by_id = {}
    for gi, g in enumerate(opt.param_groups):
        for p in g["params"]:
            assert id(p) not in by_id
            by_id[id(p)] = (gi, g["weight_decay"], g["initial_lr"] if "initial_lr" in g else g["lr"])
    groups = {n: by_id.get(id(p)) for n, p in model.named_parameters()}
    assert all(v is not None for v in groups.values())
    scales = []
    for _ in range(50):
        scales.append(sched.get_last_lr()[0] / 3e-5)
        opt.step()
        sched.step()
    torch.save(dict(groups=groups, lr_scales=scales, lr=3e-5, weight_decay=0.01, lr_mult_head=2.0, lr_mult_cross_modal=4.0,
                    max_steps=50, warmup_steps=5), os.path.join(OUT, "optim_groups.pt"))
    print("optimiser groups: %d parameters in %d groups" % (len(groups), len(opt.param_groups)))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load(use_checkpoint=False)
    if "--dual-only" in sys.argv:
        golden_dual(ref)
        return
    if "--optim-only" in sys.argv:
        golden_optim(ref)
        return
    golden_egonce(ref)
    golden_blocks_fullwidth(ref)
    golden_tiny_step(ref)
    golden_dual(ref)
    golden_optim(ref)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
