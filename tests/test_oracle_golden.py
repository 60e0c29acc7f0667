"""The CPU oracle replayed against the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  fp32, eval mode; tolerance 2e-4 relative to
max(1, max|ref|) -- observed agreement is ~1e-6 (different op order only)."""
import os

import pytest
import torch

from oracle import egovlp_oracle as O

TOL = 2e-4


def _close(a, b, tol=TOL):
    assert a.shape == b.shape
    err = (a.float() - b.float()).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), err


def test_egonce_golden(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "egonce.pt"))
    sim = O.sim_matrix(fx["t"], fx["v"])
    _close(sim, fx["sim"])
    loss, mask = O.egonce(sim, O.sim_matrix(fx["verb"], fx["verb"]), O.sim_matrix(fx["noun"], fx["noun"]))
    _close(loss, fx["loss"])
    assert torch.equal(mask, fx["mask"])
    assert fx["temperature"] == 0.05


def test_fullwidth_blocks_golden(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "blocks_fullwidth.pt"))
    C, h, T, Nf = fx["C"], fx["heads"], fx["T"], fx["Nf"]
    shapes = O.key_shapes(C=C, heads=h, depth=7, n_fuse=1, T=T, img=48, patch=16, vocab=64, proj=64)
    sd = O.seeded_state(shapes, seed=fx["weight_seed"])
    vp, tp = "video_model.blocks.6.", "text_model.encoder.layer.6."
    m = O.extended_mask(fx["attention_mask"])
    _close(O.space_time_block(fx["x"], sd, vp, h, T, Nf), fx["video_plain"])
    _close(O.space_time_block(fx["x"], sd, vp, h, T, Nf, y=fx["y"], y_mask=m), fx["video_fused"])
    _close(O.roberta_layer(fx["y"], m, sd, tp, h), fx["text_plain"])
    _close(O.roberta_layer(fx["y"], m, sd, tp, h, video=fx["x"]), fx["text_fused"])


@pytest.fixture(scope="module")
def tiny(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "tiny_step.pt"))
    c = fx["cfg"]
    shapes = O.key_shapes(C=c["C"], heads=c["heads"], depth=c["depth"], n_fuse=c["n_fuse"], T=c["T"],
                          img=c["img"], patch=c["patch"], vocab=c["vocab"], proj=c["proj"])
    sd = O.seeded_state(shapes, fx["weight_seed"])
    data = O.synthetic_batch(c["B"], c["T"], c["img"], c["S"], seed=fx["data_seed"])
    plan = O.synthetic_itm_plan(c["B"], seed=fx["plan_seed"])
    return fx, c, sd, data, plan


def test_tiny_step_golden(tiny):
    fx, c, sd, data, plan = tiny
    out = O.pretrain_step(data, sd, c["heads"], c["depth"], c["n_fuse"], plan)
    for k in ("loss_total", "EgoNCE", "loss_mlm", "loss_itm", "sim_v2t", "text_embeds", "video_embeds"):
        _close(out[k], fx[k])
    _close(out["cross_attn_itm_logits"], fx["itm_logits"])
    mlm = out["cross_attn_mlm_logits"]
    _close(mlm[:, :, ::997], fx["mlm_logits_slice"])
    _close(torch.logsumexp(mlm, -1), fx["mlm_logits_lse"])
    _close(mlm.double().sum(-1).float(), fx["mlm_logits_sum"], tol=1e-3)


def test_tiny_step_gradients_golden(tiny, golden_dir):
    fx, c, sd, data, plan = tiny
    gfx = torch.load(os.path.join(golden_dir, "tiny_step_grads.pt"))
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.pretrain_step(data, sdg, c["heads"], c["depth"], c["n_fuse"], plan)["loss_total"].backward()
    for k, g in gfx["grads"].items():
        mine = sdg[k].grad
        if mine.numel() > 70000:
            mine = mine.flatten()[::37]
        err = (mine - g).abs().max().item()
        assert err <= 5e-4 * max(g.abs().max().item(), 1e-6) + 1e-9, (k, err)


def test_synthetic_batch_properties():
    d = O.synthetic_batch(6, 2, 32, 16, seed=3)
    ids, am = d["input_ids"], d["attention_mask"]
    assert (ids[:, 0] == 0).all()
    lens = am.sum(1)
    for b in range(6):
        L = int(lens[b])
        assert ids[b, L - 1] == 2 and (ids[b, L:] == 1).all() and (ids[b, 1:L - 1] >= 3).all()
    lab = d["text_mlm_labels"]
    assert ((lab == -100) | (lab == ids)).all() and (lab != -100).any(1).all()
    assert (d["noun_vec"].sum(1) >= 1).all() and (d["verb_vec"].sum(1) >= 1).all()
    pos = O.roberta_embeddings  # position ids: cumsum over non-pad + pad_id (roberta.py:881-892)
    assert callable(pos)


def test_dual_losses_golden(golden_dir):
    """loss.py:13-31, 65-143 on a fixed matrix (values from the reference's own classes)"""
    fx = torch.load(os.path.join(golden_dir, "dual_step.pt"))["loss_cases"]
    x, w = fx["x"], fx["w"]
    _close(O.norm_softmax_loss(x, 0.07), fx["norm_softmax"])
    _close(O.max_margin_ranking_loss(x, 0.2), fx["max_margin"])
    _close(O.max_margin_ranking_loss(x, 0.2, fix_norm=False), fx["max_margin_nofix"])
    _close(O.max_margin_ranking_loss(x, 0.4, w), fx["adaptive"])
    _close(O.max_margin_ranking_loss(x, 0.4, w, fix_norm=False), fx["adaptive_nofix"])


@pytest.mark.parametrize("dataset", ["charades", "epic"])
def test_dual_step_golden(golden_dir, dataset):
    """model_epic_charades.FrozenInTime.forward(task_names='Dual') + backward"""
    fx = torch.load(os.path.join(golden_dir, "dual_step.pt"))
    c, g = fx["cfg"], fx[dataset]
    shapes = O.dual_key_shapes(C=c["C"], heads=c["heads"], depth=c["depth"], n_fuse=c["n_fuse"], T=c["T"], img=c["img"],
                               patch=c["patch"], vocab=c["vocab"], proj=c["proj"])
    sd = {k: v.requires_grad_(True) for k, v in O.seeded_state(shapes, fx["weight_seed"]).items()}
    data = dict(O.synthetic_batch(c["B"], c["T"], c["img"], c["S"], seed=fx["data_seed"]), relation=fx["relation"])
    out = O.dual_step(data, sd, c["heads"], c["depth"], dataset_name=dataset)
    for k in ("sim_v2t", "text_embeds", "video_embeds"):
        _close(out[k], g[k])
    _close(out["Dual"], g["loss"])
    out["Dual"].backward()
    for k, ref in g["grads"].items():
        got = sd[k].grad
        assert ((got - ref).norm() / ref.norm()).item() <= 1e-3, k


@pytest.mark.skipif(not os.path.isdir("/root/reference/EgoVLPv2"), reason="the reference tree only exists in the build container")
def test_live_differential_vs_reference():
    """oracle/live_diff.py: the drop-in module API (real host logic, exact-fp32 kernel restatement) against the UNMODIFIED
    reference imported live, on fresh seeded cases beyond the committed goldens (B = 1, odd batch, ragged captions, other
    frame counts / image sizes): losses, similarities, logits and 14 parameter gradients within 2e-3 (observed ~1e-5).
    Runs in a subprocess: the reference shim patches process-wide state."""
    import subprocess
    import sys
    import socket
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))          # a free rendezvous port for the shim's 1-rank gloo group
    port = sock.getsockname()[1]
    sock.close()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r = subprocess.run([sys.executable, "-m", "oracle.live_diff"], cwd=root, capture_output=True, text=True, timeout=900,
                       env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "live differential vs reference ok" in r.stdout


def test_reassociated_cross_attention_is_exact():
    """both gated cross-attentions evaluated around the S text tokens (no [N, C] x [C, C] query /
    key / value / output projections) equal the reference formulation -- values and gradients -- in fp32."""
    C, h, T, Nf, S, B = 128, 2, 2, 4, 8, 3
    shapes = O.key_shapes(C=C, heads=h, depth=7, n_fuse=1, T=T, img=32, patch=16, vocab=64, proj=64)
    base = O.seeded_state(shapes, seed=4)
    g = torch.Generator().manual_seed(8)
    a0 = torch.randn(B, 1 + T * Nf, C, generator=g)
    y0 = torch.randn(B, S, C, generator=g)
    am = torch.ones(B, S, dtype=torch.int64)
    am[1, 5:] = 0
    am[2, 3:] = 0
    m = O.extended_mask(am)
    vp, tp = "video_model.blocks.6.attn.", "text_model.encoder.layer.6.crossattention_t2i."
    outs = []
    for f_i2t, f_t2i in ((O.cross_attention_i2t, lambda hq, v, sd: O._bert_attention(hq, v, None, sd, tp, h)),
                         (O.cross_attention_i2t_reassociated, lambda hq, v, sd: O.cross_attention_t2i_reassociated(hq, v, sd, tp, h))):
        sd = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        a, y = a0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
        o1 = f_i2t(a, y, m, sd, vp, h)
        o2 = f_t2i(y, a, sd)
        (o1.square().sum() + o2.square().sum()).backward()
        keys = [k for k in sd if (k.startswith(vp) or k.startswith(tp)) and sd[k].grad is not None]
        outs.append((o1.detach(), o2.detach(), a.grad, y.grad, {k: sd[k].grad for k in keys}))
    (r1, r2, ra, ry, rg), (n1, n2, na, ny, ng) = outs
    _close(n1, r1, 1e-5), _close(n2, r2, 1e-5), _close(na, ra, 1e-4), _close(ny, ry, 1e-4)
    # the key bias of the text->video attention has an analytically zero gradient: the re-associated form never reads it
    assert set(rg) - set(ng) <= {tp + "self.key.bias"} and len(ng) >= 12
    for k in ng:
        _close(ng[k], rg[k], 1e-4)
