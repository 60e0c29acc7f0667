"""Host logic on CPU: the hand-written forward/backward kernel sequences of egovlpv2_b200.functional, executed
over the torch restatement of the kernel interface (tests/fake_kernels.py), against autograd through the oracle.

Two modes: 'exact' runs the very same sequencing with fp32 buffers (tolerance 2e-4: only op order differs), which
pins the backward derivations; 'bf16' keeps the product's bf16 activation/operand rounding (tolerance from
SURVEY.md section 8d: outputs rel-L2 <= 1e-2, gradients rel-L2 <= 3e-2)."""
import pytest
import torch

from egovlpv2_b200 import functional as Fn
from egovlpv2_b200.lib import ACT_GELU, ACT_NONE, ACT_RELU, ACT_TANH
from oracle import egovlp_oracle as O
from tests.fake_kernels import FakeKernels

C, HEADS, T, IMG, PATCH, S, B = 128, 2, 2, 32, 16, 8, 3
NF = (IMG // PATCH) ** 2
N = 1 + T * NF


@pytest.fixture(params=["exact", "bf16"])
def mode(request):
    old = Fn.BF16
    Fn.BF16 = torch.float32 if request.param == "exact" else torch.bfloat16
    yield request.param
    Fn.BF16 = old


def tol(mode, grad=False):
    if mode == "exact":
        return 3e-4
    return 3e-2 if grad else 1e-2


def rel(a, b):
    """relative L2 error; gradients that are analytically zero (e.g. the key bias under softmax) compare absolutely"""
    a, b = a.float(), b.float()
    if b.norm().item() < 1e-5:
        return (a - b).norm().item() * (1.0 if a.norm().item() > 2e-3 else 0.0)
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(scope="module")
def sd():
    shapes = O.key_shapes(C=C, heads=HEADS, depth=7, n_fuse=1, T=T, img=IMG, patch=PATCH, vocab=97, proj=256)
    return O.seeded_state(shapes, seed=3)


def block_params(sd, prefix, names):
    return {n: sd[prefix + n].clone() for n in names if prefix + n in sd}


def operand_copies(p):
    return {k: v.to(Fn.BF16) for k, v in p.items() if v.dim() >= 2}


def grads_of(sd_req, prefix, names):
    return {n: sd_req[prefix + n].grad for n in names if sd_req[prefix + n].grad is not None}


def mask_bias(lens):
    am = (torch.arange(S)[None] < torch.tensor(lens)[:, None]).long()
    return am, O.extended_mask(am).reshape(len(lens), S).contiguous()


@pytest.mark.parametrize("fused", [False, True, "reassoc", "legacy32"])
def test_video_block(sd, mode, fused):
    """fused = True: text length 8 (round-1 formulation of the gated cross-attention); 'reassoc': the BASELINE text length
    32, which takes the re-associated batched-GEMM path (xattn_reassoc.py); 'legacy32': the same with that path off."""
    K = FakeKernels()
    prefix = "video_model.blocks.6."
    S = 32 if fused in ("reassoc", "legacy32") else globals()["S"]
    Fn.set_reassoc(fused != "legacy32")
    names = Fn.VIDEO_BLOCK_PARAMS + (Fn.VIDEO_FUSE_PARAMS if fused else [])
    p = block_params(sd, prefix, names)
    w = operand_copies(p)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, N, C, generator=g)
    y = torch.randn(B, S, C, generator=g) if fused else None
    am = (torch.arange(S)[None] < torch.tensor([S, 5, 3])[:, None]).long()
    yb = O.extended_mask(am).reshape(B, S).contiguous()
    d_out = torch.randn(B, N, C, generator=g)
    out, saved = Fn.video_block_fwd(K, x, p, w, HEADS, T, NF, y=y, y_bias=yb if fused else None)
    assert (not fused) or saved.reassoc == (fused == "reassoc")
    Fn.set_reassoc(True)
    dx, dy, grads = Fn.video_block_bwd(K, saved, d_out, p, w, HEADS, T, NF)
    # oracle
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    yr = y.clone().requires_grad_(True) if fused else None
    ref = O.space_time_block(xr, sdr, prefix, HEADS, T, NF, y=yr, y_mask=O.extended_mask(am) if fused else None)
    ref.backward(d_out)
    assert rel(out, ref) <= tol(mode), rel(out, ref)
    assert rel(dx, xr.grad) <= tol(mode, True), rel(dx, xr.grad)
    if fused:
        assert rel(dy, yr.grad) <= tol(mode, True), rel(dy, yr.grad)
    for n in names:
        r = rel(grads[n].reshape(-1), sdr[prefix + n].grad.reshape(-1))
        assert r <= tol(mode, True), (n, r)


class DropoutReplay:
    """oracle.DROPOUT hook: replays the Philox keep masks of the product's dropout sites (functional.drop_site) at the
    reference's dropout sites"""

    def __init__(self, seed_dev, p, base=None):
        self.seed_dev, self.p, self.base, self.pass_id, self.calls = seed_dev, p, base, 0, 0

    def __call__(self, kind, slot, x):
        if kind == 0:
            self.pass_id += 1
        base = self.base if self.base is not None else self.pass_id << 16
        keep = FakeKernels().keep_mask(tuple(x.shape), self.p, self.seed_dev, base + slot * 16 + kind, x.device)
        self.calls += 1
        return x * keep / (1.0 - self.p)


@pytest.mark.parametrize("fused", [False, True, "dropout", "dropout_fused"])
def test_text_layer(sd, mode, fused):
    """'dropout*': train mode -- dropout (p = 0.1) on the attention probabilities and after the dense outputs, the
    cross-attention included; the oracle replays the very same Philox masks at the reference's dropout sites."""
    import types
    K = FakeKernels()
    drop = None
    if isinstance(fused, str):
        drop = types.SimpleNamespace(p=0.1, p_attn=0.1, seed=torch.tensor([424242]), base=3 << 16)
        fused = fused == "dropout_fused"
    prefix = "text_model.encoder.layer.6."
    names = Fn.TEXT_LAYER_PARAMS + (Fn.TEXT_FUSE_PARAMS if fused else [])
    p = block_params(sd, prefix, names)
    w = operand_copies(p)
    cat = lambda ns, sfx: torch.cat([p[n + sfx] for n in ns], 0)  # noqa: E731
    sa = ["attention.self.query", "attention.self.key", "attention.self.value"]
    p["qkv.bias"] = cat(sa, ".bias")
    w["qkv"] = cat(sa, ".weight").to(Fn.BF16)
    if fused:
        ca = ["crossattention_t2i.self.key", "crossattention_t2i.self.value"]
        p["cross.kv.bias"] = cat(ca, ".bias")
        w["cross.kv"] = cat(ca, ".weight").to(Fn.BF16)
    g = torch.Generator().manual_seed(1)
    h = torch.randn(B, S, C, generator=g)
    video = torch.randn(B, N, C, generator=g) if fused else None
    am, kb = mask_bias([S, 6, 2])
    d_out = torch.randn(B, S, C, generator=g)
    out, saved = Fn.text_layer_fwd(K, h, kb, p, w, HEADS, video=video, drop=drop, layer=6)
    dh, dvid, grads = Fn.text_layer_bwd(K, saved, d_out, p, w, HEADS)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    hr = h.clone().requires_grad_(True)
    vr = video.clone().requires_grad_(True) if fused else None
    if drop is not None:
        O.DROPOUT = DropoutReplay(drop.seed, drop.p, base=drop.base)
    try:
        ref = O.roberta_layer(hr, O.extended_mask(am), sdr, prefix, HEADS, video=vr)
    finally:
        hook, O.DROPOUT = O.DROPOUT, None
    assert drop is None or hook.calls == (5 if fused else 3)
    ref.backward(d_out)
    assert rel(out, ref) <= tol(mode), rel(out, ref)
    assert rel(dh, hr.grad) <= tol(mode, True), rel(dh, hr.grad)
    if fused:
        assert rel(dvid, vr.grad) <= tol(mode, True), rel(dvid, vr.grad)
    ref_g = {n: sdr[prefix + n].grad for n in names}
    got = dict(grads)
    got["attention.self.query.weight"], got["attention.self.key.weight"], got["attention.self.value.weight"] = got["qkv"].chunk(3, 0)
    got["attention.self.query.bias"], got["attention.self.key.bias"], got["attention.self.value.bias"] = got["qkv.bias"].chunk(3, 0)
    if fused:
        got["crossattention_t2i.self.key.weight"], got["crossattention_t2i.self.value.weight"] = got["cross.kv"].chunk(2, 0)
        got["crossattention_t2i.self.key.bias"], got["crossattention_t2i.self.value.bias"] = got["cross.kv.bias"].chunk(2, 0)
    for n in names:
        r = rel(got[n].reshape(-1), ref_g[n].reshape(-1))
        assert r <= tol(mode, True), (n, r)


def test_video_tokens_and_text_embeddings(sd, mode):
    K = FakeKernels()
    g = torch.Generator().manual_seed(2)
    video = torch.randn(B, T, 3, IMG, IMG, generator=g)
    vp = "video_model."
    p = {k: sd[vp + k] for k in ("patch_embed.proj.bias", "pos_embed", "temporal_embed")}
    w = {"patch_embed.proj.weight": sd[vp + "patch_embed.proj.weight"].reshape(C, -1).to(Fn.BF16)}
    tokens, saved = Fn.video_tokens_fwd(K, video, p, w, sd["cls_token"], PATCH)
    d_tok = torch.randn(B, N, C, generator=g)
    grads = Fn.video_tokens_bwd(K, saved, d_tok)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.video_tokens(video, sdr, sdr["cls_token"])
    ref.backward(d_tok)
    assert rel(tokens, ref) <= tol(mode)
    assert rel(grads["patch_embed.proj.weight"].reshape(-1), sdr[vp + "patch_embed.proj.weight"].grad.reshape(-1)) <= tol(mode, True)
    for k, name in (("patch_embed.proj.bias", vp + "patch_embed.proj.bias"), ("pos_embed", vp + "pos_embed"),
                    ("temporal_embed", vp + "temporal_embed"), ("cls_token", "cls_token")):
        assert rel(grads[k].reshape(-1), sdr[name].grad.reshape(-1)) <= tol(mode, True), k
    # text embeddings
    ids = torch.tensor([[0, 5, 9, 11, 2, 1, 1, 1], [0, 7, 7, 7, 8, 9, 10, 2], [0, 3, 2, 1, 1, 1, 1, 1]])
    ep = "text_model.embeddings."
    pe = {k: sd[ep + k] for k in ("word_embeddings.weight", "position_embeddings.weight", "token_type_embeddings.weight",
                                  "LayerNorm.weight", "LayerNorm.bias")}
    out, saved = Fn.text_embeddings_fwd(K, ids, pe)
    d = torch.randn(B, S, C, generator=g)
    ge = Fn.text_embeddings_bwd(K, saved, d, pe)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.roberta_embeddings(ids, sdr)
    ref.backward(d)
    assert rel(out, ref) <= 1e-5
    for k in pe:
        assert rel(ge[k].reshape(-1), sdr[ep + k].grad.reshape(-1)) <= 1e-4, k


def test_mlp_chain_and_mlm_head(sd, mode):
    K = FakeKernels()
    g = torch.Generator().manual_seed(4)
    # projection head: Linear(no bias)-ReLU-Linear-ReLU-Linear (model.py:105-115)
    x = torch.randn(B, C, generator=g)
    layers = [(sd["vid_proj.0.weight"].to(Fn.BF16), None, ACT_RELU), (sd["vid_proj.2.weight"].to(Fn.BF16), sd["vid_proj.2.bias"], ACT_RELU),
              (sd["vid_proj.4.weight"].to(Fn.BF16), sd["vid_proj.4.bias"], ACT_NONE)]
    out, saved = Fn.mlp_chain_fwd(K, x, layers)
    d = torch.randn(B, 256, generator=g)
    sc = torch.tensor([0.7])
    dx, grads = Fn.mlp_chain_bwd(K, saved, d, layers, scale_dev=sc)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    ref = O.projection(xr, sdr, "vid_proj")
    ref.backward(d * 0.7)
    assert rel(out, ref) <= tol(mode)
    # ReLU gates computed from bf16 operands flip for ~0.15 % of the units whose pre-activation is ~0; each flip is a
    # 100 % error on that unit's gradient, i.e. ~5 % rel-L2 per ReLU (measured; intrinsic to bf16 operands, the exact
    # mode above pins the derivation).  Hence the wider gradient tolerance for this 2-ReLU head in bf16 mode.
    gt = tol(mode, True) if mode == "exact" else 0.15
    assert rel(dx, xr.grad) <= gt
    for i, n in enumerate(("vid_proj.0", "vid_proj.2", "vid_proj.4")):
        assert rel(grads[i][0], sdr[n + ".weight"].grad) <= gt, n
        if grads[i][1] is not None:
            assert rel(grads[i][1], sdr[n + ".bias"].grad) <= gt, n
    # pooler-style chain ending in tanh + trailing GELU handled by act_grad
    layers = [(sd["cross_modal_text_transform.weight"].to(Fn.BF16), sd["cross_modal_text_transform.bias"], ACT_NONE),
              (sd["cross_modal_text_pooler.dense.weight"].to(Fn.BF16), sd["cross_modal_text_pooler.dense.bias"], ACT_TANH)]
    out, saved = Fn.mlp_chain_fwd(K, x, layers)
    d = torch.randn(B, C, generator=g)
    dx, grads = Fn.mlp_chain_bwd(K, saved, d, layers)
    xr = x.clone().requires_grad_(True)
    ref = torch.tanh(O._lin(O._lin(xr, sdr, "cross_modal_text_transform"), sdr, "cross_modal_text_pooler.dense"))
    ref.backward(d)
    assert rel(out, ref) <= tol(mode) and rel(dx, xr.grad) <= tol(mode, True)
    # MLM head + CE
    h = torch.randn(B, S, C, generator=g)
    labels = torch.randint(0, 97, (B, S), generator=g)
    labels[:, ::3] = -100
    names = ["cross_modal_text_transform.weight", "cross_modal_text_transform.bias", "mlm_score.transform.dense.weight",
             "mlm_score.transform.dense.bias", "mlm_score.transform.LayerNorm.weight", "mlm_score.transform.LayerNorm.bias",
             "mlm_score.decoder.weight", "mlm_score.bias"]
    p = {n: sd[n] for n in names}
    w = operand_copies(p)
    logits, loss_sum, count, saved = Fn.mlm_head_fwd(K, h, labels, p, w)
    inv = 1.0 / count
    dh, grads = Fn.mlm_head_bwd(K, saved, inv, p, w)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    hr = h.clone().requires_grad_(True)
    ref_logits = O.mlm_logits(hr, sdr)
    ref_loss = torch.nn.functional.cross_entropy(ref_logits.view(-1, 97), labels.view(-1), ignore_index=-100)
    ref_loss.backward()
    assert rel(logits, ref_logits.view(-1, 97)) <= tol(mode)
    assert abs((loss_sum / count).item() - ref_loss.item()) <= tol(mode) * max(1.0, abs(ref_loss.item()))
    assert rel(dh, hr.grad) <= tol(mode, True)
    for n in names:
        assert rel(grads[n].reshape(-1), sdr[n].grad.reshape(-1)) <= tol(mode, True), n


@pytest.mark.parametrize("fused", [False, True])
def test_video_block_cls_only(sd, mode, fused):
    """The CLS-only variant of the block (last block of the EgoNCE / ITM passes) against autograd through
    the oracle's FULL block with the loss reading only the CLS row."""
    K = FakeKernels()
    prefix = "video_model.blocks.6."
    names = Fn.VIDEO_BLOCK_PARAMS + (Fn.VIDEO_FUSE_PARAMS if fused else [])
    p = block_params(sd, prefix, names)
    w = operand_copies(p)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, N, C, generator=g)
    y = torch.randn(B, S, C, generator=g) if fused else None
    am, yb = mask_bias([S, 5, 3])
    d_cls = torch.randn(B, C, generator=g)
    out, saved = Fn.video_block_cls_fwd(K, x, p, w, HEADS, T, NF, y=y, y_bias=yb if fused else None)
    dx, dy, grads = Fn.video_block_cls_bwd(K, saved, d_cls, p, w, HEADS, T, NF)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    yr = y.clone().requires_grad_(True) if fused else None
    ref = O.space_time_block(xr, sdr, prefix, HEADS, T, NF, y=yr, y_mask=O.extended_mask(am) if fused else None)
    ref[:, 0].backward(d_cls)
    assert out.shape == (B, C)
    assert rel(out, ref[:, 0]) <= tol(mode), rel(out, ref[:, 0])
    assert rel(dx, xr.grad) <= tol(mode, True), rel(dx, xr.grad)
    if fused:
        assert rel(dy, yr.grad) <= tol(mode, True), rel(dy, yr.grad)
    for n in names:
        r = rel(grads[n].reshape(-1), sdr[prefix + n].grad.reshape(-1))
        assert r <= tol(mode, True), (n, r)
    # and the full kernel sequence gives the same CLS row
    full, _ = Fn.video_block_fwd(K, x, p, w, HEADS, T, NF, y=y, y_bias=yb if fused else None, save=False)
    assert rel(out, full[:, 0]) <= tol(mode)


def test_video_tokens_from_uint8_frames(sd):
    """uint8 frames through the fused normalise + im2col == the reference's host pipeline
    (frames.float() / 255, base_dataset.py:248; NormalizeVideo, transforms.py:49) followed by the fp32 path."""
    K = FakeKernels()
    g = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (B, T, 3, IMG, IMG), generator=g, dtype=torch.uint8)
    mean, std = Fn.VIDEO_NORM_MEAN, Fn.VIDEO_NORM_STD
    host = (frames.float() / 255 - torch.tensor(mean).view(1, 1, 3, 1, 1)) / torch.tensor(std).view(1, 1, 3, 1, 1)
    vp = "video_model."
    p = {k: sd[vp + k] for k in ("patch_embed.proj.bias", "pos_embed", "temporal_embed")}
    w = {"patch_embed.proj.weight": sd[vp + "patch_embed.proj.weight"].reshape(C, -1).to(Fn.BF16)}
    t_u8, _ = Fn.video_tokens_fwd(K, frames, p, w, sd["cls_token"], PATCH)
    t_f32, _ = Fn.video_tokens_fwd(K, host, p, w, sd["cls_token"], PATCH)
    assert torch.equal(t_u8, t_f32)
    assert rel(t_u8, O.video_tokens(host, sd, sd["cls_token"])) <= 1e-2
    # caller-provided statistics
    t2, _ = Fn.video_tokens_fwd(K, frames, p, w, sd["cls_token"], PATCH, norm=((0.5, 0.5, 0.5), (0.25, 0.5, 1.0)))
    host2 = (frames.float() / 255 - 0.5) / torch.tensor((0.25, 0.5, 1.0)).view(1, 1, 3, 1, 1)
    t3, _ = Fn.video_tokens_fwd(K, host2, p, w, sd["cls_token"], PATCH)
    assert torch.equal(t2, t3)


def test_fused_blocks_at_large_width(monkeypatch):
    """BASELINE cfg 5 widths (C = 1024, 16 heads of 64; SURVEY.md Q12: parity pinned at block level): the hand-derived
    forward / backward of the fused video block and the fused text layer are width-agnostic."""
    import tests.test_functional_cpu as me
    monkeypatch.setattr(me, "C", 1024)
    monkeypatch.setattr(me, "HEADS", 16)
    shapes = O.key_shapes(C=1024, heads=16, depth=7, n_fuse=1, T=T, img=IMG, patch=PATCH, vocab=97, proj=256)
    keep = {k: v for k, v in shapes.items() if "blocks.6." in k or "layer.6." in k}
    sd_l = O.seeded_state(keep, seed=5)
    old = Fn.BF16
    Fn.BF16 = torch.float32
    try:
        test_video_block(sd_l, "exact", True)
        test_text_layer(sd_l, "exact", True)
    finally:
        Fn.BF16 = old


def test_reassociated_i2t(sd, mode):
    """egovlpv2_b200/xattn_reassoc.py: the hand-derived forward / backward of the re-associated gated video->text
    cross-attention (batched GEMMs + group-softmax epilogues) == autograd through the reference formulation (q projection,
    32-key masked softmax attention, output projection, gate: video_transformer.py:165-185), for every input and
    parameter gradient.  The re-associated kernels need the BASELINE text length S = 32."""
    from egovlpv2_b200 import xattn_reassoc as XR
    K = FakeKernels()
    prefix = "video_model.blocks.6.attn."
    d = C // HEADS
    S32 = 32
    g = torch.Generator().manual_seed(17)
    ln = torch.randn(B, N, C, generator=g)
    kv = torch.randn(B, S32, 2 * C, generator=g)
    xa = torch.randn(B, N, C, generator=g)
    am = (torch.arange(S32)[None] < torch.tensor([S32, 5, 19])[:, None]).long()
    mask = O.extended_mask(am).reshape(B, S32).contiguous()
    dout = torch.randn(B, N, C, generator=g)
    wq, bq = sd[prefix + "qkv_i2t.weight"], sd[prefix + "qkv_i2t.bias"]
    wp, bp = sd[prefix + "proj_i2t.weight"], sd[prefix + "proj_i2t.bias"]
    alpha = torch.tensor([0.7])
    cast = lambda t: t.to(Fn.BF16)  # noqa: E731
    out = torch.empty(B * N, C)
    saved = XR.i2t_fwd(K, cast(ln).reshape(B * N, C), cast(kv).reshape(B * S32, 2 * C), mask, cast(wq), bq, cast(wp), bp, alpha,
                       xa.reshape(B * N, C), out, B, N, HEADS)
    dalpha, dwq, dbq, dwp = torch.zeros(1), torch.zeros(C, C), torch.zeros(C), torch.zeros(C, C)
    dln, dkv = XR.i2t_bwd(K, saved, cast(dout).reshape(B * N, C), dout.reshape(B * N, C).sum(0), cast(wq), bq, cast(wp), bp, alpha,
                          dalpha, dwq, dbq, dwp)
    # reference formulation under autograd
    r = {k: v.clone().requires_grad_(True) for k, v in dict(ln=ln, kv=kv, wq=wq, bq=bq, wp=wp, bp=bp, alpha=alpha).items()}
    q = O._heads(torch.nn.functional.linear(r["ln"], r["wq"], r["bq"]), HEADS) * d ** -0.5
    k_t, v_t = (O._heads(t, HEADS) for t in r["kv"].split(C, dim=-1))
    sc = q @ k_t.transpose(-1, -2) + mask.reshape(B, 1, 1, S32)
    ref = xa + r["alpha"] * torch.nn.functional.linear(O._merge(torch.softmax(sc, dim=-1) @ v_t), r["wp"], r["bp"])
    ref.backward(dout)
    assert rel(out.reshape(B, N, C), ref) <= tol(mode), rel(out.reshape(B, N, C), ref)
    for name, got in (("ln", dln), ("kv", dkv), ("wq", dwq), ("bq", dbq), ("wp", dwp), ("alpha", dalpha)):
        e = rel(got.reshape(-1), r[name].grad.reshape(-1))
        assert e <= tol(mode, True), (name, e)


@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_reassociated_t2i(sd, mode, p_drop):
    """xattn_reassoc.t2i_*: text queries over all N video tokens without projecting the tokens to keys / values
    == autograd through the reference formulation (roberta.py:257-327 with encoder_hidden_states = video, no mask),
    with and without the dropout on the attention probabilities (roberta.py:313; the reference formulation replays the
    Philox mask of the kernels)."""
    from egovlpv2_b200 import xattn_reassoc as XR
    K = FakeKernels()
    prefix = "text_model.encoder.layer.6.crossattention_t2i.self."
    d = C // HEADS
    g = torch.Generator().manual_seed(23)
    q = torch.randn(B, S, C, generator=g)
    x = torch.randn(B, N, C, generator=g)
    dctx = torch.randn(B, S, C, generator=g)
    wk, bk = sd[prefix + "key.weight"], sd[prefix + "key.bias"]
    wv, bv = sd[prefix + "value.weight"], sd[prefix + "value.bias"]
    cast = lambda t: t.to(Fn.BF16)  # noqa: E731
    ox = torch.empty(B * S, C, dtype=Fn.BF16)
    saved = XR.t2i_fwd(K, cast(q).reshape(B * S, C), cast(x).reshape(B * N, C), cast(wk), cast(wv), bv, ox, B, N, HEADS,
                       p_drop=p_drop, seed_dev=None, site=77)
    dwk, dwv, dbv, dx = torch.zeros(C, C), torch.zeros(C, C), torch.zeros(C), torch.empty(B * N, C)
    dq = XR.t2i_bwd(K, saved, cast(dctx).reshape(B * S, C), cast(wk), cast(wv), bv, dwk, dwv, dbv, dx)
    r = {k: v.clone().requires_grad_(True) for k, v in dict(q=q, x=x, wk=wk, bk=bk, wv=wv, bv=bv).items()}
    lin = torch.nn.functional.linear
    qh = O._heads(r["q"], HEADS)
    kh, vh = O._heads(lin(r["x"], r["wk"], r["bk"]), HEADS), O._heads(lin(r["x"], r["wv"], r["bv"]), HEADS)
    pr = torch.softmax(qh @ kh.transpose(-1, -2) / d ** 0.5, dim=-1)          # [B, H, S, N]
    if p_drop > 0:
        rows = torch.arange(B * HEADS * S).reshape(B, HEADS, S, 1)                 # kernel row order: (clip, head, query)
        keep = FakeKernels.philox_keep(FakeKernels.philox_key(None, 77), rows * N + torch.arange(N), p_drop)
        pr = pr * keep / (1 - p_drop)
    ref = O._merge(pr @ vh)
    ref.backward(dctx)
    assert rel(ox.reshape(B, S, C), ref) <= tol(mode), rel(ox.reshape(B, S, C), ref)
    for name, got in (("q", dq), ("x", dx), ("wk", dwk), ("wv", dwv), ("bv", dbv)):
        e = rel(got.reshape(-1), r[name].grad.reshape(-1))
        assert e <= tol(mode, True), (name, e)
    assert r["bk"].grad.abs().max().item() <= 1e-5 or p_drop > 0    # the key bias is softmax-invariant without dropout


def test_video_tokens_patch14_padded_im2col(mode):
    """TimeSformer-L/14 (BASELINE cfg 5): patch 14 gives an im2col depth of 3 * 14^2 = 588, not a multiple of the 8 elements
    a TMA row stride needs; the patch GEMM runs over the depth padded to 592 against a zero-padded weight
    (egv_patchify_padded).  Tokens and every gradient against the oracle's Conv2d."""
    K = FakeKernels()
    T14, IMG14, P14 = 2, 28, 14
    shapes = O.key_shapes(C=C, heads=HEADS, depth=7, n_fuse=1, T=T14, img=IMG14, patch=P14, vocab=97, proj=256)
    sd14 = O.seeded_state(shapes, seed=5)
    g = torch.Generator().manual_seed(4)
    video = torch.randn(B, T14, 3, IMG14, IMG14, generator=g)
    vp = "video_model."
    p = {k: sd14[vp + k] for k in ("patch_embed.proj.bias", "pos_embed", "temporal_embed")}
    w = {"patch_embed.proj.weight": sd14[vp + "patch_embed.proj.weight"].reshape(C, -1).to(Fn.BF16)}
    assert w["patch_embed.proj.weight"].shape[1] == 588
    tokens, saved = Fn.video_tokens_fwd(K, video, p, w, sd14["cls_token"], P14)
    assert saved.cols.shape[1] == 592 and float(saved.cols[:, 588:].float().abs().max()) == 0.0
    n14 = 1 + T14 * (IMG14 // P14) ** 2
    d_tok = torch.randn(B, n14, C, generator=g)
    grads = Fn.video_tokens_bwd(K, saved, d_tok)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd14.items()}
    ref = O.video_tokens(video, sdr, sdr["cls_token"])
    ref.backward(d_tok)
    assert rel(tokens, ref) <= tol(mode)
    assert rel(grads["patch_embed.proj.weight"].reshape(-1), sdr[vp + "patch_embed.proj.weight"].grad.reshape(-1)) <= tol(mode, True)
    for k, name in (("patch_embed.proj.bias", vp + "patch_embed.proj.bias"), ("pos_embed", vp + "pos_embed"),
                    ("temporal_embed", vp + "temporal_embed"), ("cls_token", "cls_token")):
        assert rel(grads[k].reshape(-1), sdr[name].grad.reshape(-1)) <= tol(mode, True), k
