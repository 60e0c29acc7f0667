"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the EgoVLPv2 pre-training hot path.

A plain-PyTorch fp32, *functional* restatement of the reference's algorithm
(facebookresearch/EgoVLPv2 @ 550c0596).  Every function takes the flat
``state_dict`` of the reference model (Appendix B of SURVEY.md: the same 558
key names) and explicit tensors; there are no nn.Modules, no einops, no
checkpointing and no autocast.  Each function cites the reference lines it
restates (paths relative to EgoVLPv2/).

Pinning status: the reference ships no tests, golden vectors or known-answer
files for this path, so the oracle is pinned the only way
available -- against the reference's own code executed in the build container
(`oracle/make_golden.py` imports /root/reference through `oracle/ref_shim.py`
and writes tests/golden/*.pt; `tests/test_oracle_golden.py` replays them).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this file.  The product package never does.
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- utils


def _lin(x, sd, name, bias=True):
    b = sd.get(name + ".bias") if bias else None
    return F.linear(x, sd[name + ".weight"], b)


def _ln(x, sd, name, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _heads(t, h):
    """[B, L, h*d] -> [B, h, L, d]"""
    B, L, C = t.shape
    return t.view(B, L, h, C // h).permute(0, 2, 1, 3)


def _merge(t):
    """[B, h, L, d] -> [B, L, h*d]"""
    B, h, L, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(B, L, h * d)


# --------------------------------------------------------------------------- video side


def video_tokens(video, sd, cls_token, vprefix="video_model."):
    """Patch embedding + token assembly.

    video_transformer.py:78-83 (Conv2d k=s=p on [B*T,3,H,W]) and :354-372
    (flatten, CLS, tiled pos_embed + repeat-interleaved temporal_embed);
    duplicated in model.py:211-232 / 295-318 with FrozenInTime.cls_token.
    Returns [B, 1+T*Nf, C]; token 1+f*Nf+n is patch n (row-major) of frame f.
    """
    B, T, Cin, H, W = video.shape
    w = sd[vprefix + "patch_embed.proj.weight"]
    b = sd[vprefix + "patch_embed.proj.bias"]
    C, p = w.shape[0], w.shape[-1]
    gh, gw = H // p, W // p
    # non-overlapping conv == per-patch matmul over (c, i, j)
    x = video.reshape(B * T, Cin, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(B * T, gh * gw, Cin * p * p)
    x = x @ w.reshape(C, -1).t() + b
    Nf = gh * gw
    x = x.reshape(B, T * Nf, C)
    pos = sd[vprefix + "pos_embed"][0]          # [1+Nf, C]
    tem = sd[vprefix + "temporal_embed"][0]     # [T_max, C]
    patch_pos = (pos[1:][None, :, :] + tem[:T][:, None, :]).reshape(T * Nf, C)
    x = x + patch_pos[None]
    cls = (cls_token.reshape(1, 1, C) + pos[0].reshape(1, 1, C)).expand(B, 1, C)
    return torch.cat([cls, x], dim=1)


def divided_attention(x, sd, prefix, heads, T, Nf, mode):
    """VarAttention self-attention core incl. qkv/proj (video_transformer.py:117-153).

    mode 'time'  : patch (f, n) attends CLS + patch n of every frame   (T+1 keys)
    mode 'space' : patch (f, n) attends CLS + every patch of frame f   (Nf+1 keys)
    CLS query attends all 1+T*Nf tokens.  q is scaled by d^-1/2 *before* QK^T (:123).
    """
    B, N, C = x.shape
    d = C // heads
    qkv = _lin(x, sd, prefix + "qkv")
    q, k, v = (_heads(t, heads) for t in qkv.split(C, dim=-1))     # [B,h,N,d]
    q = q * d ** -0.5
    # CLS query over everything
    cls_out = torch.softmax(q[:, :, :1] @ k.transpose(-1, -2), dim=-1) @ v     # [B,h,1,d]
    qp = q[:, :, 1:].reshape(B, heads, T, Nf, d)
    kp = k[:, :, 1:].reshape(B, heads, T, Nf, d)
    vp = v[:, :, 1:].reshape(B, heads, T, Nf, d)
    kc = k[:, :, :1]
    vc = v[:, :, :1]
    if mode == "time":
        qp, kp, vp = (t.transpose(2, 3) for t in (qp, kp, vp))     # [B,h,Nf,T,d]
    G = qp.shape[2]
    kk = torch.cat([kc[:, :, None].expand(B, heads, G, 1, d), kp], dim=3)
    vv = torch.cat([vc[:, :, None].expand(B, heads, G, 1, d), vp], dim=3)
    o = torch.softmax(qp @ kk.transpose(-1, -2), dim=-1) @ vv          # [B,h,G,L,d]
    if mode == "time":
        o = o.transpose(2, 3)
    o = o.reshape(B, heads, T * Nf, d)
    o = _merge(torch.cat([cls_out, o], dim=2))
    return _lin(o, sd, prefix + "proj")


def cross_attention_i2t(a, y, y_mask, sd, prefix, heads):
    """Gated video->text cross-attention (video_transformer.py:155-185).

    a      : [B,N,C]  output of the space self-attention branch (after proj)
    y      : [B,S,Ct] text hidden states entering text layer i
    y_mask : additive mask broadcastable to [B,1,1,S] (0 / finfo.min)
    returns a + alpha_i2t * proj_i2t(softmax(q k^T + mask) v)
    """
    B, N, C = a.shape
    d = C // heads
    kv = _lin(y, sd, prefix + "qkv_text_i2t")
    k_t, v_t = (_heads(t, heads) for t in kv.split(C, dim=-1))       # [B,h,S,d]
    q = _heads(_lin(_ln(a, sd, prefix + "norm_i2t_i", 1e-5), sd, prefix + "qkv_i2t"), heads) * d ** -0.5
    s = q @ k_t.transpose(-1, -2)
    if y_mask is not None:
        s = s + y_mask.reshape(B, 1, 1, -1)
    c = _merge(torch.softmax(s, dim=-1) @ v_t)
    c = _lin(c, sd, prefix + "proj_i2t")
    return a + sd[prefix + "alpha_i2t"] * c


def cross_attention_i2t_reassociated(a, y, y_mask, sd, prefix, heads):
    """The same function as cross_attention_i2t with the products re-associated around the S (<= 64) text keys
    (DESIGN.md section 7): per clip and head h,
        scores_h = LN(a) . M_h + c_h      M_h = d^-1/2 Wq_h^T k_h^T  [C, S],  c_h = d^-1/2 bq_h k_h^T  [S]
        out      = sum_h P_h . U_h + bp   U_h = v_h Wp[:, h]^T       [S, C]
    i.e. two [N, C] x [C, h*S] / [N, h*S] x [h*S, C] products (h*S = 384 at the BASELINE shapes) instead of the
    [N, C] x [C, C] query and output projections: half the FLOPs, no separate attention launch."""
    B, N, C = a.shape
    d = C // heads
    kv = _lin(y, sd, prefix + "qkv_text_i2t")
    k_t, v_t = (_heads(t, heads) for t in kv.split(C, dim=-1))                       # [B,h,S,d]
    wq = sd[prefix + "qkv_i2t.weight"].view(heads, d, C)                              # rows of head h
    bq = sd[prefix + "qkv_i2t.bias"].view(heads, d)
    wp = sd[prefix + "proj_i2t.weight"].view(C, heads, d)                             # columns of head h
    M = torch.einsum("hdc,bhsd->bhcs", wq, k_t) * d ** -0.5                           # [B,h,C,S]
    c0 = torch.einsum("hd,bhsd->bhs", bq, k_t) * d ** -0.5                            # [B,h,S]
    U = torch.einsum("bhsd,chd->bhsc", v_t, wp)                                       # [B,h,S,C]
    ln = _ln(a, sd, prefix + "norm_i2t_i", 1e-5)
    s = torch.einsum("bnc,bhcs->bhns", ln, M) + c0[:, :, None, :]
    if y_mask is not None:
        s = s + y_mask.reshape(B, 1, 1, -1)
    out = torch.einsum("bhns,bhsc->bnc", torch.softmax(s, dim=-1), U) + sd[prefix + "proj_i2t.bias"]
    return a + sd[prefix + "alpha_i2t"] * out


def cross_attention_t2i_reassociated(hq, video, sd, prefix, heads):
    """_bert_attention(hq, video, None, ...) (the text->video cross-attention, roberta.py:470-486) re-associated
    around the S text queries: per clip and head h,
        scores_h = Q'_h . x^T          Q'_h = d^-1/2 q_h Wk_h  [S, C]   (q_h . bk_h is constant over keys: softmax drops it)
        ctx_h    = (P_h . x) Wv_h^T + bv_h                              (rows of P_h sum to 1)
    i.e. [h*S, C] x [C, N] and [h*S, N] x [N, C] products instead of the [N, C] x [C, 2C] key / value projection of
    every video token: half the FLOPs and no K / V tensors in HBM."""
    B, S, C = hq.shape
    d = C // heads
    q = _heads(_lin(hq, sd, prefix + "self.query"), heads)                            # [B,h,S,d]
    wk = sd[prefix + "self.key.weight"].view(heads, d, -1)                            # [h,d,Cv]
    wv = sd[prefix + "self.value.weight"].view(heads, d, -1)
    bv = sd[prefix + "self.value.bias"].view(heads, d)
    qp = torch.einsum("bhsd,hdc->bhsc", q, wk) / math.sqrt(d)                         # [B,h,S,Cv]
    p = torch.softmax(torch.einsum("bhsc,bnc->bhsn", qp, video), dim=-1)
    z = torch.einsum("bhsn,bnc->bhsc", p, video)                                      # [B,h,S,Cv]
    ctx = torch.einsum("bhsc,hdc->bhsd", z, wv) + bv[None, :, None, :]
    return _lin(_merge(ctx), sd, prefix + "output.dense")


def space_time_block(x, sd, prefix, heads, T, Nf, y=None, y_mask=None, eps=1e-5):
    """SpaceTimeBlock.forward (video_transformer.py:214-228).  Note the space
    residual is added to the block *input* x, not to x+time (:218-222)."""
    t = divided_attention(_ln(x, sd, prefix + "norm3", eps), sd, prefix + "timeattn.", heads, T, Nf, "time")
    tr = x + t
    s = divided_attention(_ln(tr, sd, prefix + "norm1", eps), sd, prefix + "attn.", heads, T, Nf, "space")
    if y is not None:
        s = cross_attention_i2t(s, y, y_mask, sd, prefix + "attn.", heads)
    sr = x + s
    m = _lin(F.gelu(_lin(_ln(sr, sd, prefix + "norm2", eps), sd, prefix + "mlp.fc1")), sd, prefix + "mlp.fc2")
    return sr + m


def video_features(video, sd, heads, depth, vprefix="video_model."):
    """SpaceTimeTransformer.forward_features (video_transformer.py:353-394): PASS 1 video CLS."""
    T = video.shape[1]
    x = video_tokens(video, sd, sd[vprefix + "cls_token"], vprefix)
    Nf = (x.shape[1] - 1) // T
    for i in range(depth):
        x = space_time_block(x, sd, f"{vprefix}blocks.{i}.", heads, T, Nf)
    return _ln(x, sd, vprefix + "norm", 1e-5)[:, 0]


# --------------------------------------------------------------------------- text side


def extended_mask(attention_mask):
    """PreTrainedModel.get_extended_attention_mask (transformers): (1-m)*finfo(fp32).min, [B,1,1,S]."""
    m = attention_mask[:, None, None, :].to(torch.float32)
    return (1.0 - m) * torch.finfo(torch.float32).min


# Train-mode dropout of the text tower (roberta.py:203 embeddings, :313 attention probabilities, :342 / :422 dense outputs).
# None = eval mode.  Tests install a callable (kind, layer_slot, tensor) -> tensor that replays a given keep mask at the
# reference's dropout sites (kinds: 0 embeddings, 1 self-attention probabilities, 2 self-attention output dense,
# 3 cross-attention probabilities, 4 cross-attention output dense, 5 feed-forward output dense; layer_slot 0 = embeddings,
# i + 1 = encoder layer i), in the reference's call order.
DROPOUT = None


def _drop(kind, slot, x):
    return x if DROPOUT is None else DROPOUT(kind, slot, x)


def _layer_slot(prefix):
    import re
    m = re.search(r"layer\.(\d+)\.", prefix)
    return int(m.group(1)) + 1 if m else 1


def roberta_embeddings(input_ids, sd, tprefix="text_model.", pad_id=1, eps=1e-5):
    """RobertaEmbeddings.forward (roberta.py:174-204) + create_position_ids_from_input_ids (:881-892); dropout (:203)
    through the DROPOUT hook."""
    nonpad = input_ids.ne(pad_id).to(torch.int64)
    pos_ids = torch.cumsum(nonpad, dim=1) * nonpad + pad_id
    e = tprefix + "embeddings."
    x = sd[e + "word_embeddings.weight"][input_ids] + sd[e + "token_type_embeddings.weight"][0] \
        + sd[e + "position_embeddings.weight"][pos_ids]
    return _drop(0, 0, _ln(x, sd, e + "LayerNorm", eps))


def _bert_attention(hq, hkv, mask, sd, prefix, heads, kinds=(1, 2)):
    """RobertaSelfAttention + RobertaSelfOutput.dense (roberta.py:257-343); scores are
    scaled by 1/sqrt(d) *after* QK^T (:303)."""
    q = _heads(_lin(hq, sd, prefix + "self.query"), heads)
    k = _heads(_lin(hkv, sd, prefix + "self.key"), heads)
    v = _heads(_lin(hkv, sd, prefix + "self.value"), heads)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    if mask is not None:
        s = s + mask
    slot = _layer_slot(prefix)
    p = _drop(kinds[0], slot, torch.softmax(s, dim=-1))                                   # roberta.py:313
    return _drop(kinds[1], slot, _lin(_merge(p @ v), sd, prefix + "output.dense"))        # roberta.py:341-342


def roberta_layer(h, ext_mask, sd, prefix, heads, video=None, eps=1e-5):
    """RobertaLayer.forward (roberta.py:444-505) with the gated text->video
    cross-attention (:470-486; K,V from the *un-normalised* video stream, no mask)."""
    s = _bert_attention(h, h, ext_mask, sd, prefix + "attention.", heads)
    if video is not None:
        c = _bert_attention(s, video, None, sd, prefix + "crossattention_t2i.", heads, kinds=(3, 4))
        s = sd[prefix + "alpha_t2i"] * c + s
    a = _ln(s + h, sd, prefix + "attention.output.LayerNorm", eps)
    f = _drop(5, _layer_slot(prefix), _lin(F.gelu(_lin(a, sd, prefix + "intermediate.dense")), sd, prefix + "output.dense"))   # :421-422
    return _ln(f + a, sd, prefix + "output.LayerNorm", eps)


def text_features(input_ids, attention_mask, sd, heads, depth, tprefix="text_model."):
    """RobertaModel.forward -> last_hidden_state (roberta.py:761-878), no cross-attention (PASS 1)."""
    h = roberta_embeddings(input_ids, sd, tprefix)
    m = extended_mask(attention_mask)
    for i in range(depth):
        h = roberta_layer(h, m, sd, f"{tprefix}encoder.layer.{i}.", heads)
    return h


# --------------------------------------------------------------------------- heads / losses


def projection(x, sd, name):
    """txt_proj / vid_proj 'minimal' head (model.py:105-115): Linear(no bias)-ReLU-Linear-ReLU-Linear."""
    x = F.relu(F.linear(x, sd[name + ".0.weight"]))
    x = F.relu(_lin(x, sd, name + ".2"))
    return _lin(x, sd, name + ".4")


def sim_matrix(a, b, eps=1e-8):
    """model.py:576-584."""
    an = a / a.norm(dim=1, keepdim=True).clamp_min(eps)
    bn = b / b.norm(dim=1, keepdim=True).clamp_min(eps)
    return an @ bn.t()


def egonce(x, sim_v, sim_n, temperature=0.05):
    """loss.EgoNCE.forward with noun=verb=True (loss.py:40-61). Returns (loss, mask_bool)."""
    mask = (sim_v * sim_n + torch.eye(x.shape[0], dtype=x.dtype, device=x.device)) > 0
    i_sm = torch.softmax(x / temperature, dim=1)
    j_sm = torch.softmax(x.t() / temperature, dim=1)
    li = torch.log((i_sm * mask).sum(1)).mean()
    lj = torch.log((j_sm * mask).sum(1)).mean()
    return -li - lj, mask


def norm_softmax_loss(x, temperature=0.05):
    """loss.NormSoftmaxLoss.forward (loss.py:19-31)."""
    i = torch.log_softmax(x / temperature, dim=1)
    j = torch.log_softmax(x.t() / temperature, dim=1)
    return -torch.diag(i).sum() / x.shape[0] - torch.diag(j).sum() / x.shape[0]


def max_margin_ranking_loss(x, margin, weight=None, fix_norm=True):
    """loss.MaxMarginRankingLoss.forward (loss.py:73-100) and, with `weight` [n], AdaptiveMaxMarginRankingLoss.forward
    (loss.py:110-143): pairs (x_ii, x_ij) then (x_ii, x_ji) for all i, j; fix_norm removes i == j."""
    n = x.shape[0]
    terms = []
    for i in range(n):
        m = margin * (float(weight[i]) if weight is not None else 1.0)
        for j in range(n):
            if fix_norm and i == j:
                continue
            terms.append(torch.relu(m - (x[i, i] - x[i, j])))
            terms.append(torch.relu(m - (x[i, i] - x[j, i])))
    return torch.stack(terms).mean()


def dual_step(data, sd, heads, depth, dataset_name="charades", temperature=0.05, margin=0.2):
    """model_epic_charades.FrozenInTime.forward(task_names='Dual') for world size 1 (model_epic_charades.py:408-445):
    txt_proj = Linear(ReLU(cls)) and vid_proj = Linear(cls), 256 wide (:118-119); NormSoftmaxLoss for 'charades',
    AdaptiveMaxMarginRankingLoss weighted by data['relation'] for 'epic' (:425-430)."""
    t = text_features(data["input_ids"], data["attention_mask"], sd, heads, depth)[:, 0]
    t = _lin(F.relu(t), sd, "txt_proj.1")
    v = _lin(video_features(data["video"], sd, heads, depth), sd, "vid_proj.0")
    sim = sim_matrix(t, v)
    if dataset_name == "epic":
        loss = max_margin_ranking_loss(sim, margin, data["relation"])
    elif dataset_name == "charades":
        loss = norm_softmax_loss(sim, temperature)
    else:
        raise NameError()
    return dict(text_embeds=t, video_embeds=v, sim_v2t=sim, Dual=loss)


def mlm_head(x, sd, name="mlm_score", eps=1e-12):
    """heads.MLMHead (heads.py:38-50): BertPredictionHeadTransform (dense, erf-GELU, LN) -> decoder + bias."""
    t = _ln(F.gelu(_lin(x, sd, name + ".transform.dense")), sd, name + ".transform.LayerNorm", eps)
    return F.linear(t, sd[name + ".decoder.weight"]) + sd[name + ".bias"]


def fused_stack(video, input_ids, attention_mask, sd, heads, depth, n_fuse):
    """The fused MLM/ITM trunk of FrozenInTime.infer (model.py:211-271 / 295-357).

    depth-n_fuse plain blocks/layers per modality, then n_fuse fused pairs in which
    video block i gets the text entering layer i and text layer i gets the video
    *entering* block i.  Returns (video_tokens_out [B,N,C], text_out [B,S,C])."""
    T = video.shape[1]
    x = video_tokens(video, sd, sd["cls_token"])
    Nf = (x.shape[1] - 1) // T
    unf = depth - n_fuse
    for i in range(unf):
        x = space_time_block(x, sd, f"video_model.blocks.{i}.", heads, T, Nf)
    h = roberta_embeddings(input_ids, sd)
    m = extended_mask(attention_mask)
    for i in range(unf):
        h = roberta_layer(h, m, sd, f"text_model.encoder.layer.{i}.", heads)
    for i in range(unf, depth):
        x_next = space_time_block(x, sd, f"video_model.blocks.{i}.", heads, T, Nf, y=h, y_mask=m)
        h = roberta_layer(h, m, sd, f"text_model.encoder.layer.{i}.", heads, video=x)
        x = x_next
    return x, h


def itm_logits(x, h, sd):
    """ITM head (model.py:275-290, heads.py:15-35); FrozenInTime.norm has eps 1e-6 (model.py:154)."""
    v = _ln(x, sd, "norm", 1e-6)[:, 0]
    t = torch.tanh(_lin(_lin(h[:, 0], sd, "cross_modal_text_transform"), sd, "cross_modal_text_pooler.dense"))
    v = torch.tanh(_lin(_lin(v, sd, "cross_modal_video_transform"), sd, "cross_modal_video_pooler.dense"))
    return _lin(torch.cat([t, v], dim=-1), sd, "itm_score.fc")


def mlm_logits(h, sd):
    """MLM branch tail (model.py:360-365)."""
    return mlm_head(_lin(h, sd, "cross_modal_text_transform"), sd)


def itm_negative_weights(sim_t2v_rows, mask_rows, temperature=0.05):
    """Hard-negative sampling weights (model.py:442-447). `sim` rows = this rank's rows
    of ret['sim_v2t'] (= text x video sims) or its transpose."""
    w = torch.softmax(sim_t2v_rows / temperature, dim=1)
    return w.masked_fill(mask_rows, 0.0)


def build_itm_batch(data, itm_labels, swap_video, neg_idx):
    """Deterministic restatement of the python loop model.py:449-468 (single rank):
    for a label-0 row either the clip (swap_video[i]) or the caption is replaced by
    sample neg_idx[i]; label-1 rows keep their own pair."""
    video = data["video"].clone()
    ids = data["input_ids"].clone()
    am = data["attention_mask"].clone()
    for i in range(video.shape[0]):
        if int(itm_labels[i]) == 1:
            continue
        j = int(neg_idx[i])
        if bool(swap_video[i]):
            video[i] = data["video"][j]
        else:
            ids[i] = data["input_ids"][j]
            am[i] = data["attention_mask"][j]
    return video, ids, am


def pretrain_step(data, sd, heads, depth, n_fuse, itm_plan=None, tasks="EgoNCE_MLM_ITM"):
    """FrozenInTime.forward for one rank / world size 1 (model.py:370-487).

    data: video [B,T,3,H,W], input_ids, attention_mask, text_mlm_ids, text_mlm_labels,
          noun_vec [B,582], verb_vec [B,118]
    itm_plan: dict(labels [B], swap_video [B] bool, neg_idx [B]) making the host-RNG
          part of the ITM pass (:434-468) explicit.
    Returns dict of losses and the tensors the reference puts in `ret`.
    """
    out = {}
    t = text_features(data["input_ids"], data["attention_mask"], sd, heads, depth)[:, 0]
    t = projection(t, sd, "txt_proj")
    v = projection(video_features(data["video"], sd, heads, depth), sd, "vid_proj")
    sim = sim_matrix(t, v)
    loss_nce, mask = egonce(sim, sim_matrix(data["verb_vec"], data["verb_vec"]),
                            sim_matrix(data["noun_vec"], data["noun_vec"]))
    out.update(text_embeds=t, video_embeds=v, sim_v2t=sim, mask_bool=mask, EgoNCE=loss_nce)
    total = loss_nce
    if "MLM" in tasks:
        _, h = fused_stack(data["video"], data["text_mlm_ids"], data["attention_mask"], sd, heads, depth, n_fuse)
        logits = mlm_logits(h, sd)
        loss_mlm = F.cross_entropy(logits.view(-1, logits.shape[-1]), data["text_mlm_labels"].view(-1),
                                   ignore_index=-100)
        out.update(cross_attn_mlm_logits=logits, loss_mlm=loss_mlm)
        total = total + loss_mlm
    if "ITM" in tasks:
        video, ids, am = build_itm_batch(data, itm_plan["labels"], itm_plan["swap_video"], itm_plan["neg_idx"])
        x, h = fused_stack(video, ids, am, sd, heads, depth, n_fuse)
        logits = itm_logits(x, h, sd)
        loss_itm = F.cross_entropy(logits, itm_plan["labels"].long())
        out.update(cross_attn_itm_logits=logits, loss_itm=loss_itm)
        total = total + 2 * loss_itm
    out["loss_total"] = total
    return out


# --------------------------------------------------------------------------- synthetic inputs / weights


def key_shapes(C=768, heads=12, depth=12, n_fuse=6, T=16, img=224, patch=16, S_max=514, vocab=50265,
               proj=4096, mlp_ratio=4, fuse_from_video=None):
    """state_dict schema of FrozenInTime as {key: shape}."""
    Nf = (img // patch) ** 2
    H = C * mlp_ratio
    unf = depth - n_fuse if fuse_from_video is None else fuse_from_video
    ks = {"cls_token": (1, 1, C), "norm.weight": (C,), "norm.bias": (C,)}
    v = "video_model."
    ks.update({v + "cls_token": (1, 1, C), v + "pos_embed": (1, Nf + 1, C), v + "temporal_embed": (1, T, C),
               v + "patch_embed.proj.weight": (C, 3, patch, patch), v + "patch_embed.proj.bias": (C,),
               v + "norm.weight": (C,), v + "norm.bias": (C,)})

    def lin(name, o, i, bias=True):
        ks[name + ".weight"] = (o, i)
        if bias:
            ks[name + ".bias"] = (o,)

    def ln(name):
        ks[name + ".weight"] = (C,)
        ks[name + ".bias"] = (C,)

    for i in range(depth):
        b = f"{v}blocks.{i}."
        for n in ("norm1", "norm2", "norm3"):
            ln(b + n)
        for a in ("attn.", "timeattn."):
            lin(b + a + "qkv", 3 * C, C)
            lin(b + a + "proj", C, C)
        lin(b + "mlp.fc1", H, C)
        lin(b + "mlp.fc2", C, H)
        if i >= unf:
            ks[b + "attn.alpha_i2t"] = (1,)
            lin(b + "attn.qkv_text_i2t", 2 * C, C)
            lin(b + "attn.qkv_i2t", C, C)
            lin(b + "attn.proj_i2t", C, C)
            ln(b + "attn.norm_i2t_i")
    t = "text_model."
    e = t + "embeddings."
    ks.update({e + "word_embeddings.weight": (vocab, C), e + "position_embeddings.weight": (S_max, C),
               e + "token_type_embeddings.weight": (1, C)})
    ln(e + "LayerNorm")
    for i in range(depth):
        b = f"{t}encoder.layer.{i}."
        for n in ("query", "key", "value"):
            lin(b + "attention.self." + n, C, C)
        lin(b + "attention.output.dense", C, C)
        ln(b + "attention.output.LayerNorm")
        lin(b + "intermediate.dense", H, C)
        lin(b + "output.dense", C, H)
        ln(b + "output.LayerNorm")
        if i >= depth - n_fuse:
            ks[b + "alpha_t2i"] = (1,)
            for n in ("query", "key", "value"):
                lin(b + "crossattention_t2i.self." + n, C, C)
            lin(b + "crossattention_t2i.output.dense", C, C)
    for p in ("txt_proj", "vid_proj"):
        lin(p + ".0", proj, C, bias=False)
        lin(p + ".2", proj, proj)
        lin(p + ".4", proj, proj)
    for n in ("cross_modal_text_transform", "cross_modal_video_transform",
              "cross_modal_text_pooler.dense", "cross_modal_video_pooler.dense"):
        lin(n, C, C)
    ks["mlm_score.bias"] = (vocab,)
    lin("mlm_score.transform.dense", C, C)
    ln("mlm_score.transform.LayerNorm")
    ks["mlm_score.decoder.weight"] = (vocab, C)
    lin("itm_score.fc", 2, 2 * C)
    return ks


def dual_key_shapes(**kw):
    """state_dict schema of model_epic_charades.FrozenInTime: the pre-training schema with the 256-wide projections
    txt_proj = Sequential(ReLU, Linear), vid_proj = Sequential(Linear) (model_epic_charades.py:118-119)."""
    shapes = {k: v for k, v in key_shapes(**kw).items() if not k.startswith(("txt_proj.", "vid_proj."))}
    C = kw.get("C", 768)
    shapes.update({"txt_proj.1.weight": (256, C), "txt_proj.1.bias": (256,),
                   "vid_proj.0.weight": (256, C), "vid_proj.0.bias": (256,)})
    return shapes


def seeded_state(shapes, seed=0, device="cpu"):
    """Deterministic weights for parity runs: N(0, .02) matrices, LN weight ~ 1+N(0,.1),
    biases N(0,.02), gates alpha_* = 0.5 (the reference inits them to 0, which would hide
    the fusion branches -- SURVEY.md Q1/Q2).  Keys are visited in sorted order so the
    recipe is reproducible from (shapes, seed) alone; fixtures never store weights."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = shapes[k]
        if k.endswith("alpha_i2t") or k.endswith("alpha_t2i"):
            t = torch.full(shp, 0.5)
        elif ("norm" in k.lower()) and k.endswith(".weight") and len(shp) == 1:
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(("pos_embed", "temporal_embed", "cls_token")):
            t = 0.1 * torch.randn(shp, generator=g)
        else:
            t = 0.02 * torch.randn(shp, generator=g)
        sd[k] = t.to(device)
    return sd


def synthetic_batch(B, T, img, S, seed=1234, vocab=50265, n_noun=582, n_verb=118, device="cpu"):
    """Seeded synthetic batch of SURVEY.md section 8(d)."""
    g = torch.Generator().manual_seed(seed)
    video = torch.randn(B, T, 3, img, img, generator=g)
    lens = torch.randint(min(8, S), S + 1, (B,), generator=g)
    ids = torch.randint(3, vocab - 1, (B, S), generator=g)
    pos = torch.arange(S)[None]
    ids[:, 0] = 0
    ids = torch.where(pos == (lens[:, None] - 1), torch.full_like(ids, 2), ids)
    ids = torch.where(pos >= lens[:, None], torch.full_like(ids, 1), ids)
    am = (pos < lens[:, None]).to(torch.int64)
    special = (ids <= 2)
    pick = (torch.rand(B, S, generator=g) < 0.15) & ~special
    for b in range(B):           # at least one label per row keeps the CE well-defined
        if not pick[b].any():
            pick[b, 1] = True
    labels = torch.where(pick, ids, torch.full_like(ids, -100))
    r = torch.rand(B, S, generator=g)
    rnd = torch.randint(3, vocab - 1, (B, S), generator=g)
    mlm_ids = ids.clone()
    mlm_ids[pick & (r < 0.8)] = vocab - 1
    sel = pick & (r >= 0.8) & (r < 0.9)
    mlm_ids[sel] = rnd[sel]
    noun = (torch.rand(B, n_noun, generator=g) < 0.01).float()
    verb = (torch.rand(B, n_verb, generator=g) < 0.02).float()
    noun[torch.arange(B), torch.randint(0, n_noun, (B,), generator=g)] = 1.0
    verb[torch.arange(B), torch.randint(0, n_verb, (B,), generator=g)] = 1.0
    d = dict(video=video, input_ids=ids, attention_mask=am, text_mlm_ids=mlm_ids, text_mlm_labels=labels,
             noun_vec=noun, verb_vec=verb)
    return {k: v.to(device) for k, v in d.items()}


def synthetic_itm_plan(B, seed=4321):
    """Fixed ITM plan for throughput / parity runs: first ceil(B/2)... the reference
    shuffles floor(B/2) positives; we fix the permutation with a seed and draw the
    negatives uniformly from the other rows (the sampling *rule* is tested separately)."""
    g = torch.Generator().manual_seed(seed)
    labels = torch.cat([torch.ones(B // 2), torch.zeros(B - B // 2)])[torch.randperm(B, generator=g)]
    swap = torch.rand(B, generator=g) > 0.5
    neg = (torch.arange(B) + torch.randint(1, max(B, 2), (B,), generator=g)) % B
    return dict(labels=labels, swap_video=swap, neg_idx=neg)
