"""Host-side sequencing of the CUDA kernels for every module on the EgoVLPv2 hot path.

Each `*_fwd` launches the forward kernels of one reference module and returns (output, saved);
each `*_bwd` takes the saved activations and the output gradient and launches the hand-derived
backward kernels, returning input and parameter gradients.  No math is done in torch here: torch
allocates buffers and makes views, `K` (egovlpv2_b200.lib.Kernels) does the arithmetic.

Precision model (SURVEY.md Appendix D with bf16 in place of fp16): GEMM operands bf16, fp32
accumulation; residual streams, LayerNorm statistics, softmax and losses fp32; saved activations bf16.

Parameter dictionaries `p` map the reference's parameter names (relative to the module) to fp32
tensors; `w` maps weight names to their bf16 operand copies (see weights.WeightCache).
"""
import math
import types

import torch

from . import xattn_reassoc as XR
from .lib import (ACT_GELU, ACT_GELU_BWD, ACT_GELU_DG, ACT_MUL_AUX, ACT_NONE, ACT_RELU, ACT_RELU_BWD, ACT_TANH, ACT_TANH_BWD,
                  GEMM_NN, GEMM_NT, GEMM_TN, AttnSpec)

BF16, F32 = torch.bfloat16, torch.float32


def _e(like, shape, dtype):
    return torch.empty(shape, dtype=dtype, device=like.device)


def _z(like, shape, dtype=F32):
    return torch.zeros(shape, dtype=dtype, device=like.device)


_REASSOC = None


def reassoc_enabled():
    """The re-associated cross-attention (xattn_reassoc.py) is the default; EGV_XATTN_REASSOC=0 selects the round-1
    formulation (separate query / key / value projections + the strided attention kernels), kept for shapes the
    re-associated kernels do not cover (video->text: text length != 32; text->video: text length > 128)."""
    global _REASSOC
    if _REASSOC is None:
        import os
        _REASSOC = os.environ.get("EGV_XATTN_REASSOC", "1") != "0"
    return _REASSOC


def set_reassoc(on):
    global _REASSOC
    _REASSOC = bool(on)


# ----------------------------------------------------------------------------------------------- parameter gradients
class Grads:
    """Collects parameter gradients of one backward call.

    Without a sink every gradient is a fresh tensor returned to autograd (`self.g[name]`).  With `sink`
    (name -> fp32 tensor, normally views of the flat gradient arena, already holding the running sum of this step)
    the kernels accumulate straight into it -- weight-gradient GEMMs with atomic split-K, bias / LayerNorm gradients
    through the atomics they use anyway -- and nothing is returned for that parameter: no temporaries, no zero fills,
    no `grad += g` launches."""

    def __init__(self, K, like, sink=None):
        self.K, self.like, self.sink, self.g = K, like, sink or {}, {}

    def weight(self, name, dy_bf, x_bf, scale_dev=None, scale=1.0):
        """dW [N, Kd] (+)= scale * dy^T x"""
        N, Kd = dy_bf.shape[1], x_bf.shape[1]
        if name in self.sink:
            self.K.gemm(GEMM_TN, dy_bf, x_bf, out_f32=self.sink[name].view(N, Kd), scale=scale, scale_dev=scale_dev,
                        accumulate=True)
        else:
            dW = _e(dy_bf, (N, Kd), F32)
            self.K.gemm(GEMM_TN, dy_bf, x_bf, out_f32=dW, scale=scale, scale_dev=scale_dev)
            self.g[name] = dW

    def vec(self, name, n):
        """fp32 [n] accumulator for kernels that add with atomics (LayerNorm dgamma/dbeta, fused column sums)."""
        if name in self.sink:
            return self.sink[name].view(n)
        t = _z(self.like, (n,))
        self.g[name] = t
        return t

    def bias(self, name, dy_bf, scale_dev=None, scale=1.0):
        """db [N] (+)= scale * colsum(dy)"""
        N = dy_bf.shape[1]
        if name in self.sink:
            self.K.colsum(dy_bf, self.sink[name].view(N), accumulate=True, scale=scale, scale_dev=scale_dev)
        else:
            db = _e(dy_bf, (N,), F32)
            self.K.colsum(dy_bf, db, scale=scale, scale_dev=scale_dev)
            self.g[name] = db

    def bias_from(self, name, cs, scale_dev=None, scale=1.0):
        """db (+)= scale * cs, cs = column sums produced by another kernel's epilogue"""
        N = cs.numel()
        if name in self.sink:
            t = self.sink[name].view(N)
            self.K.axpy(t, cs, scale, scale_dev, y=t)
        elif scale_dev is None and scale == 1.0:
            self.g[name] = cs
        else:
            db = _e(cs, (N,), F32)
            self.K.axpy(None, cs, scale, scale_dev, y=db)
            self.g[name] = db

    def scalar_dot(self, name, a, b):
        """g (+)= sum(a * b)   (the scalar fusion gates)"""
        if name in self.sink:
            self.K.dot(a, b, self.sink[name].view(1), accumulate=True)
        else:
            t = _e(self.like, (1,), F32)
            self.K.dot(a, b, t)
            self.g[name] = t

    def full(self, name, shape):
        """zero-initialised (or sink) tensor for kernels that scatter-add (embedding tables)."""
        if name in self.sink:
            return self.sink[name].view(shape)
        t = _z(self.like, shape)
        self.g[name] = t
        return t


# ----------------------------------------------------------------------------------------------- divided attention
def _divided_specs(H, T, Nf, scale):
    N = 1 + T * Nf
    time = AttnSpec(H=H, G=Nf, Lq=T, Lk=T, q_row0=1, q_gstride=1, q_istride=Nf, k_row0=1, k_gstride=1, k_istride=Nf,
                    has_cls_key=True, cls_row=0, scale=scale)
    space = AttnSpec(H=H, G=T, Lq=Nf, Lk=Nf, q_row0=1, q_gstride=Nf, q_istride=1, k_row0=1, k_gstride=Nf, k_istride=1,
                     has_cls_key=True, cls_row=0, scale=scale)
    cls = AttnSpec(H=H, G=1, Lq=1, Lk=N - 1, q_row0=0, k_row0=1, k_istride=1, has_cls_key=True, cls_row=0, scale=scale)
    return time, space, cls


def divided_attention_fwd(K, qkv, H, T, Nf, mode):
    """VarAttention core (video_transformer.py:123-150) on qkv [B, N, 3C] bf16 -> o [B, N, C] bf16.
    Patch queries attend CLS + their time/space group; the CLS query attends every token."""
    B, N, C3 = qkv.shape
    C = C3 // 3
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    time, space, cls = _divided_specs(H, T, Nf, (C // H) ** -0.5)
    spec = time if mode == "time" else space
    o = _e(qkv, (B, N, C), BF16)
    lse = _e(qkv, (B * H * spec.G * spec.Lq,), F32)
    lse_cls = _e(qkv, (B * H,), F32)
    # the CLS query rides along in the space-attention kernel when it can take it (csrc/attention_tc.cu: per-frame partials
    # + the single-query combine); else the separate single-query pass over all keys
    folded = False
    if mode == "space" and getattr(K, "supports_cls_fold", False):
        folded = K.attention_fwd(spec, q, k, v, o, lse, lse_cls=lse_cls)
    else:
        K.attention_fwd(spec, q, k, v, o, lse)
    if not folded:
        K.attention_fwd(cls, q, k, v, o, lse_cls)
    return o, (lse, lse_cls)


def divided_attention_bwd(K, qkv, o, lses, d_o, H, T, Nf, mode):
    """-> d_qkv [B, N, 3C] bf16 (every element written)."""
    B, N, C3 = qkv.shape
    C = C3 // 3
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    time, space, cls = _divided_specs(H, T, Nf, (C // H) ** -0.5)
    spec = time if mode == "time" else space
    lse, lse_cls = lses
    d_qkv = _e(qkv, (B, N, C3), BF16)
    dq, dk, dv = d_qkv[:, :, :C], d_qkv[:, :, C:2 * C], d_qkv[:, :, 2 * C:]
    # the CLS query (video_transformer.py:134-150) rides along in the space-attention backward when the kernel can take it
    # (csrc/attention_tc_bwd.cu); its dq accumulator shares the zero fill of the CLS key's accumulators
    try_fold = mode == "space" and getattr(K, "supports_cls_fold", False)
    acc = _z(qkv, (B * H * (192 if try_fold else 128),))
    dkv_cls = acc[:B * H * 128]
    folded = False
    if try_fold:
        dq_cls = acc[B * H * 128:]
        folded = K.attention_bwd(spec, q, k, v, o, lse, d_o, dq, dk, dv, _e(qkv, lse.shape, F32), dkv_cls=dkv_cls,
                                 lse_cls=lse_cls, dq_cls=dq_cls)
    else:
        K.attention_bwd(spec, q, k, v, o, lse, d_o, dq, dk, dv, _e(qkv, lse.shape, F32), dkv_cls=dkv_cls)
    if folded:
        K.attention_cls_query_finalize(dq_cls, dq, H, cls_row=0)
    else:
        K.attention_bwd(cls, q, k, v, o, lse_cls, d_o, dq, dk, dv, _e(qkv, lse_cls.shape, F32), dkv_cls=dkv_cls,
                        dkv_accumulate=True)
    K.attention_cls_finalize(dkv_cls, dk, dv, H, cls_row=0, accumulate=False)
    return d_qkv


# ----------------------------------------------------------------------------------------------- SpaceTimeBlock
VIDEO_BLOCK_PARAMS = ["norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias", "norm3.weight", "norm3.bias",
                      "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
                      "timeattn.qkv.weight", "timeattn.qkv.bias", "timeattn.proj.weight", "timeattn.proj.bias",
                      "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias"]
VIDEO_FUSE_PARAMS = ["attn.alpha_i2t", "attn.qkv_text_i2t.weight", "attn.qkv_text_i2t.bias", "attn.qkv_i2t.weight",
                     "attn.qkv_i2t.bias", "attn.proj_i2t.weight", "attn.proj_i2t.bias", "attn.norm_i2t_i.weight",
                     "attn.norm_i2t_i.bias"]


I2T_PREP_PARAMS = ["attn.qkv_text_i2t.weight", "attn.qkv_text_i2t.bias", "attn.qkv_i2t.weight", "attn.qkv_i2t.bias",
                   "attn.proj_i2t.weight"]


def i2t_prep_fwd(K, y, y_bias, p, w, H):
    """Text side of the re-associated gated video->text cross-attention (xattn_reassoc.i2t_prep_fwd): a function of the
    text states and weights only, so the model runs it on the text tower's stream.  y [B,S,Ct] f32; y_bias [B,S] f32.
    -> ((Mt, U, bias1), saved)."""
    B, S, Ct = y.shape
    C = w["attn.qkv_i2t.weight"].shape[0]
    K.mark("xattn_i2t_fwd_prep")
    s = types.SimpleNamespace(B=B, S=S, Ct=Ct)
    s.y_bf = _e(y, (B * S, Ct), BF16)
    K.cast(y.reshape(B * S, Ct).contiguous(), s.y_bf)
    s.kv_t = _e(y, (B * S, 2 * C), BF16)
    K.gemm(GEMM_NT, s.y_bf, w["attn.qkv_text_i2t.weight"], bias=p["attn.qkv_text_i2t.bias"], out_bf16=s.kv_t)
    prep = XR.i2t_prep_fwd(K, s.kv_t, y_bias, w["attn.qkv_i2t.weight"], p["attn.qkv_i2t.bias"], w["attn.proj_i2t.weight"], B, H)
    return prep, s


def i2t_prep_bwd(K, s, dM_b, dU_b, dbias1, p, w, H, G):
    """Backward of i2t_prep_fwd; parameter gradients go into the Grads collector G.  -> dy [B,S,Ct] f32."""
    B, S, Ct = s.B, s.S, s.Ct
    C = w["attn.qkv_i2t.weight"].shape[0]
    K.mark("xattn_i2t_bwd_prep")
    dkv_f = XR.i2t_prep_bwd(K, s.kv_t, w["attn.qkv_i2t.weight"], p["attn.qkv_i2t.bias"], w["attn.proj_i2t.weight"], dM_b, dU_b,
                            dbias1, G.full("attn.qkv_i2t.weight", (C, C)), G.full("attn.qkv_i2t.bias", (C,)),
                            G.full("attn.proj_i2t.weight", (C, C)), B, H)
    dkv2 = _e(dkv_f, (B * S, 2 * C), BF16)
    K.cast(dkv_f, dkv2)
    G.weight("attn.qkv_text_i2t.weight", dkv2, s.y_bf)
    G.bias("attn.qkv_text_i2t.bias", dkv2)
    dy = _e(dkv_f, (B, S, Ct), F32)
    K.gemm(GEMM_NN, dkv2, w["attn.qkv_text_i2t.weight"], out_f32=dy.view(B * S, Ct))
    return dy


def video_block_fwd(K, x, p, w, H, T, Nf, y=None, y_bias=None, eps=1e-5, save=True, prep=None):
    """SpaceTimeBlock.forward (video_transformer.py:214-228), optionally with the gated video->text
    cross-attention of VarAttention (video_transformer.py:155-185).
    x [B,N,C] f32; y [B,S,Ct] f32 text states; y_bias [B,S] f32 additive key mask.  `prep`: (Mt, U, bias1) from
    i2t_prep_fwd when the caller prepared the text side itself (the model does, on the text stream); their gradients
    then come back from video_block_bwd instead of dy.  Returns (out f32, saved)."""
    B, N, C = x.shape
    M = B * N
    x2 = x.reshape(M, C)
    s = types.SimpleNamespace(fused=y is not None or prep is not None, shape=(B, N, C), x=x2)
    K.mark("video_block")
    # ---- time attention branch: t = proj(attn_time(qkv(norm3(x))))
    s.ln3, s.mean3, s.rstd3 = _e(x, (M, C), BF16), _e(x, (M,), F32), _e(x, (M,), F32)
    K.layernorm_fwd(x2, p["norm3.weight"], p["norm3.bias"], eps, y_bf16=s.ln3, mean=s.mean3, rstd=s.rstd3)
    s.qkv_t = _e(x, (M, 3 * C), BF16)
    K.gemm(GEMM_NT, s.ln3, w["timeattn.qkv.weight"], bias=p["timeattn.qkv.bias"], out_bf16=s.qkv_t)
    s.o_t, s.lse_t = divided_attention_fwd(K, s.qkv_t.view(B, N, 3 * C), H, T, Nf, "time")
    s.tr = _e(x, (M, C), F32)
    K.gemm(GEMM_NT, s.o_t.view(M, C), w["timeattn.proj.weight"], bias=p["timeattn.proj.bias"], residual=x2, out_f32=s.tr)
    # ---- space attention branch on norm1(x + t); residual goes back to the block INPUT x (:218-222)
    s.ln1, s.mean1, s.rstd1 = _e(x, (M, C), BF16), _e(x, (M,), F32), _e(x, (M,), F32)
    K.layernorm_fwd(s.tr, p["norm1.weight"], p["norm1.bias"], eps, y_bf16=s.ln1, mean=s.mean1, rstd=s.rstd1)
    s.qkv_s = _e(x, (M, 3 * C), BF16)
    K.gemm(GEMM_NT, s.ln1, w["attn.qkv.weight"], bias=p["attn.qkv.bias"], out_bf16=s.qkv_s)
    s.o_s, s.lse_s = divided_attention_fwd(K, s.qkv_s.view(B, N, 3 * C), H, T, Nf, "space")
    s.sr = _e(x, (M, C), F32)
    if not s.fused:
        K.gemm(GEMM_NT, s.o_s.view(M, C), w["attn.proj.weight"], bias=p["attn.proj.bias"], residual=x2, out_f32=s.sr)
    else:
        # a = proj(o_s) (bf16 copy feeds the cross-attention LN);  xa = x + a (fp32)
        s.a = _e(x, (M, C), BF16)
        xa = _e(x, (M, C), F32)
        K.gemm(GEMM_NT, s.o_s.view(M, C), w["attn.proj.weight"], bias=p["attn.proj.bias"], residual=x2, out_f32=xa,
               out_pre=s.a)
        Sx = 32 if prep is not None else y.shape[1]
        Ctx = C if prep is not None else y.shape[2]
        s.alg = float(B) * (4.0 * N * C * C + 4.0 * Sx * Ctx * C + 4.0 * N * Sx * C)   # xattn_i2t per launch
        K.mark("xattn_i2t_fwd", alg=s.alg)
        s.lnc, s.meanc, s.rstdc = _e(x, (M, C), BF16), _e(x, (M,), F32), _e(x, (M,), F32)
        K.layernorm_fwd(s.a, p["attn.norm_i2t_i.weight"], p["attn.norm_i2t_i.bias"], eps, y_bf16=s.lnc, mean=s.meanc,
                        rstd=s.rstdc)
        s.y_bias = y_bias
        s.reassoc = prep is not None or (reassoc_enabled() and XR.i2t_supported(y.shape[1], C, H))
        s.prep_saved = None
        if s.reassoc:
            # re-associated around the S text keys (xattn_reassoc.py): no [M, C] query / output projections
            if prep is None:
                prep, s.prep_saved = i2t_prep_fwd(K, y, y_bias, p, w, H)
                K.mark("xattn_i2t_fwd")
            s.Mt, s.U, bias1 = prep
            s.P = XR.i2t_apply_fwd(K, s.lnc, s.Mt, s.U, bias1, p["attn.proj_i2t.bias"], p["attn.alpha_i2t"], xa, s.sr, B, N, H)
        else:
            S, Ct = y.shape[1], y.shape[2]
            s.y_bf = _e(x, (B * S, Ct), BF16)
            K.cast(y.reshape(B * S, Ct), s.y_bf)
            s.kv_t = _e(x, (B * S, 2 * C), BF16)
            K.gemm(GEMM_NT, s.y_bf, w["attn.qkv_text_i2t.weight"], bias=p["attn.qkv_text_i2t.bias"], out_bf16=s.kv_t)
            s.q_c = _e(x, (M, C), BF16)
            K.gemm(GEMM_NT, s.lnc, w["attn.qkv_i2t.weight"], bias=p["attn.qkv_i2t.bias"], out_bf16=s.q_c)
            s.spec_c = AttnSpec(H=H, G=1, Lq=N, Lk=S, scale=(C // H) ** -0.5)
            kv3 = s.kv_t.view(B, S, 2 * C)
            s.o_c, s.lse_c = _e(x, (B, N, C), BF16), _e(x, (B * H * N,), F32)
            K.attention_fwd(s.spec_c, s.q_c.view(B, N, C), kv3[:, :, :C], kv3[:, :, C:], s.o_c, s.lse_c, key_bias=y_bias)
            # sr = xa + alpha * (proj_i2t(o_c));  c (pre-gate) kept in bf16 for d_alpha
            s.c = _e(x, (M, C), BF16)
            K.gemm(GEMM_NT, s.o_c.view(M, C), w["attn.proj_i2t.weight"], bias=p["attn.proj_i2t.bias"],
                   scale_dev=p["attn.alpha_i2t"], residual=xa, out_f32=s.sr, out_pre=s.c)
    # ---- MLP
    K.mark("video_block")
    s.ln2, s.mean2, s.rstd2 = _e(x, (M, C), BF16), _e(x, (M,), F32), _e(x, (M,), F32)
    K.layernorm_fwd(s.sr, p["norm2.weight"], p["norm2.bias"], eps, y_bf16=s.ln2, mean=s.mean2, rstd=s.rstd2)
    Hd = w["mlp.fc1.weight"].shape[0]
    s.h_pre, s.h_act = _e(x, (M, Hd), BF16), _e(x, (M, Hd), BF16)
    # h_pre holds gelu'(fc1 output): the backward then multiplies instead of re-deriving Phi and phi (ACT_GELU_DG)
    K.gemm(GEMM_NT, s.ln2, w["mlp.fc1.weight"], bias=p["mlp.fc1.bias"], act=ACT_GELU_DG, out_bf16=s.h_act, out_pre=s.h_pre)
    out = _e(x, (M, C), F32)
    K.gemm(GEMM_NT, s.h_act, w["mlp.fc2.weight"], bias=p["mlp.fc2.bias"], residual=s.sr, out_f32=out)
    return out.view(B, N, C), (s if save else None)


def video_block_bwd(K, s, d_out, p, w, H, T, Nf, need_dx=True, sink=None):
    """Backward of video_block_fwd.  Returns (dx [B,N,C] f32 or None, dy [B,S,Ct] f32 or None, grads dict); when the
    forward was given `prep`, grads["_prep"] = (dMt, dU, dbias1), the gradients of those three tensors (and dy is None)."""
    B, N, C = s.shape
    M = B * N
    G = Grads(K, d_out, sink)
    K.mark("video_block_bwd")
    d_out = d_out.reshape(M, C)
    d_out_bf = _e(d_out, (M, C), BF16)
    K.cast(d_out.contiguous(), d_out_bf)
    # ---- MLP: out = sr + fc2(gelu(fc1(ln2)))
    G.weight("mlp.fc2.weight", d_out_bf, s.h_act)
    G.bias("mlp.fc2.bias", d_out_bf)
    d_hpre = _e(d_out, s.h_pre.shape, BF16)
    K.gemm(GEMM_NN, d_out_bf, w["mlp.fc2.weight"], aux=s.h_pre, act=ACT_MUL_AUX, out_bf16=d_hpre,
           colsum=G.vec("mlp.fc1.bias", s.h_pre.shape[1]))
    G.weight("mlp.fc1.weight", d_hpre, s.ln2)
    d_ln2 = _e(d_out, (M, C), BF16)
    K.gemm(GEMM_NN, d_hpre, w["mlp.fc1.weight"], out_bf16=d_ln2)
    del d_hpre
    # d_sr = d_out + LN2'(d_ln2)
    d_sr, d_sr_bf = _e(d_out, (M, C), F32), _e(d_out, (M, C), BF16)
    # column sums of d_sr = bias gradient of attn.proj (plain block) or, times alpha, of proj_i2t (fused block)
    cs_sr = _z(d_out, (C,)) if s.fused else G.vec("attn.proj.bias", C)
    K.layernorm_bwd(d_ln2, s.sr, p["norm2.weight"], s.mean2, s.rstd2, add=d_out, dx=d_sr, dx_bf16=d_sr_bf,
                    bf16_total=True, dgamma=G.vec("norm2.weight", C), dbeta=G.vec("norm2.bias", C), out_colsum=cs_sr)
    dy = None
    d_s_bf = d_sr_bf
    if s.fused:
        alpha = p["attn.alpha_i2t"]
        K.mark("xattn_i2t_bwd", alg=2.0 * s.alg)
        # sr = x + a + alpha * c,  c = proj_i2t(attention)
        G.bias_from("attn.proj_i2t.bias", cs_sr, scale_dev=alpha)
        if s.reassoc:
            d_lnc, dM_b, dU_b, dbias1 = XR.i2t_apply_bwd(K, s.lnc, s.Mt, s.U, s.P, d_sr_bf, cs_sr, p["attn.proj_i2t.bias"], alpha,
                                                         G.full("attn.alpha_i2t", (1,)), B, N, H)
            if s.prep_saved is not None:
                dy = i2t_prep_bwd(K, s.prep_saved, dM_b, dU_b, dbias1, p, w, H, G)
                K.mark("xattn_i2t_bwd")
            else:
                G.g["_prep"] = (dM_b, dU_b, dbias1)
        else:
            S, Ct = s.y_bf.shape[0] // B, s.y_bf.shape[1]
            G.scalar_dot("attn.alpha_i2t", d_sr, s.c)
            G.weight("attn.proj_i2t.weight", d_sr_bf, s.o_c.view(M, C), scale_dev=alpha)
            d_oc = _e(d_out, (B, N, C), BF16)
            K.gemm(GEMM_NN, d_sr_bf, w["attn.proj_i2t.weight"], scale_dev=alpha, out_bf16=d_oc.view(M, C))
            kv3 = s.kv_t.view(B, S, 2 * C)
            dq_c = _e(d_out, (B, N, C), BF16)
            dkv = _e(d_out, (B, S, 2 * C), BF16)
            K.attention_bwd(s.spec_c, s.q_c.view(B, N, C), kv3[:, :, :C], kv3[:, :, C:], s.o_c, s.lse_c, d_oc, dq_c,
                            dkv[:, :, :C], dkv[:, :, C:], _e(d_out, s.lse_c.shape, F32), key_bias=s.y_bias)
            dq2, dkv2 = dq_c.view(M, C), dkv.view(B * S, 2 * C)
            G.weight("attn.qkv_i2t.weight", dq2, s.lnc)
            G.bias("attn.qkv_i2t.bias", dq2)
            d_lnc = _e(d_out, (M, C), BF16)
            K.gemm(GEMM_NN, dq2, w["attn.qkv_i2t.weight"], out_bf16=d_lnc)
            G.weight("attn.qkv_text_i2t.weight", dkv2, s.y_bf)
            G.bias("attn.qkv_text_i2t.bias", dkv2)
            dy = _e(d_out, (B, S, Ct), F32)
            K.gemm(GEMM_NN, dkv2, w["attn.qkv_text_i2t.weight"], out_f32=dy.view(B * S, Ct))
        # d_a = d_sr + LNc'(d_lnc)
        d_a_bf = _e(d_out, (M, C), BF16)
        K.layernorm_bwd(d_lnc, s.a, p["attn.norm_i2t_i.weight"], s.meanc, s.rstdc, add=d_sr, dx=None, dx_bf16=d_a_bf,
                        bf16_total=True, dgamma=G.vec("attn.norm_i2t_i.weight", C), dbeta=G.vec("attn.norm_i2t_i.bias", C),
                        out_colsum=G.vec("attn.proj.bias", C))
        d_s_bf = d_a_bf
        K.mark("video_block_bwd")
    # ---- space attention: s = proj(attn_space(qkv(ln1)))
    G.weight("attn.proj.weight", d_s_bf, s.o_s.view(M, C))
    d_os = _e(d_out, (B, N, C), BF16)
    K.gemm(GEMM_NN, d_s_bf, w["attn.proj.weight"], out_bf16=d_os.view(M, C))
    d_qkv = divided_attention_bwd(K, s.qkv_s.view(B, N, 3 * C), s.o_s, s.lse_s, d_os, H, T, Nf, "space").view(M, 3 * C)
    G.weight("attn.qkv.weight", d_qkv, s.ln1)
    G.bias("attn.qkv.bias", d_qkv)
    d_ln1 = _e(d_out, (M, C), BF16)
    K.gemm(GEMM_NN, d_qkv, w["attn.qkv.weight"], out_bf16=d_ln1)
    # d_tr = LN1'(d_ln1);  running d_x = d_sr + d_tr   (x feeds sr directly and tr directly)
    d_x, d_tr_bf = _e(d_out, (M, C), F32), _e(d_out, (M, C), BF16)
    K.layernorm_bwd(d_ln1, s.tr, p["norm1.weight"], s.mean1, s.rstd1, add=d_sr, dx=d_x, dx_bf16=d_tr_bf,
                    bf16_total=False, dgamma=G.vec("norm1.weight", C), dbeta=G.vec("norm1.bias", C),
                    out_colsum=G.vec("timeattn.proj.bias", C))
    # ---- time attention: t = proj(attn_time(qkv(ln3)))
    G.weight("timeattn.proj.weight", d_tr_bf, s.o_t.view(M, C))
    d_ot = _e(d_out, (B, N, C), BF16)
    K.gemm(GEMM_NN, d_tr_bf, w["timeattn.proj.weight"], out_bf16=d_ot.view(M, C))
    d_qkv = divided_attention_bwd(K, s.qkv_t.view(B, N, 3 * C), s.o_t, s.lse_t, d_ot, H, T, Nf, "time").view(M, 3 * C)
    G.weight("timeattn.qkv.weight", d_qkv, s.ln3)
    G.bias("timeattn.qkv.bias", d_qkv)
    d_ln3 = _e(d_out, (M, C), BF16)
    K.gemm(GEMM_NN, d_qkv, w["timeattn.qkv.weight"], out_bf16=d_ln3)
    K.layernorm_bwd(d_ln3, s.x, p["norm3.weight"], s.mean3, s.rstd3, add=d_x, dx=d_x if need_dx else None,
                    dgamma=G.vec("norm3.weight", C), dbeta=G.vec("norm3.bias", C))
    return (d_x.view(B, N, C) if need_dx else None), dy, G.g


# ----------------------------------------------------------------------------------------------- SpaceTimeBlock, CLS row only
def video_block_cls_fwd(K, x, p, w, H, T, Nf, y=None, y_bias=None, eps=1e-5, save=True):
    """SpaceTimeBlock.forward for a block whose output is consumed at the CLS row only -- the LAST block of the
    EgoNCE pass (video_transformer.py:391 `norm(x)[:, 0]`) and of the ITM pass (model.py:275-278), SURVEY.md Q6.
    The time branch and LN1 still run on every token (the CLS query of the space attention attends the keys / values
    of all tokens); everything after the space attention -- query projection, out-projection, gated video->text
    cross-attention, MLP -- runs on the B CLS rows.  The CLS row of the result is identical to video_block_fwd's.
    x [B,N,C] f32 -> (out [B,C] f32, saved)."""
    B, N, C = x.shape
    M = B * N
    x2 = x.reshape(M, C)
    x3 = x2.view(B, N, C)
    s = types.SimpleNamespace(fused=y is not None, shape=(B, N, C), x=x2)
    K.mark("video_block")
    # ---- time attention branch on every token (identical to video_block_fwd)
    s.ln3, s.mean3, s.rstd3 = _e(x, (M, C), BF16), _e(x, (M,), F32), _e(x, (M,), F32)
    K.layernorm_fwd(x2, p["norm3.weight"], p["norm3.bias"], eps, y_bf16=s.ln3, mean=s.mean3, rstd=s.rstd3)
    s.qkv_t = _e(x, (M, 3 * C), BF16)
    K.gemm(GEMM_NT, s.ln3, w["timeattn.qkv.weight"], bias=p["timeattn.qkv.bias"], out_bf16=s.qkv_t)
    s.o_t, s.lse_t = divided_attention_fwd(K, s.qkv_t.view(B, N, 3 * C), H, T, Nf, "time")
    s.tr = _e(x, (M, C), F32)
    K.gemm(GEMM_NT, s.o_t.view(M, C), w["timeattn.proj.weight"], bias=p["timeattn.proj.bias"], residual=x2, out_f32=s.tr)
    # ---- space attention, CLS query only: keys / values of every token, the query of the B CLS rows
    s.ln1, s.mean1, s.rstd1 = _e(x, (M, C), BF16), _e(x, (M,), F32), _e(x, (M,), F32)
    K.layernorm_fwd(s.tr, p["norm1.weight"], p["norm1.bias"], eps, y_bf16=s.ln1, mean=s.mean1, rstd=s.rstd1)
    wqkv, bqkv = w["attn.qkv.weight"], p["attn.qkv.bias"]
    s.kv_s = _e(x, (M, 2 * C), BF16)
    K.gemm(GEMM_NT, s.ln1, wqkv[C:], bias=bqkv[C:], out_bf16=s.kv_s)
    ln1_cls = s.ln1.view(B, N, C)[:, 0]                       # [B, C] strided view (row stride N*C)
    s.q_cls = _e(x, (B, 1, C), BF16)
    K.gemm(GEMM_NT, ln1_cls, wqkv[:C], bias=bqkv[:C], out_bf16=s.q_cls.view(B, C))
    s.spec_cls = AttnSpec(H=H, G=1, Lq=1, Lk=N - 1, q_row0=0, k_row0=1, k_istride=1, has_cls_key=True, cls_row=0,
                          scale=(C // H) ** -0.5)
    kv3 = s.kv_s.view(B, N, 2 * C)
    s.o_cls, s.lse_cls = _e(x, (B, 1, C), BF16), _e(x, (B * H,), F32)
    K.attention_fwd(s.spec_cls, s.q_cls, kv3[:, :, :C], kv3[:, :, C:], s.o_cls, s.lse_cls)
    x_cls = x3[:, 0]                                          # [B, C] f32 strided view
    s.sr = _e(x, (B, C), F32)
    if y is None:
        K.gemm(GEMM_NT, s.o_cls.view(B, C), w["attn.proj.weight"], bias=p["attn.proj.bias"], residual=x_cls, out_f32=s.sr)
    else:
        S, Ct = y.shape[1], y.shape[2]
        s.a = _e(x, (B, C), BF16)
        xa = _e(x, (B, C), F32)
        K.gemm(GEMM_NT, s.o_cls.view(B, C), w["attn.proj.weight"], bias=p["attn.proj.bias"], residual=x_cls, out_f32=xa,
               out_pre=s.a)
        K.mark("xattn_i2t_fwd")
        s.lnc, s.meanc, s.rstdc = _e(x, (B, C), BF16), _e(x, (B,), F32), _e(x, (B,), F32)
        K.layernorm_fwd(s.a, p["attn.norm_i2t_i.weight"], p["attn.norm_i2t_i.bias"], eps, y_bf16=s.lnc, mean=s.meanc,
                        rstd=s.rstdc)
        s.q_c = _e(x, (B, 1, C), BF16)
        K.gemm(GEMM_NT, s.lnc, w["attn.qkv_i2t.weight"], bias=p["attn.qkv_i2t.bias"], out_bf16=s.q_c.view(B, C))
        s.y_bf = _e(x, (B * S, Ct), BF16)
        K.cast(y.reshape(B * S, Ct), s.y_bf)
        s.kv_t = _e(x, (B * S, 2 * C), BF16)
        K.gemm(GEMM_NT, s.y_bf, w["attn.qkv_text_i2t.weight"], bias=p["attn.qkv_text_i2t.bias"], out_bf16=s.kv_t)
        s.spec_c = AttnSpec(H=H, G=1, Lq=1, Lk=S, scale=(C // H) ** -0.5)
        s.y_bias = y_bias
        kvt3 = s.kv_t.view(B, S, 2 * C)
        s.o_c, s.lse_c = _e(x, (B, 1, C), BF16), _e(x, (B * H,), F32)
        K.attention_fwd(s.spec_c, s.q_c, kvt3[:, :, :C], kvt3[:, :, C:], s.o_c, s.lse_c, key_bias=y_bias)
        s.c = _e(x, (B, C), BF16)
        K.gemm(GEMM_NT, s.o_c.view(B, C), w["attn.proj_i2t.weight"], bias=p["attn.proj_i2t.bias"],
               scale_dev=p["attn.alpha_i2t"], residual=xa, out_f32=s.sr, out_pre=s.c)
    # ---- MLP on the CLS rows
    K.mark("video_block")
    s.ln2, s.mean2, s.rstd2 = _e(x, (B, C), BF16), _e(x, (B,), F32), _e(x, (B,), F32)
    K.layernorm_fwd(s.sr, p["norm2.weight"], p["norm2.bias"], eps, y_bf16=s.ln2, mean=s.mean2, rstd=s.rstd2)
    Hd = w["mlp.fc1.weight"].shape[0]
    s.h_pre, s.h_act = _e(x, (B, Hd), BF16), _e(x, (B, Hd), BF16)
    # h_pre holds gelu'(fc1 output): the backward then multiplies instead of re-deriving Phi and phi (ACT_GELU_DG)
    K.gemm(GEMM_NT, s.ln2, w["mlp.fc1.weight"], bias=p["mlp.fc1.bias"], act=ACT_GELU_DG, out_bf16=s.h_act, out_pre=s.h_pre)
    out = _e(x, (B, C), F32)
    K.gemm(GEMM_NT, s.h_act, w["mlp.fc2.weight"], bias=p["mlp.fc2.bias"], residual=s.sr, out_f32=out)
    return out, (s if save else None)


def video_block_cls_bwd(K, s, d_out, p, w, H, T, Nf, need_dx=True, sink=None):
    """Backward of video_block_cls_fwd.  d_out [B,C] f32 -> (dx [B,N,C] f32 or None, dy [B,S,Ct] f32 or None, grads)."""
    B, N, C = s.shape
    M = B * N
    G = Grads(K, d_out, sink)
    K.mark("video_block_bwd")
    d_out = d_out.reshape(B, C).contiguous()
    d_out_bf = _e(d_out, (B, C), BF16)
    K.cast(d_out, d_out_bf)
    # ---- MLP (B rows)
    G.weight("mlp.fc2.weight", d_out_bf, s.h_act)
    G.bias("mlp.fc2.bias", d_out_bf)
    d_hpre = _e(d_out, s.h_pre.shape, BF16)
    K.gemm(GEMM_NN, d_out_bf, w["mlp.fc2.weight"], aux=s.h_pre, act=ACT_MUL_AUX, out_bf16=d_hpre,
           colsum=G.vec("mlp.fc1.bias", s.h_pre.shape[1]))
    G.weight("mlp.fc1.weight", d_hpre, s.ln2)
    d_ln2 = _e(d_out, (B, C), BF16)
    K.gemm(GEMM_NN, d_hpre, w["mlp.fc1.weight"], out_bf16=d_ln2)
    d_sr, d_sr_bf = _e(d_out, (B, C), F32), _e(d_out, (B, C), BF16)
    cs_sr = _z(d_out, (C,)) if s.fused else G.vec("attn.proj.bias", C)
    K.layernorm_bwd(d_ln2, s.sr, p["norm2.weight"], s.mean2, s.rstd2, add=d_out, dx=d_sr, dx_bf16=d_sr_bf,
                    bf16_total=True, dgamma=G.vec("norm2.weight", C), dbeta=G.vec("norm2.bias", C), out_colsum=cs_sr)
    dy = None
    d_s_bf = d_sr_bf
    if s.fused:
        S, Ct = s.y_bf.shape[0] // B, s.y_bf.shape[1]
        alpha = p["attn.alpha_i2t"]
        K.mark("xattn_i2t_bwd")
        G.scalar_dot("attn.alpha_i2t", d_sr, s.c)
        G.weight("attn.proj_i2t.weight", d_sr_bf, s.o_c.view(B, C), scale_dev=alpha)
        G.bias_from("attn.proj_i2t.bias", cs_sr, scale_dev=alpha)
        d_oc = _e(d_out, (B, 1, C), BF16)
        K.gemm(GEMM_NN, d_sr_bf, w["attn.proj_i2t.weight"], scale_dev=alpha, out_bf16=d_oc.view(B, C))
        kvt3 = s.kv_t.view(B, S, 2 * C)
        dq_c = _e(d_out, (B, 1, C), BF16)
        dkv = _e(d_out, (B, S, 2 * C), BF16)
        K.attention_bwd(s.spec_c, s.q_c, kvt3[:, :, :C], kvt3[:, :, C:], s.o_c, s.lse_c, d_oc, dq_c, dkv[:, :, :C],
                        dkv[:, :, C:], _e(d_out, s.lse_c.shape, F32), key_bias=s.y_bias)
        dq2, dkv2 = dq_c.view(B, C), dkv.view(B * S, 2 * C)
        G.weight("attn.qkv_i2t.weight", dq2, s.lnc)
        G.bias("attn.qkv_i2t.bias", dq2)
        d_lnc = _e(d_out, (B, C), BF16)
        K.gemm(GEMM_NN, dq2, w["attn.qkv_i2t.weight"], out_bf16=d_lnc)
        G.weight("attn.qkv_text_i2t.weight", dkv2, s.y_bf)
        G.bias("attn.qkv_text_i2t.bias", dkv2)
        dy = _e(d_out, (B, S, Ct), F32)
        K.gemm(GEMM_NN, dkv2, w["attn.qkv_text_i2t.weight"], out_f32=dy.view(B * S, Ct))
        d_a_bf = _e(d_out, (B, C), BF16)
        K.layernorm_bwd(d_lnc, s.a, p["attn.norm_i2t_i.weight"], s.meanc, s.rstdc, add=d_sr, dx=None, dx_bf16=d_a_bf,
                        bf16_total=True, dgamma=G.vec("attn.norm_i2t_i.weight", C), dbeta=G.vec("attn.norm_i2t_i.bias", C),
                        out_colsum=G.vec("attn.proj.bias", C))
        d_s_bf = d_a_bf
        K.mark("video_block_bwd")
    # ---- space attention of the CLS query
    G.weight("attn.proj.weight", d_s_bf, s.o_cls.view(B, C))
    d_ocls = _e(d_out, (B, 1, C), BF16)
    K.gemm(GEMM_NN, d_s_bf, w["attn.proj.weight"], out_bf16=d_ocls.view(B, C))
    kv3 = s.kv_s.view(B, N, 2 * C)
    d_qkv_cls = _e(d_out, (B, 1, 3 * C), BF16)                # [dq | dk | dv] of the CLS rows
    d_kv = _e(d_out, (B, N, 2 * C), BF16)                     # every row written by the CLS-query backward
    dkv_cls = _z(d_out, (B * H * 128,))
    K.attention_bwd(s.spec_cls, s.q_cls, kv3[:, :, :C], kv3[:, :, C:], s.o_cls, s.lse_cls, d_ocls, d_qkv_cls[:, :, :C],
                    d_kv[:, :, :C], d_kv[:, :, C:], _e(d_out, s.lse_cls.shape, F32), dkv_cls=dkv_cls)
    K.attention_cls_finalize(dkv_cls, d_kv[:, :, :C], d_kv[:, :, C:], H, cls_row=0, accumulate=False)
    d_qkv_cls[:, 0, C:].copy_(d_kv[:, 0])                     # data movement only (B rows)
    wqkv = w["attn.qkv.weight"]
    ln1_cls = s.ln1.view(B, N, C)[:, 0]
    dq2, d_kv2 = d_qkv_cls.view(B, 3 * C)[:, :C], d_kv.view(M, 2 * C)
    gw = G.full("attn.qkv.weight", (3 * C, C))
    K.gemm(GEMM_TN, dq2, ln1_cls, out_f32=gw[:C], accumulate=True)
    K.gemm(GEMM_TN, d_kv2, s.ln1, out_f32=gw[C:], accumulate=True)
    gb = G.full("attn.qkv.bias", (3 * C,))
    K.colsum(dq2, gb[:C], accumulate=True)
    K.colsum(d_kv2, gb[C:], accumulate=True)
    d_ln1 = _e(d_out, (M, C), BF16)
    K.gemm(GEMM_NN, d_kv2, wqkv[C:], out_bf16=d_ln1)
    # the CLS rows also carry the query path: d_ln1[cls] = [dq | dk | dv][cls] @ W_qkv (overwrites the kv-only value)
    K.gemm(GEMM_NN, d_qkv_cls.view(B, 3 * C), wqkv, out_bf16=d_ln1.view(B, N, C)[:, 0])
    # d_tr = LN1'(d_ln1);  d_x = d_tr (+ d_sr on the CLS rows: x feeds sr there directly)
    d_x, d_tr_bf = _e(d_out, (M, C), F32), _e(d_out, (M, C), BF16)
    K.layernorm_bwd(d_ln1, s.tr, p["norm1.weight"], s.mean1, s.rstd1, add=None, dx=d_x, dx_bf16=d_tr_bf,
                    bf16_total=False, dgamma=G.vec("norm1.weight", C), dbeta=G.vec("norm1.bias", C),
                    out_colsum=G.vec("timeattn.proj.bias", C))
    d_x_cls = d_x.view(B, N, C)[:, 0]
    K.axpy_rows(d_x_cls, d_sr)
    # ---- time attention (identical to video_block_bwd)
    G.weight("timeattn.proj.weight", d_tr_bf, s.o_t.view(M, C))
    d_ot = _e(d_out, (B, N, C), BF16)
    K.gemm(GEMM_NN, d_tr_bf, w["timeattn.proj.weight"], out_bf16=d_ot.view(M, C))
    d_qkv = divided_attention_bwd(K, s.qkv_t.view(B, N, 3 * C), s.o_t, s.lse_t, d_ot, H, T, Nf, "time").view(M, 3 * C)
    G.weight("timeattn.qkv.weight", d_qkv, s.ln3)
    G.bias("timeattn.qkv.bias", d_qkv)
    d_ln3 = _e(d_out, (M, C), BF16)
    K.gemm(GEMM_NN, d_qkv, w["timeattn.qkv.weight"], out_bf16=d_ln3)
    K.layernorm_bwd(d_ln3, s.x, p["norm3.weight"], s.mean3, s.rstd3, add=d_x, dx=d_x if need_dx else None,
                    dgamma=G.vec("norm3.weight", C), dbeta=G.vec("norm3.bias", C))
    return (d_x.view(B, N, C) if need_dx else None), dy, G.g


# ----------------------------------------------------------------------------------------------- RobertaLayer
TEXT_LAYER_PARAMS = ["attention.self.query.weight", "attention.self.query.bias", "attention.self.key.weight",
                     "attention.self.key.bias", "attention.self.value.weight", "attention.self.value.bias",
                     "attention.output.dense.weight", "attention.output.dense.bias", "attention.output.LayerNorm.weight",
                     "attention.output.LayerNorm.bias", "intermediate.dense.weight", "intermediate.dense.bias",
                     "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias"]
TEXT_FUSE_PARAMS = ["alpha_t2i", "crossattention_t2i.self.query.weight", "crossattention_t2i.self.query.bias",
                    "crossattention_t2i.self.key.weight", "crossattention_t2i.self.key.bias",
                    "crossattention_t2i.self.value.weight", "crossattention_t2i.self.value.bias",
                    "crossattention_t2i.output.dense.weight", "crossattention_t2i.output.dense.bias"]


def text_layer_fwd(K, h, key_bias, p, w, H, video=None, eps=1e-5, save=True, drop=None, layer=0):
    """RobertaLayer.forward (roberta.py:444-505), `last_norm=True`.  `drop` (train mode, see drop_site; `layer` = encoder
    layer index): dropout on the attention probabilities (:313) and after the three dense outputs (:342, :422), the
    cross-attention included; None = eval mode.
    h [B,S,C] f32; key_bias [B,S] f32 additive mask; video [B,N,Cv] f32 = un-normalised video stream entering
    video block i (roberta.py:470-486: K,V = Linear(video), no mask).  `w['qkv']` = cat(query,key,value) weights,
    `p['qkv.bias']` the concatenated bias; likewise `w['cross.kv']`, `p['cross.kv.bias']`.
    Returns (out [B,S,C] f32, saved)."""
    B, S, C = h.shape
    M = B * S
    h2 = h.reshape(M, C)
    d = C // H
    s = types.SimpleNamespace(fused=video is not None, shape=(B, S, C), key_bias=key_bias, drop=drop, layer=layer)
    site = (lambda kind: drop_site(drop, layer + 1, kind)) if drop is not None else None
    K.mark("text_layer")
    s.h_bf = _e(h, (M, C), BF16)
    K.cast(h2.contiguous(), s.h_bf)
    s.qkv = _e(h, (M, 3 * C), BF16)
    K.gemm(GEMM_NT, s.h_bf, w["qkv"], bias=p["qkv.bias"], out_bf16=s.qkv)
    s.spec = AttnSpec(H=H, G=1, Lq=S, Lk=S, scale=1.0 / math.sqrt(d))
    qkv3 = s.qkv.view(B, S, 3 * C)
    s.o, s.lse = _e(h, (B, S, C), BF16), _e(h, (B * H * S,), F32)
    if drop is None:
        K.attention_fwd(s.spec, qkv3[:, :, :C], qkv3[:, :, C:2 * C], qkv3[:, :, 2 * C:], s.o, s.lse, key_bias=key_bias)
    else:
        K.text_attention_fwd(qkv3[:, :, :C], qkv3[:, :, C:2 * C], qkv3[:, :, 2 * C:], key_bias, s.spec.scale, drop.p_attn, drop.seed,
                             site(DROP_SELF_PROBS), H, s.o, s.lse)
    # so = attention.output.dense(o) (no residual / LN inside RobertaSelfOutput: roberta.py:331-343)
    # sh = so + h  (fp32), so_bf = bf16(so) feeds the cross-attention query
    sh = _e(h, (M, C), F32)
    s.so_bf = _e(h, (M, C), BF16) if s.fused else None
    if drop is None:
        K.gemm(GEMM_NT, s.o.view(M, C), w["attention.output.dense.weight"], bias=p["attention.output.dense.bias"],
               residual=h2, out_f32=sh, out_pre=s.so_bf)
    else:
        so_f = _e(h, (M, C), F32)
        K.gemm(GEMM_NT, s.o.view(M, C), w["attention.output.dense.weight"], bias=p["attention.output.dense.bias"], out_f32=so_f)
        K.dropout_add(so_f, h2.contiguous(), drop.p, drop.seed, site(DROP_SELF_OUT), out_f32=sh, out_bf16=s.so_bf)
    if s.fused:
        Bv, N, Cv = video.shape
        s.alg = float(Bv) * (4.0 * S * C * C + 4.0 * N * Cv * C + 4.0 * S * N * C)   # xattn_t2i per launch
        K.mark("xattn_t2i_fwd", alg=s.alg)
        s.vshape = (Bv, N, Cv)
        s.x_bf = _e(h, (Bv * N, Cv), BF16)
        K.cast(video.reshape(Bv * N, Cv).contiguous(), s.x_bf)
        s.qx = _e(h, (M, C), BF16)
        K.gemm(GEMM_NT, s.so_bf, w["crossattention_t2i.self.query.weight"], bias=p["crossattention_t2i.self.query.bias"],
               out_bf16=s.qx)
        s.ox = _e(h, (B, S, C), BF16)
        s.reassoc = reassoc_enabled() and Bv == B and XR.t2i_supported(S, N, C, H)
        if s.reassoc:
            # re-associated around the S text queries (xattn_reassoc.py): K = V = the video stream itself, no K/V projection
            s.xr = XR.t2i_fwd(K, s.qx, s.x_bf, w["cross.kv"][:C], w["cross.kv"][C:], p["cross.kv.bias"][C:], s.ox.view(M, C),
                              B, N, H, p_drop=0.0 if drop is None else drop.p_attn, seed_dev=None if drop is None else drop.seed,
                              site=0 if drop is None else site(DROP_CROSS_PROBS))
        elif drop is not None:
            raise NotImplementedError("train-mode dropout of the text->video cross-attention needs the re-associated path "
                                      "(text length <= 128, <= 65536 video tokens per clip)")
        else:
            s.kv = _e(h, (Bv * N, 2 * C), BF16)
            K.gemm(GEMM_NT, s.x_bf, w["cross.kv"], bias=p["cross.kv.bias"], out_bf16=s.kv)
            s.spec_x = AttnSpec(H=H, G=1, Lq=S, Lk=N, scale=1.0 / math.sqrt(d))
            kv3 = s.kv.view(Bv, N, 2 * C)
            s.lse_x = _e(h, (B * H * S,), F32)
            K.attention_fwd(s.spec_x, s.qx.view(B, S, C), kv3[:, :, :C], kv3[:, :, C:], s.ox, s.lse_x)
        # attn_out + h = alpha * c + so + h
        s.c = _e(h, (M, C), BF16)
        sh2 = _e(h, (M, C), F32)
        if drop is None:
            K.gemm(GEMM_NT, s.ox.view(M, C), w["crossattention_t2i.output.dense.weight"],
                   bias=p["crossattention_t2i.output.dense.bias"], scale_dev=p["alpha_t2i"], residual=sh, out_f32=sh2,
                   out_pre=s.c)
        else:   # c = dropout(dense(ox));  sh2 = sh + alpha * c
            c_f = _e(h, (M, C), F32)
            K.gemm(GEMM_NT, s.ox.view(M, C), w["crossattention_t2i.output.dense.weight"],
                   bias=p["crossattention_t2i.output.dense.bias"], out_f32=c_f)
            K.dropout_add(c_f, sh, drop.p, drop.seed, site(DROP_CROSS_OUT), out_f32=sh2, out_bf16=s.c, scale_dev=p["alpha_t2i"])
        sh = sh2
        K.mark("text_layer")
    s.sh = sh
    s.a, s.a_bf = _e(h, (M, C), F32), _e(h, (M, C), BF16)
    s.mean_a, s.rstd_a = _e(h, (M,), F32), _e(h, (M,), F32)
    K.layernorm_fwd(sh, p["attention.output.LayerNorm.weight"], p["attention.output.LayerNorm.bias"], eps, y_bf16=s.a_bf,
                    y_f32=s.a, mean=s.mean_a, rstd=s.rstd_a)
    Hd = w["intermediate.dense.weight"].shape[0]
    s.f_pre, s.f_act = _e(h, (M, Hd), BF16), _e(h, (M, Hd), BF16)
    K.gemm(GEMM_NT, s.a_bf, w["intermediate.dense.weight"], bias=p["intermediate.dense.bias"], act=ACT_GELU,
           out_bf16=s.f_act, out_pre=s.f_pre)
    s.fa = _e(h, (M, C), F32)
    if drop is None:
        K.gemm(GEMM_NT, s.f_act, w["output.dense.weight"], bias=p["output.dense.bias"], residual=s.a, out_f32=s.fa)
    else:
        fo_f = _e(h, (M, C), F32)
        K.gemm(GEMM_NT, s.f_act, w["output.dense.weight"], bias=p["output.dense.bias"], out_f32=fo_f)
        K.dropout_add(fo_f, s.a, drop.p, drop.seed, site(DROP_FFN_OUT), out_f32=s.fa)
    out = _e(h, (M, C), F32)
    s.mean_o, s.rstd_o = _e(h, (M,), F32), _e(h, (M,), F32)
    K.layernorm_fwd(s.fa, p["output.LayerNorm.weight"], p["output.LayerNorm.bias"], eps, y_f32=out, mean=s.mean_o,
                    rstd=s.rstd_o)
    return out.view(B, S, C), (s if save else None)


def text_layer_bwd(K, s, d_out, p, w, H, need_dh=True, sink=None):
    """Backward of text_layer_fwd -> (dh [B,S,C] f32, dvideo [B,N,Cv] f32 or None, grads dict with the
    concatenated 'qkv' / 'cross.kv' gradients)."""
    B, S, C = s.shape
    M = B * S
    G = Grads(K, d_out, sink)
    drop = s.drop
    site = (lambda kind: drop_site(drop, s.layer + 1, kind)) if drop is not None else None
    K.mark("text_layer_bwd")
    d_out = d_out.reshape(M, C).contiguous()
    # out = LN_o(fa)
    d_fa, d_fa_bf = _e(d_out, (M, C), F32), _e(d_out, (M, C), BF16)
    K.layernorm_bwd(d_out, s.fa, p["output.LayerNorm.weight"], s.mean_o, s.rstd_o, dx=d_fa, dx_bf16=d_fa_bf,
                    dgamma=G.vec("output.LayerNorm.weight", C), dbeta=G.vec("output.LayerNorm.bias", C))
    if drop is not None:   # the dense branch sees the masked gradient, the residual branch (d_fa) the plain one
        K.dropout_bwd(d_fa, drop.p, drop.seed, site(DROP_FFN_OUT), out_bf16=d_fa_bf)
    # fa = a + dense2(gelu(dense1(a)))
    G.weight("output.dense.weight", d_fa_bf, s.f_act)
    G.bias("output.dense.bias", d_fa_bf)
    d_fpre = _e(d_out, s.f_pre.shape, BF16)
    K.gemm(GEMM_NN, d_fa_bf, w["output.dense.weight"], aux=s.f_pre, act=ACT_GELU_BWD, out_bf16=d_fpre)
    G.weight("intermediate.dense.weight", d_fpre, s.a_bf)
    G.bias("intermediate.dense.bias", d_fpre)
    d_a = _e(d_out, (M, C), F32)
    K.gemm(GEMM_NN, d_fpre, w["intermediate.dense.weight"], residual=d_fa, out_f32=d_a)   # d_a = d_fa + dense1'(...)
    # a = LN_a(sh)
    d_sh, d_sh_bf = _e(d_out, (M, C), F32), _e(d_out, (M, C), BF16)
    K.layernorm_bwd(d_a, s.sh, p["attention.output.LayerNorm.weight"], s.mean_a, s.rstd_a, dx=d_sh, dx_bf16=d_sh_bf,
                    dgamma=G.vec("attention.output.LayerNorm.weight", C), dbeta=G.vec("attention.output.LayerNorm.bias", C))
    # sh = h + so + alpha * c
    dvideo = None
    d_so_bf = d_sh_bf
    if s.fused:
        Bv, N, Cv = s.vshape
        alpha = p["alpha_t2i"]
        K.mark("xattn_t2i_bwd", alg=2.0 * s.alg)
        G.scalar_dot("alpha_t2i", d_sh, s.c)
        d_c_bf = d_sh_bf
        if drop is not None:
            d_c_bf = _e(d_out, (M, C), BF16)
            K.dropout_bwd(d_sh, drop.p, drop.seed, site(DROP_CROSS_OUT), out_bf16=d_c_bf)
        G.weight("crossattention_t2i.output.dense.weight", d_c_bf, s.ox.view(M, C), scale_dev=alpha)
        G.bias("crossattention_t2i.output.dense.bias", d_c_bf, scale_dev=alpha)
        d_ox = _e(d_out, (B, S, C), BF16)
        K.gemm(GEMM_NN, d_c_bf, w["crossattention_t2i.output.dense.weight"], scale_dev=alpha, out_bf16=d_ox.view(M, C))
        dvideo = _e(d_out, (Bv, N, Cv), F32)
        if s.reassoc:
            gw, gb = G.full("cross.kv", (2 * C, Cv)), G.full("cross.kv.bias", (2 * C,))
            dqx2 = XR.t2i_bwd(K, s.xr, d_ox.view(M, C), w["cross.kv"][:C], w["cross.kv"][C:], p["cross.kv.bias"][C:], gw[:C],
                              gw[C:], gb[C:], dvideo.view(Bv * N, Cv))
        else:
            kv3 = s.kv.view(Bv, N, 2 * C)
            dqx = _e(d_out, (B, S, C), BF16)
            dkv = _e(d_out, (Bv, N, 2 * C), BF16)
            K.attention_bwd(s.spec_x, s.qx.view(B, S, C), kv3[:, :, :C], kv3[:, :, C:], s.ox, s.lse_x, d_ox, dqx, dkv[:, :, :C],
                            dkv[:, :, C:], _e(d_out, s.lse_x.shape, F32))
            dqx2, dkv2 = dqx.view(M, C), dkv.view(Bv * N, 2 * C)
            G.weight("cross.kv", dkv2, s.x_bf)
            G.bias("cross.kv.bias", dkv2)
            K.gemm(GEMM_NN, dkv2, w["cross.kv"], out_f32=dvideo.view(Bv * N, Cv))
        G.weight("crossattention_t2i.self.query.weight", dqx2, s.so_bf)
        G.bias("crossattention_t2i.self.query.bias", dqx2)
        # d_so = d_sh + Wq_x'(dqx)
        d_so_bf = _e(d_out, (M, C), BF16)
        K.gemm(GEMM_NN, dqx2, w["crossattention_t2i.self.query.weight"], residual=d_sh, out_bf16=d_so_bf)
        K.mark("text_layer_bwd")
    if drop is not None:   # so entered sh (and the cross-attention query) through the dropout of roberta.py:342
        d_so_m = _e(d_out, (M, C), BF16)
        K.dropout_bwd(d_so_bf, drop.p, drop.seed, site(DROP_SELF_OUT), out_bf16=d_so_m)
        d_so_bf = d_so_m
    G.weight("attention.output.dense.weight", d_so_bf, s.o.view(M, C))
    G.bias("attention.output.dense.bias", d_so_bf)
    d_o = _e(d_out, (B, S, C), BF16)
    K.gemm(GEMM_NN, d_so_bf, w["attention.output.dense.weight"], out_bf16=d_o.view(M, C))
    qkv3 = s.qkv.view(B, S, 3 * C)
    d_qkv = _e(d_out, (B, S, 3 * C), BF16)
    if drop is None:
        K.attention_bwd(s.spec, qkv3[:, :, :C], qkv3[:, :, C:2 * C], qkv3[:, :, 2 * C:], s.o, s.lse, d_o, d_qkv[:, :, :C],
                        d_qkv[:, :, C:2 * C], d_qkv[:, :, 2 * C:], _e(d_out, s.lse.shape, F32), key_bias=s.key_bias)
    else:
        K.text_attention_bwd(qkv3[:, :, :C], qkv3[:, :, C:2 * C], qkv3[:, :, 2 * C:], s.key_bias, s.spec.scale, drop.p_attn, drop.seed,
                             site(DROP_SELF_PROBS), H, s.lse, d_o, d_qkv[:, :, :C], d_qkv[:, :, C:2 * C], d_qkv[:, :, 2 * C:])
    d_qkv2 = d_qkv.view(M, 3 * C)
    G.weight("qkv", d_qkv2, s.h_bf)
    G.bias("qkv.bias", d_qkv2)
    dh = None
    if need_dh:
        dh = _e(d_out, (M, C), F32)
        K.gemm(GEMM_NN, d_qkv2, w["qkv"], residual=d_sh, out_f32=dh)   # dh = d_sh (residual path) + qkv'(...)
        dh = dh.view(B, S, C)
    return dh, dvideo, G.g


# ----------------------------------------------------------------------------------------------- embeddings
# NormalizeVideo defaults of the reference's transforms (data_loader/transforms.py:39-49), used when frames arrive as uint8
VIDEO_NORM_MEAN, VIDEO_NORM_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def video_tokens_fwd(K, video, p, w, cls_token, patch, save=True, norm=None):
    """VideoPatchEmbed + token assembly (video_transformer.py:78-83, 354-372; model.py:211-232).
    video [B,T,3,H,W] f32 -> tokens [B, 1+T*Nf, C] f32.  uint8 video: the loader's `/ 255` and
    NormalizeVideo (base_dataset.py:248, transforms.py:49; norm = (mean, std)) run inside the im2col kernel."""
    B, T, Cin, Hh, Ww = video.shape
    Nf = (Hh // patch) * (Ww // patch)
    C = w["patch_embed.proj.weight"].shape[0]
    K.mark("embed")
    Kc = Cin * patch * patch
    Kp = (Kc + 7) // 8 * 8          # TMA row strides are multiples of 8 elements: patch 14 (TimeSformer-L/14) pads 588 -> 592
    cols = _e(video, (B * T * Nf, Kp), BF16)
    wpe = w["patch_embed.proj.weight"]
    if Kp != Kc:
        wpad = _z(video, (C, Kp), wpe.dtype)
        wpad[:, :Kc].copy_(wpe)      # data movement only: the zero-padded operand copy of the convolution weight
        wpe = wpad
    if video.dtype == torch.uint8:
        if Kp != Kc:
            raise NotImplementedError("uint8 frames need a patch size that is a multiple of 8")
        mean, std = norm if norm is not None else (VIDEO_NORM_MEAN, VIDEO_NORM_STD)
        K.patchify_u8(video.reshape(B * T, Cin, Hh, Ww).contiguous(), patch, cols, mean, std)
    else:
        K.patchify(video.reshape(B * T, Cin, Hh, Ww).contiguous(), patch, cols)
    pe = _e(video, (B * T * Nf, C), F32)
    K.gemm(GEMM_NT, cols, wpe, bias=p["patch_embed.proj.bias"], out_f32=pe)
    tokens = _e(video, (B, 1 + T * Nf, C), F32)
    K.assemble_tokens(pe, cls_token.reshape(-1).contiguous(), p["pos_embed"].reshape(1 + Nf, C),
                      p["temporal_embed"].reshape(-1, C)[:T].contiguous(), B, T, Nf, tokens)
    s = types.SimpleNamespace(cols=cols, B=B, T=T, Nf=Nf, C=C, Kc=Kc, t_max=p["temporal_embed"].shape[1]) if save else None
    return tokens, s


def video_tokens_bwd(K, s, d_tokens, sink=None):
    """-> grads dict: patch_embed.proj.{weight [C, 3*p*p] flattened, bias}, pos_embed, temporal_embed, cls_token."""
    B, T, Nf, C = s.B, s.T, s.Nf, s.C
    K.mark("embed_bwd")
    G = Grads(K, d_tokens, sink)
    d_patch = _e(d_tokens, (B * T * Nf, C), BF16)
    K.assemble_tokens_bwd(d_tokens.contiguous(), B, T, Nf, d_patch, G.full("cls_token", (C,)), G.full("pos_embed", (1 + Nf, C)),
                          G.full("temporal_embed", (s.t_max, C)))
    if s.cols.shape[1] == s.Kc:
        G.weight("patch_embed.proj.weight", d_patch, s.cols)
    else:   # padded im2col depth: the gradient of the real weight is the leading Kc columns
        dwp = _e(d_tokens, (C, s.cols.shape[1]), F32)
        K.gemm(GEMM_TN, d_patch, s.cols, out_f32=dwp)
        K.axpy_rows(G.full("patch_embed.proj.weight", (C, s.Kc)), dwp[:, :s.Kc])
    G.bias("patch_embed.proj.bias", d_patch)
    return G.g


# dropout sites of one text-tower invocation (roberta.py:203,313,342,422): site id = drop.base + layer slot * 16 + kind
DROP_EMB, DROP_SELF_PROBS, DROP_SELF_OUT, DROP_CROSS_PROBS, DROP_CROSS_OUT, DROP_FFN_OUT = range(6)


def drop_site(drop, layer_slot, kind):
    """drop: namespace(p (dense outputs / embeddings), p_attn (attention probabilities), seed (device int64 [1]), base (int:
    unique per text-tower invocation)) -- rng.drop_cfg; layer_slot 0 = embeddings, i + 1 = encoder layer i"""
    return drop.base + layer_slot * 16 + kind


def text_embeddings_fwd(K, ids, p, eps=1e-5, pad_id=1, save=True, drop=None):
    """RobertaEmbeddings.forward (roberta.py:174-204); `drop` (train mode): the dropout after the LayerNorm (:203)."""
    B, S = ids.shape
    C = p["word_embeddings.weight"].shape[1]
    ids = ids.contiguous()
    K.mark("embed")
    pre = _e(p["LayerNorm.weight"], (B, S, C), F32)
    K.text_embed(ids, p["word_embeddings.weight"], p["position_embeddings.weight"],
                 p["token_type_embeddings.weight"].reshape(-1)[:C].contiguous(), pre, pad_id)
    out = _e(pre, (B, S, C), F32)
    mean, rstd = _e(pre, (B * S,), F32), _e(pre, (B * S,), F32)
    K.layernorm_fwd(pre, p["LayerNorm.weight"], p["LayerNorm.bias"], eps, y_f32=out, mean=mean, rstd=rstd)
    if drop is not None:
        dropped = _e(pre, (B, S, C), F32)
        K.dropout_add(out, None, drop.p, drop.seed, drop_site(drop, 0, DROP_EMB), out_f32=dropped)
        out = dropped
    s = types.SimpleNamespace(ids=ids, pre=pre, mean=mean, rstd=rstd, pad_id=pad_id, drop=drop) if save else None
    return out, s


def text_embeddings_bwd(K, s, d_out, p, sink=None):
    C = s.pre.shape[-1]
    K.mark("embed_bwd")
    G = Grads(K, d_out, sink)
    d_pre = _e(d_out, s.pre.shape, F32)
    d_out = d_out.contiguous()
    if s.drop is not None:
        d_ln = _e(d_out, d_out.shape, F32)
        K.dropout_bwd(d_out, s.drop.p, s.drop.seed, drop_site(s.drop, 0, DROP_EMB), out_f32=d_ln)
        d_out = d_ln
    K.layernorm_bwd(d_out, s.pre, p["LayerNorm.weight"], s.mean, s.rstd, dx=d_pre,
                    dgamma=G.vec("LayerNorm.weight", C), dbeta=G.vec("LayerNorm.bias", C))
    K.text_embed_bwd(d_pre, s.ids, G.full("word_embeddings.weight", tuple(p["word_embeddings.weight"].shape)),
                     G.full("position_embeddings.weight", tuple(p["position_embeddings.weight"].shape)),
                     G.full("token_type_embeddings.weight", tuple(p["token_type_embeddings.weight"].shape)).view(-1)[:C],
                     s.pad_id)
    return G.g


# ----------------------------------------------------------------------------------------------- small heads
def layernorm_rows_fwd(K, x, gamma, beta, eps, save=True):
    """LayerNorm on [rows, C] f32 -> f32 (final norms on the CLS rows only: video_transformer.py:391, model.py:275)."""
    rows, C = x.shape
    x = x.contiguous()
    y, mean, rstd = _e(x, (rows, C), F32), _e(x, (rows,), F32), _e(x, (rows,), F32)
    K.layernorm_fwd(x, gamma, beta, eps, y_f32=y, mean=mean, rstd=rstd)
    return y, (types.SimpleNamespace(x=x, mean=mean, rstd=rstd) if save else None)


def layernorm_rows_bwd(K, s, dy, gamma, sink=None):
    """sink: {'weight': t, 'bias': t} to accumulate into; returns (dx, dgamma or None, dbeta or None)."""
    C = s.x.shape[1]
    G = Grads(K, dy, sink)
    dx = _e(dy, s.x.shape, F32)
    K.layernorm_bwd(dy.contiguous(), s.x, gamma, s.mean, s.rstd, dx=dx, dgamma=G.vec("weight", C), dbeta=G.vec("bias", C))
    return dx, G.g.get("weight"), G.g.get("bias")


def relu_rows_fwd(K, x):
    """Leading nn.ReLU of the fine-tuning text projection (model_epic_charades.py:118): f32 [M, C] -> bf16 relu(x), the
    operand format of the Linear that follows.  relu(x) = x * relu'(x): one pass of the activation-gradient kernel."""
    xb = _e(x, x.shape, BF16)
    K.cast(x.contiguous(), xb)
    y = _e(x, x.shape, BF16)
    K.act_grad(x.contiguous(), xb, ACT_RELU_BWD, y)
    return y


def relu_rows_bwd(K, y, dy):
    """dx (f32) = dy * [y > 0]"""
    db = _e(dy, dy.shape, BF16)
    K.act_grad(dy.contiguous(), y, ACT_RELU_BWD, db)
    dx = _e(dy, dy.shape, F32)
    K.cast(db, dx)
    return dx


def mlp_chain_fwd(K, x, layers, save=True):
    """A chain of Linear(+activation) layers on [M, Kd] f32 input (projection heads model.py:105-115, poolers
    heads.py:15-27, cross_modal transforms model.py:145-148, ITM fc heads.py:30-35).
    layers: list of (w_bf16 [N,Kd], bias f32 or None, act in {ACT_NONE, ACT_RELU, ACT_TANH, ACT_GELU}).
    Returns (out f32 [M, N_last], saved)."""
    M = x.shape[0]
    K.mark("heads")
    if x.dtype == BF16:
        cur = x.contiguous()
    else:
        cur = _e(x, x.shape, BF16)
        K.cast(x.contiguous(), cur)
    saved = []
    out = None
    for i, (wt, b, act) in enumerate(layers):
        last = i == len(layers) - 1
        N = wt.shape[0]
        nxt = _e(x, (M, N), BF16)
        pre = _e(x, (M, N), BF16) if act == ACT_GELU else None
        out = _e(x, (M, N), F32) if last else None
        K.gemm(GEMM_NT, cur, wt, bias=b, act=act, out_bf16=nxt, out_f32=out, out_pre=pre)
        saved.append((cur, nxt, pre))
        cur = nxt
    return out, (saved if save else None)


_ACT_BWD = {ACT_NONE: ACT_NONE, ACT_GELU: ACT_GELU_BWD, ACT_RELU: ACT_RELU_BWD, ACT_TANH: ACT_TANH_BWD}


def mlp_chain_bwd(K, saved, d_out, layers, scale_dev=None, need_dx=True, sinks=None):
    """-> (dx f32 [M, Kd] or None, [(dW, db)] per layer; None where the gradient went to `sinks[i]` = (dW sink, db sink)).
    `scale_dev` scales d_out (device scalar).
    The gradient of every activation is folded into the epilogue of the GEMM that produces it; only a
    trailing activation of the LAST layer needs the stand-alone act_grad kernel."""
    grads = [None] * len(layers)
    K.mark("heads_bwd")
    last = len(layers) - 1
    act = layers[last][2]
    aux = None if act == ACT_NONE else (saved[last][2] if act == ACT_GELU else saved[last][1])
    d_cur = _e(d_out, d_out.shape, BF16)          # gradient w.r.t. the PRE-activation output of layer i
    K.act_grad(d_out.contiguous(), aux, _ACT_BWD[act], d_cur, scale_dev=scale_dev)
    dx = None
    for i in range(last, -1, -1):
        wt, b, _ = layers[i]
        x_in = saved[i][0]
        sw, sb = sinks[i] if sinks is not None else (None, None)
        G = Grads(K, d_out, {k: v for k, v in (("w", sw), ("b", sb)) if v is not None})
        G.weight("w", d_cur, x_in)
        if b is not None:
            G.bias("b", d_cur)
        grads[i] = (G.g.get("w"), G.g.get("b"))
        if i > 0:
            pact = layers[i - 1][2]
            paux = None if pact == ACT_NONE else (saved[i - 1][2] if pact == ACT_GELU else saved[i - 1][1])
            d_prev = _e(d_out, (d_cur.shape[0], wt.shape[1]), BF16)
            K.gemm(GEMM_NN, d_cur, wt, aux=paux, act=_ACT_BWD[pact], out_bf16=d_prev)
            d_cur = d_prev
        elif need_dx:
            dx = _e(d_out, (d_cur.shape[0], wt.shape[1]), F32)
            K.gemm(GEMM_NN, d_cur, wt, out_f32=dx)
    return dx, grads


# ----------------------------------------------------------------------------------------------- MLM head + CE
def mlm_head_fwd(K, h, labels, p, w, save=True):
    """cross_modal_text_transform -> MLMHead (heads.py:38-50: dense, erf-GELU, LayerNorm 1e-12, decoder + bias)
    -> mean cross-entropy over labels != -100 (model.py:408-418).  h [B,S,C] f32.
    Returns (logits f32 view [B*S, V] with padded row stride, loss_sum [1], count [1], saved)."""
    B, S, C = h.shape
    M = B * S
    V = w["mlm_score.decoder.weight"].shape[0]
    ldv = (V + 7) // 8 * 8
    layers = [(w["cross_modal_text_transform.weight"], p["cross_modal_text_transform.bias"], ACT_NONE),
              (w["mlm_score.transform.dense.weight"], p["mlm_score.transform.dense.bias"], ACT_GELU)]
    t, chain = mlp_chain_fwd(K, h.reshape(M, C), layers)
    K.mark("mlm_head")
    tn = _e(h, (M, C), BF16)
    mean, rstd = _e(h, (M,), F32), _e(h, (M,), F32)
    K.layernorm_fwd(t, p["mlm_score.transform.LayerNorm.weight"], p["mlm_score.transform.LayerNorm.bias"], 1e-12, y_bf16=tn,
                    mean=mean, rstd=rstd)
    logits = _e(h, (M, ldv), F32)[:, :V]
    K.gemm(GEMM_NT, tn, w["mlm_score.decoder.weight"], bias=p["mlm_score.bias"], out_f32=logits)
    loss_sum, count = _z(h, (1,)), _z(h, (1,))
    dlogits = None
    if save:
        dlogits = _e(h, (M, ldv), BF16)
        if ldv != V:
            dlogits[:, V:].zero_()
        dlogits = dlogits[:, :V]
    K.softmax_xent(logits, labels.reshape(-1).contiguous(), V, loss_sum, count, dlogits=dlogits)
    s = types.SimpleNamespace(layers=layers, chain=chain, t=t, tn=tn, mean=mean, rstd=rstd, dlogits=dlogits,
                              shape=(B, S, C)) if save else None
    return logits, loss_sum, count, s


def mlm_head_bwd(K, s, scale_dev, p, w, sink=None):
    """scale_dev: device scalar = d(loss)/d(loss_sum) (grad_out / global count).  -> (dh [B,S,C] f32, grads dict)."""
    B, S, C = s.shape
    M = B * S
    sink = sink or {}
    G = Grads(K, s.t, sink)
    K.mark("mlm_head_bwd")
    G.weight("mlm_score.decoder.weight", s.dlogits, s.tn, scale_dev=scale_dev)
    G.bias("mlm_score.bias", s.dlogits, scale_dev=scale_dev)
    d_tn = _e(s.t, (M, C), BF16)
    K.gemm(GEMM_NN, s.dlogits, w["mlm_score.decoder.weight"], scale_dev=scale_dev, out_bf16=d_tn)
    d_t = _e(s.t, (M, C), F32)
    K.layernorm_bwd(d_tn, s.t, p["mlm_score.transform.LayerNorm.weight"], s.mean, s.rstd, dx=d_t,
                    dgamma=G.vec("mlm_score.transform.LayerNorm.weight", C), dbeta=G.vec("mlm_score.transform.LayerNorm.bias", C))
    names = [("cross_modal_text_transform.weight", "cross_modal_text_transform.bias"),
             ("mlm_score.transform.dense.weight", "mlm_score.transform.dense.bias")]
    dh, grads = mlp_chain_bwd(K, s.chain, d_t, s.layers, sinks=[(sink.get(a), sink.get(b)) for a, b in names])
    g = G.g
    for (a, b), (dW, db) in zip(names, grads):
        if dW is not None:
            g[a] = dW
        if db is not None:
            g[b] = db
    return dh.view(B, S, C), g


def xent_rows_fwd(K, logits, labels, save=True):
    """Plain CE on small [rows, V] f32 logits (ITM: model.py:478).  -> (loss_sum, count, dlogits bf16)."""
    rows, V = logits.shape
    loss_sum, count = _z(logits, (1,)), _z(logits, (1,))
    dl = _e(logits, (rows, V), BF16) if save else None
    K.softmax_xent(logits.contiguous(), labels.contiguous(), V, loss_sum, count, dlogits=dl)
    return loss_sum, count, dl
