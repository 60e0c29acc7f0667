"""ROUND-2 PATH, NOT WIRED INTO THE MODEL YET: the core of the gated video->text cross-attention
(video_transformer.py:155-185: q = qkv_i2t(LN(x)), softmax(q k^T + mask) v, proj_i2t) re-associated around the S text keys
(oracle/egovlp_oracle.py::cross_attention_i2t_reassociated):

    scores[b,n,h,s] = LN(x)[b,n,:] . Mt[h,b,s,:] + c0[b,s,h] + mask[b,s]     Mt = d^-1/2 k_h Wq_h        [H, B*S, C]
    c[b,n,:]        = sum_{h,s} P[b,n,h,s] U[h,b,s,:]                        U  = v_h Wp[:,h]^T + bp/H   [H, B*S, C]

i.e. the [M, C] x [C, C] query and output projections (M = B*N = 25 096 rows at the BASELINE shapes) and the 32-key
attention launch become two [M, C] x [C, H*S] products with a 32-wide group softmax between them: half the FLOPs.
(The text->video direction, re-associated the same way, follows below.)
What is on the M rows runs in FOUR fused device kernels that do not exist in libegovlp_b200.so yet (batched-per-clip
variants of gemm_tc_kernel with softmax / softmax-backward epilogues):

    K.xattn_scores_softmax(ln, Mt, c0, mask, P)     K.xattn_weighted_sum(P_or_dS, U_or_Mt, out)
    K.xattn_dscores(dc, U, P, dS, dbias)            K.xattn_tn(P_or_dS, dc_or_ln, out)

Everything on the B*S = 256 text rows uses the existing GEMM entry point on strided head views.  The sequencing and the
hand-derived backward below are pinned against autograd through the oracle on CPU (tests/test_functional_cpu.py::
test_reassociated_i2t_core) with the torch restatement of those four kernels (tests/fake_kernels.py); calling this module
with the real library raises AttributeError -- there is no fallback."""
import types

import torch

from . import functional as Fn
from .lib import GEMM_NN, GEMM_NT, GEMM_TN

_e = Fn._e


def _blockdiag_rows(vec, H):
    """[C] -> [H, C] with vec's head-h slice in row h (the query bias as a block-diagonal operand)."""
    C = vec.numel()
    d = C // H
    out = vec.new_zeros(H, C)
    for h in range(H):
        out[h, h * d:(h + 1) * d] = vec[h * d:(h + 1) * d]
    return out


def i2t_core_fwd(K, ln, kv, mask, wq, bq, wp, bp, H):
    """ln [B,N,C] operand dtype; kv [B,S,2C] operand dtype (k | v); mask [B,S] f32 additive; wq / wp [C,C] operand
    dtype; bq / bp [C] f32.  Returns (c [B,N,C] operand dtype = proj_i2t(attention) incl. its bias, saved)."""
    BF16, F32 = Fn.BF16, torch.float32
    B, N, C = ln.shape
    S = kv.shape[1]
    d = C // H
    sc = d ** -0.5
    kv2 = kv.view(B * S, 2 * C)
    Mt, U = _e(ln, (H, B * S, C), BF16), _e(ln, (H, B * S, C), BF16)
    bp_h = (bp / H).contiguous()
    for h in range(H):
        kh, vh = kv2[:, h * d:(h + 1) * d], kv2[:, C + h * d:C + (h + 1) * d]
        K.gemm(GEMM_NN, kh, wq[h * d:(h + 1) * d, :], scale=sc, out_bf16=Mt[h])
        K.gemm(GEMM_NT, vh, wp[:, h * d:(h + 1) * d], bias=bp_h, out_bf16=U[h])
    bq_bd = _blockdiag_rows(bq, H).to(BF16)
    c0 = _e(ln, (B * S, H), F32)
    K.gemm(GEMM_NT, kv2[:, :C], bq_bd, scale=sc, out_f32=c0)
    P = _e(ln, (B, N, H, S), BF16)
    K.xattn_scores_softmax(ln, Mt.view(H, B, S, C), c0.view(B, S, H), mask, P)
    c = _e(ln, (B, N, C), BF16)
    K.xattn_weighted_sum(P, U.view(H, B, S, C), c)
    return c, types.SimpleNamespace(ln=ln, kv=kv, Mt=Mt, U=U, P=P, bq_bd=bq_bd, H=H)


def i2t_core_bwd(K, s, dc, wq, wp):
    """dc [B,N,C] operand dtype -> (dln [B,N,C], dkv [B,S,2C] f32, dwq [C,C], dbq [C], dwp [C,C], dbp [C])  (f32 grads)."""
    BF16, F32 = Fn.BF16, torch.float32
    B, N, C = dc.shape
    H, S = s.H, s.kv.shape[1]
    d = C // H
    sc = d ** -0.5
    kv2 = s.kv.view(B * S, 2 * C)
    dS = _e(dc, (B, N, H, S), BF16)
    dc0 = _e(dc, (B, S, H), F32)
    K.xattn_dscores(dc, s.U.view(H, B, S, C), s.P, dS, dc0)
    dU, dM = _e(dc, (H, B * S, C), F32), _e(dc, (H, B * S, C), F32)
    K.xattn_tn(s.P, dc, dU.view(H, B, S, C))
    K.xattn_tn(dS, s.ln, dM.view(H, B, S, C))
    dln = _e(dc, (B, N, C), BF16)
    K.xattn_weighted_sum(dS, s.Mt.view(H, B, S, C), dln)
    dU_b, dM_b = _e(dc, dU.shape, BF16), _e(dc, dM.shape, BF16)
    K.cast(dU, dU_b)
    K.cast(dM, dM_b)
    dkv = _e(dc, (B * S, 2 * C), F32)
    dwq, dwp = _e(dc, (C, C), F32), _e(dc, (C, C), F32)
    for h in range(H):
        hs = slice(h * d, (h + 1) * d)
        kh, vh = kv2[:, hs], kv2[:, C + h * d:C + (h + 1) * d]
        K.gemm(GEMM_NT, dM_b[h], wq[hs, :], scale=sc, out_f32=dkv[:, hs])                     # dk_h  = d^-1/2 dM_h Wq_h^T
        K.gemm(GEMM_TN, kh, dM_b[h], scale=sc, out_f32=dwq[hs, :])                             # dWq_h = d^-1/2 k_h^T dM_h
        K.gemm(GEMM_NN, dU_b[h], wp[:, hs], out_f32=dkv[:, C + h * d:C + (h + 1) * d])         # dv_h  = dU_h Wp[:,h]
        K.gemm(GEMM_TN, dU_b[h], vh, out_f32=dwp[:, hs])                                       # dWp[:,h] = dU_h^T v_h
    # the query bias: scores += d^-1/2 (bq_h . k_h)  ->  dk += d^-1/2 dc0 (x) bq ; dbq_h = d^-1/2 k_h^T dc0[:, h]
    dc0_b = _e(dc, (B * S, H), BF16)
    K.cast(dc0.view(B * S, H), dc0_b)
    K.gemm(GEMM_NN, dc0_b, s.bq_bd, scale=sc, out_f32=dkv[:, :C], accumulate=True)
    dbq_all = _e(dc, (C, H), F32)
    K.gemm(GEMM_TN, kv2[:, :C], dc0_b, scale=sc, out_f32=dbq_all)
    dbq = torch.cat([dbq_all[h * d:(h + 1) * d, h] for h in range(H)])
    dbp = _e(dc, (C,), F32)
    K.colsum(dU_b.view(H * B * S, C), dbp, scale=1.0 / H)
    return dln, dkv.view(B, S, 2 * C), dwq, dbq, dwp, dbp


# ----------------------------------------------------------------------------------------------------------------------
# text -> video direction (roberta.py:470-486: q from the text, k / v = Linear(x) of all N video tokens, no mask)
#
#     scores[b,h,s,n] = Qp[h,b,s,:] . x[b,n,:]            Qp = d^-1/2 q_h Wk_h   [H, B*S, Cv]   (q_h . bk_h drops out of softmax)
#     ctx_h           = (P_h x) Wv_h^T + bv_h             Z  = P x               [H, B*S, Cv]
#
# The key / value projection of the video tokens (2 * M * Cv * 2C FLOPs, two [M, C] tensors written) disappears; what is
# left on the M = B*N video rows is ONE flash-style kernel per direction with K = V = x and "head dim" Cv, still to be written:
#     K.xattn_t2i_flash(Qp, x, Z, lse)          K.xattn_t2i_flash_bwd(dZ, Qp, x, Z, lse, dQp, dx)
def t2i_core_fwd(K, q, x, wk, wv, bv, H):
    """q [B,S,C] operand dtype (query projection of the text, bias included); x [B,N,Cv] operand dtype (video stream);
    wk / wv [C,Cv] operand dtype; bv [C] f32.  Returns (ctx [B,S,C] f32 = merged heads before output.dense, saved)."""
    BF16, F32 = Fn.BF16, torch.float32
    B, S, C = q.shape
    Cv = x.shape[2]
    d = C // H
    sc = d ** -0.5
    q2 = q.view(B * S, C)
    Qp = _e(q, (H, B * S, Cv), BF16)
    for h in range(H):
        K.gemm(GEMM_NN, q2[:, h * d:(h + 1) * d], wk[h * d:(h + 1) * d, :], scale=sc, out_bf16=Qp[h])
    Z, lse = _e(q, (H, B * S, Cv), BF16), _e(q, (H, B * S), F32)
    K.xattn_t2i_flash(Qp.view(H, B, S, Cv), x, Z.view(H, B, S, Cv), lse.view(H, B, S))
    ctx = _e(q, (B * S, C), F32)
    for h in range(H):
        hs = slice(h * d, (h + 1) * d)
        K.gemm(GEMM_NT, Z[h], wv[hs, :], bias=bv[hs], out_f32=ctx[:, hs])
    return ctx.view(B, S, C), types.SimpleNamespace(q=q, x=x, Qp=Qp, Z=Z, lse=lse, H=H)


def t2i_core_bwd(K, s, dctx, wk, wv):
    """dctx [B,S,C] operand dtype -> (dq [B,S,C] f32, dx [B,N,Cv] f32, dwk [C,Cv], dwv [C,Cv], dbv [C]); the key bias has
    an analytically zero gradient."""
    BF16, F32 = Fn.BF16, torch.float32
    B, S, C = dctx.shape
    H, Cv = s.H, s.x.shape[2]
    d = C // H
    sc = d ** -0.5
    g2, q2 = dctx.view(B * S, C), s.q.view(B * S, C)
    dZ = _e(dctx, (H, B * S, Cv), BF16)
    dwk, dwv = _e(dctx, (C, Cv), F32), _e(dctx, (C, Cv), F32)
    for h in range(H):
        hs = slice(h * d, (h + 1) * d)
        K.gemm(GEMM_NN, g2[:, hs], wv[hs, :], out_bf16=dZ[h])                                   # dZ_h  = dctx_h Wv_h
        K.gemm(GEMM_TN, g2[:, hs], s.Z[h], out_f32=dwv[hs, :])                                   # dWv_h = dctx_h^T Z_h
    dbv = _e(dctx, (C,), F32)
    K.colsum(g2, dbv)
    dQp, dx = _e(dctx, (H, B * S, Cv), F32), _e(dctx, s.x.shape, F32)
    K.xattn_t2i_flash_bwd(dZ.view(H, B, S, Cv), s.Qp.view(H, B, S, Cv), s.x, s.Z.view(H, B, S, Cv), s.lse.view(H, B, S),
                          dQp.view(H, B, S, Cv), dx)
    dQp_b = _e(dctx, dQp.shape, BF16)
    K.cast(dQp, dQp_b)
    dq = _e(dctx, (B * S, C), F32)
    for h in range(H):
        hs = slice(h * d, (h + 1) * d)
        K.gemm(GEMM_NT, dQp_b[h], wk[hs, :], scale=sc, out_f32=dq[:, hs])                        # dq_h  = d^-1/2 dQp_h Wk_h^T
        K.gemm(GEMM_TN, q2[:, hs], dQp_b[h], scale=sc, out_f32=dwk[hs, :])                       # dWk_h = d^-1/2 q_h^T dQp_h
    return dq.view(B, S, C), dx, dwk, dwv, dbv
