"""The two gated cross-attentions of the fusion layers, re-associated around the S text tokens
(oracle/egovlp_oracle.py::cross_attention_{i2t,t2i}_reassociated prove the algebra exact).

video -> text (video_transformer.py:155-185: q = qkv_i2t(LN(a)), softmax(q k^T + mask) v, proj_i2t, gate):

    P[b,n,h,:] = softmax_s( LN(a)[b,n,:] . Mt[b,h,s,:] + c0[b,h,s] + mask[b,s] )      Mt = d^-1/2 k_h Wq_h       [B, H, S, C]
    out[b,n,:] = xa[b,n,:] + alpha * ( sum_{h,s} P[b,n,h,s] U[b,h,s,:] + bp )          U  = v_h Wp[:,h]^T         [B, H, S, C]

    -- the [M, C] x [C, C] query and output projections (M = B*N = 25 096 rows at the BASELINE shapes) and the 32-key attention
    launch become two [N, C] x [C, H*S] products per clip with a 32-wide group softmax in the first one's epilogue.

text -> video (roberta.py:470-486: q from the text, k / v = Linear(x) of all N video tokens, no mask):

    scores[b,h,s,n] = Qp[b,h,s,:] . x[b,n,:]        Qp = d^-1/2 q_h Wk_h   (q_h . bk_h is constant over keys: drops out)
    ctx_h           = (P_h x) Wv_h^T + bv_h         Z  = P x               [B, H, S, Cv]

    -- the key / value projection of the video tokens (2 * M * Cv * 2C FLOPs, two [M, C] tensors written, and the same again
    twice in the backward) disappears; the video stream x itself is the key and the value.

Everything runs on `K.bgemm` (csrc/xgemm.cu: batched tcgen05 GEMM over rank-4 TMA maps, batches = clips or (clip, head)
pairs) plus four small row kernels (csrc/xattn.cu).  The hand-derived backward is pinned against autograd through the
reference formulation (tests/test_functional_cpu.py::test_reassociated_*), on the GPU against the oracle at C = 768."""
import types

import torch

from .lib import BV, EPI_DSOFTMAX32, EPI_SOFTMAX32, GEMM_NN, GEMM_NT, GEMM_TN

F32 = torch.float32
HD = 64          # head dim of every attention on this path
GROUP = 32       # softmax group width of the bgemm epilogues = text tokens per clip


def _e(like, shape, dtype):
    return torch.empty(shape, dtype=dtype, device=like.device)


def _z(like, shape, dtype=F32):
    return torch.zeros(shape, dtype=dtype, device=like.device)


def i2t_supported(S, C, H):
    return S == GROUP and C == H * HD


def t2i_supported(S, N, C, H):
    return C == H * HD and N <= 65536 and S <= 128     # rows of more than 4093 video tokens take the streaming row kernels


# ----------------------------------------------------------------------------------------------------------------------
# video -> text.  Two halves: the TEXT-side preparation (functions of the text states and weights only: [B*S] = 256 rows;
# small, launch-latency-bound products that the model runs on the text tower's stream, next to the video tower) and the
# VIDEO-side application (two per-clip GEMMs over the M = B*N token rows in each direction).
def i2t_prep_fwd(K, kv, mask, wq, bq, wp, B, H):
    """kv [B*S, 2C] (k | v of the text, operand dtype); mask [B, S] f32 additive or None; wq / wp [C, C]; bq [C] f32.
    -> Mt [B, H, S, C] = d^-1/2 k_h Wq_h,  U [B, H, S, C] = v_h Wp[:, h]^T,  bias1 [B, H*S] f32 = d^-1/2 bq_h . k_h + mask."""
    C = kv.shape[1] // 2
    S = kv.shape[0] // B
    HS = H * S
    sc = HD ** -0.5
    dt = kv.dtype
    Mt, U = _e(kv, (B, H, S, C), dt), _e(kv, (B, H, S, C), dt)
    ldkv = kv.stride(0)
    K.bgemm(GEMM_NN, S, C, HD, BV(kv, ldkv, HD, S * ldkv), BV(wq, wq.stride(0), HD * wq.stride(0), 0), nb=(H, B), scale=sc,
            out_bf16=BV(Mt, C, S * C, HS * C))
    K.bgemm(GEMM_NT, S, C, HD, BV(kv[:, C:], ldkv, HD, S * ldkv), BV(wp, wp.stride(0), HD, 0), nb=(H, B),
            out_bf16=BV(U, C, S * C, HS * C))
    bias1 = _e(kv, (B, HS), F32)
    K.xattn_qbias_fwd(kv, ldkv, bq, mask, sc, B, S, H, bias1)
    return Mt, U, bias1


def i2t_apply_fwd(K, lnc, Mt, U, bias1, bp, alpha, xa, out, B, N, H):
    """lnc [B*N, C] = LN(a); xa [B*N, C] f32 (residual x + a); out [B*N, C] f32 <- xa + alpha * (P U + bp) with
    P [B, N, H*S] = group-softmax(lnc Mt^T + bias1), which is returned (saved for the backward)."""
    C = lnc.shape[1]
    HS = Mt.shape[1] * Mt.shape[2]
    P = _e(lnc, (B, N, HS), lnc.dtype)
    K.bgemm(GEMM_NT, N, HS, C, BV(lnc, C, N * C), BV(Mt, C, HS * C), nb=(B, 1), bias=BV(bias1, 0, HS), out_bf16=BV(P, HS, N * HS),
            epilogue=EPI_SOFTMAX32)
    K.bgemm(GEMM_NN, N, C, HS, BV(P, HS, N * HS), BV(U, C, HS * C), nb=(B, 1), bias=BV(bp, 0, 0), scale_dev=alpha,
            residual=BV(xa, C, N * C), out_f32=BV(out, C, N * C))
    return P


def i2t_apply_bwd(K, lnc, Mt, U, P, d_out_bf, cs_out, bp, alpha, dalpha, B, N, H):
    """d_out_bf [B*N, C] = gradient of `out`; cs_out [C] f32 = its column sums.  ACCUMULATES d alpha into dalpha [1];
    returns (d_lnc [B*N, C], dM [B, H, S, C], dU [B, H, S, C] (operand dtype), dbias1 [B, H*S] f32)."""
    C = lnc.shape[1]
    S = Mt.shape[2]
    HS = H * S
    dt = d_out_bf.dtype
    # dS = alpha * P * (dP - rowdot(P, dP)),  dP = d_out U^T ;  d alpha = sum rowdot ;  dbias1 = column sums of dS
    dS = _e(d_out_bf, (B, N, HS), dt)
    dbias1 = _z(d_out_bf, (B, HS))
    K.bgemm(GEMM_NT, N, HS, C, BV(d_out_bf, C, N * C), BV(U, C, HS * C), nb=(B, 1), scale_dev=alpha, aux=BV(P, HS, N * HS),
            out_bf16=BV(dS, HS, N * HS), colsum=BV(dbias1, 0, HS), dot_out=dalpha, epilogue=EPI_DSOFTMAX32)
    K.dot(cs_out, bp, dalpha, accumulate=True)        # the bias part of c = P U + bp:  d alpha += sum_n d_out[n] . bp
    d_lnc = _e(d_out_bf, (B * N, C), dt)
    K.bgemm(GEMM_NN, N, C, HS, BV(dS, HS, N * HS), BV(Mt, C, HS * C), nb=(B, 1), out_bf16=BV(d_lnc, C, N * C))
    # dU[b] = alpha * P[b]^T d_out[b] ;  dM[b] = dS[b]^T LN(a)[b]      (K = N tokens per clip, split-K + fp32 atomics)
    dU, dM = _z(d_out_bf, (B, H, S, C)), _z(d_out_bf, (B, H, S, C))
    K.bgemm(GEMM_TN, HS, C, N, BV(P, HS, N * HS), BV(d_out_bf, C, N * C), nb=(B, 1), scale_dev=alpha,
            out_f32=BV(dU, C, HS * C), accumulate=True)
    K.bgemm(GEMM_TN, HS, C, N, BV(dS, HS, N * HS), BV(lnc, C, N * C), nb=(B, 1), out_f32=BV(dM, C, HS * C), accumulate=True)
    dU_b, dM_b = _e(dU, dU.shape, dt), _e(dM, dM.shape, dt)
    K.cast(dU, dU_b)
    K.cast(dM, dM_b)
    return d_lnc, dM_b, dU_b, dbias1


def i2t_prep_bwd(K, kv, wq, bq, wp, dM_b, dU_b, dbias1, dwq, dbq, dwp, B, H):
    """Backward of i2t_prep_fwd: ACCUMULATES into dwq [C, C], dbq [C], dwp [C, C] (f32) and returns dkv [B*S, 2C] f32."""
    C = kv.shape[1] // 2
    S = kv.shape[0] // B
    HS = H * S
    sc = HD ** -0.5
    ldkv = kv.stride(0)
    dkv = _e(kv, (B * S, 2 * C), F32)
    bh = dict(nb=(H, B))
    K.bgemm(GEMM_NT, S, HD, C, BV(dM_b, C, S * C, HS * C), BV(wq, wq.stride(0), HD * wq.stride(0), 0), scale=sc,
            out_f32=BV(dkv, 2 * C, HD, S * 2 * C), **bh)                                   # dk_h  = d^-1/2 dM_h Wq_h^T
    K.bgemm(GEMM_NN, S, HD, C, BV(dU_b, C, S * C, HS * C), BV(wp, wp.stride(0), HD, 0),
            out_f32=BV(dkv[:, C:], 2 * C, HD, S * 2 * C), **bh)                            # dv_h  = dU_h Wp[:, h]
    K.bgemm(GEMM_TN, HD, C, S, BV(kv, ldkv, HD, S * ldkv), BV(dM_b, C, S * C, HS * C), scale=sc,
            out_f32=BV(dwq, C, HD * C, 0), accumulate=True, **bh)                          # dWq_h += d^-1/2 k_h^T dM_h
    K.bgemm(GEMM_TN, C, HD, S, BV(dU_b, C, S * C, HS * C), BV(kv[:, C:], ldkv, HD, S * ldkv),
            out_f32=BV(dwp, C, HD, 0), accumulate=True, **bh)                              # dWp[:, h] += dU_h^T v_h
    # the query bias: scores += d^-1/2 bq_h . k_h  ->  dk_h += d^-1/2 dbias1 (x) bq_h ;  dbq_h += d^-1/2 k_h^T dbias1
    K.xattn_qbias_bwd(kv, ldkv, bq, dbias1, sc, B, S, H, dk=dkv, lddk=2 * C, dbq=dbq)
    return dkv


def i2t_fwd(K, lnc, kv, mask, wq, bq, wp, bp, alpha, xa, out, B, N, H):
    """both halves in sequence (see i2t_prep_fwd / i2t_apply_fwd); returns the saved tensors"""
    Mt, U, bias1 = i2t_prep_fwd(K, kv, mask, wq, bq, wp, B, H)
    P = i2t_apply_fwd(K, lnc, Mt, U, bias1, bp, alpha, xa, out, B, N, H)
    return types.SimpleNamespace(lnc=lnc, kv=kv, Mt=Mt, U=U, P=P, B=B, N=N, H=H, S=kv.shape[0] // B)


def i2t_bwd(K, s, d_out_bf, cs_out, wq, bq, wp, bp, alpha, dalpha, dwq, dbq, dwp):
    """-> (d_lnc [B*N, C], dkv [B*S, 2C] f32); ACCUMULATES into dalpha [1], dwq [C, C], dbq [C], dwp [C, C]."""
    d_lnc, dM_b, dU_b, dbias1 = i2t_apply_bwd(K, s.lnc, s.Mt, s.U, s.P, d_out_bf, cs_out, bp, alpha, dalpha, s.B, s.N, s.H)
    dkv = i2t_prep_bwd(K, s.kv, wq, bq, wp, dM_b, dU_b, dbias1, dwq, dbq, dwp, s.B, s.H)
    return d_lnc, dkv


# ----------------------------------------------------------------------------------------------------------------------
def t2i_fwd(K, q, x_bf, wk, wv, bv, ox, B, N, H, p_drop=0.0, seed_dev=None, site=0):
    """q [B*S, C] bf16 (text query projection, bias included); x_bf [B*N, Cv] bf16 (un-normalised video stream entering
    the video block); wk / wv [C, Cv] bf16; bv [C] f32; ox [B*S, C] bf16 is written with the merged heads' context
    (the input of crossattention_t2i.output.dense).  p_drop > 0: dropout on the probabilities (roberta.py:313)."""
    C = q.shape[1]
    Cv = x_bf.shape[1]
    S = q.shape[0] // B
    HS = H * S
    sc = HD ** -0.5
    Np = (N + 7) // 8 * 8
    BF16 = q.dtype
    # ZQ[b, 0] = dZ (backward), ZQ[b, 1] = Qp: stacked so that dx = [P; dS]^T [dZ; Qp] is ONE product
    ZQ = _e(q, (B, 2, HS, Cv), BF16)
    Qp = ZQ[:, 1]
    K.bgemm(GEMM_NN, S, Cv, HD, BV(q, C, HD, S * C), BV(wk, wk.stride(0), HD * wk.stride(0), 0), nb=(H, B), scale=sc,
            out_bf16=BV(Qp, Cv, S * Cv, 2 * HS * Cv))
    Sc = _e(q, (B, HS, Np), F32)
    K.bgemm(GEMM_NT, HS, N, Cv, BV(Qp, Cv, 2 * HS * Cv), BV(x_bf, Cv, N * Cv), nb=(B, 1), out_f32=BV(Sc, Np, HS * Np))
    PdS = _e(q, (B, 2, HS, Np), BF16)              # [b, 0] = P (after dropout), [b, 1] = dS (backward)
    lse, rsum = _e(q, (B * HS,), F32), _e(q, (B * HS,), F32)
    K.xattn_row_softmax(Sc, Np, B * HS, HS, HS * Np, N, PdS, Np, 2 * HS * Np, lse, p_drop, seed_dev, site, rsum)
    Zf = _z(q, (B, H, S, Cv))
    K.bgemm(GEMM_NN, HS, Cv, N, BV(PdS, Np, 2 * HS * Np), BV(x_bf, Cv, N * Cv), nb=(B, 1), out_f32=BV(Zf, Cv, HS * Cv),
            accumulate=True)
    Z = _e(q, (B, H, S, Cv), BF16)
    K.cast(Zf, Z)
    # ctx[b*S+s, h] = Z[b,h,s,:] Wv_h^T + rowsum(P)[b,h,s] * bv_h   (row sum = 1 without dropout)
    if p_drop > 0.0:
        K.bgemm(GEMM_NT, S, HD, Cv, BV(Z, Cv, S * Cv, HS * Cv), BV(wv, wv.stride(0), HD * wv.stride(0), 0), nb=(H, B),
                out_bf16=BV(ox, C, HD, S * C))
        K.xattn_rowscale_bias(ox, rsum, bv, B, S, H)
    else:
        K.bgemm(GEMM_NT, S, HD, Cv, BV(Z, Cv, S * Cv, HS * Cv), BV(wv, wv.stride(0), HD * wv.stride(0), 0), nb=(H, B),
                bias=BV(bv, 0, HD, 0), out_bf16=BV(ox, C, HD, S * C))
    return types.SimpleNamespace(q=q, x=x_bf, ZQ=ZQ, Sc=Sc, PdS=PdS, lse=lse, rsum=rsum, Z=Z, B=B, N=N, H=H, S=S, Np=Np,
                                 p_drop=p_drop, seed_dev=seed_dev, site=site)


def t2i_bwd(K, s, d_ox, wk, wv, bv, dwk, dwv, dbv, dvideo):
    """d_ox [B*S, C] bf16.  ACCUMULATES into dwk / dwv [C, Cv] and dbv [C] (f32; the key bias has an analytically zero
    gradient), writes dvideo [B*N, Cv] f32 and returns dq [B*S, C] bf16."""
    B, N, H, S, Np = s.B, s.N, s.H, s.S, s.Np
    C = d_ox.shape[1]
    Cv = s.x.shape[1]
    HS = H * S
    sc = HD ** -0.5
    bh = dict(nb=(H, B))
    BF16 = d_ox.dtype
    dZ, Qp = s.ZQ[:, 0], s.ZQ[:, 1]
    K.bgemm(GEMM_NN, S, Cv, HD, BV(d_ox, C, HD, S * C), BV(wv, wv.stride(0), HD * wv.stride(0), 0),
            out_bf16=BV(dZ, Cv, S * Cv, 2 * HS * Cv), **bh)                                 # dZ_h  = d_ox_h Wv_h
    K.bgemm(GEMM_TN, HD, Cv, S, BV(d_ox, C, HD, S * C), BV(s.Z, Cv, S * Cv, HS * Cv), out_f32=BV(dwv, Cv, HD * Cv, 0),
            accumulate=True, **bh)                                                         # dWv_h += d_ox_h^T Z_h
    if s.p_drop > 0.0:
        K.xattn_rowscale_bias_bwd(d_ox, s.rsum, dbv, B, S, H)
    else:
        K.colsum(d_ox, dbv, accumulate=True)
    dPf = _e(d_ox, (B, HS, Np), F32)
    K.bgemm(GEMM_NT, HS, N, Cv, BV(dZ, Cv, 2 * HS * Cv), BV(s.x, Cv, N * Cv), nb=(B, 1), out_f32=BV(dPf, Np, HS * Np))
    dS = s.PdS[:, 1]
    crow = None
    if s.p_drop > 0.0:
        # with dropout the value bias no longer cancels in the softmax backward: dP += d_ox_h . bv_h (a row constant)
        crow = _e(d_ox, (B, HS), F32)
        K.xattn_qbias_fwd(d_ox, C, bv, None, 1.0, B, S, H, crow)
    K.xattn_row_dsoftmax(s.Sc, Np, B * HS, HS, HS * Np, N, s.lse, dPf, Np, HS * Np, dS, Np, 2 * HS * Np, s.p_drop, s.seed_dev,
                         s.site, row_const=crow)
    dQpf = _z(d_ox, (B, H, S, Cv))
    K.bgemm(GEMM_NN, HS, Cv, N, BV(dS, Np, 2 * HS * Np), BV(s.x, Cv, N * Cv), nb=(B, 1), out_f32=BV(dQpf, Cv, HS * Cv),
            accumulate=True)
    dQp = _e(d_ox, (B, H, S, Cv), BF16)
    K.cast(dQpf, dQp)
    # dx[b] = P[b]^T dZ[b] + dS[b]^T Qp[b] = [P; dS][b]^T [dZ; Qp][b]
    K.bgemm(GEMM_TN, N, Cv, 2 * HS, BV(s.PdS, Np, 2 * HS * Np), BV(s.ZQ, Cv, 2 * HS * Cv), nb=(B, 1), out_f32=BV(dvideo, Cv, N * Cv))
    dq = _e(d_ox, (B * S, C), BF16)
    K.bgemm(GEMM_NT, S, HD, Cv, BV(dQp, Cv, S * Cv, HS * Cv), BV(wk, wk.stride(0), HD * wk.stride(0), 0), scale=sc,
            out_bf16=BV(dq, C, HD, S * C), **bh)                                            # dq_h  = d^-1/2 dQp_h Wk_h^T
    K.bgemm(GEMM_TN, HD, Cv, S, BV(s.q, C, HD, S * C), BV(dQp, Cv, S * Cv, HS * Cv), scale=sc, out_f32=BV(dwk, Cv, HD * Cv, 0),
            accumulate=True, **bh)                                                         # dWk_h += d^-1/2 q_h^T dQp_h
    return dq
