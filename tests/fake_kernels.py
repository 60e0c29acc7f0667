"""TEST INFRASTRUCTURE ONLY -- plain-torch restatement of the `egovlpv2_b200.lib.Kernels` interface.

Two uses:
  * `-m gpu` tests: per-kernel reference (same inputs, fp32 math) for the CUDA kernels;
  * `-m "not gpu"` tests: installed with `lib.set_kernels(FakeKernels())` so the host-side logic
    (hand-written backward sequences, module API, state_dict handling) can be checked on CPU against
    the oracle.  The product package never imports this file and has no CPU path of its own.
bf16 outputs are rounded to bf16 like the kernels do; everything else is fp32.
"""
import math

import torch
import torch.nn.functional as F

from egovlpv2_b200.lib import (ACT_GELU, ACT_GELU_BWD, ACT_GELU_DG, ACT_MUL_AUX, ACT_NONE, ACT_RELU, ACT_RELU_BWD, ACT_TANH,
                               ACT_TANH_BWD, GEMM_NN, GEMM_NT, GEMM_TN)


def attn_indices(spec, B, device):
    """Row indices (within one batch element) of queries [G, Lq] and keys [G, LkT]."""
    g = torch.arange(spec.G, device=device)[:, None]
    qi = spec.q_row0 + g * spec.q_gstride + torch.arange(spec.Lq, device=device)[None] * spec.q_istride
    ki = spec.k_row0 + g * spec.k_gstride + torch.arange(spec.Lk, device=device)[None] * spec.k_istride
    if spec.has_cls_key:
        ki = torch.cat([torch.full((spec.G, 1), spec.cls_row, device=device, dtype=ki.dtype), ki], 1)
    return qi, ki


def attn_reference(spec, q, k, v, key_bias=None):
    """fp32 attention on [B, rows, H*64] tensors; returns o_groups [B,H,G,Lq,64] and lse (log2 domain)."""
    B = q.shape[0]
    H = spec.H
    qi, ki = attn_indices(spec, B, q.device)
    qh = q.float().reshape(B, q.shape[1], H, 64)
    kh = k.float().reshape(B, k.shape[1], H, 64)
    vh = v.float().reshape(B, v.shape[1], H, 64)
    Q = qh[:, qi].permute(0, 3, 1, 2, 4)      # [B,H,G,Lq,64]
    K = kh[:, ki].permute(0, 3, 1, 2, 4)
    V = vh[:, ki].permute(0, 3, 1, 2, 4)
    s = spec.scale * (Q @ K.transpose(-1, -2))
    if key_bias is not None:
        s = s + key_bias.reshape(B, 1, 1, 1, -1).clamp_min(-1e30)
    lse = torch.logsumexp(s, -1)
    p = torch.softmax(s, -1)
    return p @ V, lse * math.log2(math.e), qi, ki


class FakeKernels:
    def __init__(self):
        self._launches = 0

    def launch_count(self):
        return self._launches

    def sm_count(self):
        return 148

    def force_simt(self, on):
        pass

    def mark(self, region):
        pass

    # ------------------------------------------------------------------ GEMM
    def gemm(self, layout, A, B, *, bias=None, aux=None, act=ACT_NONE, scale=1.0, scale_dev=None, residual=None,
             out_f32=None, out_bf16=None, out_pre=None, accumulate=False, split_k=1, colsum=None):
        self._launches += 1
        a, b = A.float(), B.float()
        if layout == GEMM_NT:
            v = a @ b.t()
        elif layout == GEMM_NN:
            v = a @ b
        else:
            v = a.t() @ b
        if bias is not None:
            v = v + bias
        if out_pre is not None:
            if act == ACT_GELU_DG:   # the derivative is saved instead of the pre-activation
                cdf = 0.5 * (1 + torch.erf(v / math.sqrt(2)))
                out_pre.copy_(cdf + v * torch.exp(-0.5 * v * v) / math.sqrt(2 * math.pi))
            else:
                out_pre.copy_(v)
        if act in (ACT_GELU, ACT_GELU_DG):
            v = F.gelu(v)
        elif act == ACT_MUL_AUX:
            v = v * aux.float()
        elif act == ACT_RELU:
            v = F.relu(v)
        elif act == ACT_TANH:
            v = torch.tanh(v)
        elif act == ACT_GELU_BWD:
            x = aux.float()
            cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
            pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
            v = v * (cdf + x * pdf)
        elif act == ACT_RELU_BWD:
            v = v * (aux.float() > 0)
        elif act == ACT_TANH_BWD:
            v = v * (1 - aux.float() ** 2)
        s = scale * (float(scale_dev.item()) if scale_dev is not None else 1.0)
        v = v * s
        if residual is not None:
            v = v + residual
        if out_f32 is not None:
            if accumulate:
                out_f32.add_(v)
            else:
                out_f32.copy_(v)
        if out_bf16 is not None:
            out_bf16.copy_(v)
        if colsum is not None:
            colsum.add_(v.sum(0))

    # ------------------------------------------------------------------ LayerNorm
    def layernorm_fwd(self, x, gamma, beta, eps, y_bf16=None, y_f32=None, mean=None, rstd=None):
        self._launches += 1
        xf = x.float()
        mu = xf.mean(-1, keepdim=True)
        var = ((xf - mu) ** 2).mean(-1, keepdim=True)
        rs = torch.rsqrt(var + eps)
        y = (xf - mu) * rs * gamma + beta
        if y_f32 is not None:
            y_f32.copy_(y.reshape(y_f32.shape))
        if y_bf16 is not None:
            y_bf16.copy_(y.reshape(y_bf16.shape))
        if mean is not None:
            mean.copy_(mu.reshape(mean.shape))
        if rstd is not None:
            rstd.copy_(rs.reshape(rstd.shape))

    def layernorm_bwd(self, dy, x, gamma, mean, rstd, add=None, dx=None, dx_bf16=None, bf16_total=True, dgamma=None,
                      dbeta=None, out_colsum=None):
        self._launches += 1
        Cd = x.shape[-1]
        xf, d = x.float().reshape(-1, Cd), dy.float().reshape(-1, Cd)
        xh = (xf - mean.reshape(-1, 1)) * rstd.reshape(-1, 1)
        gy = d * gamma
        val = rstd.reshape(-1, 1) * (gy - gy.mean(-1, keepdim=True) - xh * (gy * xh).mean(-1, keepdim=True))
        tot = val if add is None else val + add.reshape(-1, Cd)
        if dx is not None:
            dx.copy_(tot.reshape(dx.shape))
        if dx_bf16 is not None:
            dx_bf16.copy_((tot if bf16_total else val).reshape(dx_bf16.shape))
        if out_colsum is not None:
            out_colsum.add_((tot if bf16_total else val).sum(0))
        if dgamma is not None:
            dgamma.add_((d * xh).sum(0))
        if dbeta is not None:
            dbeta.add_(d.sum(0))

    # ------------------------------------------------------------------ attention
    def attention_fwd(self, spec, q, k, v, o, lse, key_bias=None):
        self._launches += 1
        og, l2, qi, _ = attn_reference(spec, q, k, v, key_bias)
        B, H = q.shape[0], spec.H
        og = og.permute(0, 2, 3, 1, 4).reshape(B, spec.G * spec.Lq, H * 64)  # [B, G*Lq, H*64]
        o[:, qi.reshape(-1)] = og.to(o.dtype)
        lse.copy_(l2.reshape(lse.shape))

    def attention_bwd(self, spec, q, k, v, o, lse, d_o, dq, dk, dv, delta, dkv_cls=None, dkv_accumulate=False,
                      key_bias=None):
        self._launches += 2
        B, H = q.shape[0], spec.H
        qf = q.float().detach().clone().requires_grad_(True)
        kf = k.float().detach().clone().requires_grad_(True)
        vf = v.float().detach().clone().requires_grad_(True)
        with torch.enable_grad():
            og, _, qi, ki = attn_reference(spec, qf, kf, vf, key_bias)
        dog = d_o.float().reshape(B, d_o.shape[1], H, 64)[:, qi].permute(0, 3, 1, 2, 4)
        with torch.enable_grad():
            og.backward(dog)
        delta.copy_((dog * og.detach()).sum(-1).reshape(delta.shape))

        def scatter_rows(dst, grad, rows, accumulate):
            """write grad rows (only the addressed rows/columns are touched, like the kernel)"""
            g = grad.reshape(B, grad.shape[1], H * 64)
            sel = torch.zeros(grad.shape[1], dtype=torch.bool, device=grad.device)
            sel[rows.reshape(-1)] = True
            cur = dst[:, sel].float() if accumulate else 0
            dst[:, sel] = (cur + g[:, sel]).to(dst.dtype)

        scatter_rows(dq, qf.grad, qi, False)
        krows = ki[:, 1:] if spec.has_cls_key else ki
        scatter_rows(dk, kf.grad, krows, dkv_accumulate)
        scatter_rows(dv, vf.grad, krows, dkv_accumulate)
        if spec.has_cls_key:
            dkv_cls.view(B, H, 2, 64)[:, :, 0] += kf.grad[:, spec.cls_row].reshape(B, H, 64)
            dkv_cls.view(B, H, 2, 64)[:, :, 1] += vf.grad[:, spec.cls_row].reshape(B, H, 64)

    def attention_cls_finalize(self, dkv_cls, dk, dv, H, cls_row=0, accumulate=False):
        self._launches += 1
        B = dk.shape[0]
        g = dkv_cls.view(B, H, 2, 64)
        for dst, j in ((dk, 0), (dv, 1)):
            cur = dst[:, cls_row].float() if accumulate else 0
            dst[:, cls_row] = (cur + g[:, :, j].reshape(B, H * 64)).to(dst.dtype)

    # ------------------------------------------------------------------ elementwise
    def cast(self, x, y):
        self._launches += 1
        y.copy_(x.to(y.dtype))

    def axpy(self, a, b, alpha=1.0, alpha_dev=None, y=None, y_bf16=None):
        self._launches += 1
        al = alpha * (float(alpha_dev.item()) if alpha_dev is not None else 1.0)
        r = al * b if a is None else a + al * b
        if y is not None:
            y.copy_(r)
        if y_bf16 is not None:
            y_bf16.copy_(r.reshape(y_bf16.shape))

    def axpy_rows(self, y, x):
        self._launches += 1
        y.add_(x)

    def act_grad(self, dy, aux, act, out_bf16, scale=1.0, scale_dev=None):
        self._launches += 1
        v = dy.float()
        if act == ACT_GELU_BWD:
            x = aux.float()
            v = v * (0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi))
        elif act == ACT_RELU_BWD:
            v = v * (aux.float() > 0)
        elif act == ACT_TANH_BWD:
            v = v * (1 - aux.float() ** 2)
        else:
            assert act == ACT_NONE
        v = v * scale * (float(scale_dev.item()) if scale_dev is not None else 1.0)
        out_bf16.copy_(v.reshape(out_bf16.shape))

    def zero(self, p):
        self._launches += 1
        p.zero_()

    def colsum(self, x, out, accumulate=False, scale=1.0, scale_dev=None):
        self._launches += 1
        s = x.float().sum(0) * scale * (float(scale_dev.item()) if scale_dev is not None else 1.0)
        if accumulate:
            out.add_(s.reshape(out.shape))
        else:
            out.copy_(s.reshape(out.shape))

    def dot(self, a, b, out, accumulate=False):
        self._launches += 1
        s = (a.float().reshape(-1) * b.float().reshape(-1)).sum()
        if accumulate:
            out.add_(s)
        else:
            out.fill_(s)

    # ------------------------------------------------------------------ embeddings
    def patchify(self, video, p, out):
        self._launches += 1
        BT, Cin, H, W = video.shape
        gh, gw = H // p, W // p
        x = video.reshape(BT, Cin, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(BT * gh * gw, Cin * p * p)
        out.copy_(x.reshape(out.shape))

    def patchify_u8(self, video, p, out, mean, std):
        # the reference's host pipeline: frames.float() / 255 (base_dataset.py:248), NormalizeVideo (transforms.py:49)
        # evaluated on the CPU as a 256-entry table per channel (torch's CUDA `x / 255` multiplies by a rounded reciprocal)
        m = torch.tensor(mean, dtype=torch.float32).view(-1, 1)
        s = torch.tensor(std, dtype=torch.float32).view(-1, 1)
        lut = ((torch.arange(256, dtype=torch.float32).view(1, -1) / 255 - m) / s).to(video.device)     # [Cin, 256]
        ch = torch.arange(video.shape[1], device=video.device).view(1, -1, 1, 1).expand_as(video)
        self.patchify(lut[ch, video.long()], p, out)

    def assemble_tokens(self, patch, cls, pos, temporal, B, T, Nf, tokens):
        self._launches += 1
        Cd = tokens.shape[-1]
        pos2 = pos.reshape(1 + Nf, Cd)
        tem = temporal.reshape(-1, Cd)[:T]
        x = patch.reshape(B, T, Nf, Cd) + pos2[1:][None, None] + tem[None, :, None]
        c = (cls.reshape(1, 1, Cd) + pos2[0].reshape(1, 1, Cd)).expand(B, 1, Cd)
        tokens.copy_(torch.cat([c, x.reshape(B, T * Nf, Cd)], 1).reshape(tokens.shape))

    def assemble_tokens_bwd(self, d_tokens, B, T, Nf, d_patch_bf16=None, d_cls=None, d_pos=None, d_temporal=None):
        self._launches += 1
        Cd = d_tokens.shape[-1]
        d = d_tokens.reshape(B, 1 + T * Nf, Cd)
        dp = d[:, 1:].reshape(B, T, Nf, Cd)
        if d_patch_bf16 is not None:
            d_patch_bf16.copy_(dp.reshape(d_patch_bf16.shape))
        if d_cls is not None:
            d_cls.add_(d[:, 0].sum(0).reshape(d_cls.shape))
        if d_pos is not None:
            d_pos.reshape(1 + Nf, Cd)[0] += d[:, 0].sum(0)
            d_pos.reshape(1 + Nf, Cd)[1:] += dp.sum((0, 1))
        if d_temporal is not None:
            d_temporal.reshape(-1, Cd)[:T] += dp.sum((0, 2))

    @staticmethod
    def _pos_ids(ids, pad_id):
        nonpad = ids.ne(pad_id).to(torch.int64)
        return torch.cumsum(nonpad, 1) * nonpad + pad_id

    def text_embed(self, ids, word, pos, type0, out, pad_id=1):
        self._launches += 1
        out.copy_((word[ids] + pos[self._pos_ids(ids, pad_id)] + type0.reshape(1, 1, -1)).reshape(out.shape))

    def text_embed_bwd(self, d_out, ids, d_word=None, d_pos=None, d_type0=None, pad_id=1):
        self._launches += 1
        Cd = d_out.shape[-1]
        d = d_out.reshape(-1, Cd)
        if d_word is not None:
            d_word.index_add_(0, ids.reshape(-1), d)
        if d_pos is not None:
            d_pos.index_add_(0, self._pos_ids(ids, pad_id).reshape(-1), d)
        if d_type0 is not None:
            d_type0.add_(d.sum(0).reshape(d_type0.shape))

    # ------------------------------------------------------------------ losses
    def softmax_xent(self, logits, labels, V, loss_sum, count, dlogits=None, ignore_index=-100):
        self._launches += 1
        x = logits[:, :V].float()
        valid = labels != ignore_index
        lab = labels.clamp_min(0)
        lse = torch.logsumexp(x, -1)
        per = (lse - x.gather(1, lab[:, None])[:, 0]) * valid
        loss_sum.add_(per.sum())
        count.add_(valid.sum().float())
        if dlogits is not None:
            p = torch.softmax(x, -1)
            p[torch.arange(x.shape[0], device=x.device), lab] -= 1.0
            p = p * valid[:, None]
            dlogits[:, :V] = p.to(dlogits.dtype)

    def xent_finalize(self, loss_sum, count, loss=None, inv_count=None):
        self._launches += 1
        c = count.clamp_min(1.0)
        if loss is not None:
            loss.copy_((loss_sum / c).reshape(loss.shape))
        if inv_count is not None:
            inv_count.copy_((1.0 / c).reshape(inv_count.shape))

    def egonce(self, t, v, noun, verb, temperature, sim, mask, loss, grad_row0=0, grad_rows=0, dt=None, dv=None):
        self._launches += 9
        tt = t.detach().clone().requires_grad_(True)
        vv = v.detach().clone().requires_grad_(True)

        def sm(a, b):
            an = a / a.norm(dim=1, keepdim=True).clamp_min(1e-8)
            bn = b / b.norm(dim=1, keepdim=True).clamp_min(1e-8)
            return an @ bn.t()

        with torch.enable_grad():
            s = sm(tt, vv)
            m = (sm(verb, verb) * sm(noun, noun) + torch.eye(t.shape[0], device=t.device)) > 0
            i_sm = torch.softmax(s / temperature, 1)
            j_sm = torch.softmax(s.t() / temperature, 1)
            L = -torch.log((i_sm * m).sum(1)).mean() - torch.log((j_sm * m).sum(1)).mean()
            L.backward()
        sim.copy_(s.detach())
        mask.copy_(m.to(torch.uint8))
        loss.copy_(L.detach().reshape(loss.shape))
        if dt is not None:
            dt.copy_(tt.grad[grad_row0:grad_row0 + grad_rows])
        if dv is not None:
            dv.copy_(vv.grad[grad_row0:grad_row0 + grad_rows])

    def dual_loss(self, t, v, kind, param, sim, loss, weight=None, fix_norm=True, grad_row0=0, grad_rows=0, dt=None, dv=None):
        # restatement of model_epic_charades.py:542-550 (sim_matrix) + loss.py:13-31 / 65-100 / 102-143
        self._launches += 5
        tt = t.detach().clone().requires_grad_(True)
        vv = v.detach().clone().requires_grad_(True)
        G = t.shape[0]
        with torch.enable_grad():
            an = tt / tt.norm(dim=1, keepdim=True).clamp_min(1e-8)
            bn = vv / vv.norm(dim=1, keepdim=True).clamp_min(1e-8)
            x = an @ bn.t()
            if kind == 0:
                L = -torch.log_softmax(x / param, 1).diag().mean() - torch.log_softmax(x.t() / param, 1).diag().mean()
            else:
                w = weight.reshape(G, 1) if kind == 2 else torch.ones(G, 1, device=t.device)
                d = x.diag().reshape(G, 1)
                terms = torch.stack([torch.relu(w * param - (d - x)), torch.relu(w * param - (d - x.t()))])
                if fix_norm:
                    off = ~torch.eye(G, dtype=torch.bool, device=t.device)
                    L = terms[:, off].mean()
                else:
                    L = terms.mean()
            L.backward()
        sim.copy_(x.detach())
        loss.copy_(L.detach().reshape(loss.shape))
        if dt is not None:
            dt.copy_(tt.grad[grad_row0:grad_row0 + grad_rows])
        if dv is not None:
            dv.copy_(vv.grad[grad_row0:grad_row0 + grad_rows])

    # ------------------------------------------------------------------ re-associated cross-attention (round-2 kernels)
    def xattn_scores_softmax(self, ln, Mt, c0, mask, P):
        """P[b,n,h,:] = softmax_s(ln[b,n,:] . Mt[h,b,s,:] + c0[b,s,h] + mask[b,s])"""
        self._launches += 1
        s = torch.einsum("bnc,hbsc->bnhs", ln.float(), Mt.float()) + c0.float().permute(0, 2, 1)[:, None] + mask.float()[:, None, None, :]
        P.copy_(torch.softmax(s, dim=-1))

    def xattn_weighted_sum(self, P, U, out):
        """out[b,n,:] = sum_{h,s} P[b,n,h,s] U[h,b,s,:]"""
        self._launches += 1
        out.copy_(torch.einsum("bnhs,hbsc->bnc", P.float(), U.float()))

    def xattn_dscores(self, dc, U, P, dS, dbias):
        """dP = dc . U^T; dS = P * (dP - sum_s dP * P); dbias[b,s,h] = sum_n dS[b,n,h,s]"""
        self._launches += 1
        p = P.float()
        dP = torch.einsum("bnc,hbsc->bnhs", dc.float(), U.float())
        ds = p * (dP - (dP * p).sum(-1, keepdim=True))
        dS.copy_(ds)
        dbias.copy_(ds.sum(1).permute(0, 2, 1))

    def xattn_tn(self, X, Y, out):
        """out[h,b,s,:] = sum_n X[b,n,h,s] Y[b,n,:]"""
        self._launches += 1
        out.copy_(torch.einsum("bnhs,bnc->hbsc", X.float(), Y.float()))

    def xattn_t2i_flash(self, Qp, x, Z, lse):
        """P[h,b,s,:] = softmax_n(Qp[h,b,s,:] . x[b,n,:]); Z = P x; lse = logsumexp of the scores"""
        self._launches += 1
        sc = torch.einsum("hbsc,bnc->hbsn", Qp.float(), x.float())
        lse.copy_(torch.logsumexp(sc, dim=-1))
        Z.copy_(torch.einsum("hbsn,bnc->hbsc", torch.softmax(sc, dim=-1), x.float()))

    def xattn_t2i_flash_bwd(self, dZ, Qp, x, Z, lse, dQp, dx):
        """dP = dZ x^T; dS = P * (dP - rowsum(dZ * Z)); dQp = dS x; dx = P^T dZ + dS^T Qp"""
        self._launches += 1
        xf, qf, dz = x.float(), Qp.float(), dZ.float()
        p = torch.exp(torch.einsum("hbsc,bnc->hbsn", qf, xf) - lse.float()[..., None])
        dP = torch.einsum("hbsc,bnc->hbsn", dz, xf)
        dS = p * (dP - (dz * Z.float()).sum(-1, keepdim=True))
        dQp.copy_(torch.einsum("hbsn,bnc->hbsc", dS, xf))
        dx.copy_(torch.einsum("hbsn,hbsc->bnc", p, dz) + torch.einsum("hbsn,hbsc->bnc", dS, qf))

    # ------------------------------------------------------------------ optimiser
    def adamw(self, p, g, m, v, p_bf16, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, hyper_dev=None):
        self._launches += 1
        if hyper_dev is not None:
            lr = lr * float(hyper_dev[0])
            c1, c2 = float(hyper_dev[1]), float(hyper_dev[2])
        else:
            c1, c2 = 1 - beta1 ** step, 1 - beta2 ** step
        gi = g * grad_scale
        m.mul_(beta1).add_(gi, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(gi, gi, value=1 - beta2)
        step_size = lr * math.sqrt(c2) / c1
        p.addcdiv_(m, v.sqrt().add_(eps), value=-step_size)
        p.add_(p, alpha=-lr * weight_decay)
        if p_bf16 is not None:
            p_bf16.copy_(p)
