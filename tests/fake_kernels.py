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

    def mark(self, region, alg=0.0):
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
        if out.shape[-1] == Cin * p * p:
            out.copy_(x.reshape(out.shape))
        else:   # padded im2col depth (patch sizes that are not multiples of 8): zero tail columns
            out.zero_()
            out[:, :Cin * p * p] = x.to(out.dtype)

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

    # ------------------------------------------------------------------ batched GEMM / re-associated cross-attention
    @staticmethod
    def _bview(v, rows, cols, z0, z1):
        t = v.t
        return t.as_strided((rows, cols), (v.ld, 1), t.storage_offset() + z0 * v.s0 + z1 * v.s1)

    def bgemm(self, layout, M, N, K, A, B, nb=(1, 1), bias=None, scale=1.0, scale_dev=None, residual=None, out_f32=None,
              out_bf16=None, aux=None, colsum=None, dot_out=None, accumulate=False, split_k=1, epilogue=0):
        """restatement of egv_bgemm_bf16 (include/egovlp_b200.h): one small matmul per batch"""
        self._launches += 1
        sc = scale * (float(scale_dev.item()) if scale_dev is not None else 1.0)
        for z1 in range(nb[1]):
            for z0 in range(nb[0]):
                if layout == GEMM_NT:
                    v = self._bview(A, M, K, z0, z1).float() @ self._bview(B, N, K, z0, z1).float().t()
                elif layout == GEMM_NN:
                    v = self._bview(A, M, K, z0, z1).float() @ self._bview(B, K, N, z0, z1).float()
                else:
                    v = self._bview(A, K, M, z0, z1).float().t() @ self._bview(B, K, N, z0, z1).float()
                if bias is not None:
                    bt = bias.t
                    v = v + bt.as_strided((N,), (1,), bt.storage_offset() + z0 * bias.s0 + z1 * bias.s1)
                if epilogue == 1:
                    v = torch.softmax(v.reshape(M, N // 32, 32), -1).reshape(M, N)
                elif epilogue == 2:
                    pr = self._bview(aux, M, N, z0, z1).float().reshape(M, N // 32, 32)
                    dp = v.reshape(M, N // 32, 32)
                    t = (pr * dp).sum(-1, keepdim=True)
                    if dot_out is not None:
                        dot_out.add_(t.sum())
                    v = (pr * (dp - t)).reshape(M, N)
                v = v * sc
                if residual is not None:
                    v = v + self._bview(residual, M, N, z0, z1)
                if out_f32 is not None:
                    o = self._bview(out_f32, M, N, z0, z1)
                    if accumulate:
                        o.add_(v)
                    else:
                        o.copy_(v)
                if out_bf16 is not None:
                    self._bview(out_bf16, M, N, z0, z1).copy_(v)
                if colsum is not None:
                    ct = colsum.t
                    ct.as_strided((N,), (1,), ct.storage_offset() + z0 * colsum.s0 + z1 * colsum.s1).add_(v.sum(0))

    @staticmethod
    def philox_keep(seed, idx, p_drop):
        """Philox4x32-10 keep decisions for the int64 element indices `idx` (any shape): the stream of csrc/xattn.cu"""
        import numpy as np
        i = idx.cpu().numpy().astype(np.uint64)
        ctr = i >> np.uint64(2)
        m32 = np.uint64(0xFFFFFFFF)
        c0, c1 = ctr & m32, ctr >> np.uint64(32)
        c2, c3 = np.zeros_like(c0), np.zeros_like(c0)
        k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
        for _ in range(10):
            p0, p1 = np.uint64(0xD2511F53) * c0, np.uint64(0xCD9E8D57) * c2
            hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & m32, p1 >> np.uint64(32), p1 & m32
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m32, (k1 + np.uint64(0xBB67AE85)) & m32
        lane = i & np.uint64(3)
        w = np.where(lane == 0, c0, np.where(lane == 1, c1, np.where(lane == 2, c2, c3)))
        u = (w >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
        return torch.from_numpy(u >= np.float32(p_drop)).to(idx.device)

    @staticmethod
    def philox_key(seed_dev, site):
        """key of a dropout site (csrc/philox.cuh::philox_key)"""
        base = 0 if seed_dev is None else int(seed_dev.item()) & 0xFFFFFFFFFFFFFFFF
        return (base + int(site) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF

    def _keep_rows(self, rows, n, p_drop, seed_dev, site, device):
        idx = torch.arange(rows, dtype=torch.int64, device=device)[:, None] * n + torch.arange(n, dtype=torch.int64, device=device)[None]
        return self.philox_keep(self.philox_key(seed_dev, site), idx, p_drop)

    def keep_mask(self, shape, p_drop, seed_dev, site, device):
        """keep decisions of a dropout site over a tensor of `shape` (row-major element index)"""
        n = 1
        for d in shape:
            n *= d
        return self.philox_keep(self.philox_key(seed_dev, site), torch.arange(n, dtype=torch.int64, device=device), p_drop).reshape(shape)

    def rng_advance(self, state):
        self._launches += 1
        v = (int(state.item()) + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        state.fill_(v - (1 << 64) if v >= (1 << 63) else v)

    def dropout_add(self, x, res, p_drop, seed_dev, site, out_f32=None, out_bf16=None, scale=1.0, scale_dev=None):
        self._launches += 1
        y = x.float() * self.keep_mask(x.shape, p_drop, seed_dev, site, x.device) / (1.0 - p_drop)
        sc = scale * (float(scale_dev.item()) if scale_dev is not None else 1.0)
        if out_f32 is not None:
            out_f32.copy_((0 if res is None else res) + sc * y)
        if out_bf16 is not None:
            out_bf16.copy_(y)

    def dropout_bwd(self, dy, p_drop, seed_dev, site, out_f32=None, out_bf16=None):
        self._launches += 1
        y = dy.float() * self.keep_mask(dy.shape, p_drop, seed_dev, site, dy.device) / (1.0 - p_drop)
        if out_f32 is not None:
            out_f32.copy_(y)
        if out_bf16 is not None:
            out_bf16.copy_(y)

    def _text_attn(self, q, k, v, key_bias, scale, p_drop, seed_dev, site, H):
        B, S, _ = q.shape
        qh, kh, vh = (t.float().reshape(B, S, H, 64).permute(0, 2, 1, 3) for t in (q, k, v))
        sc = scale * (qh @ kh.transpose(-1, -2))
        if key_bias is not None:
            sc = sc + key_bias.reshape(B, 1, 1, S).clamp_min(-3.0e38)
        pr = torch.softmax(sc, -1)
        keep = None
        if p_drop > 0:
            keep = self.keep_mask((B, H, S, S), p_drop, seed_dev, site, q.device) / (1.0 - p_drop)
        return qh, kh, vh, sc, pr, keep

    def text_attention_fwd(self, q, k, v, key_bias, scale, p_drop, seed_dev, site, H, o, lse):
        self._launches += 1
        B, S, _ = q.shape
        qh, kh, vh, sc, pr, keep = self._text_attn(q, k, v, key_bias, scale, p_drop, seed_dev, site, H)
        lse.copy_(torch.logsumexp(sc, -1).reshape(lse.shape))
        pd = pr if keep is None else pr * keep
        o.copy_((pd @ vh).permute(0, 2, 1, 3).reshape(B, S, H * 64))

    def text_attention_bwd(self, q, k, v, key_bias, scale, p_drop, seed_dev, site, H, lse, d_o, dq, dk, dv):
        self._launches += 1
        B, S, _ = q.shape
        qh, kh, vh, sc, pr, keep = self._text_attn(q, k, v, key_bias, scale, p_drop, seed_dev, site, H)
        g = d_o.float().reshape(B, S, H, 64).permute(0, 2, 1, 3)
        pd = pr if keep is None else pr * keep
        dpd = g @ vh.transpose(-1, -2)
        dp = dpd if keep is None else dpd * keep
        ds = pr * (dp - (pr * dp).sum(-1, keepdim=True)) * scale
        back = lambda t: t.permute(0, 2, 1, 3).reshape(B, S, H * 64)  # noqa: E731
        dq.copy_(back(ds @ kh))
        dk.copy_(back(ds.transpose(-1, -2) @ qh))
        dv.copy_(back(pd.transpose(-1, -2) @ g))

    def xattn_row_softmax(self, scores, ld_s, rows, rows_per_batch, s_bstride, n, P, ld_p, p_bstride, lse, p_drop=0.0,
                          seed_dev=None, site=0, rsum=None):
        self._launches += 1
        nb = rows // rows_per_batch
        sv = scores.as_strided((nb, rows_per_batch, n), (s_bstride, ld_s, 1), scores.storage_offset()).float()
        lse.reshape(-1)[:rows].copy_(torch.logsumexp(sv, -1).reshape(-1))
        pr = torch.softmax(sv, -1)
        if p_drop > 0:
            keep = self._keep_rows(rows, n, p_drop, seed_dev, site, scores.device).reshape(nb, rows_per_batch, n)
            pr = pr * keep / (1.0 - p_drop)
        P.as_strided((nb, rows_per_batch, n), (p_bstride, ld_p, 1), P.storage_offset()).copy_(pr)
        if rsum is not None:
            rsum.reshape(-1)[:rows].copy_((pr.sum(-1) if p_drop > 0 else torch.ones_like(pr[..., 0])).reshape(-1))

    def xattn_row_dsoftmax(self, scores, ld_s, rows, rows_per_batch, s_bstride, n, lse, dP, ld_dp, dp_bstride, dS, ld_ds,
                           ds_bstride, p_drop=0.0, seed_dev=None, site=0, row_const=None):
        self._launches += 1
        nb = rows // rows_per_batch
        sv = scores.as_strided((nb, rows_per_batch, n), (s_bstride, ld_s, 1), scores.storage_offset()).float()
        pr = torch.exp(sv - lse.reshape(-1)[:rows].reshape(nb, rows_per_batch, 1))
        g = dP.as_strided((nb, rows_per_batch, n), (dp_bstride, ld_dp, 1), dP.storage_offset()).float()
        if row_const is not None:
            g = g + row_const.reshape(nb, rows_per_batch, 1)
        if p_drop > 0:
            keep = self._keep_rows(rows, n, p_drop, seed_dev, site, scores.device).reshape(nb, rows_per_batch, n)
            g = g * keep / (1.0 - p_drop)
        ds = pr * (g - (pr * g).sum(-1, keepdim=True))
        dS.as_strided((nb, rows_per_batch, n), (ds_bstride, ld_ds, 1), dS.storage_offset()).copy_(ds)

    def xattn_rowscale_bias(self, ox, rsum, bv, B, S, H):
        self._launches += 1
        r = rsum.reshape(B, H, S).permute(0, 2, 1).reshape(B * S, H, 1)
        ox.copy_((ox.float().reshape(B * S, H, 64) + r * bv.reshape(1, H, 64)).reshape(ox.shape))

    def xattn_rowscale_bias_bwd(self, d_ox, rsum, dbv, B, S, H):
        self._launches += 1
        r = rsum.reshape(B, H, S).permute(0, 2, 1).reshape(B * S, H, 1)
        dbv.add_((r * d_ox.float().reshape(B * S, H, 64)).sum(0).reshape(-1))

    def xattn_qbias_fwd(self, k, ldk, bq, mask, scale, B, S, H, out):
        self._launches += 1
        kk = k.as_strided((B * S, H, 64), (ldk, 64, 1), k.storage_offset()).float()
        v = scale * (kk * bq.reshape(1, H, 64)).sum(-1)            # [B*S, H]
        v = v.reshape(B, S, H).permute(0, 2, 1)
        if mask is not None:
            v = v + mask.reshape(B, 1, S)
        out.reshape(B, H, S).copy_(v)

    def xattn_qbias_bwd(self, k, ldk, bq, dbias, scale, B, S, H, dk=None, lddk=0, dbq=None):
        self._launches += 1
        g = scale * dbias.reshape(B, H, S).permute(0, 2, 1).reshape(B * S, H, 1)      # [B*S, H, 1]
        if dk is not None:
            dk.as_strided((B * S, H, 64), (lddk, 64, 1), dk.storage_offset()).add_(g * bq.reshape(1, H, 64))
        if dbq is not None:
            kk = k.as_strided((B * S, H, 64), (ldk, 64, 1), k.storage_offset()).float()
            dbq.reshape(H, 64).add_((g * kk).sum(0))

    # ------------------------------------------------------------------ optimiser
    def adamw_schedule(self, step_dev, hyper_dev, warmup_steps, max_steps, beta1, beta2):
        self._launches += 1
        s = int(step_dev.item())
        scale = 1.0
        if max_steps and max_steps > 0:
            if s < warmup_steps:
                scale = s / max(1, warmup_steps)
            else:
                prog = (s - warmup_steps) / max(1, max_steps - warmup_steps)
                scale = max(0.0, 0.5 * (1.0 + math.cos(math.pi * min(1.0, prog))))
        s += 1
        step_dev.fill_(s)
        hyper_dev.copy_(torch.tensor([scale, 1.0 - beta1 ** s, 1.0 - beta2 ** s], dtype=torch.float32))

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
