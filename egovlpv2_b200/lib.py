"""ctypes binding of libegovlp_b200.so (include/egovlp_b200.h) over torch tensors.

torch is used here only for device memory and the current CUDA stream; every method validates its
tensors, passes raw device pointers through the C ABI and raises on a non-zero return code.  There is
no CPU fallback: if the library cannot be loaded the import of `Kernels` users fails loudly.
"""
import ctypes as C
import dataclasses
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libegovlp_b200.so")

GEMM_NT, GEMM_NN, GEMM_TN = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH, ACT_GELU_BWD, ACT_RELU_BWD, ACT_TANH_BWD, ACT_GELU_DG, ACT_MUL_AUX = range(9)
DUAL_NORM_SOFTMAX, DUAL_MAX_MARGIN, DUAL_ADAPTIVE_MAX_MARGIN = range(3)

c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_float


class GemmArgs(C.Structure):
    _fields_ = [("layout", c_int), ("M", c_int), ("N", c_int), ("K", c_int),
                ("A", c_void_p), ("lda", c_int64), ("B", c_void_p), ("ldb", c_int64),
                ("bias", c_void_p), ("aux", c_void_p), ("ld_aux", c_int64),
                ("scale_dev", c_void_p), ("scale", c_float),
                ("residual", c_void_p), ("ld_res", c_int64),
                ("out_f32", c_void_p), ("ld_out_f32", c_int64),
                ("out_bf16", c_void_p), ("ld_out_bf16", c_int64),
                ("out_pre_bf16", c_void_p), ("ld_out_pre", c_int64),
                ("act", c_int), ("accumulate", c_int), ("split_k", c_int), ("colsum", c_void_p)]


class BGemmArgs(C.Structure):
    _fields_ = [("layout", c_int), ("M", c_int), ("N", c_int), ("K", c_int), ("nb0", c_int), ("nb1", c_int),
                ("A", c_void_p), ("lda", c_int64), ("sa0", c_int64), ("sa1", c_int64),
                ("B", c_void_p), ("ldb", c_int64), ("sb0", c_int64), ("sb1", c_int64),
                ("bias", c_void_p), ("sbias0", c_int64), ("sbias1", c_int64),
                ("scale", c_float), ("scale_dev", c_void_p),
                ("residual", c_void_p), ("ld_res", c_int64), ("sres0", c_int64), ("sres1", c_int64),
                ("out_f32", c_void_p), ("ld_out_f32", c_int64), ("so32_0", c_int64), ("so32_1", c_int64),
                ("out_bf16", c_void_p), ("ld_out_bf16", c_int64), ("so16_0", c_int64), ("so16_1", c_int64),
                ("aux", c_void_p), ("ld_aux", c_int64), ("saux0", c_int64), ("saux1", c_int64),
                ("colsum", c_void_p), ("scol0", c_int64), ("scol1", c_int64),
                ("dot_out", c_void_p),
                ("accumulate", c_int), ("split_k", c_int), ("epilogue", c_int)]


class TextAttnArgs(C.Structure):
    _fields_ = [("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("ld", c_int64), ("key_bias", c_void_p),
                ("scale", c_float), ("p_drop", c_float), ("seed_dev", c_void_p), ("site", C.c_uint64),
                ("B", c_int), ("H", c_int), ("S", c_int), ("o", c_void_p), ("ldo", c_int64), ("lse", c_void_p),
                ("d_o", c_void_p), ("dq", c_void_p), ("dk", c_void_p), ("dv", c_void_p), ("ldd", c_int64)]


class BV:
    """Batched view of a tensor for `Kernels.bgemm`: base tensor `t` (its data_ptr is element (z0=0, z1=0, row 0, col 0)),
    row stride `ld` and the two batch strides `s0`, `s1` in elements (0 = shared by that batch level)."""
    __slots__ = ("t", "ld", "s0", "s1")

    def __init__(self, t, ld=None, s0=0, s1=0):
        self.t, self.ld, self.s0, self.s1 = t, (t.stride(-2) if ld is None else ld), s0, s1


EPI_NONE, EPI_SOFTMAX32, EPI_DSOFTMAX32 = 0, 1, 2


class AttnArgs(C.Structure):
    _fields_ = [("B", c_int), ("H", c_int), ("G", c_int), ("Lq", c_int), ("Lk", c_int),
                ("q", c_void_p), ("ldq", c_int64), ("q_bstride", c_int64),
                ("q_row0", c_int), ("q_gstride", c_int), ("q_istride", c_int),
                ("k", c_void_p), ("v", c_void_p), ("ldkv", c_int64), ("kv_bstride", c_int64),
                ("k_row0", c_int), ("k_gstride", c_int), ("k_istride", c_int),
                ("has_cls_key", c_int), ("cls_row", c_int),
                ("key_bias", c_void_p), ("scale", c_float),
                ("o", c_void_p), ("ldo", c_int64), ("o_bstride", c_int64),
                ("lse", c_void_p),
                ("d_o", c_void_p), ("dq", c_void_p), ("lddq", c_int64),
                ("dk", c_void_p), ("dv", c_void_p), ("lddkv", c_int64),
                ("delta", c_void_p), ("dkv_cls", c_void_p), ("dkv_accumulate", c_int),
                ("workspace", c_void_p), ("workspace_bytes", c_int64),
                ("lse_cls", c_void_p), ("dq_cls", c_void_p), ("cls_query_folded", C.POINTER(c_int))]


@dataclasses.dataclass(frozen=True)
class AttnSpec:
    """Group addressing of one attention call (see egv_attn_args in include/egovlp_b200.h)."""
    H: int
    G: int
    Lq: int
    Lk: int                 # regular keys per group (the shared CLS key is extra)
    q_row0: int = 0
    q_gstride: int = 0
    q_istride: int = 1
    k_row0: int = 0
    k_gstride: int = 0
    k_istride: int = 1
    has_cls_key: bool = False
    cls_row: int = 0
    scale: float = 1.0


def _load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "egovlpv2_b200: %s is missing -- build it with `python -m egovlpv2_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.egv_last_error.restype = C.c_char_p
    lib.egv_launch_count.restype = C.c_longlong
    lib.egv_egonce_scratch_floats.restype = c_int64
    lib.egv_dual_loss_scratch_floats.restype = c_int64
    lib.egv_attention_workspace_bytes.restype = c_int64
    return lib


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk2d(t, dtype, name):
    if t.dim() != 2 or t.stride(1) != 1 or t.dtype != dtype or not t.is_cuda:
        raise ValueError("%s: expected 2-D %s CUDA tensor with unit inner stride, got %s %s %s" %
                         (name, dtype, tuple(t.shape), t.stride(), t.dtype))
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


class Kernels:
    """The device kernels of the hot path.  One instance per process (CUDA context of the current device)."""

    def __init__(self):
        self.lib = _load()
        self.profile = None      # list of (region, kind, work, start_event, stop_event) when profiling is on
        self.alg = {}            # region -> algorithmic FLOPs of the reference formulation (see mark)
        self._region = "other"

    # ------------------------------------------------------------------ per-launch timing (bench.py roofline)
    def mark(self, region, alg=0.0):
        """Label the kernels launched from here on (cheap; only read while profiling).  `alg`: ALGORITHMIC FLOPs of the
        reference formulation of the region being entered, summed per region while profiling."""
        self._region = region
        if self.profile is not None and alg:
            self.alg[region] = self.alg.get(region, 0.0) + alg

    def start_profile(self):
        self.profile = []
        self.alg = {}

    def stop_profile(self):
        """-> {(region, kind): (seconds, work, launches)}; call after a stream synchronize."""
        out = {}
        for region, kind, work, e0, e1 in self.profile:
            t, w, n = out.get((region, kind), (0.0, 0.0, 0))
            out[(region, kind)] = (t + e0.elapsed_time(e1) * 1e-3, w + work, n + 1)
        self.profile = None
        return out

    def _timed(self, kind, work, fn):
        if self.profile is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        self.profile.append((self._region, kind, work, e0, e1))
        return r

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("egovlp_b200 kernel error %d: %s" % (rc, self.lib.egv_last_error().decode()))

    def launch_count(self):
        return int(self.lib.egv_launch_count())

    def sm_count(self):
        return int(self.lib.egv_device_sm_count())

    def force_simt(self, on):
        self.lib.egv_gemm_force_simt(int(bool(on)))

    def set_cluster(self, mode):
        """0 / False: single CTAs; 1 / True: CTA pairs wherever possible; 2: pairs for the TN layout only (default)."""
        self.lib.egv_gemm_set_cluster(int(mode))

    def set_attention_tiny(self, mode):
        self.lib.egv_attention_set_tiny(int(mode))

    def set_attention_tc(self, mode):
        self.lib.egv_attention_set_tc(int(mode))

    def set_plan(self, mode):
        self.lib.egv_gemm_set_plan(int(mode))

    # ------------------------------------------------------------------ GEMM
    def gemm(self, layout, A, B, *, bias=None, aux=None, act=ACT_NONE, scale=1.0, scale_dev=None,
             residual=None, out_f32=None, out_bf16=None, out_pre=None, accumulate=False, split_k=1, colsum=None):
        lda = _chk2d(A, torch.bfloat16, "A")
        ldb = _chk2d(B, torch.bfloat16, "B")
        if layout == GEMM_NT:
            (M, K), (N, K2) = A.shape, B.shape
        elif layout == GEMM_NN:
            (M, K), (K2, N) = A.shape, B.shape
        else:
            (K, M), (K2, N) = A.shape, B.shape
        if K != K2:
            raise ValueError("gemm: inner dimensions differ (%d vs %d)" % (K, K2))
        a = GemmArgs()
        a.layout, a.M, a.N, a.K = layout, M, N, K
        a.A, a.lda, a.B, a.ldb = _p(A), lda, _p(B), ldb
        if bias is not None:
            assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
            a.bias = _p(bias)
        for name, t, dt in (("aux", aux, torch.bfloat16), ("residual", residual, torch.float32),
                            ("out_f32", out_f32, torch.float32), ("out_bf16", out_bf16, torch.bfloat16),
                            ("out_pre", out_pre, torch.bfloat16)):
            if t is None:
                continue
            ld = _chk2d(t, dt, name)
            if tuple(t.shape) != (M, N):
                raise ValueError("gemm: %s has shape %s, expected %s" % (name, tuple(t.shape), (M, N)))
            if name == "aux":
                a.aux, a.ld_aux = _p(t), ld
            elif name == "residual":
                a.residual, a.ld_res = _p(t), ld
            elif name == "out_f32":
                a.out_f32, a.ld_out_f32 = _p(t), ld
            elif name == "out_bf16":
                a.out_bf16, a.ld_out_bf16 = _p(t), ld
            else:
                a.out_pre_bf16, a.ld_out_pre = _p(t), ld
        if scale_dev is not None:
            assert scale_dev.dtype == torch.float32 and scale_dev.numel() == 1
            a.scale_dev = _p(scale_dev)
        a.scale = float(scale)
        if colsum is not None:
            assert colsum.dtype == torch.float32 and colsum.numel() == N and colsum.is_contiguous()
            a.colsum = _p(colsum)
        a.act, a.accumulate, a.split_k = int(act), int(bool(accumulate)), int(split_k)
        self._timed("gemm", 2.0 * M * N * K, lambda: self._check(self.lib.egv_gemm_bf16(C.byref(a), self._stream())))

    # ------------------------------------------------------------------ batched GEMM / re-associated cross-attention
    def bgemm(self, layout, M, N, K, A, B, nb=(1, 1), bias=None, scale=1.0, scale_dev=None, residual=None, out_f32=None,
              out_bf16=None, aux=None, colsum=None, dot_out=None, accumulate=False, split_k=1, epilogue=EPI_NONE):
        """D[z] = epilogue(A[z] x B[z]) over the batch grid nb = (nb0, nb1); every tensor argument is a `BV`
        (bias / colsum: BV whose `ld` is ignored).  See egv_bgemm_bf16 in include/egovlp_b200.h."""
        a = BGemmArgs()
        a.layout, a.M, a.N, a.K, a.nb0, a.nb1 = layout, M, N, K, nb[0], nb[1]
        for v, dt in ((A, torch.bfloat16), (B, torch.bfloat16), (bias, torch.float32), (residual, torch.float32),
                      (out_f32, torch.float32), (out_bf16, torch.bfloat16), (aux, torch.bfloat16), (colsum, torch.float32)):
            if v is not None and (v.t.dtype != dt or not v.t.is_cuda):
                raise ValueError("bgemm: expected a CUDA %s tensor, got %s" % (dt, v.t.dtype))
        a.A, a.lda, a.sa0, a.sa1 = _p(A.t), A.ld, A.s0, A.s1
        a.B, a.ldb, a.sb0, a.sb1 = _p(B.t), B.ld, B.s0, B.s1
        if bias is not None:
            a.bias, a.sbias0, a.sbias1 = _p(bias.t), bias.s0, bias.s1
        a.scale = float(scale)
        if scale_dev is not None:
            assert scale_dev.dtype == torch.float32 and scale_dev.numel() == 1
            a.scale_dev = _p(scale_dev)
        if residual is not None:
            a.residual, a.ld_res, a.sres0, a.sres1 = _p(residual.t), residual.ld, residual.s0, residual.s1
        if out_f32 is not None:
            a.out_f32, a.ld_out_f32, a.so32_0, a.so32_1 = _p(out_f32.t), out_f32.ld, out_f32.s0, out_f32.s1
        if out_bf16 is not None:
            a.out_bf16, a.ld_out_bf16, a.so16_0, a.so16_1 = _p(out_bf16.t), out_bf16.ld, out_bf16.s0, out_bf16.s1
        if aux is not None:
            a.aux, a.ld_aux, a.saux0, a.saux1 = _p(aux.t), aux.ld, aux.s0, aux.s1
        if colsum is not None:
            a.colsum, a.scol0, a.scol1 = _p(colsum.t), colsum.s0, colsum.s1
        if dot_out is not None:
            assert dot_out.dtype == torch.float32 and dot_out.numel() == 1
            a.dot_out = _p(dot_out)
        a.accumulate, a.split_k, a.epilogue = int(bool(accumulate)), int(split_k), int(epilogue)
        self._timed("gemm", 2.0 * M * N * K * nb[0] * nb[1], lambda: self._check(self.lib.egv_bgemm_bf16(C.byref(a), self._stream())))

    @staticmethod
    def _seed(seed_dev):
        assert seed_dev is None or (seed_dev.dtype == torch.int64 and seed_dev.numel() == 1 and seed_dev.is_cuda)
        return _p(seed_dev)

    def xattn_row_softmax(self, scores, ld_s, rows, rows_per_batch, s_bstride, n, P, ld_p, p_bstride, lse, p_drop=0.0,
                          seed_dev=None, site=0, rsum=None):
        assert scores.dtype == torch.float32 and P.dtype == torch.bfloat16 and lse.dtype == torch.float32
        self._timed("softmax", 0.0, lambda: self._check(self.lib.egv_xattn_row_softmax(
            _p(scores), c_int64(ld_s), c_int64(rows), rows_per_batch, c_int64(s_bstride), n, _p(P), c_int64(ld_p),
            c_int64(p_bstride), _p(lse), c_float(p_drop), self._seed(seed_dev), C.c_uint64(site), _p(rsum), self._stream())))

    def xattn_row_dsoftmax(self, scores, ld_s, rows, rows_per_batch, s_bstride, n, lse, dP, ld_dp, dp_bstride, dS, ld_ds,
                           ds_bstride, p_drop=0.0, seed_dev=None, site=0, row_const=None):
        assert scores.dtype == torch.float32 and dP.dtype == torch.float32 and dS.dtype == torch.bfloat16
        self._timed("softmax", 0.0, lambda: self._check(self.lib.egv_xattn_row_dsoftmax(
            _p(scores), c_int64(ld_s), c_int64(rows), rows_per_batch, c_int64(s_bstride), n, _p(lse), _p(dP), c_int64(ld_dp),
            c_int64(dp_bstride), _p(dS), c_int64(ld_ds), c_int64(ds_bstride), c_float(p_drop), self._seed(seed_dev),
            C.c_uint64(site), _p(row_const), self._stream())))

    def xattn_rowscale_bias(self, ox, rsum, bv, B, S, H):
        assert ox.dtype == torch.bfloat16 and ox.stride(1) == 1 and rsum.dtype == torch.float32 and bv.dtype == torch.float32
        self._check(self.lib.egv_xattn_rowscale_bias(_p(ox), c_int64(ox.stride(0)), _p(rsum), _p(bv), B, S, H, self._stream()))

    def xattn_rowscale_bias_bwd(self, d_ox, rsum, dbv, B, S, H):
        assert d_ox.dtype == torch.bfloat16 and d_ox.stride(1) == 1 and dbv.dtype == torch.float32
        self._check(self.lib.egv_xattn_rowscale_bias_bwd(_p(d_ox), c_int64(d_ox.stride(0)), _p(rsum), _p(dbv), B, S, H,
                                                         self._stream()))

    def xattn_qbias_fwd(self, k, ldk, bq, mask, scale, B, S, H, out):
        assert k.dtype == torch.bfloat16 and bq.dtype == torch.float32 and out.dtype == torch.float32 and out.is_contiguous()
        assert mask is None or (mask.dtype == torch.float32 and mask.is_contiguous() and mask.numel() == B * S)
        self._check(self.lib.egv_xattn_qbias_fwd(_p(k), c_int64(ldk), _p(bq), _p(mask), c_float(scale), B, S, H, _p(out),
                                                 self._stream()))

    def xattn_qbias_bwd(self, k, ldk, bq, dbias, scale, B, S, H, dk=None, lddk=0, dbq=None):
        assert k.dtype == torch.bfloat16 and dbias.dtype == torch.float32 and dbias.is_contiguous()
        self._check(self.lib.egv_xattn_qbias_bwd(_p(k), c_int64(ldk), _p(bq), _p(dbias), c_float(scale), B, S, H, _p(dk),
                                                 c_int64(lddk), _p(dbq), self._stream()))

    # ------------------------------------------------------------------ train-mode dropout of the text tower
    def rng_advance(self, state):
        """bump the device-resident step seed (int64 [1]); part of the captured step, so every replay draws new masks"""
        assert state.dtype == torch.int64 and state.numel() == 1 and state.is_cuda
        self._check(self.lib.egv_rng_advance(_p(state), self._stream()))

    def dropout_add(self, x, res, p_drop, seed_dev, site, out_f32=None, out_bf16=None, scale=1.0, scale_dev=None):
        """y = dropout(x); out_f32 = res + scale * (*scale_dev) * y; out_bf16 = bf16(y)"""
        assert x.is_contiguous() and x.dtype in (torch.float32, torch.bfloat16)
        for t, dt in ((res, torch.float32), (out_f32, torch.float32), (out_bf16, torch.bfloat16)):
            assert t is None or (t.dtype == dt and t.is_contiguous() and t.numel() == x.numel())
        self._check(self.lib.egv_dropout_add(_p(x), int(x.dtype == torch.bfloat16), _p(res), c_float(scale), _p(scale_dev),
                                             c_float(p_drop), self._seed(seed_dev), C.c_uint64(site), _p(out_f32), _p(out_bf16),
                                             c_int64(x.numel()), self._stream()))

    def dropout_bwd(self, dy, p_drop, seed_dev, site, out_f32=None, out_bf16=None):
        assert dy.is_contiguous() and dy.dtype in (torch.float32, torch.bfloat16)
        for t, dt in ((out_f32, torch.float32), (out_bf16, torch.bfloat16)):
            assert t is None or (t.dtype == dt and t.is_contiguous() and t.numel() == dy.numel())
        self._check(self.lib.egv_dropout_bwd(_p(dy), int(dy.dtype == torch.bfloat16), c_float(p_drop), self._seed(seed_dev),
                                             C.c_uint64(site), _p(out_f32), _p(out_bf16), c_int64(dy.numel()), self._stream()))

    def _text_attn_args(self, q, k, v, key_bias, scale, p_drop, seed_dev, site, H, lse):
        a = TextAttnArgs()
        B, S, _ = q.shape
        ld, bs = self._rows3(q, "q")
        assert bs == S and self._rows3(k, "k") == (ld, bs) and self._rows3(v, "v") == (ld, bs)
        a.q, a.k, a.v, a.ld = _p(q), _p(k), _p(v), ld
        if key_bias is not None:
            assert key_bias.dtype == torch.float32 and key_bias.is_contiguous() and key_bias.numel() == B * S
            a.key_bias = _p(key_bias)
        a.scale, a.p_drop, a.seed_dev, a.site = float(scale), float(p_drop), self._seed(seed_dev), int(site)
        a.B, a.H, a.S = B, H, S
        assert lse.dtype == torch.float32 and lse.numel() == B * H * S
        a.lse = _p(lse)
        return a

    def text_attention_fwd(self, q, k, v, key_bias, scale, p_drop, seed_dev, site, H, o, lse):
        """RobertaSelfAttention core with dropout on the probabilities; q / k / v / o: [B, S, H*64] bf16 views"""
        a = self._text_attn_args(q, k, v, key_bias, scale, p_drop, seed_dev, site, H, lse)
        a.ldo, bs = self._rows3(o, "o")
        assert bs == q.shape[1]
        a.o = _p(o)
        self._check(self.lib.egv_text_attention_fwd(C.byref(a), self._stream()))

    def text_attention_bwd(self, q, k, v, key_bias, scale, p_drop, seed_dev, site, H, lse, d_o, dq, dk, dv):
        a = self._text_attn_args(q, k, v, key_bias, scale, p_drop, seed_dev, site, H, lse)
        a.ldo, bs = self._rows3(d_o, "d_o")
        a.ldd, bs2 = self._rows3(dq, "dq")
        assert bs == q.shape[1] and bs2 == bs and self._rows3(dk, "dk") == (a.ldd, bs) and self._rows3(dv, "dv") == (a.ldd, bs)
        a.d_o, a.dq, a.dk, a.dv = _p(d_o), _p(dq), _p(dk), _p(dv)
        self._check(self.lib.egv_text_attention_bwd(C.byref(a), self._stream()))

    # ------------------------------------------------------------------ LayerNorm
    def layernorm_fwd(self, x, gamma, beta, eps, y_bf16=None, y_f32=None, mean=None, rstd=None):
        Cdim = x.shape[-1]
        rows = x.numel() // Cdim
        assert x.is_contiguous() and x.dtype in (torch.float32, torch.bfloat16)
        for t in (y_bf16, y_f32):
            assert t is None or (t.is_contiguous() and t.numel() == x.numel())
        self._check(self.lib.egv_layernorm_fwd(_p(x), int(x.dtype == torch.bfloat16), _p(gamma), _p(beta), c_float(eps),
                                               c_int64(rows), Cdim, _p(y_bf16), _p(y_f32), _p(mean), _p(rstd),
                                               self._stream()))

    def layernorm_bwd(self, dy, x, gamma, mean, rstd, add=None, dx=None, dx_bf16=None, bf16_total=True, dgamma=None,
                      dbeta=None, out_colsum=None):
        Cdim = x.shape[-1]
        rows = x.numel() // Cdim
        assert dy.is_contiguous() and x.is_contiguous() and dy.numel() == x.numel()
        for t in (add, dx, dx_bf16):
            assert t is None or (t.is_contiguous() and t.numel() == x.numel())
        assert add is None or add.dtype == torch.float32
        self._check(self.lib.egv_layernorm_bwd(_p(dy), int(dy.dtype == torch.bfloat16), _p(x),
                                               int(x.dtype == torch.bfloat16), _p(gamma), _p(mean), _p(rstd),
                                               c_int64(rows), Cdim, _p(add), _p(dx), _p(dx_bf16), int(bool(bf16_total)),
                                               _p(dgamma), _p(dbeta), _p(out_colsum), self._stream()))

    # ------------------------------------------------------------------ attention
    @staticmethod
    def _rows3(t, name):
        """[B, rows, width] view with unit inner stride -> (ld, batch stride in rows)."""
        if t.dim() != 3 or t.stride(2) != 1 or t.dtype != torch.bfloat16:
            raise ValueError("%s: expected [B, rows, W] bf16 view with unit inner stride" % name)
        ld = t.stride(1)
        if t.stride(0) % ld:
            raise ValueError("%s: batch stride is not a whole number of rows" % name)
        return ld, t.stride(0) // ld

    def _attn_args(self, spec, q, k, v, o, lse, key_bias):
        a = AttnArgs()
        a.B, a.H, a.G, a.Lq, a.Lk = q.shape[0], spec.H, spec.G, spec.Lq, spec.Lk
        a.ldq, a.q_bstride = self._rows3(q, "q")
        a.ldkv, a.kv_bstride = self._rows3(k, "k")
        assert self._rows3(v, "v") == (a.ldkv, a.kv_bstride)
        a.ldo, a.o_bstride = self._rows3(o, "o")
        a.q, a.k, a.v, a.o, a.lse = _p(q), _p(k), _p(v), _p(o), _p(lse)
        a.q_row0, a.q_gstride, a.q_istride = spec.q_row0, spec.q_gstride, spec.q_istride
        a.k_row0, a.k_gstride, a.k_istride = spec.k_row0, spec.k_gstride, spec.k_istride
        a.has_cls_key, a.cls_row = int(spec.has_cls_key), spec.cls_row
        a.scale = float(spec.scale)
        assert lse.dtype == torch.float32 and lse.numel() == a.B * spec.H * spec.G * spec.Lq
        if key_bias is not None:
            assert key_bias.dtype == torch.float32 and key_bias.is_contiguous()
            assert key_bias.numel() == a.B * (spec.Lk + int(spec.has_cls_key))
            a.key_bias = _p(key_bias)
        nbytes = int(self.lib.egv_attention_workspace_bytes(C.byref(a)))
        if nbytes > 0:   # split-stream scratch (kept alive by the caller's frame until the launches are enqueued)
            a._ws = torch.empty(nbytes // 4, dtype=torch.float32, device=q.device)
            a.workspace, a.workspace_bytes = _p(a._ws), nbytes
        return a

    def attention_fwd(self, spec, q, k, v, o, lse, key_bias=None, lse_cls=None):
        """lse_cls [B*H] f32 given: the kernel may take the clip's CLS query along (o row cls_row and lse_cls written);
        -> True when it did (the caller then skips the single-query forward)."""
        a = self._attn_args(spec, q, k, v, o, lse, key_bias)
        folded = c_int(0)
        if lse_cls is not None:
            assert lse_cls.dtype == torch.float32 and lse_cls.numel() == a.B * spec.H and lse_cls.is_contiguous()
            a.lse_cls, a.cls_query_folded = _p(lse_cls), C.pointer(folded)
        work = 4.0 * a.B * spec.H * spec.G * spec.Lq * (spec.Lk + int(spec.has_cls_key)) * 64
        self._timed("attn_fwd", work, lambda: self._check(self.lib.egv_attention_fwd(C.byref(a), self._stream())))
        return bool(folded.value)

    supports_cls_fold = True    # attention_bwd(..., lse_cls=, dq_cls=) may take the clip's CLS query along (egv_attn_args)

    def attention_bwd(self, spec, q, k, v, o, lse, d_o, dq, dk, dv, delta, dkv_cls=None, dkv_accumulate=False,
                      key_bias=None, lse_cls=None, dq_cls=None):
        """-> True when the kernel also handled the CLS query (lse_cls [B*H] f32 from the single-query forward, dq_cls
        [B*H*64] f32 zeroed accumulator given): the caller then skips the single-query backward and calls
        attention_cls_query_finalize."""
        a = self._attn_args(spec, q, k, v, o, lse, key_bias)
        assert self._rows3(d_o, "d_o") == (a.ldo, a.o_bstride)
        a.lddq, bq = self._rows3(dq, "dq")
        assert bq == a.q_bstride
        a.lddkv, bk = self._rows3(dk, "dk")
        assert bk == a.kv_bstride and self._rows3(dv, "dv") == (a.lddkv, bk)
        assert delta.dtype == torch.float32 and delta.numel() == lse.numel()
        a.d_o, a.dq, a.dk, a.dv, a.delta = _p(d_o), _p(dq), _p(dk), _p(dv), _p(delta)
        if dkv_cls is not None:
            assert dkv_cls.dtype == torch.float32 and dkv_cls.numel() == a.B * spec.H * 128
            a.dkv_cls = _p(dkv_cls)
        a.dkv_accumulate = int(bool(dkv_accumulate))
        folded = c_int(0)
        if lse_cls is not None and dq_cls is not None:
            assert lse_cls.dtype == torch.float32 and lse_cls.numel() == a.B * spec.H and lse_cls.is_contiguous()
            assert dq_cls.dtype == torch.float32 and dq_cls.numel() == a.B * spec.H * 64 and dq_cls.is_contiguous()
            a.lse_cls, a.dq_cls, a.cls_query_folded = _p(lse_cls), _p(dq_cls), C.pointer(folded)
        work = 8.0 * a.B * spec.H * spec.G * spec.Lq * (spec.Lk + int(spec.has_cls_key)) * 64
        self._timed("attn_bwd", work, lambda: self._check(self.lib.egv_attention_bwd(C.byref(a), self._stream())))
        return bool(folded.value)

    def attention_cls_query_finalize(self, dq_cls, dq, H, cls_row=0):
        ld, bs = self._rows3(dq, "dq")
        self._check(self.lib.egv_attention_cls_query_finalize(_p(dq_cls), _p(dq), c_int64(ld), c_int64(bs), cls_row, dq.shape[0], H,
                                                              self._stream()))

    def attention_cls_finalize(self, dkv_cls, dk, dv, H, cls_row=0, accumulate=False):
        ld, bs = self._rows3(dk, "dk")
        assert self._rows3(dv, "dv") == (ld, bs)
        self._check(self.lib.egv_attention_cls_finalize(_p(dkv_cls), _p(dk), _p(dv), c_int64(ld), c_int64(bs), cls_row,
                                                        dk.shape[0], H, int(bool(accumulate)), self._stream()))

    # ------------------------------------------------------------------ elementwise / reductions
    def cast(self, x, y):
        assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()
        if x.dtype == torch.float32 and y.dtype == torch.bfloat16:
            self._check(self.lib.egv_cast_f32_bf16(_p(x), _p(y), c_int64(x.numel()), self._stream()))
        elif x.dtype == torch.bfloat16 and y.dtype == torch.float32:
            self._check(self.lib.egv_cast_bf16_f32(_p(x), _p(y), c_int64(x.numel()), self._stream()))
        else:
            raise ValueError("cast: unsupported dtypes %s -> %s" % (x.dtype, y.dtype))

    def axpy(self, a, b, alpha=1.0, alpha_dev=None, y=None, y_bf16=None):
        assert b.dtype == torch.float32 and b.is_contiguous()
        for t in (a, y):
            assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() == b.numel())
        assert y_bf16 is None or (y_bf16.dtype == torch.bfloat16 and y_bf16.is_contiguous())
        self._check(self.lib.egv_axpy_f32(_p(a), _p(b), c_float(alpha), _p(alpha_dev), _p(y), _p(y_bf16),
                                          c_int64(b.numel()), self._stream()))

    def axpy_rows(self, y, x):
        """y[r, :] += x[r, :] on 2-D f32 views with unit inner stride (rows may be strided)."""
        assert y.dim() == 2 and x.dim() == 2 and y.shape == x.shape and y.stride(1) == 1 and x.stride(1) == 1
        assert y.dtype == torch.float32 and x.dtype == torch.float32
        self._check(self.lib.egv_add_rows_f32(_p(y), c_int64(y.stride(0)), _p(x), c_int64(x.stride(0)), y.shape[0],
                                              y.shape[1], self._stream()))

    def act_grad(self, dy, aux, act, out_bf16, scale=1.0, scale_dev=None):
        assert dy.is_contiguous() and out_bf16.is_contiguous() and out_bf16.dtype == torch.bfloat16
        assert dy.numel() == out_bf16.numel() and (aux is None or (aux.is_contiguous() and aux.dtype == torch.bfloat16))
        self._check(self.lib.egv_act_grad(_p(dy), int(dy.dtype == torch.bfloat16), _p(aux), int(act), c_float(scale),
                                          _p(scale_dev), _p(out_bf16), c_int64(dy.numel()), self._stream()))

    def zero(self, p):
        assert p.dtype == torch.float32 and p.is_contiguous()
        self._check(self.lib.egv_zero_f32(_p(p), c_int64(p.numel()), self._stream()))

    def colsum(self, x, out, accumulate=False, scale=1.0, scale_dev=None):
        ld = x.stride(0)
        assert x.dim() == 2 and x.stride(1) == 1 and out.dtype == torch.float32 and out.numel() == x.shape[1]
        self._check(self.lib.egv_colsum(_p(x), int(x.dtype == torch.bfloat16), c_int64(x.shape[0]), x.shape[1], c_int64(ld),
                                        _p(out), int(bool(accumulate)), c_float(scale), _p(scale_dev), self._stream()))

    def dot(self, a, b, out, accumulate=False):
        assert a.is_contiguous() and b.is_contiguous() and a.numel() == b.numel() and out.numel() == 1
        self._check(self.lib.egv_dot(_p(a), int(a.dtype == torch.bfloat16), _p(b), int(b.dtype == torch.bfloat16),
                                     c_int64(a.numel()), _p(out), int(bool(accumulate)), self._stream()))

    # ------------------------------------------------------------------ embeddings
    def patchify(self, video, p, out):
        """im2col of [BT, Cin, H, W] fp32 frames into out [BT*gh*gw, ld] bf16 (ld = Cin*p*p, or that rounded up to 8 with
        zero-filled tail columns when the patch size is not a multiple of 8)"""
        BT, Cin, H, W = video.shape
        assert video.dtype == torch.float32 and video.is_contiguous() and out.dtype == torch.bfloat16 and out.is_contiguous()
        Kc = Cin * p * p
        rows = BT * (H // p) * (W // p)
        assert out.dim() == 2 and out.shape[0] == rows and out.shape[1] >= Kc
        if out.shape[1] == Kc and p % 8 == 0:
            self._check(self.lib.egv_patchify(_p(video), BT, Cin, H, W, p, _p(out), self._stream()))
        else:
            self._check(self.lib.egv_patchify_padded(_p(video), BT, Cin, H, W, p, c_int64(out.shape[1]), _p(out), self._stream()))

    def patchify_u8(self, video, p, out, mean, std):
        """uint8 frames -> normalised bf16 patches; mean / std: per-channel sequences (transforms.py:49)."""
        BT, Cin, H, W = video.shape
        assert video.dtype == torch.uint8 and video.is_contiguous() and out.dtype == torch.bfloat16
        assert out.is_contiguous() and out.numel() == video.numel() and len(mean) == Cin and len(std) == Cin
        fa = C.c_float * Cin
        self._check(self.lib.egv_patchify_u8(_p(video), BT, Cin, H, W, p, fa(*[float(m) for m in mean]),
                                             fa(*[float(v) for v in std]), _p(out), self._stream()))

    def assemble_tokens(self, patch, cls, pos, temporal, B, T, Nf, tokens):
        Cd = tokens.shape[-1]
        for t in (patch, cls, pos, temporal, tokens):
            assert t.dtype == torch.float32 and t.is_contiguous()
        assert tokens.numel() == B * (1 + T * Nf) * Cd and patch.numel() == B * T * Nf * Cd
        self._check(self.lib.egv_assemble_tokens(_p(patch), _p(cls), _p(pos), _p(temporal), B, T, Nf, Cd, _p(tokens),
                                                 self._stream()))

    def assemble_tokens_bwd(self, d_tokens, B, T, Nf, d_patch_bf16=None, d_cls=None, d_pos=None, d_temporal=None):
        Cd = d_tokens.shape[-1]
        assert d_tokens.dtype == torch.float32 and d_tokens.is_contiguous()
        self._check(self.lib.egv_assemble_tokens_bwd(_p(d_tokens), B, T, Nf, Cd, _p(d_patch_bf16), _p(d_cls), _p(d_pos),
                                                     _p(d_temporal), self._stream()))

    def text_embed(self, ids, word, pos, type0, out, pad_id=1):
        B, S = ids.shape
        assert ids.dtype == torch.int64 and ids.is_contiguous() and out.is_contiguous()
        self._check(self.lib.egv_text_embed(_p(ids), B, S, out.shape[-1], pad_id, _p(word), _p(pos), _p(type0), _p(out),
                                            self._stream()))

    def text_embed_bwd(self, d_out, ids, d_word=None, d_pos=None, d_type0=None, pad_id=1):
        B, S = ids.shape
        assert d_out.dtype == torch.float32 and d_out.is_contiguous()
        self._check(self.lib.egv_text_embed_bwd(_p(d_out), _p(ids), B, S, d_out.shape[-1], pad_id, _p(d_word), _p(d_pos),
                                                _p(d_type0), self._stream()))

    # ------------------------------------------------------------------ losses
    def softmax_xent(self, logits, labels, V, loss_sum, count, dlogits=None, ignore_index=-100):
        assert logits.dim() == 2 and logits.stride(1) == 1 and logits.dtype == torch.float32
        assert labels.dtype == torch.int64 and labels.is_contiguous() and labels.numel() == logits.shape[0]
        ld_d = 0
        if dlogits is not None:
            assert dlogits.dtype in (torch.bfloat16, torch.float32) and dlogits.stride(1) == 1
            assert dlogits.shape[0] == logits.shape[0]
            ld_d = dlogits.stride(0)
        is_f32 = int(dlogits is not None and dlogits.dtype == torch.float32)
        self._check(self.lib.egv_softmax_xent(_p(logits), c_int64(logits.stride(0)), _p(labels), c_int64(logits.shape[0]),
                                              V, ignore_index, _p(loss_sum), _p(count), _p(dlogits), is_f32, c_int64(ld_d),
                                              self._stream()))

    def xent_finalize(self, loss_sum, count, loss=None, inv_count=None):
        self._check(self.lib.egv_xent_finalize(_p(loss_sum), _p(count), _p(loss), _p(inv_count), self._stream()))

    def egonce(self, t, v, noun, verb, temperature, sim, mask, loss, grad_row0=0, grad_rows=0, dt=None, dv=None):
        G, P = t.shape
        for x in (t, v, noun, verb, sim):
            assert x.dtype == torch.float32 and x.is_contiguous()
        assert mask.dtype == torch.uint8 and mask.numel() == G * G
        n = int(self.lib.egv_egonce_scratch_floats(G, P, noun.shape[1], verb.shape[1]))
        scratch = torch.empty(n, dtype=torch.float32, device=t.device)
        self._check(self.lib.egv_egonce(_p(t), _p(v), G, P, _p(noun), noun.shape[1], _p(verb), verb.shape[1],
                                        c_float(temperature), _p(sim), _p(mask), _p(loss), grad_row0, grad_rows, _p(dt),
                                        _p(dv), _p(scratch), self._stream()))

    def dual_loss(self, t, v, kind, param, sim, loss, weight=None, fix_norm=True, grad_row0=0, grad_rows=0, dt=None, dv=None):
        """sim_matrix + NormSoftmax / (Adaptive)MaxMarginRanking loss on gathered embeddings (model_epic_charades.py:408-445)."""
        G, P = t.shape
        for x in (t, v, sim) + (() if weight is None else (weight,)):
            assert x.dtype == torch.float32 and x.is_contiguous()
        assert weight is None or weight.numel() == G
        scratch = torch.empty(int(self.lib.egv_dual_loss_scratch_floats(G, P)), dtype=torch.float32, device=t.device)
        self._check(self.lib.egv_dual_loss(_p(t), _p(v), G, P, int(kind), c_float(param), _p(weight), int(bool(fix_norm)),
                                           _p(sim), _p(loss), grad_row0, grad_rows, _p(dt), _p(dv), _p(scratch),
                                           self._stream()))

    # ------------------------------------------------------------------ optimiser
    def adamw(self, p, g, m, v, p_bf16, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, hyper_dev=None):
        for x in (p, g, m, v):
            assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() == p.numel()
        self._check(self.lib.egv_adamw(_p(p), _p(g), _p(m), _p(v), _p(p_bf16), c_int64(p.numel()), c_float(lr),
                                       c_float(beta1), c_float(beta2), c_float(eps), c_float(weight_decay),
                                       c_float(1.0 - beta1 ** step), c_float(1.0 - beta2 ** step), c_float(grad_scale),
                                       _p(hyper_dev), self._stream()))

    def adamw_schedule(self, step_dev, hyper_dev, warmup_steps, max_steps, beta1, beta2):
        """device-side step counter += 1; hyper_dev <- {LR multiplier, 1 - beta1^t, 1 - beta2^t} (capturable)"""
        assert step_dev.dtype == torch.int64 and step_dev.numel() == 1 and hyper_dev.dtype == torch.float32 and hyper_dev.numel() == 3
        self._check(self.lib.egv_adamw_schedule(_p(step_dev), _p(hyper_dev), int(warmup_steps), int(max_steps or 0),
                                                c_float(beta1), c_float(beta2), self._stream()))

    # ------------------------------------------------------------------ NVSwitch P2P all-gather
    def p2p_alloc(self, nbytes):
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._check(self.lib.egv_p2p_alloc(c_int64(nbytes), C.byref(ptr), handle))
        return ptr.value, handle.raw

    def p2p_open(self, handle):
        ptr = C.c_void_p()
        self._check(self.lib.egv_p2p_open(C.c_char_p(handle), C.byref(ptr)))
        return ptr.value

    def p2p_error(self, flags_ptr):
        """0, or 0x100 | peer rank when a gather of this rank timed out waiting for that peer (sticky)"""
        err = c_int(0)
        self._check(self.lib.egv_p2p_error(c_void_p(flags_ptr), C.byref(err)))
        return err.value

    def p2p_allgather(self, src, nbytes, slot_bytes, slots, flags, rank, world, out):
        arr_s = (C.c_void_p * world)(*slots)
        arr_f = (C.c_void_p * world)(*flags)
        self._check(self.lib.egv_p2p_allgather(_p(src), c_int64(nbytes), c_int64(slot_bytes), arr_s, arr_f, rank, world,
                                               _p(out), self._stream()))


_KERNELS = None


def kernels():
    """Process-wide Kernels instance (raises if the CUDA library is missing)."""
    global _KERNELS
    if _KERNELS is None:
        _KERNELS = Kernels()
    return _KERNELS


def set_kernels(k):
    """Test hook: install a different implementation of the Kernels interface (tests/fake_kernels.py)."""
    global _KERNELS
    _KERNELS = k
