"""-m gpu: every CUDA kernel behind the C ABI against its fp32 torch restatement (tests/fake_kernels.py)
and, where the oracle has the function, against oracle/egovlp_oracle.py.

Tolerances (stated): GEMM / attention outputs are bf16-rounded products of bf16 operands with fp32
accumulation -> relative-L2 <= 4e-3 (bf16 has 8 mantissa bits: eps = 3.9e-3) and max-abs within
2 bf16 ulps of the largest magnitude; fp32-only kernels (LayerNorm, losses, reductions) <= 2e-5 relative.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from egovlpv2_b200 import functional as Fn  # noqa: E402
from egovlpv2_b200 import lib as L  # noqa: E402
from tests.fake_kernels import FakeKernels  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def K():
    return L.Kernels()


@pytest.fixture(scope="module")
def R():
    return FakeKernels()


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def check(a, b, tol, what="", atol=0.0):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.isfinite(a.float()).all(), what + ": non-finite output"
    if atol and (a.float() - b.float()).abs().max().item() <= atol:
        return
    r = rel_l2(a, b)
    assert r <= tol, "%s: rel-L2 %.3e > %.1e (max abs diff %.3e)" % (what, r, tol, (a.float() - b.float()).abs().max().item())


def rnd(*shape, dtype=torch.bfloat16, scale=1.0, seed=None):
    g = torch.Generator(device="cpu").manual_seed(seed if seed is not None else (hash(shape) & 0xFFFF))
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


# ------------------------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [
    (768, 768, 6280),     # weight-gradient shape: few output tiles, long K -> automatic split-K (TN layout)
    # M, N, K
    (128, 64, 64),        # one tile, BN=64
    (300, 200, 136),      # tails in every dimension
    (1000, 768, 768),     # BN=128 (few tiles)
    (4096, 768, 768),     # BN=256
    (4100, 2304, 768),    # BN=256, > 148 tiles: persistent loop + TMEM double buffering, M tail
    (256, 3072, 768),     # text-tower shape
    (8, 4096, 768),       # projection head shape (M << tile)
    (520, 50265, 128),    # MLM decoder: N tail, not a multiple of 8
]


def _operands(layout, M, N, Kd, seed=0):
    pad = lambda n: (n + 7) // 8 * 8  # noqa: E731  (row strides must be multiples of 8 elements for TMA)
    if layout == L.GEMM_NT:
        A = rnd(M, pad(Kd), seed=seed)[:, :Kd]
        B = rnd(N, pad(Kd), seed=seed + 1)[:, :Kd]
    elif layout == L.GEMM_NN:
        A = rnd(M, pad(Kd), seed=seed)[:, :Kd]
        B = rnd(Kd, pad(N), seed=seed + 1)[:, :N]
    else:
        A = rnd(Kd, pad(M), seed=seed)[:, :M]
        B = rnd(Kd, pad(N), seed=seed + 1)[:, :N]
    return A, B


@pytest.mark.parametrize("layout", [L.GEMM_NT, L.GEMM_NN, L.GEMM_TN])
@pytest.mark.parametrize("shape", GEMM_SHAPES)
def test_gemm_plain(K, R, layout, shape):
    M, N, Kd = shape
    A, B = _operands(layout, M, N, Kd)
    out, ref = torch.full((M, N), float("nan"), device=DEV), torch.empty(M, N, device=DEV)
    K.gemm(layout, A, B, out_f32=out)
    R.gemm(layout, A, B, out_f32=ref)
    check(out, ref, 1e-5 * math.sqrt(Kd) + 1e-5, "gemm layout %d %s" % (layout, shape))


@pytest.mark.parametrize("act", [L.ACT_NONE, L.ACT_GELU, L.ACT_RELU, L.ACT_TANH, L.ACT_GELU_BWD, L.ACT_RELU_BWD,
                                 L.ACT_TANH_BWD, L.ACT_GELU_DG, L.ACT_MUL_AUX])
def test_gemm_epilogue(K, R, act):
    M, N, Kd = 700, 520, 264
    A, B = _operands(L.GEMM_NT, M, N, Kd, seed=5)
    A, B = A * 0.2, B * 0.2
    bias = rnd(N, dtype=torch.float32, seed=9)
    aux = rnd(M, N, seed=10) if act >= L.ACT_GELU_BWD else None
    if act == L.ACT_TANH_BWD:
        aux = torch.tanh(aux.float()).to(torch.bfloat16)
    res = rnd(M, N, dtype=torch.float32, seed=11)
    sdev = torch.tensor([0.37], device=DEV)
    outs = []
    for impl in (K, R):
        o32 = torch.full((M, N), float("nan"), device=DEV)
        o16 = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        opre = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        cs = torch.ones(N, device=DEV)
        impl.gemm(L.GEMM_NT, A, B, bias=bias, aux=aux, act=act, scale=1.5, scale_dev=sdev, residual=res, out_f32=o32,
                  out_bf16=o16, out_pre=opre, colsum=cs)
        # same problem through the scalar epilogue (odd N: 519 columns)
        o32s = torch.full((M, N - 1), float("nan"), device=DEV)
        css = torch.zeros(N - 1, device=DEV)
        impl.gemm(L.GEMM_NT, A, B[:N - 1], bias=bias[:N - 1].contiguous(), aux=None if aux is None else aux[:, :N - 1],
                  act=act, scale=1.5, scale_dev=sdev, residual=res[:, :N - 1], out_f32=o32s, colsum=css)
        outs.append((o32, o16, opre, cs, o32s, css))
    check(outs[0][0], outs[1][0], 2e-5, "epilogue f32 act %d" % act)
    check(outs[0][1], outs[1][1], 4e-3, "epilogue bf16 act %d" % act)
    check(outs[0][2], outs[1][2], 4e-3, "epilogue pre act %d" % act)
    check(outs[0][3], outs[1][3], 1e-4, "epilogue colsum act %d" % act)
    check(outs[0][4], outs[1][4], 2e-5, "scalar epilogue f32 act %d" % act)
    check(outs[0][5], outs[1][5], 1e-4, "scalar epilogue colsum act %d" % act)


@pytest.mark.parametrize("layout", [L.GEMM_NT, L.GEMM_NN])
def test_gemm_mlp_fast_epilogues(K, R, layout):
    """The lean epilogues of the MLP's activation GEMMs (fc1 forward: bias + pre + GELU outputs; its backward: GELU' of
    the saved pre-activation + column sums), full and partial tiles, strided outputs."""
    M, N, Kd = 128 * 3 + 45, 512 + 64, 200
    A, B = _operands(layout, M, N, Kd, seed=15)
    A, B = A * 0.25, B * 0.25
    bias = rnd(N, dtype=torch.float32, seed=16)
    aux = rnd(M, N, seed=17)
    outs = []
    for impl in (K, R):
        wide = torch.zeros(M, 2 * N, dtype=torch.bfloat16, device=DEV)
        o16, opre = wide[:, :N], wide[:, N:]                      # row stride 2N
        impl.gemm(layout, A, B, bias=bias, act=L.ACT_GELU, out_bf16=o16, out_pre=opre)
        g16 = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        cs = torch.ones(N, device=DEV)
        impl.gemm(layout, A, B, aux=aux, act=L.ACT_GELU_BWD, out_bf16=g16, colsum=cs)
        g16b = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        impl.gemm(layout, A, B, aux=aux, act=L.ACT_GELU_BWD, out_bf16=g16b)
        p16 = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        impl.gemm(layout, A, B, bias=bias, out_bf16=p16, scale=0.5, scale_dev=torch.tensor([1.7], device=DEV))  # lean bias, scale -> bf16
        res = rnd(M, N, dtype=torch.float32, seed=18)
        r32 = res.clone()
        rpre = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        impl.gemm(layout, A, B, bias=bias, scale_dev=torch.tensor([0.37], device=DEV), residual=r32, out_f32=r32,
                  out_pre=rpre)                                                            # lean gated residual, in place
        wide2 = torch.zeros(M, 2 * N, dtype=torch.bfloat16, device=DEV)
        impl.gemm(layout, A, B, bias=bias, act=L.ACT_GELU_DG, out_bf16=wide2[:, :N], out_pre=wide2[:, N:])   # value + derivative
        m16 = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        cs2 = torch.zeros(N, device=DEV)
        impl.gemm(layout, A, B, aux=wide2[:, N:], act=L.ACT_MUL_AUX, out_bf16=m16, colsum=cs2)
        outs.append((o16.clone(), opre.clone(), g16, cs, g16b, p16, r32, rpre, wide2[:, :N].clone(), wide2[:, N:].clone(), m16, cs2))
    names = "gelu pre gelu_bwd colsum gelu_bwd_nocs plain_bf16 residual_f32 residual_pre dg_value dg_deriv mul_aux colsum2"
    for n, a, r in zip(names.split(), outs[0], outs[1]):
        check(a, r, 2e-3 if n.startswith("colsum") else (2e-5 if n == "residual_f32" else 5e-3), "fast epilogue " + n)


def test_gemm_inplace_residual_and_strided_outputs(K, R):
    M, N, Kd = 390, 256, 128
    A, B = _operands(L.GEMM_NT, M, N, Kd, seed=21)
    big = rnd(M, 3 * N, dtype=torch.float32, seed=22)
    big_ref = big.clone()
    K.gemm(L.GEMM_NT, A, B, residual=big[:, N:2 * N], out_f32=big[:, N:2 * N])
    R.gemm(L.GEMM_NT, A, B, residual=big_ref[:, N:2 * N], out_f32=big_ref[:, N:2 * N])
    check(big, big_ref, 1e-5, "in-place residual into a column slice")


@pytest.mark.parametrize("layout", [L.GEMM_NT, L.GEMM_NN, L.GEMM_TN])
def test_gemm_split_k_accumulate(K, R, layout):
    M, N, Kd = 96, 768, 4104
    A, B = _operands(layout, M, N, Kd, seed=31)
    bias = rnd(N, dtype=torch.float32, seed=32)
    init = rnd(M, N, dtype=torch.float32, seed=33)
    out, ref = init.clone(), init.clone()
    K.gemm(layout, A, B, bias=bias, out_f32=out, accumulate=True, split_k=7)
    R.gemm(layout, A, B, bias=bias, out_f32=ref, accumulate=True)
    check(out, ref, 2e-5 * math.sqrt(Kd / 64), "split-k accumulate layout %d" % layout)


@pytest.mark.parametrize("layout", [L.GEMM_NT, L.GEMM_NN, L.GEMM_TN])
def test_gemm_simt_fallback(K, R, layout):
    # odd row strides cannot be described by a TMA tensor map -> SIMT kernel, same semantics
    M, N, Kd = 37, 50, 131
    if layout == L.GEMM_NT:
        A, B = rnd(M, Kd, seed=41), rnd(N, Kd, seed=42)
    elif layout == L.GEMM_NN:
        A, B = rnd(M, Kd, seed=41), rnd(Kd, N, seed=42)
    else:
        A, B = rnd(Kd, M, seed=41), rnd(Kd, N, seed=42)
    bias = rnd(N, dtype=torch.float32, seed=43)
    out, ref = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    K.gemm(layout, A, B, bias=bias, act=L.ACT_GELU, out_f32=out)
    R.gemm(layout, A, B, bias=bias, act=L.ACT_GELU, out_f32=ref)
    check(out, ref, 2e-5, "simt layout %d" % layout)
    # and the forced-SIMT route must agree with the tensor-core route on an aligned problem
    A, B = _operands(layout, 300, 200, 136, seed=44)
    o1, o2 = torch.empty(300, 200, device=DEV), torch.empty(300, 200, device=DEV)
    K.gemm(layout, A, B, out_f32=o1)
    K.force_simt(True)
    try:
        K.gemm(layout, A, B, out_f32=o2)
    finally:
        K.force_simt(False)
    check(o1, o2, 1e-5, "tcgen05 vs simt layout %d" % layout)


def test_gemm_linearity_full_size(K):
    """Size-independent property at the BASELINE cfg-3 shape (M = 8*3137): (A1+A2) W^T == A1 W^T + A2 W^T
    up to bf16 rounding of the summed operand, and rows of a zero A block are exactly bias."""
    M, N, Kd = 8 * 3137, 2304, 768
    A = rnd(M, Kd, seed=51)
    A[1000:1128] = 0
    W = rnd(N, Kd, seed=52, scale=0.05)
    bias = rnd(N, dtype=torch.float32, seed=53)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    K.gemm(L.GEMM_NT, A, W, bias=bias, out_bf16=out)
    assert torch.equal(out[1000:1128].float(), bias.to(torch.bfloat16).float().expand(128, N))
    idx = torch.tensor([0, 1, 127, 128, 5000, 12547, 12548, M - 9, M - 8, M - 1], device=DEV)
    ref = A[idx].float() @ W.float().t() + bias
    check(out[idx], ref.to(torch.bfloat16), 4e-3, "sampled rows of the full-size GEMM")


@pytest.mark.parametrize("layout", [L.GEMM_NT, L.GEMM_NN, L.GEMM_TN])
def test_gemm_cluster_pairs_match_single_cta(K, layout):
    """tcgen05 CTA pairs (cta_group::2) must give bit-identical tiles to the single-CTA kernel (odd number of m-tiles:
    the last pair has one out-of-range tile)."""
    M, N, Kd = 128 * 37 + 5, 1536, 328
    A, B = _operands(layout, M, N, Kd, seed=61)
    bias = rnd(N, dtype=torch.float32, seed=62)
    o1 = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    o2 = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    K.set_cluster(1)
    try:
        K.gemm(layout, A, B, bias=bias, act=L.ACT_GELU, out_bf16=o1)
        K.set_cluster(0)
        K.gemm(layout, A, B, bias=bias, act=L.ACT_GELU, out_bf16=o2)
    finally:
        K.set_cluster(2)
    assert torch.equal(o1, o2)


@pytest.mark.parametrize("shape", [(768, 768, 6277), (2304, 768, 3137), (768, 3072, 4100)])
def test_gemm_pair_weight_gradient(K, R, shape):
    """dW (+)= dy^T x through the default path: tile/split planner + CTA pairs + fp32 vector reductions into a gradient
    sink that already holds a value."""
    N, Kd, M = shape            # dW [N, Kd], reduction over M tokens
    dy, x = rnd(M, N, seed=91), rnd(M, Kd, seed=92)
    base = rnd(N, Kd, dtype=torch.float32, seed=93)
    outs = []
    for impl in (K, R):
        o = base.clone()
        impl.gemm(L.GEMM_TN, dy, x, out_f32=o, accumulate=True, scale=0.5)
        outs.append(o)
    check(outs[0], outs[1], 2e-3, "pair weight gradient")


def test_gemm_argument_errors(K):
    A, B = rnd(64, 64), rnd(64, 64)
    with pytest.raises(RuntimeError):
        K.gemm(L.GEMM_NT, A, B)  # no output
    with pytest.raises(RuntimeError):
        K.gemm(L.GEMM_NT, A, B, act=L.ACT_GELU_BWD, out_f32=torch.empty(64, 64, device=DEV))  # aux missing
    with pytest.raises(RuntimeError):
        K.gemm(L.GEMM_NT, A, B, out_f32=torch.empty(64, 64, device=DEV), split_k=2)  # split-k without accumulate


# ------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,C", [(5, 128), (1000, 768), (8 * 3137, 768), (77, 1024)])
@pytest.mark.parametrize("xbf", [False, True])
def test_layernorm(K, R, rows, C, xbf):
    x = rnd(rows, C, dtype=torch.bfloat16 if xbf else torch.float32, seed=61, scale=2.0) + 0.5
    g, b = rnd(C, dtype=torch.float32, seed=62) * 0.1 + 1, rnd(C, dtype=torch.float32, seed=63) * 0.1
    dy = rnd(rows, C, dtype=torch.float32, seed=64)
    res = []
    for impl in (K, R):
        y16 = torch.empty(rows, C, dtype=torch.bfloat16, device=DEV)
        y32 = torch.empty(rows, C, device=DEV)
        mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
        impl.layernorm_fwd(x, g, b, 1e-5, y_bf16=y16, y_f32=y32, mean=mean, rstd=rstd)
        dx = torch.ones(rows, C, device=DEV)
        dx16 = torch.empty(rows, C, dtype=torch.bfloat16, device=DEV)
        dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        cs1, cs2 = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        impl.layernorm_bwd(dy, x, g, mean, rstd, add=dx, dx=dx, dx_bf16=dx16, bf16_total=True, dgamma=dg, dbeta=db,
                           out_colsum=cs1)
        dx2 = torch.empty(rows, C, device=DEV)
        dx216 = torch.empty(rows, C, dtype=torch.bfloat16, device=DEV)
        impl.layernorm_bwd(dy.to(torch.bfloat16), x, g, mean, rstd, add=dx, dx=dx2, dx_bf16=dx216, bf16_total=False,
                           out_colsum=cs2)
        res.append((y16, y32, mean, rstd, dx, dx16, dg, db, dx2, dx216, cs1, cs2))
    names = "y16 y32 mean rstd dx dx16 dgamma dbeta dx_from_bf16 dx16_partial colsum_total colsum_partial".split()
    for n, a, r in zip(names, res[0], res[1]):
        check(a, r, 4e-3 if a.dtype == torch.bfloat16 else 3e-5, "layernorm " + n)
    ref = torch.nn.functional.layer_norm(x.float(), (C,), g, b, 1e-5)
    check(res[0][1], ref, 1e-5, "layernorm vs torch")


# ------------------------------------------------------------------------------------------- attention
def _attn_case(name, B=2, H=2):
    T, Nf = 4, 9
    N = 1 + T * Nf
    sc = 64 ** -0.5
    if name == "time":
        return N, N, L.AttnSpec(H=H, G=Nf, Lq=T, Lk=T, q_row0=1, q_gstride=1, q_istride=Nf, k_row0=1, k_gstride=1,
                                k_istride=Nf, has_cls_key=True, cls_row=0, scale=sc), False
    if name == "space":
        return N, N, L.AttnSpec(H=H, G=T, Lq=Nf, Lk=Nf, q_row0=1, q_gstride=Nf, q_istride=1, k_row0=1, k_gstride=Nf,
                                k_istride=1, has_cls_key=True, cls_row=0, scale=sc), False
    if name == "cls":
        return N, N, L.AttnSpec(H=H, G=1, Lq=1, Lk=N - 1, q_row0=0, k_row0=1, has_cls_key=True, cls_row=0, scale=sc), False
    if name == "cls_h12":   # 12 heads, 785 tokens: the all-heads-per-warp single-query kernels
        n = 1 + 4 * 196
        return n, n, L.AttnSpec(H=H, G=1, Lq=1, Lk=n - 1, q_row0=0, k_row0=1, has_cls_key=True, cls_row=0, scale=sc), False
    if name == "i2t":      # many queries, 11 keys with a pad mask
        return 200, 11, L.AttnSpec(H=H, G=1, Lq=200, Lk=11, scale=sc), True
    if name == "t2i":      # few queries, many keys
        return 11, 333, L.AttnSpec(H=H, G=1, Lq=11, Lk=333, scale=sc), False
    if name == "t2i_long":   # 32 text queries x 3137 video keys: split-stream forward / dQ
        return 32, 3137, L.AttnSpec(H=H, G=1, Lq=32, Lk=3137, scale=sc), False
    if name == "t2i_11":     # <= 16 rows, long stream
        return 11, 1000, L.AttnSpec(H=H, G=1, Lq=11, Lk=1000, scale=sc), False
    if name == "i2t_long":   # 3137 video queries x 32 text keys with a pad mask: split-stream dK / dV
        return 3137, 32, L.AttnSpec(H=H, G=1, Lq=3137, Lk=32, scale=sc), True
    if name == "text":
        return 32, 32, L.AttnSpec(H=H, G=1, Lq=32, Lk=32, scale=sc), True
    if name == "space196":  # real per-frame size, 2 frames
        Nf2, T2 = 196, 2
        n = 1 + T2 * Nf2
        return n, n, L.AttnSpec(H=H, G=T2, Lq=Nf2, Lk=Nf2, q_row0=1, q_gstride=Nf2, q_istride=1, k_row0=1,
                                k_gstride=Nf2, k_istride=1, has_cls_key=True, cls_row=0, scale=sc), False
    if name.startswith("grp"):   # group-resident kernels (attention_group.cu): contiguous groups of 64..224 rows
        Lq, Lk, cls, G = {"grp_nocls": (100, 150, False, 3), "grp112": (112, 112, True, 2), "grp64": (64, 70, True, 3),
                          "grp224": (224, 207, True, 2), "grp_q196_k40": (196, 40, False, 2),
                          # edges of the tcgen05 backward (attention_tc_bwd.cu): full tiles / a 2-row second query tile
                          "grp_tc256": (256, 255, True, 2), "grp_tc130": (130, 140, True, 2)}[name]
        nq, nk = 1 + G * Lq, 1 + G * Lk
        return nq, nk, L.AttnSpec(H=H, G=G, Lq=Lq, Lk=Lk, q_row0=1, q_gstride=Lq, q_istride=1, k_row0=1, k_gstride=Lk,
                                  k_istride=1, has_cls_key=cls, cls_row=0, scale=sc), False
    if name.startswith("tiny"):   # fused tiny-group kernels (attention_tiny.cu): frames Nf rows apart, adjacent groups
        T2, Nf2, cls = {"tiny12_nocls": (12, 100, False), "tiny8_cls": (8, 150, True), "tiny16_g8": (16, 8, True),
                        "tiny_time16": (16, 196, True)}[name]
        n = 1 + T2 * Nf2
        return n, n, L.AttnSpec(H=H, G=Nf2, Lq=T2, Lk=T2, q_row0=1, q_gstride=1, q_istride=Nf2, k_row0=1, k_gstride=1,
                                k_istride=Nf2, has_cls_key=cls, cls_row=0, scale=sc), False
    if name == "time16":    # 16 frames: 16 queries x 17 keys per group, many groups -> warp-per-group kernel
        Nf2, T2 = 196, 16
        n = 1 + T2 * Nf2
        return n, n, L.AttnSpec(H=H, G=Nf2, Lq=T2, Lk=T2, q_row0=1, q_gstride=1, q_istride=Nf2, k_row0=1, k_gstride=1,
                                k_istride=Nf2, has_cls_key=True, cls_row=0, scale=sc), False
    raise KeyError(name)


@pytest.mark.parametrize("name", ["time", "space", "cls", "cls_h12", "i2t", "t2i", "text", "space196", "time16", "grp_nocls",
                                  "grp112", "grp64", "grp224", "grp_q196_k40", "grp_tc256", "grp_tc130", "tiny12_nocls", "tiny8_cls", "tiny16_g8", "tiny_time16", "t2i_long", "t2i_11", "i2t_long"])
def test_attention_fwd_bwd(K, R, name):
    B, H = (3, 3) if name == "time16" else ((3, 12) if name == "cls_h12" else (2, 2))
    if name.startswith("tiny"):
        B, H = (12, 12) if name == "tiny16_g8" else (3, 4)
    Nq, Nk, spec, masked = _attn_case(name, B, H)
    C = H * 64
    fused = Nq == Nk and spec.has_cls_key
    if fused:   # q, k, v are column slices of one [B, N, 3C] tensor, like the video tower
        qkv = rnd(B, Nq, 3 * C, seed=71)
        q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    else:
        q, k, v = rnd(B, Nq, C, seed=72), rnd(B, Nk, C, seed=73), rnd(B, Nk, C, seed=74)
    kb = None
    if masked:
        lens = torch.tensor([Nk - 3, Nk][:B] if B == 2 else [Nk] * B, device=DEV)
        kb = torch.where(torch.arange(Nk, device=DEV)[None] < lens[:, None], 0.0, torch.finfo(torch.float32).min)
        kb = kb.float().contiguous()
    d_o = rnd(B, Nq, C, seed=75)
    outs = []
    K.set_attention_tiny(3 if name.startswith("tiny") else 0)   # "time*" cases keep the generic kernels covered
    for impl in (K, R):
        o = torch.zeros(B, Nq, C, dtype=torch.bfloat16, device=DEV)
        lse = torch.zeros(B * H * spec.G * spec.Lq, device=DEV)
        impl.attention_fwd(spec, q, k, v, o, lse, key_bias=kb)
        if fused:
            dqkv = torch.zeros(B, Nq, 3 * C, dtype=torch.bfloat16, device=DEV)
            dq, dk, dv = dqkv[:, :, :C], dqkv[:, :, C:2 * C], dqkv[:, :, 2 * C:]
        else:
            dq = torch.zeros(B, Nq, C, dtype=torch.bfloat16, device=DEV)
            dk = torch.zeros(B, Nk, C, dtype=torch.bfloat16, device=DEV)
            dv = torch.zeros(B, Nk, C, dtype=torch.bfloat16, device=DEV)
        delta = torch.zeros_like(lse)
        cls = torch.zeros(B * H * 128, device=DEV) if spec.has_cls_key else None
        impl.attention_bwd(spec, q, k, v, o, lse, d_o, dq, dk, dv, delta, dkv_cls=cls, key_bias=kb)
        if spec.has_cls_key:
            impl.attention_cls_finalize(cls, dk, dv, H, cls_row=0, accumulate=False)
        outs.append((o, lse, dq, dk, dv, delta))
    K.set_attention_tiny(3)
    for n, a, r in zip("o lse dq dk dv delta".split(), outs[0], outs[1]):
        tol = {"dq": 1.2e-2, "dk": 1.2e-2, "dv": 1.2e-2, "o": 8e-3, "delta": 2e-2, "lse": 1e-4}[n]
        check(a, r, tol, "attention %s %s" % (name, n))


@pytest.mark.parametrize("shape", [(2, 2, 3, 196), (2, 12, 2, 196), (1, 4, 2, 170)])
def test_space_attention_backward_folds_cls_query(K, R, shape):
    """functional.divided_attention_bwd in space mode: the tcgen05 backward (csrc/attention_tc_bwd.cu) takes the clip's CLS
    query along (its q / dO / O rows in the first unused query slot of tile 1, its global lse in the statistics) instead of
    a separate pass over all keys -- against the restatement, which runs the two backward calls of the reference
    formulation (video_transformer.py:134-150).  Every element of d_qkv incl. the CLS row."""
    B, H, T, Nf = shape
    C = H * 64
    N = 1 + T * Nf
    qkv = rnd(B, N, 3 * C, seed=91)
    d_o = rnd(B, N, C, seed=92)
    outs = []
    for impl in (K, R):
        n0 = impl.launch_count() if impl is K else 0
        o, lses = Fn.divided_attention_fwd(impl, qkv, H, T, Nf, "space")
        if impl is K:   # tcgen05 forward with the CLS query folded + the combine of its per-frame partials
            assert K.launch_count() - n0 == 2, K.launch_count() - n0
        n0 = impl.launch_count() if impl is K else 0
        d_qkv = Fn.divided_attention_bwd(impl, qkv, o, lses, d_o, H, T, Nf, "space")
        if impl is K:
            # tcgen05 backward + CLS-query finalize + CLS-key finalize (+ nothing else): the single-query backward is gone
            assert K.launch_count() - n0 == 3, K.launch_count() - n0
        outs.append((o, d_qkv, lses[1]))
    check(outs[0][0][:, 1:], outs[1][0][:, 1:], 8e-3, "space attention o")
    check(outs[0][0][:, :1], outs[1][0][:, :1], 8e-3, "o of the CLS query (folded)")
    check(outs[0][2], outs[1][2], 1e-4, "lse of the CLS query (folded)")
    dq, dk, dv = (outs[0][1][:, :, i * C:(i + 1) * C] for i in range(3))
    rq, rk, rv = (outs[1][1][:, :, i * C:(i + 1) * C] for i in range(3))
    check(dq[:, 1:], rq[:, 1:], 1.2e-2, "dq patches")
    check(dq[:, :1], rq[:, :1], 1.2e-2, "dq of the CLS query (folded)")
    check(dk[:, 1:], rk[:, 1:], 1.2e-2, "dk patches (incl. the CLS query's share)")
    check(dv[:, 1:], rv[:, 1:], 1.2e-2, "dv patches (incl. the CLS query's share)")
    check(dk[:, :1], rk[:, :1], 1.2e-2, "dk of the CLS key")
    check(dv[:, :1], rv[:, :1], 1.2e-2, "dv of the CLS key")


@pytest.mark.parametrize("case", ["cls", "cls_h12"])
def test_attention_dkv_accumulate(K, R, case):
    """The CLS-query pass adds its dk/dv onto what the grouped pass wrote (read-modify-write rows)."""
    B, H = (2, 2) if case == "cls" else (2, 12)
    Nq, Nk, spec, _ = _attn_case(case, B, H)
    C = H * 64
    qkv = rnd(B, Nq, 3 * C, seed=81)
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    d_o = rnd(B, Nq, C, seed=82)
    base = rnd(B, Nq, 3 * C, seed=83)
    outs = []
    for impl in (K, R):
        o = torch.zeros(B, Nq, C, dtype=torch.bfloat16, device=DEV)
        lse = torch.zeros(B * H, device=DEV)
        impl.attention_fwd(spec, q, k, v, o, lse)
        dqkv = base.clone()
        cls = torch.zeros(B * H * 128, device=DEV)
        impl.attention_bwd(spec, q, k, v, o, lse, d_o, dqkv[:, :, :C], dqkv[:, :, C:2 * C], dqkv[:, :, 2 * C:],
                           torch.zeros_like(lse), dkv_cls=cls, dkv_accumulate=True)
        impl.attention_cls_finalize(cls, dqkv[:, :, C:2 * C], dqkv[:, :, 2 * C:], H, cls_row=0, accumulate=True)
        outs.append(dqkv)
    check(outs[0], outs[1], 8e-3, "dkv accumulate")


# ------------------------------------------------------------------------------------------- elementwise & embeddings
def test_elementwise(K, R):
    x = rnd(1000003, dtype=torch.float32, seed=91)
    y1, y2 = torch.empty_like(x, dtype=torch.bfloat16), torch.empty_like(x, dtype=torch.bfloat16)
    K.cast(x, y1)
    R.cast(x, y2)
    assert torch.equal(y1, y2)
    z1, z2 = torch.empty_like(x), torch.empty_like(x)
    K.cast(y1, z1)
    R.cast(y1, z2)
    assert torch.equal(z1, z2)
    b = rnd(1000003, dtype=torch.float32, seed=92)
    al = torch.tensor([-0.25], device=DEV)
    for impl, (o, o16) in ((K, (z1, y1)), (R, (z2, y2))):
        impl.axpy(x, b, 2.0, al, y=o, y_bf16=o16)
    check(z1, z2, 1e-6, "axpy")
    assert torch.equal(y1, y2)
    m = rnd(3001, 770, seed=93)
    for dt in (torch.bfloat16, torch.float32):
        mm = m.to(dt)
        c1, c2 = torch.ones(770, device=DEV), torch.ones(770, device=DEV)
        K.colsum(mm, c1, accumulate=True, scale=0.5, scale_dev=al)
        R.colsum(mm, c2, accumulate=True, scale=0.5, scale_dev=al)
        check(c1, c2, 2e-5, "colsum")
        c1, c2 = torch.ones(768, device=DEV), torch.ones(768, device=DEV)
        K.colsum(mm[:, 1:769], c1)
        R.colsum(mm[:, 1:769], c2)
        check(c1, c2, 2e-5, "colsum odd offset")
    wide = rnd(3000, 2304, seed=95)        # streaming bf16x8 path (aligned rows, C % 8 == 0)
    c1, c2 = torch.ones(2304, device=DEV), torch.ones(2304, device=DEV)
    K.colsum(wide, c1, accumulate=True, scale=2.0)
    R.colsum(wide, c2, accumulate=True, scale=2.0)
    check(c1, c2, 2e-5, "colsum bf16x8")
    c1, c2 = torch.zeros(768, device=DEV), torch.zeros(768, device=DEV)
    K.colsum(wide[:, 768:1536], c1)
    R.colsum(wide[:, 768:1536], c2)
    check(c1, c2, 2e-5, "colsum bf16x8 column slice")
    d1, d2 = torch.ones(1, device=DEV), torch.ones(1, device=DEV)
    K.dot(x, y1, d1, accumulate=True)
    R.dot(x, y1, d2, accumulate=True)
    check(d1, d2, 1e-4, "dot")
    K.zero(z1)
    assert z1.abs().max().item() == 0
    aux = rnd(1000003, seed=94)
    for act in (L.ACT_NONE, L.ACT_GELU_BWD, L.ACT_RELU_BWD, L.ACT_TANH_BWD):
        a = torch.tanh(aux.float()).to(torch.bfloat16) if act == L.ACT_TANH_BWD else aux
        for dy in (x, y1):
            K.act_grad(dy, None if act == L.ACT_NONE else a, act, y1.new_empty(0) if False else y2, scale=0.5, scale_dev=al)
            ref = torch.empty_like(y2)
            R.act_grad(dy, None if act == L.ACT_NONE else a, act, ref, scale=0.5, scale_dev=al)
            check(y2, ref, 4e-3, "act_grad %d" % act)


def test_patch_embed_and_text_embed(K, R):
    B, T, Nf, C, p = 2, 3, 4, 128, 16
    video = rnd(B * T, 3, 2 * p, 2 * p, dtype=torch.float32, seed=101)
    o1, o2 = torch.empty(B * T * Nf, 3 * p * p, dtype=torch.bfloat16, device=DEV), torch.empty(B * T * Nf, 3 * p * p, dtype=torch.bfloat16, device=DEV)
    K.patchify(video, p, o1)
    R.patchify(video, p, o2)
    assert torch.equal(o1, o2)
    # uint8 frames: fused / 255 + NormalizeVideo + im2col is bit-identical to the host pipeline + fp32 im2col
    g = torch.Generator().manual_seed(107)
    for shape, pp in (((B * T, 3, 2 * p, 2 * p), p), ((5, 3, 48, 24), 8), ((8 * 16, 3, 224, 224), 16), ((3, 1, 32, 32), 16)):
        u8_host = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
        mean, std = ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)) if shape[1] == 3 else ((0.45,), (0.225,))
        # the reference's loader runs on the CPU (true fp32 divisions)
        host = (((u8_host.float() / 255) - torch.tensor(mean).view(1, -1, 1, 1)) / torch.tensor(std).view(1, -1, 1, 1)).to(DEV)
        u8 = u8_host.to(DEV)
        a = torch.empty(u8.numel() // (shape[1] * pp * pp), shape[1] * pp * pp, dtype=torch.bfloat16, device=DEV)
        b = torch.empty_like(a)
        K.patchify_u8(u8, pp, a, mean, std)
        K.patchify(host.contiguous(), pp, b)
        assert torch.equal(a, b), shape
        R.patchify_u8(u8, pp, b, mean, std)
        assert torch.equal(a, b), shape
    with pytest.raises(RuntimeError):
        K.patchify_u8(u8, 16, a, (0.45,), (0.0,))
    patch = rnd(B * T * Nf, C, dtype=torch.float32, seed=102)
    cls, pos, tem = rnd(C, dtype=torch.float32, seed=103), rnd(1 + Nf, C, dtype=torch.float32, seed=104), rnd(T, C, dtype=torch.float32, seed=105)
    t1, t2 = torch.empty(B, 1 + T * Nf, C, device=DEV), torch.empty(B, 1 + T * Nf, C, device=DEV)
    K.assemble_tokens(patch, cls, pos, tem, B, T, Nf, t1)
    R.assemble_tokens(patch, cls, pos, tem, B, T, Nf, t2)
    check(t1, t2, 1e-7, "assemble")
    d = rnd(B, 1 + T * Nf, C, dtype=torch.float32, seed=106)
    g = []
    for impl in (K, R):
        dp = torch.empty(B * T * Nf, C, dtype=torch.bfloat16, device=DEV)
        dc, dpos, dt = torch.ones(C, device=DEV), torch.ones(1 + Nf, C, device=DEV), torch.ones(T, C, device=DEV)
        impl.assemble_tokens_bwd(d, B, T, Nf, dp, dc, dpos, dt)
        g.append((dp, dc, dpos, dt))
    for a, r in zip(*g):
        check(a, r, 1e-5 if a.dtype == torch.float32 else 4e-3, "assemble bwd")
    ids = torch.tensor([[0, 5, 9, 2, 1, 1], [0, 7, 7, 7, 8, 2]], device=DEV)
    word, posw, ty = rnd(12, C, dtype=torch.float32, seed=107), rnd(10, C, dtype=torch.float32, seed=108), rnd(C, dtype=torch.float32, seed=109)
    e1, e2 = torch.empty(2, 6, C, device=DEV), torch.empty(2, 6, C, device=DEV)
    K.text_embed(ids, word, posw, ty, e1)
    R.text_embed(ids, word, posw, ty, e2)
    check(e1, e2, 1e-7, "text embed")
    de = rnd(2, 6, C, dtype=torch.float32, seed=110)
    g = []
    for impl in (K, R):
        dw, dp, dt = torch.zeros(12, C, device=DEV), torch.zeros(10, C, device=DEV), torch.zeros(C, device=DEV)
        impl.text_embed_bwd(de, ids, dw, dp, dt)
        g.append((dw, dp, dt))
    for a, r in zip(*g):
        check(a, r, 1e-5, "text embed bwd")


# ------------------------------------------------------------------------------------------- losses / optimiser
def test_softmax_xent(K, R):
    rows, V = 37, 50265
    ld = 50272
    logits = rnd(rows, ld, dtype=torch.float32, seed=121, scale=3.0)[:, :V]
    labels = torch.randint(0, V, (rows,), generator=torch.Generator().manual_seed(1)).to(DEV)
    labels[::5] = -100
    r = []
    for impl in (K, R):
        ls, cnt = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
        dl = torch.full((rows, ld), 7.0, dtype=torch.bfloat16, device=DEV)
        impl.softmax_xent(logits, labels, V, ls, cnt, dlogits=dl[:, :V])
        dl32 = torch.full((rows, V), 7.0, device=DEV)
        ls2, cnt2 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
        impl.softmax_xent(logits, labels, V, ls2, cnt2, dlogits=dl32)
        assert torch.equal(ls, ls2) or abs((ls - ls2).item()) < 1e-3
        loss, inv = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
        impl.xent_finalize(ls, cnt, loss, inv)
        r.append((loss, inv, dl, dl32))
    check(r[0][0], r[1][0], 1e-5, "xent loss")
    check(r[0][1], r[1][1], 1e-6, "xent inv count")
    check(r[0][2], r[1][2], 4e-3, "xent dlogits")
    check(r[0][3], r[1][3], 1e-5, "xent dlogits f32")
    ref = torch.nn.functional.cross_entropy(logits, labels, ignore_index=-100)
    check(r[0][0], ref.reshape(1), 1e-5, "xent vs torch")


@pytest.mark.parametrize("G", [4, 8, 64])
def test_egonce_vs_oracle(K, R, G):
    from oracle import egovlp_oracle as O
    P = 4096 if G > 4 else 256
    t, v = rnd(G, P, dtype=torch.float32, seed=131), rnd(G, P, dtype=torch.float32, seed=132)
    v = v + 0.5 * t  # make the diagonal somewhat dominant, like trained embeddings
    gen = torch.Generator().manual_seed(3)
    noun = (torch.rand(G, 582, generator=gen) < 0.02).float()
    verb = (torch.rand(G, 118, generator=gen) < 0.05).float()
    noun[torch.arange(G), torch.randint(0, 582, (G,), generator=gen)] = 1
    verb[torch.arange(G), torch.randint(0, 118, (G,), generator=gen)] = 1
    noun, verb = noun.to(DEV), verb.to(DEV)
    sim, mask, loss = torch.empty(G, G, device=DEV), torch.empty(G, G, dtype=torch.uint8, device=DEV), torch.empty(1, device=DEV)
    r0, nr = G // 4, G // 2
    dt, dv = torch.empty(nr, P, device=DEV), torch.empty(nr, P, device=DEV)
    K.egonce(t, v, noun, verb, 0.05, sim, mask, loss, r0, nr, dt, dv)
    tt, vv = t.clone().requires_grad_(True), v.clone().requires_grad_(True)
    osim = O.sim_matrix(tt, vv)
    oloss, omask = O.egonce(osim, O.sim_matrix(verb, verb), O.sim_matrix(noun, noun))
    oloss.backward()
    check(sim, osim.detach(), 1e-5, "egonce sim")
    assert torch.equal(mask.bool(), omask)
    check(loss, oloss.detach().reshape(1), 1e-5, "egonce loss", atol=2e-6)
    check(dt, tt.grad[r0:r0 + nr], 1e-4, "egonce dt", atol=1e-7)
    check(dv, vv.grad[r0:r0 + nr], 1e-4, "egonce dv", atol=1e-7)


@pytest.mark.parametrize("kind,param,fix_norm", [(0, 0.05, True), (0, 0.07, True), (1, 0.2, True), (1, 0.2, False),
                                                 (2, 0.4, True), (2, 0.2, False)])
def test_dual_loss_vs_oracle(K, kind, param, fix_norm):
    """egv_dual_loss (sim_matrix + NormSoftmax / MaxMarginRanking / AdaptiveMaxMarginRanking + local-slice gradients)
    vs autograd through the oracle's restatement of loss.py:13-31, 65-143 (pinned to the reference's classes in
    tests/test_oracle_golden.py)."""
    from oracle import egovlp_oracle as O
    for G, P in ((1, 64), (6, 256), (32, 256), (257, 64)):
        if G == 1 and kind != 0 and fix_norm:
            continue    # mean over zero off-diagonal terms: NaN in the reference too
        if G == 257 and kind != 0:
            continue    # the pure-Python oracle loops over pairs
        t, v = rnd(G, P, dtype=torch.float32, seed=151), rnd(G, P, dtype=torch.float32, seed=152)
        v = v + 0.7 * t
        w = torch.rand(G, generator=torch.Generator().manual_seed(4)).to(DEV)
        sim, loss = torch.empty(G, G, device=DEV), torch.empty(1, device=DEV)
        r0, nr = G // 4, max(G // 2, 1)
        dt, dv = torch.empty(nr, P, device=DEV), torch.empty(nr, P, device=DEV)
        K.dual_loss(t, v, kind, param, sim, loss, weight=w if kind == 2 else None, fix_norm=fix_norm, grad_row0=r0,
                    grad_rows=nr, dt=dt, dv=dv)
        tt, vv = t.clone().requires_grad_(True), v.clone().requires_grad_(True)
        osim = O.sim_matrix(tt, vv)
        if kind == 0:
            oloss = O.norm_softmax_loss(osim, param)
        else:
            oloss = O.max_margin_ranking_loss(osim, param, w.cpu() if kind == 2 else None, fix_norm)
        oloss.backward()
        check(sim, osim.detach(), 1e-5, "dual sim")
        check(loss, oloss.detach().reshape(1), 2e-5, "dual loss", atol=2e-6)
        check(dt, tt.grad[r0:r0 + nr], 2e-4, "dual dt", atol=1e-7)
        check(dv, vv.grad[r0:r0 + nr], 2e-4, "dual dv", atol=1e-7)
    with pytest.raises(RuntimeError):
        K.dual_loss(t, v, 2, 0.2, sim, loss)          # adaptive margin without weights
    with pytest.raises(RuntimeError):
        K.dual_loss(t, v, 7, 0.2, sim, loss)


def test_adamw(K, R):
    n = 100003
    p0, g = rnd(n, dtype=torch.float32, seed=141), rnd(n, dtype=torch.float32, seed=142)
    st = []
    for impl in (K, R):
        p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
        p16 = torch.empty(n, dtype=torch.bfloat16, device=DEV)
        for step in (1, 2, 3):
            impl.adamw(p, g * step, m, v, p16, 1e-3, 0.9, 0.98, 1e-8, 0.01, step, grad_scale=0.5)
        st.append((p, m, v, p16))
    for a, r in zip(*st):
        check(a, r, 4e-3 if a.dtype == torch.bfloat16 else 1e-5, "adamw")


# ------------------------------------------------------------------------------------------- batched GEMM (csrc/xgemm.cu)
from egovlpv2_b200.lib import BV, EPI_DSOFTMAX32, EPI_NONE, EPI_SOFTMAX32, GEMM_NN, GEMM_NT, GEMM_TN  # noqa: E402


def _both(K, R, fn):
    """run `fn(kernels)` (which returns the output tensors it allocated) on the CUDA library and on the restatement"""
    return fn(K), fn(R)


def test_bgemm_group_softmax_epilogue_per_clip(K, R):
    """P[b] = softmax32(A[b] Mt[b]^T + bias[b]): per-clip operands, 300 rows per clip (row tail inside every batch:
    rows past a clip's own M must come back as TMA zero fill, not as the next clip's rows), H*S = 384 columns."""
    B, N, HS, C = 3, 300, 384, 768
    A, Mt = rnd(B * N, C, seed=1), rnd(B, HS, C, scale=0.05, seed=2)
    bias = rnd(B, HS, dtype=torch.float32, seed=3)
    bias[1, 40:64] = torch.finfo(torch.float32).min        # masked keys of one clip (Q13)

    def run(k):
        P = torch.full((B, N + 5, HS), 7.0, dtype=torch.bfloat16, device=DEV)   # 5 guard rows per clip must stay untouched
        k.bgemm(GEMM_NT, N, HS, C, BV(A, C, N * C), BV(Mt, C, HS * C), nb=(B, 1), bias=BV(bias, 0, HS),
                out_bf16=BV(P, HS, (N + 5) * HS), epilogue=EPI_SOFTMAX32)
        return P
    got, ref = _both(K, R, run)
    check(got[:, :N], ref[:, :N], 6e-3, "softmax32")
    assert (got[:, N:] == 7.0).all(), "rows past a batch's M were written"
    s = got[:, :N].float().reshape(B, N, HS // 32, 32).sum(-1)
    assert (s - 1).abs().max().item() < 2e-2
    assert got[1, :N, 40:64].abs().max().item() == 0.0


def test_bgemm_softmax_backward_epilogue(K, R):
    B, N, HS, C = 2, 261, 384, 768
    dO, U = rnd(B * N, C, seed=4), rnd(B, HS, C, scale=0.05, seed=5)
    P = torch.softmax(rnd(B, N, HS // 32, 32, dtype=torch.float32, seed=6), -1).reshape(B, N, HS).to(torch.bfloat16)
    alpha = torch.tensor([0.37], device=DEV)

    def run(k):
        dS = torch.zeros(B, N, HS, dtype=torch.bfloat16, device=DEV)
        dbias, dalpha = torch.zeros(B, HS, device=DEV), torch.zeros(1, device=DEV)
        k.bgemm(GEMM_NT, N, HS, C, BV(dO, C, N * C), BV(U, C, HS * C), nb=(B, 1), scale_dev=alpha, aux=BV(P, HS, N * HS),
                out_bf16=BV(dS, HS, N * HS), colsum=BV(dbias, 0, HS), dot_out=dalpha, epilogue=EPI_DSOFTMAX32)
        return dS, dbias, dalpha
    (dS, db, da), (dS_r, db_r, da_r) = _both(K, R, run)
    check(dS, dS_r, 8e-3, "dsoftmax32 dS")
    check(db, db_r, 2e-2, "dsoftmax32 colsum", atol=2e-2)
    check(da, da_r, 1e-3, "dsoftmax32 dot")


def test_bgemm_head_batches_and_residual(K, R):
    """batches = (clip, head) pairs addressed as column slices of a [B*S, 2C] tensor (batch stride 64 elements along a
    row), shared weights (batch stride 0 in the clip level), K = 64 / K = 32 reductions, strided column-slice outputs,
    accumulation over the clip level into a shared output."""
    B, S, H, C = 2, 32, 12, 768
    kv, wq = rnd(B * S, 2 * C, seed=7), rnd(C, C, scale=0.05, seed=8)
    dM = rnd(B, H, S, C, seed=9)
    bv = rnd(C, dtype=torch.float32, seed=10)

    def run(k):
        Mt = torch.zeros(B, H, S, C, dtype=torch.bfloat16, device=DEV)
        k.bgemm(GEMM_NN, S, C, 64, BV(kv, 2 * C, 64, S * 2 * C), BV(wq, C, 64 * C, 0), nb=(H, B), scale=0.125,
                out_bf16=BV(Mt, C, S * C, H * S * C))
        U = torch.zeros(B, H, S, C, dtype=torch.bfloat16, device=DEV)
        k.bgemm(GEMM_NT, S, C, 64, BV(kv[:, C:], 2 * C, 64, S * 2 * C), BV(wq, C, 64, 0), nb=(H, B), out_bf16=BV(U, C, S * C, H * S * C))
        dk = torch.zeros(B * S, 2 * C, device=DEV)
        k.bgemm(GEMM_NT, S, 64, C, BV(dM, C, S * C, H * S * C), BV(wq, C, 64 * C, 0), nb=(H, B), scale=0.125,
                bias=BV(bv, 0, 64, 0), out_f32=BV(dk, 2 * C, 64, S * 2 * C))
        k.bgemm(GEMM_NN, S, 64, C, BV(dM, C, S * C, H * S * C), BV(wq, C, 64, 0), nb=(H, B), out_f32=BV(dk[:, C:], 2 * C, 64, S * 2 * C))
        dwq = torch.zeros(C, C, device=DEV)
        k.bgemm(GEMM_TN, 64, C, S, BV(kv, 2 * C, 64, S * 2 * C), BV(dM, C, S * C, H * S * C), nb=(H, B), scale=0.125,
                out_f32=BV(dwq, C, 64 * C, 0), accumulate=True)
        dwp = torch.zeros(C, C, device=DEV)
        k.bgemm(GEMM_TN, C, 64, S, BV(dM, C, S * C, H * S * C), BV(kv[:, C:], 2 * C, 64, S * 2 * C), nb=(H, B),
                out_f32=BV(dwp, C, 64, 0), accumulate=True)
        return Mt, U, dk, dwq, dwp
    for name, a, b in zip(("Mt", "U", "dk|dv", "dwq", "dwp"), *_both(K, R, run)):
        check(a, b, 6e-3, "head batches " + name)


def test_bgemm_per_clip_reductions_over_tokens(K, R):
    """K = tokens of one clip (3137 is odd: the reduction tail of clip b must not read clip b+1), odd N with a padded
    row stride, split-K accumulation, residual + gate epilogue."""
    B, N, HS, C = 2, 1201, 384, 768
    Np = (N + 7) // 8 * 8
    P = torch.softmax(rnd(B, N, HS // 32, 32, dtype=torch.float32, seed=11), -1).reshape(B, N, HS).to(torch.bfloat16)
    dO, x = rnd(B * N, C, seed=12), rnd(B * N, C, seed=13)
    Qp = rnd(B, HS, C, scale=0.05, seed=14)
    xa = rnd(B * N, C, dtype=torch.float32, seed=15)
    bp = rnd(C, dtype=torch.float32, seed=16)
    alpha = torch.tensor([0.5], device=DEV)

    def run(k):
        dU = torch.zeros(B, HS, C, device=DEV)
        k.bgemm(GEMM_TN, HS, C, N, BV(P, HS, N * HS), BV(dO, C, N * C), nb=(B, 1), scale_dev=alpha, out_f32=BV(dU, C, HS * C),
                accumulate=True)
        Sc = torch.full((B, HS, Np), -5.0, device=DEV)
        k.bgemm(GEMM_NT, HS, N, C, BV(Qp, C, HS * C), BV(x, C, N * C), nb=(B, 1), out_f32=BV(Sc, Np, HS * Np))
        Pt = torch.zeros(B, HS, Np, dtype=torch.bfloat16, device=DEV)
        Pt[:, :, :N] = torch.softmax(Sc[:, :, :N], -1)
        Z = torch.zeros(B, HS, C, device=DEV)
        k.bgemm(GEMM_NN, HS, C, N, BV(Pt, Np, HS * Np), BV(x, C, N * C), nb=(B, 1), out_f32=BV(Z, C, HS * C), accumulate=True)
        dx = torch.zeros(B * N, C, device=DEV)
        k.bgemm(GEMM_TN, N, C, HS, BV(Pt, Np, HS * Np), BV(Qp, C, HS * C), nb=(B, 1), out_f32=BV(dx, C, N * C))
        out = torch.zeros(B * N, C, device=DEV)
        k.bgemm(GEMM_NN, N, C, HS, BV(P, HS, N * HS), BV(Qp, C, HS * C), nb=(B, 1), bias=BV(bp, 0, 0), scale_dev=alpha,
                residual=BV(xa, C, N * C), out_f32=BV(out, C, N * C))
        return dU, Sc[:, :, :N].contiguous(), Z, dx, out
    for name, a, b in zip(("dU", "scores", "Z", "dx", "gated residual"), *_both(K, R, run)):
        check(a, b, 6e-3, "per-clip " + name)


@pytest.mark.parametrize("N", [3137, 18433])   # cfg 3 rows (register-resident kernels); cfg 5 rows (streaming variants)
def test_xattn_row_kernels(K, R, N):
    B, HS = (2, 384) if N < 4094 else (1, 96)
    Np = (N + 7) // 8 * 8
    Sc = rnd(B, HS, Np, dtype=torch.float32, scale=3.0, seed=21)
    dP = rnd(B, HS, Np, dtype=torch.float32, seed=22)
    rc = rnd(B, HS, dtype=torch.float32, seed=23)
    seed_dev = torch.tensor([987654321012345], dtype=torch.int64, device=DEV)
    for p_drop in (0.0, 0.1):
        def run(k):
            PdS = torch.zeros(B, 2, HS, Np, dtype=torch.bfloat16, device=DEV)
            lse, rsum = torch.zeros(B * HS, device=DEV), torch.zeros(B * HS, device=DEV)
            k.xattn_row_softmax(Sc, Np, B * HS, HS, HS * Np, N, PdS, Np, 2 * HS * Np, lse, p_drop, seed_dev, 1234, rsum)
            k.xattn_row_dsoftmax(Sc, Np, B * HS, HS, HS * Np, N, lse, dP, Np, HS * Np, PdS[:, 1], Np, 2 * HS * Np, p_drop, seed_dev, 1234,
                                 row_const=rc if p_drop > 0 else None)
            return PdS[:, 0, :, :N].contiguous(), PdS[:, 1, :, :N].contiguous(), lse, rsum
        (P, dS, lse, rs), (P_r, dS_r, lse_r, rs_r) = _both(K, R, run)
        # the dropout masks must be IDENTICAL (same Philox stream): compare the zero patterns exactly
        assert ((P == 0) == (P_r == 0)).all(), "dropout mask differs from the Philox restatement"
        if p_drop > 0:
            frac = (P == 0).float().mean().item()
            assert abs(frac - p_drop) < 0.01, frac
        check(P, P_r, 6e-3, "row softmax p=%g" % p_drop)
        check(dS, dS_r, 8e-3, "row dsoftmax p=%g" % p_drop)
        check(lse, lse_r, 1e-5, "lse")
        check(rs, rs_r, 5e-3, "rsum")
    if N != 3137:
        return
    # query-bias kernels and the value bias under dropout
    Bc, S, H, C = 3, 32, 12, 768
    kv = rnd(Bc * S, 2 * C, seed=24)
    bq = rnd(C, dtype=torch.float32, seed=25)
    mask = torch.zeros(Bc, S, device=DEV)
    mask[1, 20:] = torch.finfo(torch.float32).min
    dbias = rnd(Bc, H * S, dtype=torch.float32, seed=26)
    rsum = rnd(Bc, H * S, dtype=torch.float32, seed=27)
    ox0 = rnd(Bc * S, C, seed=28)

    def run2(k):
        out = torch.zeros(Bc, H * S, device=DEV)
        k.xattn_qbias_fwd(kv, 2 * C, bq, mask, 0.125, Bc, S, H, out)
        dk, dbq = torch.ones(Bc * S, 2 * C, device=DEV), torch.zeros(C, device=DEV)
        k.xattn_qbias_bwd(kv, 2 * C, bq, dbias, 0.125, Bc, S, H, dk=dk, lddk=2 * C, dbq=dbq)
        ox, dbv = ox0.clone(), torch.zeros(C, device=DEV)
        k.xattn_rowscale_bias(ox, rsum, bq, Bc, S, H)
        k.xattn_rowscale_bias_bwd(ox0, rsum, dbv, Bc, S, H)
        return out, dk, dbq, ox, dbv
    for name, a, b in zip(("qbias", "dk", "dbq", "rowscale", "dbv"), *_both(K, R, run2)):
        check(a, b, 6e-3 if name == "rowscale" else 2e-4, name)


def test_reassociated_cross_attention_full_size(K, R):
    """xattn_reassoc at the BASELINE shapes (N = 3137 tokens per clip, C = 768, 12 heads, S = 32; 2 clips): the CUDA
    kernels against the restatement running the very same sequencing, forward and every gradient, both directions."""
    from egovlpv2_b200 import xattn_reassoc as XR
    B, N, C, H, S = 2, 3137, 768, 12, 32
    lnc, kv = rnd(B * N, C, seed=31), rnd(B * S, 2 * C, seed=32)
    wq, wp = rnd(C, C, scale=0.03, seed=33), rnd(C, C, scale=0.03, seed=34)
    bq, bp = rnd(C, dtype=torch.float32, scale=0.1, seed=35), rnd(C, dtype=torch.float32, scale=0.1, seed=36)
    xa, dout = rnd(B * N, C, dtype=torch.float32, seed=37), rnd(B * N, C, seed=38)
    mask = torch.zeros(B, S, device=DEV)
    mask[1, 11:] = torch.finfo(torch.float32).min
    alpha = torch.tensor([0.5], device=DEV)

    def run_i2t(k):
        out = torch.empty(B * N, C, device=DEV)
        s = XR.i2t_fwd(k, lnc, kv, mask, wq, bq, wp, bp, alpha, xa, out, B, N, H)
        da, dwq, dbq, dwp = (torch.zeros(n, device=DEV) for n in ((1,), (C, C), (C,), (C, C)))
        dln, dkv = XR.i2t_bwd(k, s, dout, dout.float().sum(0), wq, bq, wp, bp, alpha, da, dwq, dbq, dwp)
        return out, s.P, dln, dkv, da, dwq, dbq, dwp
    for name, a, b in zip(("out", "P", "dln", "dkv", "dalpha", "dwq", "dbq", "dwp"), *_both(K, R, run_i2t)):
        check(a, b, 1.5e-2, "i2t " + name)
    q, x = rnd(B * S, C, seed=41), rnd(B * N, C, seed=42)
    wk, wv = rnd(C, C, scale=0.03, seed=43), rnd(C, C, scale=0.03, seed=44)
    bv = rnd(C, dtype=torch.float32, scale=0.1, seed=45)
    dox = rnd(B * S, C, seed=46)
    for p_drop in (0.0, 0.1):
        def run_t2i(k):
            ox = torch.empty(B * S, C, dtype=torch.bfloat16, device=DEV)
            s = XR.t2i_fwd(k, q, x, wk, wv, bv, ox, B, N, H, p_drop=p_drop, seed_dev=None, site=99)
            dwk, dwv, dbv = (torch.zeros(n, device=DEV) for n in ((C, C), (C, C), (C,)))
            dx = torch.empty(B * N, C, device=DEV)
            dq = XR.t2i_bwd(k, s, dox, wk, wv, bv, dwk, dwv, dbv, dx)
            return ox, dq, dx, dwk, dwv, dbv
        for name, a, b in zip(("ox", "dq", "dx", "dwk", "dwv", "dbv"), *_both(K, R, run_t2i)):
            check(a, b, 1.5e-2, "t2i p=%g %s" % (p_drop, name))


# ------------------------------------------------------------------------------------------- train-mode dropout (csrc/dropout.cu)
def test_dropout_kernels_and_text_attention(K, R):
    """Philox masks identical to the host restatement (bit-for-bit: zero patterns compared exactly), dense-output dropout +
    residual, its backward, the step-seed advance, and the 32-token self-attention with dropout on the probabilities."""
    seed = torch.tensor([0x1234567890ABCDE], dtype=torch.int64, device=DEV)
    seed_r = seed.clone()
    K.rng_advance(seed)
    R.rng_advance(seed_r)
    assert int(seed) == int(seed_r)
    M, C = 256, 768
    x32, xb = rnd(M, C, dtype=torch.float32, seed=51), rnd(M, C, seed=52)
    res = rnd(M, C, dtype=torch.float32, seed=53)
    alpha = torch.tensor([0.25], device=DEV)
    for x in (x32, xb):
        def run(k):
            o32, o16 = torch.zeros(M, C, device=DEV), torch.zeros(M, C, dtype=torch.bfloat16, device=DEV)
            k.dropout_add(x, res, 0.1, seed, 77, out_f32=o32, out_bf16=o16, scale=2.0, scale_dev=alpha)
            g32, g16 = torch.zeros(M, C, device=DEV), torch.zeros(M, C, dtype=torch.bfloat16, device=DEV)
            k.dropout_bwd(x, 0.1, seed, 77, out_f32=g32, out_bf16=g16)
            return o32, o16, g32, g16
        (o32, o16, g32, g16), (o32r, o16r, g32r, g16r) = _both(K, R, run)
        assert ((o16 == 0) == (o16r == 0)).all() and ((g32 == 0) == (g32r == 0)).all(), "dropout mask differs from Philox restatement"
        assert abs((o16 == 0).float().mean().item() - 0.1) < 0.01
        check(o32, o32r, 1e-5, "dropout_add f32")
        check(o16, o16r, 4e-3, "dropout_add bf16")
        check(g32, g32r, 1e-5, "dropout_bwd f32")
        check(g16, g16r, 4e-3, "dropout_bwd bf16")
    for (B, S, H, p) in ((8, 32, 12, 0.1), (3, 17, 2, 0.1), (2, 64, 12, 0.0)):
        Cq = H * 64
        qkv = rnd(B, S, 3 * Cq, seed=54)
        kb = torch.zeros(B, S, device=DEV)
        kb[0, S // 2:] = torch.finfo(torch.float32).min
        d_o = rnd(B, S, Cq, seed=55)

        def run_attn(k):
            o, lse = torch.zeros(B, S, Cq, dtype=torch.bfloat16, device=DEV), torch.zeros(B * H * S, device=DEV)
            k.text_attention_fwd(qkv[:, :, :Cq], qkv[:, :, Cq:2 * Cq], qkv[:, :, 2 * Cq:], kb, 0.125, p, seed, 5, H, o, lse)
            dqkv = torch.zeros(B, S, 3 * Cq, dtype=torch.bfloat16, device=DEV)
            k.text_attention_bwd(qkv[:, :, :Cq], qkv[:, :, Cq:2 * Cq], qkv[:, :, 2 * Cq:], kb, 0.125, p, seed, 5, H, lse, d_o,
                                 dqkv[:, :, :Cq], dqkv[:, :, Cq:2 * Cq], dqkv[:, :, 2 * Cq:])
            return o, lse, dqkv
        for name, a, b in zip(("o", "lse", "dqkv"), *_both(K, R, run_attn)):
            check(a, b, 1e-5 if name == "lse" else 6e-3, "text attention S=%d p=%g %s" % (S, p, name))
