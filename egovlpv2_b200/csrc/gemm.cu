// egv_gemm_bf16: D = epilogue(A x B), bf16 operands, fp32 accumulation.
//
// Main path (sm_100a): persistent, warp-specialised tcgen05 kernel.
//   warp 0      TMA producer   (cp.async.bulk.tensor 2-D, 128B swizzle, mbarrier complete_tx)
//   warp 1      MMA issuer     (one elected thread, tcgen05.mma cta_group::1 kind::f16, M=128 x N=BN x K=16)
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4-11  epilogue       (tcgen05.ld 32x32b -> smem transpose -> coalesced global I/O, fused
//                               bias / activation / activation-gradient / scale / residual / split-K atomics)
// Operands may be K-major or MN-major (UMMA descriptor "major" bits), which gives the three layouts
// the training step needs without ever materialising a transpose:
//   NT  y  = x W^T   (A K-major, B K-major)     nn.Linear forward
//   NN  dx = dy W    (A K-major, B MN-major)
//   TN  dW = dy^T x  (A MN-major, B MN-major)
// Fallback path: a small SIMT kernel with identical semantics for shapes TMA cannot address
// (row strides that are not multiples of 16 bytes) and for tiny problems.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#define EGV_PDL_CLASS 1
#include "host_common.h"

namespace egv {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 128 + EPI_WARPS * 32;
constexpr int EPI_STAGE_FLOATS = 32 * 32;

struct GemmParams {
  int M, N, K;
  int num_n_tiles, total_items, split_k, k_blocks_total, k_blocks_per_split;
  const float* bias;
  const bf16* aux;
  long long ld_aux;
  const float* scale_dev;
  float scale;
  const float* residual;
  long long ld_res;
  float* out_f32;
  long long ld_out_f32;
  bf16* out_bf16;
  long long ld_out_bf16;
  bf16* out_pre;
  long long ld_out_pre;
  int act;
  int accumulate;
  int vec_ok;  // N % 4 == 0, every row stride % 4 == 0 and every pointer aligned for 4-element vector access
  float* colsum;  // optional [N]: += column sums of the final value (bias gradients), fp32 atomics
  int fast;       // lean full-tile epilogues: 1 / 5 fc1 forward (GELU, pre or derivative saved), 2 / 6 its backward, 3 bf16, 4 residual
  int prefetch_res;   // the producer prefetches each tile's fp32 residual block into L2 while the tile's mainloop runs
};

EGV_DEVINL float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
EGV_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Phi(x) = 0.5 (1 + erf(x / sqrt 2)) with ONE transcendental: Abramowitz-Stegun 7.1.28,
// erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16, |err| <= 3e-7 (far below bf16 resolution); the 1/sqrt(2) is folded
// into the coefficients.  The epilogue of the GELU GEMMs is bound by the XU (MUFU) pipe, hence the form.
EGV_DEVINL float gelu_cdf(float x) {
  const float ax = fabsf(x);
  float d = fmaf(ax, 5.3829750e-6f, 4.8890636e-5f);   // a6/8, a5/(4 sqrt2)
  d = fmaf(d, ax, 3.8003575e-5f);                     // a4/4
  d = fmaf(d, ax, 3.2776263e-3f);                     // a3/(2 sqrt2)
  d = fmaf(d, ax, 2.1141006e-2f);                     // a2/2
  d = fmaf(d, ax, 4.9867347e-2f);                     // a1/sqrt2
  d = fmaf(d, ax, 1.0f);
  float r = rcp_approx(d);
  r *= r;
  r *= r;
  r *= r;
  r *= r;                       // d^-16 = 1 - erf(|x|/sqrt2)
  const float half_erfc = 0.5f * r;
  return x >= 0.f ? 1.0f - half_erfc : half_erfc;
}

// x Phi(x) = relu(x) - |x| * erfc(|x| / sqrt 2) / 2: same polynomial, no select / complement (3 instructions fewer)
EGV_DEVINL float gelu_fwd(float x) {
  const float ax = fabsf(x);
  float d = fmaf(ax, 5.3829750e-6f, 4.8890636e-5f);
  d = fmaf(d, ax, 3.8003575e-5f);
  d = fmaf(d, ax, 3.2776263e-3f);
  d = fmaf(d, ax, 2.1141006e-2f);
  d = fmaf(d, ax, 4.9867347e-2f);
  d = fmaf(d, ax, 1.0f);
  float r = rcp_approx(d);
  r *= r;
  r *= r;
  r *= r;
  r *= r;
  return fmaf(-0.5f * ax, r, fmaxf(x, 0.0f));
}

// gelu'(x) = Phi(x) + x phi(x)
EGV_DEVINL float gelu_grad(float x) {
  const float pdf = 0.3989422804014327f * ex2_approx(-0.72134752044448170f * x * x);
  return fmaf(x, pdf, gelu_cdf(x));
}
EGV_DEVINL bool act_needs_aux(int act) { return (act >= EGV_ACT_GELU_BWD && act <= EGV_ACT_TANH_BWD) || act == EGV_ACT_MUL_AUX; }

template <int ACT>
EGV_DEVINL float apply_act(float v, float auxv) {
  if (ACT == EGV_ACT_GELU || ACT == EGV_ACT_GELU_DG) return gelu_fwd(v);
  if (ACT == EGV_ACT_MUL_AUX) return v * auxv;
  if (ACT == EGV_ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == EGV_ACT_TANH) return tanhf(v);
  if (ACT == EGV_ACT_GELU_BWD) {
    // gelu'(x) = Phi(x) + x phi(x),  phi(x) = exp(-x^2/2) / sqrt(2 pi)
    const float pdf = 0.3989422804014327f * ex2_approx(-0.72134752044448170f * auxv * auxv);
    return v * fmaf(auxv, pdf, gelu_cdf(auxv));
  }
  if (ACT == EGV_ACT_RELU_BWD) return auxv > 0.0f ? v : 0.0f;
  if (ACT == EGV_ACT_TANH_BWD) return v * (1.0f - auxv * auxv);
  return v;
}

// Epilogue math for one element (scalar path: SIMT kernel and odd-shaped tensor-core problems).
// `lead` is false for split-K slices > 0 (they only add their partial sum).
template <int ACT>
EGV_DEVINL float epilogue_store_t(const GemmParams& p, float v, int row, int col, bool lead, float resv, float auxv,
                                  float scale_total) {
  if (lead && p.bias) v += __ldg(p.bias + col);
  if (p.out_pre) p.out_pre[(long long)row * p.ld_out_pre + col] = __float2bfloat16(ACT == EGV_ACT_GELU_DG ? gelu_grad(v) : v);
  v = apply_act<ACT>(v, auxv) * scale_total;
  if (lead) v += resv;
  if (p.out_f32) {
    float* o = p.out_f32 + (long long)row * p.ld_out_f32 + col;
    if (p.accumulate) atomicAdd(o, v);
    else *o = v;
  }
  if (p.out_bf16) p.out_bf16[(long long)row * p.ld_out_bf16 + col] = __float2bfloat16(v);
  return v;
}
EGV_DEVINL float epilogue_store(const GemmParams& p, float v, int row, int col, bool lead, float resv, float auxv,
                                float scale_total) {
  switch (p.act) {
    case EGV_ACT_GELU: return epilogue_store_t<EGV_ACT_GELU>(p, v, row, col, lead, resv, auxv, scale_total);
    case EGV_ACT_RELU: return epilogue_store_t<EGV_ACT_RELU>(p, v, row, col, lead, resv, auxv, scale_total);
    case EGV_ACT_TANH: return epilogue_store_t<EGV_ACT_TANH>(p, v, row, col, lead, resv, auxv, scale_total);
    case EGV_ACT_GELU_BWD: return epilogue_store_t<EGV_ACT_GELU_BWD>(p, v, row, col, lead, resv, auxv, scale_total);
    case EGV_ACT_RELU_BWD: return epilogue_store_t<EGV_ACT_RELU_BWD>(p, v, row, col, lead, resv, auxv, scale_total);
    case EGV_ACT_TANH_BWD: return epilogue_store_t<EGV_ACT_TANH_BWD>(p, v, row, col, lead, resv, auxv, scale_total);
    case EGV_ACT_GELU_DG: return epilogue_store_t<EGV_ACT_GELU_DG>(p, v, row, col, lead, resv, auxv, scale_total);
    case EGV_ACT_MUL_AUX: return epilogue_store_t<EGV_ACT_MUL_AUX>(p, v, row, col, lead, resv, auxv, scale_total);
    default: return epilogue_store_t<EGV_ACT_NONE>(p, v, row, col, lead, resv, auxv, scale_total);
  }
}

// Staging tile of one epilogue warp: 32 rows x 32 fp32, row stride 32 words, 16-byte chunks XOR-swizzled with the
// row index (chunk' = chunk ^ (row & 7)): both the row-per-thread writes and the 4-rows-per-instruction reads are
// conflict-free without padding.
EGV_DEVINL int stg_off(int row, int chunk) { return row * 32 + ((chunk ^ (row & 7)) << 2); }

EGV_DEVINL uint2 pack4_bf16(float a, float b, float c, float d) {
  uint2 u;
  u.x = pack_bf16(a, b);
  u.y = pack_bf16(c, d);
  return u;
}

// One 32-row x 32-column chunk, read back 4 rows per instruction: lane l owns columns 4*(l&7)..+3 of rows 4i+(l>>3).
// FULL: the tile has no row / column tail (no predicates in the loop).
template <int ACT, bool FULL>
EGV_DEVINL void epi_rows_vec4(const GemmParams& p, const float* stg, int lane, int row_base, int gcol, bool lead,
                              float scale_total) {
  const int sub = lane >> 3, ch = lane & 7;
  const bool c_ok = FULL || gcol < p.N;   // N % 4 == 0 on this path: the whole 4-column group is in or out
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lead && p.bias && c_ok) b = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
  constexpr bool NEED_AUX = (ACT >= EGV_ACT_GELU_BWD && ACT <= EGV_ACT_TANH_BWD) || ACT == EGV_ACT_MUL_AUX;
  const bool has_res = lead && p.residual != nullptr;
  const int rows_left = p.M - row_base;   // rows of this chunk that exist (>= 32 when FULL)
  const float* res_p = p.residual ? p.residual + (long long)row_base * p.ld_res + gcol : nullptr;
  const bf16* aux_p = p.aux ? p.aux + (long long)row_base * p.ld_aux + gcol : nullptr;
  float* o32 = p.out_f32 ? p.out_f32 + (long long)row_base * p.ld_out_f32 + gcol : nullptr;
  bf16* o16 = p.out_bf16 ? p.out_bf16 + (long long)row_base * p.ld_out_bf16 + gcol : nullptr;
  bf16* opre = p.out_pre ? p.out_pre + (long long)row_base * p.ld_out_pre + gcol : nullptr;
  const int ld_res = (int)p.ld_res, ld_aux = (int)p.ld_aux, ld32 = (int)p.ld_out_f32, ld16 = (int)p.ld_out_bf16,
            ldpre = (int)p.ld_out_pre;
  // all residual / aux loads of the chunk are issued before the first use (memory-level parallelism)
  float4 rs[8];
  uint2 ax[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = 4 * i + sub;
    const bool ok = c_ok && (FULL || rl < rows_left);
    rs[i] = (has_res && ok) ? __ldg(reinterpret_cast<const float4*>(res_p + rl * ld_res)) : make_float4(0.f, 0.f, 0.f, 0.f);
    ax[i] = (NEED_AUX && ok) ? __ldg(reinterpret_cast<const uint2*>(aux_p + rl * ld_aux)) : make_uint2(0u, 0u);
  }
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = 4 * i + sub;
    const float4 a = *reinterpret_cast<const float4*>(stg + stg_off(rl, ch));
    if (!FULL && !(c_ok && rl < rows_left)) continue;
    float v0 = a.x + b.x, v1 = a.y + b.y, v2 = a.z + b.z, v3 = a.w + b.w;
    if (opre) {
      if (ACT == EGV_ACT_GELU_DG) *reinterpret_cast<uint2*>(opre + rl * ldpre) = pack4_bf16(gelu_grad(v0), gelu_grad(v1), gelu_grad(v2), gelu_grad(v3));
      else *reinterpret_cast<uint2*>(opre + rl * ldpre) = pack4_bf16(v0, v1, v2, v3);
    }
    const float2 a01 = unpack_bf16(ax[i].x), a23 = unpack_bf16(ax[i].y);
    v0 = fmaf(apply_act<ACT>(v0, a01.x), scale_total, rs[i].x);
    v1 = fmaf(apply_act<ACT>(v1, a01.y), scale_total, rs[i].y);
    v2 = fmaf(apply_act<ACT>(v2, a23.x), scale_total, rs[i].z);
    v3 = fmaf(apply_act<ACT>(v3, a23.y), scale_total, rs[i].w);
    if (o32) {
      float* o = o32 + rl * ld32;
      if (p.accumulate) {
        // one 16-byte reduction instead of four scalar ones (split-K slices / gradient accumulation)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
      } else {
        *reinterpret_cast<float4*>(o) = make_float4(v0, v1, v2, v3);
      }
    }
    if (o16) *reinterpret_cast<uint2*>(o16 + rl * ld16) = pack4_bf16(v0, v1, v2, v3);
    cs.x += v0;
    cs.y += v1;
    cs.z += v2;
    cs.w += v3;
  }
  if (p.colsum) {
    // rows 4i+sub for sub = 0..3 live in lanes l, l^8, l^16, l^24
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o);
      cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o);
      cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
    }
    if (sub == 0 && c_ok) {
      atomicAdd(p.colsum + gcol, cs.x);
      atomicAdd(p.colsum + gcol + 1, cs.y);
      atomicAdd(p.colsum + gcol + 2, cs.z);
      atomicAdd(p.colsum + gcol + 3, cs.w);
    }
  }
}

// Lean epilogues of the two activation GEMMs of the MLP (fc1 forward: 118 GF with K = 768, i.e. the epilogue has to
// keep up with a 5 us mainloop per 128 x 256 tile): full tiles only, no optional outputs, pointers advanced instead of
// re-derived, no per-row uniform branches.  p.fast == 1: out_pre = bf16(acc + bias), out_bf16 = bf16(gelu(acc + bias)).
// DG: out_pre = bf16(gelu'(acc + bias)) instead (EGV_ACT_GELU_DG): Phi is shared between the value and the derivative.
template <bool DG>
EGV_DEVINL void epi_rows_gelu_fwd_fast(const GemmParams& p, const float* stg, int lane, int row_base, int gcol) {
  const int sub = lane >> 3, ch = lane & 7;
  const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
  bf16* opre = p.out_pre + (long long)(row_base + sub) * p.ld_out_pre + gcol;
  bf16* oact = p.out_bf16 + (long long)(row_base + sub) * p.ld_out_bf16 + gcol;
  const long long spre = 4 * p.ld_out_pre, sact = 4 * p.ld_out_bf16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(stg + stg_off(4 * i + sub, ch));
    const float v0 = a.x + b.x, v1 = a.y + b.y, v2 = a.z + b.z, v3 = a.w + b.w;
    if (DG) {
      const float c0 = gelu_cdf(v0), c1 = gelu_cdf(v1), c2 = gelu_cdf(v2), c3 = gelu_cdf(v3);
      const float k = -0.72134752044448170f, n = 0.3989422804014327f;
      *reinterpret_cast<uint2*>(opre) = pack4_bf16(fmaf(v0 * n, ex2_approx(k * v0 * v0), c0), fmaf(v1 * n, ex2_approx(k * v1 * v1), c1),
                                                   fmaf(v2 * n, ex2_approx(k * v2 * v2), c2), fmaf(v3 * n, ex2_approx(k * v3 * v3), c3));
      *reinterpret_cast<uint2*>(oact) = pack4_bf16(v0 * c0, v1 * c1, v2 * c2, v3 * c3);
    } else {
      *reinterpret_cast<uint2*>(opre) = pack4_bf16(v0, v1, v2, v3);
      *reinterpret_cast<uint2*>(oact) = pack4_bf16(gelu_fwd(v0), gelu_fwd(v1), gelu_fwd(v2), gelu_fwd(v3));
    }
    opre += spre;
    oact += sact;
  }
}
// p.fast == 2: out_bf16 = bf16(acc * gelu'(aux)), colsum (optional) += column sums;  MUL (p.fast == 6): acc * aux
template <bool MUL>
EGV_DEVINL void epi_rows_gelu_bwd_fast(const GemmParams& p, const float* stg, int lane, int row_base, int gcol) {
  const int sub = lane >> 3, ch = lane & 7;
  const bf16* aux = p.aux + (long long)(row_base + sub) * p.ld_aux + gcol;
  bf16* o16 = p.out_bf16 + (long long)(row_base + sub) * p.ld_out_bf16 + gcol;
  const long long saux = 4 * p.ld_aux, s16 = 4 * p.ld_out_bf16;
  uint2 ax[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ax[i] = __ldg(reinterpret_cast<const uint2*>(aux + i * saux));
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(stg + stg_off(4 * i + sub, ch));
    const float2 a01 = unpack_bf16(ax[i].x), a23 = unpack_bf16(ax[i].y);
    constexpr int A = MUL ? EGV_ACT_MUL_AUX : EGV_ACT_GELU_BWD;
    const float v0 = apply_act<A>(a.x, a01.x), v1 = apply_act<A>(a.y, a01.y);
    const float v2 = apply_act<A>(a.z, a23.x), v3 = apply_act<A>(a.w, a23.y);
    *reinterpret_cast<uint2*>(o16) = pack4_bf16(v0, v1, v2, v3);
    o16 += s16;
    cs.x += v0;
    cs.y += v1;
    cs.z += v2;
    cs.w += v3;
  }
  if (p.colsum) {
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o);
      cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o);
      cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
    }
    if (sub == 0) {
      atomicAdd(p.colsum + gcol, cs.x);
      atomicAdd(p.colsum + gcol + 1, cs.y);
      atomicAdd(p.colsum + gcol + 2, cs.z);
      atomicAdd(p.colsum + gcol + 3, cs.w);
    }
  }
}

// p.fast == 3: out_bf16 = bf16((acc [+ bias]) * scale)   (qkv projections, input gradients)
EGV_DEVINL void epi_rows_bf16_fast(const GemmParams& p, const float* stg, int lane, int row_base, int gcol, float sc) {
  const int sub = lane >> 3, ch = lane & 7;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
  bf16* o16 = p.out_bf16 + (long long)(row_base + sub) * p.ld_out_bf16 + gcol;
  const long long s16 = 4 * p.ld_out_bf16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(stg + stg_off(4 * i + sub, ch));
    *reinterpret_cast<uint2*>(o16) = pack4_bf16((a.x + b.x) * sc, (a.y + b.y) * sc, (a.z + b.z) * sc, (a.w + b.w) * sc);
    o16 += s16;
  }
}
// p.fast == 4: v = acc [+ bias]; out_pre (optional) = bf16(v); out_f32 = v * scale + residual   (out-projections)
EGV_DEVINL void epi_rows_residual_fast(const GemmParams& p, const float* stg, int lane, int row_base, int gcol,
                                       float scale_total) {
  const int sub = lane >> 3, ch = lane & 7;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
  const float* res = p.residual + (long long)(row_base + sub) * p.ld_res + gcol;
  float* o32 = p.out_f32 + (long long)(row_base + sub) * p.ld_out_f32 + gcol;
  const long long sres = 4 * p.ld_res, s32 = 4 * p.ld_out_f32;
  float4 rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) rs[i] = __ldg(reinterpret_cast<const float4*>(res + i * sres));
  bf16* opre = p.out_pre ? p.out_pre + (long long)(row_base + sub) * p.ld_out_pre + gcol : nullptr;
  const long long spre = 4 * p.ld_out_pre;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(stg + stg_off(4 * i + sub, ch));
    const float v0 = a.x + b.x, v1 = a.y + b.y, v2 = a.z + b.z, v3 = a.w + b.w;
    if (opre) {
      *reinterpret_cast<uint2*>(opre) = pack4_bf16(v0, v1, v2, v3);
      opre += spre;
    }
    *reinterpret_cast<float4*>(o32) = make_float4(fmaf(v0, scale_total, rs[i].x), fmaf(v1, scale_total, rs[i].y),
                                                  fmaf(v2, scale_total, rs[i].z), fmaf(v3, scale_total, rs[i].w));
    o32 += s32;
  }
}

template <int ACT>
EGV_DEVINL void epi_rows_vec4_d(const GemmParams& p, const float* stg, int lane, int row_base, int gcol, bool lead,
                                float scale_total, bool full) {
  if (full) epi_rows_vec4<ACT, true>(p, stg, lane, row_base, gcol, lead, scale_total);
  else epi_rows_vec4<ACT, false>(p, stg, lane, row_base, gcol, lead, scale_total);
}

// Scalar read-back of the same staging tile: lane = column, one row per iteration (odd N / unaligned tensors).
EGV_DEVINL void epi_rows_scalar(const GemmParams& p, const float* stg, int lane, int row_base, int gcol, bool lead,
                                float scale_total) {
  const bool c_ok = gcol < p.N;
  const bool need_aux = act_needs_aux(p.act);
  float cs = 0.f;
#pragma unroll 4
  for (int r = 0; r < 32; ++r) {
    const int row = row_base + r;
    const float acc = stg[stg_off(r, lane >> 2) + (lane & 3)];
    if (!(c_ok && row < p.M)) continue;
    const float resv = (lead && p.residual) ? p.residual[(long long)row * p.ld_res + gcol] : 0.0f;
    const float auxv = need_aux ? __bfloat162float(p.aux[(long long)row * p.ld_aux + gcol]) : 0.0f;
    cs += epilogue_store(p, acc, row, gcol, lead, resv, auxv, scale_total);
  }
  if (p.colsum && c_ok) atomicAdd(p.colsum + gcol, cs);
}

template <int BN, int CL = 1>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / CL) * BK * 2;   // a CTA pair splits B's N rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = CL == 2 ? (BN == 256 ? 6 : 8) : (BN == 256 ? 4 : (BN == 192 ? 4 : (BN == 128 ? 6 : 8)));
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_STAGE_FLOATS * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SLACK = 1024;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + SLACK;
  static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);   // allocation: power of two
};

// CL = CTAs per cluster (1 or 2).  With CL == 2 the two CTAs of a cluster form a tcgen05 CTA pair (cta_group::2) on a
// 256 x BN output tile (m-tiles 2i and 2i+1 of the same n-tile): each CTA stages its own 128 rows of A and its own half
// of B's N rows, the leader (rank 0) issues M = 256 MMAs for both, each CTA's epilogue drains its own TMEM; per SM the
// operand traffic through shared memory and from L2 drops from A + B to A + B/2 per k-block.  Bit-identical to the
// single-CTA kernel (tests/test_kernels_gpu.py::test_gemm_cluster_pairs_match_single_cta).  Measured (round 1,
// tools/gemm_bench.py, M or K = 25096): weight gradients (TN) 79 / 87 / 85 us vs 85 / 95 / 95 us single-CTA (+7-11 %);
// NT / NN with K = 768 equal, with K = 3072 5 % slower -> default: pairs for the TN layout only (EGV_GEMM_CLUSTER=2).
// (First version: 2x slower -- remote mbarrier arrivals with the default .release.cluster semantics compile to
// MEMBAR + ERRBAR and wait for the epilogue's outstanding global stores; they are .relaxed now.)
template <int BN, bool A_MN, bool B_MN, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_res, const GemmParams p) {
  using Cfg = GemmCfg<BN, CL>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // keep the shared address space visible to the compiler (pointer + offset, no integer round trip)
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  if (pad > (uint32_t)Cfg::SLACK) {
    if (threadIdx.x == 0) printf("egv: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* smem = smem_raw + pad;
  float* epi_smem = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
  const int first_item = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
  const int item_stride = CL > 1 ? (int)(gridDim.x / CL) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], CL);    // pair: one arrive.expect_tx per CTA, both on the LEADER's barrier
      mbar_init(&empty_bar[s], 1);    // pair: the leader's commit is multicast to both CTAs' barriers
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], EPI_WARPS * CL);   // pair: both CTAs' epilogue warps release the leader's accumulator
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CL == 2) tmem_alloc_cg2<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) {   // barrier inits must be visible to the peer before it multicasts into / arrives on them
    cluster_arrive();
    cluster_wait();
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers initialised, tensor memory allocated: from here on the kernel touches global memory (programmatic dependent
  // launch: the predecessor's results must be complete; the successor may start its own prologue)
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = first_item; w < p.total_items; w += item_stride) {
        const int ks = w % p.split_k;
        const int tile = w / p.split_k;
        const int m0 = ((tile / p.num_n_tiles) * CL + (int)cta_rank) * BM;
        const int n0 = (tile % p.num_n_tiles) * BN;
        const int kb0 = ks * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        // The epilogue of this tile will read its fp32 residual block with ordinary loads, 4 KB in flight per warp: from
        // HBM that is latency-bound (measured 3.6 TB/s on the out-projections).  Pull the block into L2 now.
        if (p.prefetch_res && ks == 0) tma_prefetch_l2_2d(&tmap_res, n0, m0);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_sleep(&empty_bar[stage], phase ^ 1, 64);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const int k0 = kb * BK;
          if (CL == 2) {
            // CTA pair: this CTA's 128 rows of A and its half of B's N rows land in ITS shared memory; every byte is
            // accounted on the LEADER's full barrier, which gates the leader's M = 256 MMAs.
            const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0u);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            else mbar_arrive_expect_tx_cluster(fb, Cfg::STAGE_BYTES);
            if (A_MN) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d_cg2(sa + j * 8192, &tmap_a, fb, m0 + 64 * j, k0);
            } else {
              tma_load_2d_cg2(sa, &tmap_a, fb, k0, m0);
            }
            const int nh = n0 + (int)cta_rank * (BN / 2);
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) tma_load_2d_cg2(sb + j * 8192, &tmap_b, fb, nh + 64 * j, k0);
            } else {
              tma_load_2d_cg2(sb, &tmap_b, fb, k0, nh);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            if (A_MN) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &tmap_a, &full_bar[stage], m0 + 64 * j, k0);
            } else {
              tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);
            }
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[stage], n0 + 64 * j, k0);
            } else {
              tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // elect.sync, not `lane == 0`: ptxas then knows one thread runs the region and emits bare UTCHMMA / UTCBAR instead of
    // an ELECT / PLOP3 / BRA.U.ANY convergence loop around each of them (see attention_tc_bwd.cu)
    if (cta_rank == 0 && elect_one()) {   // pair: the leader issues for both CTAs
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CL, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // encoded (>>4) start-address advance per UMMA_K step
      constexpr uint32_t a_adv = A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      constexpr uint32_t b_adv = B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = first_item; w < p.total_items; w += item_stride, ++it) {
        const int ks = w % p.split_k;
        const int kb0 = ks * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        const int as = it & 1;
        const uint32_t use = (uint32_t)(it >> 1) & 1u;
        mbar_wait_sleep(&tmem_empty[as], use ^ 1, 32);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_sleep(&full_bar[stage], phase, 20);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = A_MN ? umma_desc_sw128(sa, 8192, 1024) : umma_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = B_MN ? umma_desc_sw128(sb, 8192, 1024) : umma_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if (CL == 2)
              umma_bf16_cg2(d_tmem, adesc + (uint64_t)(k * a_adv), bdesc + (uint64_t)(k * b_adv), idesc,
                            (kb > kb0 || k > 0) ? 1u : 0u);
            else
              umma_bf16(d_tmem, adesc + (uint64_t)(k * a_adv), bdesc + (uint64_t)(k * b_adv), idesc,
                        (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // smem slot reusable once these MMAs retire (pair: in both CTAs)
          if (CL == 2) umma_commit_cg2(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue (pair: of both CTAs)
        if (CL == 2) umma_commit_cg2(&tmem_full[as]);
        else umma_commit(&tmem_full[as]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int quad = warp & 3;       // TMEM lane quadrant this warp may read (warp id % 4)
    const int chalf = ew >> 2;       // which half of the BN columns
    float* stg = epi_smem + ew * EPI_STAGE_FLOATS;
    const float scale_total = p.scale * (p.scale_dev ? __ldg(p.scale_dev) : 1.0f);
    int it = 0;
    for (int w = first_item; w < p.total_items; w += item_stride, ++it) {
      const int ks = w % p.split_k;
      const int tile = w / p.split_k;
      const int m0 = ((tile / p.num_n_tiles) * CL + (int)cta_rank) * BM;
      const int n0 = (tile % p.num_n_tiles) * BN;
      const int as = it & 1;
      const uint32_t use = (uint32_t)(it >> 1) & 1u;
      const bool lead = (ks == 0);
      mbar_wait(&tmem_full[as], use);
      tc_fence_after();
      const int row_base = m0 + quad * 32;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int col0 = chalf * (BN / 2) + c * 32;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + col0), v);
        tmem_ld_wait();
        // thread = row: 32 consecutive columns -> swizzled staging tile (8 x 16-byte stores)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stg + stg_off(lane, j)) =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                          __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        if (p.act == 99) {
          // debug: mainloop-only timing (tools/gemm_bench.py) -- consume the tile without any global traffic
          if (stg[lane * 32] == 1.2345e38f) p.out_f32[0] = 0.f;
        } else if (p.vec_ok) {
          const int gcol = n0 + col0 + 4 * (lane & 7);
          const bool full = (row_base + 32 <= p.M) && (n0 + col0 + 32 <= p.N);
          if (full && p.fast == 1) {
            epi_rows_gelu_fwd_fast<false>(p, stg, lane, row_base, gcol);
          } else if (full && p.fast == 5) {
            epi_rows_gelu_fwd_fast<true>(p, stg, lane, row_base, gcol);
          } else if (full && p.fast == 2) {
            epi_rows_gelu_bwd_fast<false>(p, stg, lane, row_base, gcol);
          } else if (full && p.fast == 6) {
            epi_rows_gelu_bwd_fast<true>(p, stg, lane, row_base, gcol);
          } else if (full && p.fast == 3) {
            epi_rows_bf16_fast(p, stg, lane, row_base, gcol, scale_total);
          } else if (full && p.fast == 4 && lead) {
            epi_rows_residual_fast(p, stg, lane, row_base, gcol, scale_total);
          } else
          switch (p.act) {
            case EGV_ACT_GELU: epi_rows_vec4_d<EGV_ACT_GELU>(p, stg, lane, row_base, gcol, lead, scale_total, full); break;
            case EGV_ACT_RELU: epi_rows_vec4_d<EGV_ACT_RELU>(p, stg, lane, row_base, gcol, lead, scale_total, full); break;
            case EGV_ACT_TANH: epi_rows_vec4_d<EGV_ACT_TANH>(p, stg, lane, row_base, gcol, lead, scale_total, full); break;
            case EGV_ACT_GELU_BWD: epi_rows_vec4_d<EGV_ACT_GELU_BWD>(p, stg, lane, row_base, gcol, lead, scale_total, full); break;
            case EGV_ACT_RELU_BWD: epi_rows_vec4_d<EGV_ACT_RELU_BWD>(p, stg, lane, row_base, gcol, lead, scale_total, full); break;
            case EGV_ACT_TANH_BWD: epi_rows_vec4_d<EGV_ACT_TANH_BWD>(p, stg, lane, row_base, gcol, lead, scale_total, full); break;
            case EGV_ACT_GELU_DG: epi_rows_vec4_d<EGV_ACT_GELU_DG>(p, stg, lane, row_base, gcol, lead, scale_total, full); break;
            case EGV_ACT_MUL_AUX: epi_rows_vec4_d<EGV_ACT_MUL_AUX>(p, stg, lane, row_base, gcol, lead, scale_total, full); break;
            default: epi_rows_vec4_d<EGV_ACT_NONE>(p, stg, lane, row_base, gcol, lead, scale_total, full); break;
          }
        } else {
          epi_rows_scalar(p, stg, lane, row_base, n0 + col0 + lane, lead, scale_total);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL == 2 && cta_rank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[as]), 0u));   // the leader's barrier
        else mbar_arrive(&tmem_empty[as]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) {   // no CTA may exit while its peer can still multicast into its smem or arrive on its barriers
    cluster_arrive();
    cluster_wait();
  }
  if (warp == 2) {
    tc_fence_after();
    if (CL == 2) tmem_dealloc_cg2<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// --------------------------------------------------------------------------------- SIMT fallback
// C[m,n] = sum_k A(m,k) B(n,k) with arbitrary element strides; 32x32 tile, 16x16 threads, 2x2 per thread.
struct SimtStrides {
  long long a_m, a_k, b_n, b_k;
};
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, SimtStrides st, const GemmParams p) {
  pdl_enter();
  __shared__ float sA[32][33];
  __shared__ float sB[32][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int ks = blockIdx.z;
  const int k_lo = ks * p.k_blocks_per_split;  // here: elements per split
  const int k_hi = min(p.K, k_lo + p.k_blocks_per_split);
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = k_lo; k0 < k_hi; k0 += 32) {
    for (int i = threadIdx.x; i < 1024; i += 256) {
      int r, kk;
      // pick the loop order that keeps the contiguous global dimension on adjacent threads
      if (st.a_k == 1) { r = i >> 5; kk = i & 31; } else { kk = i >> 5; r = i & 31; }
      int m = m0 + r, k = k0 + kk;
      sA[r][kk] = (m < p.M && k < k_hi) ? __bfloat162float(A[m * st.a_m + k * st.a_k]) : 0.0f;
      if (st.b_k == 1) { r = i >> 5; kk = i & 31; } else { kk = i >> 5; r = i & 31; }
      int n = n0 + r;
      k = k0 + kk;
      sB[r][kk] = (n < p.N && k < k_hi) ? __bfloat162float(B[n * st.b_n + k * st.b_k]) : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      float a0 = sA[ty][kk], a1 = sA[ty + 16][kk];
      float b0 = sB[tx][kk], b1 = sB[tx + 16][kk];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
  const float scale_total = p.scale * (p.scale_dev ? __ldg(p.scale_dev) : 1.0f);
  const bool lead = ks == 0;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int row = m0 + ty + 16 * i, col = n0 + tx + 16 * j;
      if (row < p.M && col < p.N) {
        float resv = (p.residual && lead) ? p.residual[(long long)row * p.ld_res + col] : 0.0f;
        float auxv = act_needs_aux(p.act) ? __bfloat162float(p.aux[(long long)row * p.ld_aux + col]) : 0.0f;
        const float fv = epilogue_store(p, acc[i][j], row, col, lead, resv, auxv, scale_total);
        if (p.colsum) atomicAdd(p.colsum + col, fv);
      }
    }
}

// --------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t inner, outer, ld;
  uint32_t box_inner, box_outer;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
           box_outer == o.box_outer;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= std::hash<uint64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.ld); mix(k.box_inner); mix(k.box_outer);
    return h;
  }
};

// bf16 2-D tensor map: `inner` contiguous elements, `outer` rows `ld` elements apart, 128B swizzle.
int get_tensor_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                   uint32_t box_outer, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, inner, outer, ld, box_inner, box_outer};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return EGV_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(EGV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(EGV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%llu outer=%llu ld=%llu", (int)r, ptr,
                (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() > 8192) cache.clear();
  cache[key] = *out;
  return EGV_OK;
}

// tensor map of rank 2..4: dims[0] = contiguous elements, byte strides of dims 1.. in strides[0..rank-2], box extents per
// dimension.  f32 = 0: bf16 elements, 128B swizzle (box[0] must be 64 elements = one swizzle row: MMA operand tiles);
// f32 = 1: fp32 elements, no swizzle (L2 prefetch boxes of the residual stream).
int get_tensor_map_nd_ex(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                         int f32, CUtensorMap* out) {
  struct Key {
    const void* ptr;
    int rank, f32;
    uint64_t d[4], s[3];
    uint32_t b[4];
    bool operator==(const Key& o) const { return memcmp(this, &o, sizeof(Key)) == 0; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
      size_t h = 1469598103934665603ull;
      for (size_t i = 0; i < sizeof(Key) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
      return h;
    }
  };
  static std::mutex mu;
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  if (rank < 2 || rank > 4) return fail(EGV_ERR_ARG, "tensor map rank %d", rank);
  Key key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr;
  key.rank = rank;
  key.f32 = f32;
  for (int i = 0; i < rank; ++i) {
    key.d[i] = dims[i];
    key.b[i] = box[i];
    if (i + 1 < rank) key.s[i] = strides_bytes[i];
  }
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return EGV_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(EGV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t d[4], st[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) st[i] = strides_bytes[i];
  }
  CUresult r = enc(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                   const_cast<void*>(ptr), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   f32 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EGV_ERR_CUDA, "cuTensorMapEncodeTiled (rank %d) failed (%d) ptr=%p", rank, (int)r, ptr);
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() > 8192) cache.clear();
  cache[key] = *out;
  return EGV_OK;
}

int get_tensor_map_nd(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMap* out) {
  return get_tensor_map_nd_ex(ptr, rank, dims, strides_bytes, box, 0, out);
}

// fp32 2-D tensor map without swizzle (L2 prefetch boxes of the residual stream)
static int get_tensor_map_f32(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                              CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, inner, outer, ld, box_inner, box_outer};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return EGV_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(EGV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EGV_ERR_CUDA, "cuTensorMapEncodeTiled (f32) failed (%d)", (int)r);
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() > 8192) cache.clear();
  cache[key] = *out;
  return EGV_OK;
}

template <int BN, bool A_MN, bool B_MN, int CL>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CL>;
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, CL>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "gemm smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int units = sm_count() / CL;   // persistent: one CTA (or CTA pair) per SM (pair)
  const int grid = (p.total_items < units ? p.total_items : units) * CL;
  cudaError_t e = CL == 1 ? launch_k(kern, dim3((unsigned)grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, ta, tb, tr, p)
                          : launch_cluster_k(kern, CL, dim3((unsigned)grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, ta, tb, tr, p);
  if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
  return check_launch("gemm_tc_kernel");
}

template <int BN, int CL>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr,
                          const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_tc<BN, false, false, CL>(ta, tb, tr, p, s);
  if (!a_mn && b_mn) return launch_tc<BN, false, true, CL>(ta, tb, tr, p, s);
  if (a_mn && b_mn) return launch_tc<BN, true, true, CL>(ta, tb, tr, p, s);
  return launch_tc<BN, true, false, CL>(ta, tb, tr, p, s);
}

static int g_force_simt = 0;
static int g_plan_mode = -1;      // -1: env EGV_GEMM_PLAN (default 1); 0: round-1 heuristic; 1: cost-model planner; 2: + 192-wide tiles
static int g_cluster_mode = -1;   // -1: env EGV_GEMM_CLUSTER (default 2); 0 = single CTAs, 1 = CTA pairs wherever possible, 2 = pairs for TN (weight gradients)

}  // namespace egv

using namespace egv;

extern "C" void egv_gemm_force_simt(int on) { g_force_simt = on; }
extern "C" void egv_gemm_set_cluster(int mode) { g_cluster_mode = mode; }
extern "C" void egv_gemm_set_plan(int mode) { g_plan_mode = mode; }

extern "C" int egv_gemm_bf16(const egv_gemm_args* a, egv_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a || !a->A || !a->B) return fail(EGV_ERR_ARG, "gemm: null operand");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return fail(EGV_ERR_ARG, "gemm: bad shape %d %d %d", a->M, a->N, a->K);
  if (a->layout < 0 || a->layout > 2) return fail(EGV_ERR_ARG, "gemm: bad layout %d", a->layout);
  if (!a->out_f32 && !a->out_bf16 && !a->out_pre_bf16) return fail(EGV_ERR_ARG, "gemm: no output");
  if (((a->act >= EGV_ACT_GELU_BWD && a->act <= EGV_ACT_TANH_BWD) || a->act == EGV_ACT_MUL_AUX) && !a->aux)
    return fail(EGV_ERR_ARG, "gemm: act %d needs aux", a->act);
  int split_k = a->split_k < 1 ? 1 : a->split_k;
  if (split_k > 1 && (!a->accumulate || !a->out_f32 || a->out_bf16 || a->out_pre_bf16 || a->act != EGV_ACT_NONE))
    return fail(EGV_ERR_ARG, "gemm: split_k > 1 needs accumulate=1, f32 output only, no activation");

  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias = a->bias;
  p.aux = (const bf16*)a->aux; p.ld_aux = a->ld_aux;
  p.scale_dev = a->scale_dev; p.scale = a->scale;
  p.residual = a->residual; p.ld_res = a->ld_res;
  p.out_f32 = a->out_f32; p.ld_out_f32 = a->ld_out_f32;
  p.out_bf16 = (bf16*)a->out_bf16; p.ld_out_bf16 = a->ld_out_bf16;
  p.out_pre = (bf16*)a->out_pre_bf16; p.ld_out_pre = a->ld_out_pre;
  p.act = a->act; p.accumulate = a->accumulate;
  {
    auto al4 = [](const void* ptr, long long ld, int bytes) { return !ptr || (((uintptr_t)ptr) % bytes == 0 && ld % 4 == 0); };
    p.vec_ok = (a->N % 4 == 0) && (!a->bias || ((uintptr_t)a->bias) % 16 == 0) && al4(a->aux, a->ld_aux, 8) &&
               al4(a->residual, a->ld_res, 16) && al4(a->out_f32, a->ld_out_f32, 16) &&
               al4(a->out_bf16, a->ld_out_bf16, 8) && al4(a->out_pre_bf16, a->ld_out_pre, 8) &&
               (!a->colsum || ((uintptr_t)a->colsum) % 4 == 0);
  }
  p.colsum = a->colsum;
  p.fast = 0;
  p.prefetch_res = 0;
  static const bool no_fast = getenv("EGV_GEMM_NO_FAST") != nullptr;
  if (!no_fast && p.vec_ok && split_k == 1 && !a->accumulate) {
    const bool unit = !a->scale_dev && a->scale == 1.0f;
    if (!a->residual && !a->out_f32 && a->out_bf16) {
      if (unit && a->act == EGV_ACT_GELU && a->bias && a->out_pre_bf16 && !a->colsum) p.fast = 1;
      if (unit && a->act == EGV_ACT_GELU_BWD && !a->bias && !a->out_pre_bf16 && a->aux) p.fast = 2;
      if (unit && a->act == EGV_ACT_GELU_DG && a->bias && a->out_pre_bf16 && !a->colsum) p.fast = 5;
      if (unit && a->act == EGV_ACT_MUL_AUX && !a->bias && !a->out_pre_bf16 && a->aux) p.fast = 6;
      if (a->act == EGV_ACT_NONE && !a->out_pre_bf16 && !a->colsum) p.fast = 3;
    }
    if (a->act == EGV_ACT_NONE && a->residual && a->out_f32 && !a->out_bf16 && !a->colsum) p.fast = 4;
  }
  if (a->colsum && split_k > 1) return fail(EGV_ERR_ARG, "gemm: colsum cannot be combined with split_k > 1");

  const bool a_mn = a->layout == EGV_GEMM_TN;                           // A stored [K, M]
  const bool b_mn = a->layout == EGV_GEMM_NN || a->layout == EGV_GEMM_TN;  // B stored [K, N]
  auto aligned = [](const void* ptr, long long ld) { return (((uintptr_t)ptr) & 15) == 0 && (ld % 8) == 0; };
  bool tma_ok = aligned(a->A, a->lda) && aligned(a->B, a->ldb);
  // MN-major boxes are 64 elements wide along M/N: fine for any extent (OOB zero fill), but the contiguous
  // extent must itself be addressable, which the ld % 8 test already guarantees.
  const bool tiny = (long long)a->M * a->N * a->K < (1ll << 18);
  if (!tma_ok || tiny || g_force_simt) {
    SimtStrides st;
    st.a_m = a_mn ? 1 : a->lda; st.a_k = a_mn ? a->lda : 1;
    st.b_n = b_mn ? 1 : a->ldb; st.b_k = b_mn ? a->ldb : 1;
    // split over K in chunks of >= 256 elements when the output grid alone cannot fill the GPU
    long long tiles = cdiv(a->M, 32) * cdiv(a->N, 32);
    int sk = 1;
    if (a->accumulate && a->out_f32 && !a->out_bf16 && !a->out_pre_bf16 && a->act == EGV_ACT_NONE) {
      sk = split_k;
      if (sk == 1 && tiles < sm_count() && a->K >= 1024) sk = (int)std::min<long long>(cdiv(a->K, 256), 64);
    }
    int per = (int)cdiv(cdiv(a->K, sk), 32) * 32;
    sk = (int)cdiv(a->K, per);
    p.split_k = sk; p.k_blocks_per_split = per; p.k_blocks_total = 0; p.num_n_tiles = 0; p.total_items = 0;
    dim3 grid((unsigned)cdiv(a->N, 32), (unsigned)cdiv(a->M, 32), (unsigned)sk);
    launch_k(gemm_simt_kernel, dim3(grid), dim3(256), 0, stream, (const bf16*)a->A, (const bf16*)a->B, st, p);
    return check_launch("gemm_simt_kernel");
  }

  int BN = a->N > 128 ? 256 : (a->N > 64 ? 128 : 64);
  const int num_m_tiles = (int)cdiv(a->M, BM);
  p.k_blocks_total = (int)cdiv(a->K, BK);
  // Automatic split-K: a plain fp32 output whose tile grid cannot fill the GPU but whose reduction is long
  // (weight gradients dW = dy^T x: 768..3072-wide outputs, K = B*N tokens).  The output is zeroed (unless the caller
  // accumulates anyway) and the slices are combined with fp32 vector reductions.
  bool auto_split = false, planned = false;
  const bool splittable = split_k == 1 && a->out_f32 && !a->out_bf16 && !a->out_pre_bf16 && a->act == EGV_ACT_NONE &&
                          !a->residual && !a->colsum;
  if (g_plan_mode < 0) g_plan_mode = getenv("EGV_GEMM_PLAN") ? atoi(getenv("EGV_GEMM_PLAN")) : 1;
  if (splittable && g_plan_mode >= 1) {
    // Plan (tile width, split) together.  Cost model in units of "one 128x256 k-block": a CTA processes
    // ceil(items / SMs) items of ceil(kb / split) k-blocks each.  Narrow tiles cost less per k-block but run the
    // SS-mode MMA against the shared-memory read limit (128x128: 8 KB per 64 cycles = 128 B/clk, the whole budget),
    // hence the efficiency factors; every item also pays an un-overlapped epilogue (fp32 reductions of its tile).
    if ((long long)num_m_tiles * cdiv(a->N, 256) < 2 * sm_count() && p.k_blocks_total >= 16) {
      const int widths[3] = {256, 192, 128};
      const double kcost[3] = {1.0, 0.75 / 0.92, 0.5 / 0.72};
      const double ecost[3] = {8.0, 6.0, 4.0};
      double best = 1e30;
      int best_bn = BN, best_split = 1;
      for (int wi = 0; wi < 3; ++wi) {
        const int bn = widths[wi];
        if (bn > 128 && a->N <= 128) continue;
        if (bn == 192 && (a->N % 192 != 0 || g_plan_mode < 2)) continue;
        const long long tiles = (long long)num_m_tiles * cdiv(a->N, bn);
        const int max_split = (int)std::max<long long>(1, std::min<long long>(p.k_blocks_total / 8, 64));
        for (int sp = 1; sp <= max_split; ++sp) {
          const long long items = tiles * sp;
          const long long rounds = cdiv(items, sm_count());
          const double cost = (double)rounds * ((double)cdiv(p.k_blocks_total, sp) * kcost[wi] + ecost[wi]);
          if (cost < best * 0.999) {
            best = cost;
            best_bn = bn;
            best_split = sp;
          }
        }
      }
      BN = best_bn;
      split_k = best_split;
      auto_split = split_k > 1;
      planned = true;
    }
  } else if (splittable) {
    if (BN == 256 && (long long)num_m_tiles * cdiv(a->N, 128) <= sm_count()) BN = 128;
    const long long tiles256 = (long long)num_m_tiles * cdiv(a->N, BN);
    if (tiles256 * 2 <= sm_count() && p.k_blocks_total >= 16) {
      long long want = cdiv(sm_count(), tiles256);
      long long cap = p.k_blocks_total / 8;   // >= 8 k-blocks (512 elements) per slice
      split_k = (int)std::max<long long>(1, std::min<long long>(want, cap));
      auto_split = split_k > 1;
    }
  }
  // prefer the 128-wide tile when the 256-wide one would leave most SMs idle
  if (!planned && BN == 256 && (long long)num_m_tiles * cdiv(a->N, 256) * split_k < sm_count() / 2) BN = 128;
  p.num_n_tiles = (int)cdiv(a->N, BN);
  if (split_k > p.k_blocks_total) split_k = p.k_blocks_total;
  p.k_blocks_per_split = (int)cdiv(p.k_blocks_total, split_k);
  split_k = (int)cdiv(p.k_blocks_total, p.k_blocks_per_split);
  p.split_k = split_k;
  // CTA pairs sharing the B tile (TMA multicast) when there are enough tile rows to pair up and fill the GPU
  if (g_cluster_mode < 0) g_cluster_mode = getenv("EGV_GEMM_CLUSTER") ? atoi(getenv("EGV_GEMM_CLUSTER")) : 2;
  const bool pair_wanted = g_cluster_mode == 1 || (g_cluster_mode == 2 && a_mn && b_mn);
  const bool pair = pair_wanted && (BN == 128 || BN == 256) && num_m_tiles >= 2 &&
                    (long long)num_m_tiles * p.num_n_tiles * split_k >= sm_count() / 2;
  const int m_units = pair ? (int)cdiv(num_m_tiles, 2) : num_m_tiles;
  p.total_items = m_units * p.num_n_tiles * split_k;
  if (auto_split && !a->accumulate) {
    if (a->ld_out_f32 == a->N) {
      cudaMemsetAsync(a->out_f32, 0, (size_t)a->M * a->N * sizeof(float), stream);
    } else {
      cudaMemset2DAsync(a->out_f32, (size_t)a->ld_out_f32 * sizeof(float), 0, (size_t)a->N * sizeof(float), (size_t)a->M, stream);
    }
  }
  if (auto_split) p.accumulate = 1;

  CUtensorMap ta, tb;
  int rc;
  if (a_mn) rc = get_tensor_map(a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BK, &ta);
  else rc = get_tensor_map(a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BK, BM, &ta);
  if (rc) return rc;
  if (b_mn) rc = get_tensor_map(a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BK, &tb);
  else rc = get_tensor_map(a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, BK, (uint32_t)(pair ? BN / 2 : BN), &tb);
  if (rc) return rc;

  // L2 prefetch of the fp32 residual blocks (one TMA prefetch per tile, issued by the producer; a hint: results are
  // unaffected).  OFF by default: measured on B200 (round 2, tools/bgemm_vs_gemm.py) the out-projection shape
  // M = 25096, N = 768, K = 384 with an fp32 residual got SLOWER (49.2 -> 55.3 us) and the step lost 2 % -- the prefetch
  // competes with the operand stream for the same L2 -> SM path instead of hiding latency.  EGV_GEMM_RES_PREFETCH=1 enables it.
  CUtensorMap tr = ta;
  p.prefetch_res = 0;
  static const bool no_prefetch = getenv("EGV_GEMM_RES_PREFETCH") == nullptr;
  if (!no_prefetch && a->residual && (((uintptr_t)a->residual) & 15) == 0 && a->ld_res % 4 == 0 &&
      (long long)a->M * a->N >= (1ll << 20)) {
    rc = get_tensor_map_f32(a->residual, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ld_res, (uint32_t)(BN < a->N ? BN : a->N),
                            BM, &tr);
    if (rc == EGV_OK) p.prefetch_res = 1;
  }
  if (pair) {
    switch (BN) {
      case 256: return dispatch_major<256, 2>(a_mn, b_mn, ta, tb, tr, p, stream);
      default: return dispatch_major<128, 2>(a_mn, b_mn, ta, tb, tr, p, stream);
    }
  }
  switch (BN) {
    case 256: return dispatch_major<256, 1>(a_mn, b_mn, ta, tb, tr, p, stream);
    case 192: return dispatch_major<192, 1>(a_mn, b_mn, ta, tb, tr, p, stream);
    case 128: return dispatch_major<128, 1>(a_mn, b_mn, ta, tb, tr, p, stream);
    default: return dispatch_major<64, 1>(a_mn, b_mn, ta, tb, tr, p, stream);
  }
}
