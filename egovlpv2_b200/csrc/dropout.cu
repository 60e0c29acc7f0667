// Train-mode dropout of the RoBERTa tower (hidden_dropout_prob = attention_probs_dropout_prob = 0.1: roberta.py:162,203
// embeddings; :244,313 attention probabilities; :337,342 attention output; :418,422 feed-forward output), Philox-based:
// the keep mask is a pure function of (device-resident step seed, site id, element index), so the backward regenerates
// it and a captured CUDA graph draws fresh masks on every replay (the step seed lives in device memory and is advanced
// by a kernel of the graph).
//
//   egv_rng_advance          step seed += odd constant
//   egv_dropout_add          out_f32 = res + scale * drop(x);  out_bf16 = bf16(drop(x))        (dense -> dropout -> + residual)
//   egv_dropout_bwd          out = drop'(dy) = keep ? dy / (1 - p) : 0
//   egv_text_attention_*     RobertaSelfAttention core (roberta.py:281-321) for S <= 64 tokens with dropout on the
//                            probabilities: scores / sqrt(d) + mask -> softmax -> dropout -> . V   (fp32 math, one CTA per
//                            (clip, head): 96 problems of 32 x 32 x 64 -- launch-latency-bound, hidden on the text stream)
#include <cuda.h>

#include "common.cuh"
#include "host_common.h"
#include "philox.cuh"

namespace egv {
namespace dr {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int HD = 64;
constexpr int SMAX = 64;
constexpr int ATT_THREADS = 512;   // 16 warps: two query rows per warp at S = 32 (4 warps walking 8 rows each were latency-bound: 59 us backward)

__global__ void rng_advance_kernel(unsigned long long* state) {
  pdl_enter();
  if (threadIdx.x == 0 && blockIdx.x == 0) state[0] += 0x9E3779B97F4A7C15ull;
}

template <bool X_BF16>
__global__ void dropout_add_kernel(const void* __restrict__ x_, const float* __restrict__ res, float scale,
                                   const float* __restrict__ scale_dev, float p_drop, const unsigned long long* __restrict__ seed_dev,
                                   unsigned long long site, float* __restrict__ out_f32, bf16* __restrict__ out_bf16, long long n) {
  pdl_enter();
  const unsigned long long key = philox_key(seed_dev, site);
  const float sc = scale * (scale_dev ? __ldg(scale_dev) : 1.0f);
  const float ks = 1.0f / (1.0f - p_drop);
  // four consecutive elements per thread = one Philox block
  for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 * 4 < n; i4 += (long long)gridDim.x * blockDim.x) {
    const uint4 r = philox4((uint32_t)i4, (uint32_t)((unsigned long long)i4 >> 32), (uint32_t)key, (uint32_t)(key >> 32));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = i4 * 4 + j;
      if (i >= n) break;
      const float xv = X_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(x_)[i]) : reinterpret_cast<const float*>(x_)[i];
      const float y = philox_keep_word(w[j], p_drop) ? xv * ks : 0.f;
      if (out_f32) out_f32[i] = (res ? res[i] : 0.f) + sc * y;
      if (out_bf16) out_bf16[i] = __float2bfloat16(y);
    }
  }
}

template <bool X_BF16>
__global__ void dropout_bwd_kernel(const void* __restrict__ dy_, float p_drop, const unsigned long long* __restrict__ seed_dev,
                                   unsigned long long site, float* __restrict__ out_f32, bf16* __restrict__ out_bf16, long long n) {
  pdl_enter();
  const unsigned long long key = philox_key(seed_dev, site);
  const float ks = 1.0f / (1.0f - p_drop);
  for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 * 4 < n; i4 += (long long)gridDim.x * blockDim.x) {
    const uint4 r = philox4((uint32_t)i4, (uint32_t)((unsigned long long)i4 >> 32), (uint32_t)key, (uint32_t)(key >> 32));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = i4 * 4 + j;
      if (i >= n) break;
      const float g = X_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(dy_)[i]) : reinterpret_cast<const float*>(dy_)[i];
      const float y = philox_keep_word(w[j], p_drop) ? g * ks : 0.f;
      if (out_f32) out_f32[i] = y;
      if (out_bf16) out_bf16[i] = __float2bfloat16(y);
    }
  }
}

struct AttP {
  const bf16* q; const bf16* k; const bf16* v; long long ld;   // rows b*S + s, head h at columns h*64
  const float* key_bias;                                       // [B, S] additive or NULL
  float scale, p_drop;
  const unsigned long long* seed_dev; unsigned long long site;
  int B, H, S;
  bf16* o; long long ldo;
  float* lse;                                                  // [B, H, S] natural log
  const bf16* d_o;
  bf16* dq; bf16* dk; bf16* dv; long long ldd;
};

// shared layout (floats): Kt [64][SMAX+1] (transposed keys), V [SMAX][64], Q / dO rows, P / dS [SMAX][SMAX+1]
template <bool BWD>
__global__ void __launch_bounds__(ATT_THREADS) text_attention_kernel(const AttP a) {
  pdl_enter();
  extern __shared__ float sm[];
  const int S = a.S;
  float* Kt = sm;                          // [64][SMAX + 1]
  float* V = Kt + HD * (SMAX + 1);         // [SMAX][64]
  float* Q = V + SMAX * HD;                // [SMAX][64]
  float* Pm = Q + SMAX * HD;               // [SMAX][SMAX + 1]   fwd: unused;  bwd: dropped probabilities
  float* dSm = Pm + SMAX * (SMAX + 1);     // bwd: dS
  float* dO = dSm + SMAX * (SMAX + 1);     // bwd: [SMAX][64]
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)b * S;
  for (int i = threadIdx.x; i < S * HD; i += ATT_THREADS) {
    const int s = i / HD, d = i % HD;
    const long long g = (row0 + s) * a.ld + h * HD + d;
    Kt[d * (SMAX + 1) + s] = __bfloat162float(a.k[g]);
    V[s * HD + d] = __bfloat162float(a.v[g]);
    Q[s * HD + d] = __bfloat162float(a.q[g]);
    if (BWD) dO[s * HD + d] = __bfloat162float(a.d_o[(row0 + s) * a.ldo + h * HD + d]);
  }
  __syncthreads();
  const unsigned long long key = a.p_drop > 0.f ? philox_key(a.seed_dev, a.site) : 0ull;
  const float ks = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f;
  for (int r = warp; r < S; r += ATT_THREADS / 32) {
    // scores of query r against keys lane, lane + 32
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; ++d) {
      const float qv = Q[r * HD + d];
      s0 = fmaf(qv, Kt[d * (SMAX + 1) + lane], s0);
      s1 = fmaf(qv, Kt[d * (SMAX + 1) + lane + 32], s1);
    }
    const bool v0 = lane < S, v1 = lane + 32 < S;
    s0 = v0 ? s0 * a.scale + (a.key_bias ? fmaxf(a.key_bias[row0 + lane], -3.0e38f) : 0.f) : -3.0e38f;
    s1 = v1 ? s1 * a.scale + (a.key_bias ? fmaxf(a.key_bias[row0 + lane + 32], -3.0e38f) : 0.f) : -3.0e38f;
    float p0, p1, l;
    if (!BWD) {
      const float mx = warp_max(fmaxf(s0, s1));
      p0 = v0 ? exp2f((s0 - mx) * LOG2E) : 0.f;
      p1 = v1 ? exp2f((s1 - mx) * LOG2E) : 0.f;
      const float sum = warp_sum(p0 + p1);
      l = mx + logf(sum);
      p0 /= sum;
      p1 /= sum;
      if (lane == 0) a.lse[((long long)b * a.H + h) * S + r] = l;
    } else {
      l = a.lse[((long long)b * a.H + h) * S + r];
      p0 = v0 ? exp2f((s0 - l) * LOG2E) : 0.f;
      p1 = v1 ? exp2f((s1 - l) * LOG2E) : 0.f;
    }
    // dropout on the probabilities: element index ((b*H + h)*S + r)*S + j
    float m0 = 1.f, m1 = 1.f;
    if (a.p_drop > 0.f) {
      const unsigned long long e = (((unsigned long long)b * a.H + h) * S + r) * S;
      m0 = philox_keep(key, e + lane, a.p_drop) ? ks : 0.f;
      m1 = philox_keep(key, e + lane + 32, a.p_drop) ? ks : 0.f;
    }
    if (!BWD) {
      const float q0 = p0 * m0, q1 = p1 * m1;
      // o_r[d] = sum_j q_j V[j][d],  d = lane, lane + 32
      float o0 = 0.f, o1 = 0.f;
      for (int j = 0; j < S; ++j) {
        const float pj = __shfl_sync(0xffffffffu, j < 32 ? q0 : q1, j & 31);
        o0 = fmaf(pj, V[j * HD + lane], o0);
        o1 = fmaf(pj, V[j * HD + lane + 32], o1);
      }
      bf16* orow = a.o + (row0 + r) * a.ldo + h * HD;
      orow[lane] = __float2bfloat16(o0);
      orow[lane + 32] = __float2bfloat16(o1);
    } else {
      // dP~_j = dO_r . V_j  (gradient of the dropped probabilities), dP = dP~ * mask
      float g0 = 0.f, g1 = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) {
        const float gd = dO[r * HD + d];
        g0 = fmaf(gd, v0 ? V[lane * HD + d] : 0.f, g0);
        g1 = fmaf(gd, v1 ? V[(lane + 32) * HD + d] : 0.f, g1);
      }
      g0 *= m0;
      g1 *= m1;
      const float dot = warp_sum(p0 * g0 + p1 * g1);
      if (v0) {
        dSm[r * (SMAX + 1) + lane] = p0 * (g0 - dot) * a.scale;
        Pm[r * (SMAX + 1) + lane] = p0 * m0;
      }
      if (v1) {
        dSm[r * (SMAX + 1) + lane + 32] = p1 * (g1 - dot) * a.scale;
        Pm[r * (SMAX + 1) + lane + 32] = p1 * m1;
      }
    }
  }
  if (!BWD) return;
  __syncthreads();
  // dQ_r = sum_j dS_rj K_j ;  dK_j = sum_r dS_rj Q_r ;  dV_j = sum_r P~_rj dO_r      (rows r / j per warp, d per lane)
  for (int r = warp; r < S; r += ATT_THREADS / 32) {
    float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f, w0 = 0.f, w1 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float ds_rj = dSm[r * (SMAX + 1) + j];
      q0 = fmaf(ds_rj, Kt[lane * (SMAX + 1) + j], q0);
      q1 = fmaf(ds_rj, Kt[(lane + 32) * (SMAX + 1) + j], q1);
      const float ds_jr = dSm[j * (SMAX + 1) + r], p_jr = Pm[j * (SMAX + 1) + r];   // here r plays the key index
      k0 = fmaf(ds_jr, Q[j * HD + lane], k0);
      k1 = fmaf(ds_jr, Q[j * HD + lane + 32], k1);
      w0 = fmaf(p_jr, dO[j * HD + lane], w0);
      w1 = fmaf(p_jr, dO[j * HD + lane + 32], w1);
    }
    const long long g = (row0 + r) * a.ldd + h * HD;
    a.dq[g + lane] = __float2bfloat16(q0);
    a.dq[g + lane + 32] = __float2bfloat16(q1);
    a.dk[g + lane] = __float2bfloat16(k0);
    a.dk[g + lane + 32] = __float2bfloat16(k1);
    a.dv[g + lane] = __float2bfloat16(w0);
    a.dv[g + lane + 32] = __float2bfloat16(w1);
  }
}

constexpr int ATT_SMEM = (HD * (SMAX + 1) + 3 * SMAX * HD + 2 * SMAX * (SMAX + 1)) * 4;

}  // namespace dr
}  // namespace egv

using namespace egv;

extern "C" int egv_rng_advance(uint64_t* state, egv_stream_t stream) {
  if (!state) return fail(EGV_ERR_ARG, "rng_advance: null state");
  launch_k(dr::rng_advance_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (unsigned long long*)state);
  return check_launch("rng_advance_kernel");
}

extern "C" int egv_dropout_add(const void* x, int x_is_bf16, const float* res, float scale, const float* scale_dev, float p_drop,
                               const uint64_t* seed_dev, uint64_t site, float* out_f32, void* out_bf16, int64_t n,
                               egv_stream_t stream) {
  if (!x || !seed_dev || n <= 0 || (!out_f32 && !out_bf16)) return fail(EGV_ERR_ARG, "dropout_add: bad arguments");
  if (p_drop < 0.f || p_drop >= 1.f) return fail(EGV_ERR_ARG, "dropout_add: probability %f", p_drop);
  const unsigned grid = (unsigned)std::min<long long>(cdiv(cdiv(n, 4), 256), 148 * 8);
  if (x_is_bf16)
    launch_k(dr::dropout_add_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, res, scale, scale_dev, p_drop, (const unsigned long long*)seed_dev, site, out_f32, (bf16*)out_bf16, n);
  else
    launch_k(dr::dropout_add_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, res, scale, scale_dev, p_drop, (const unsigned long long*)seed_dev, site, out_f32, (bf16*)out_bf16, n);
  return check_launch("dropout_add_kernel");
}

extern "C" int egv_dropout_bwd(const void* dy, int dy_is_bf16, float p_drop, const uint64_t* seed_dev, uint64_t site, float* out_f32,
                               void* out_bf16, int64_t n, egv_stream_t stream) {
  if (!dy || !seed_dev || n <= 0 || (!out_f32 && !out_bf16)) return fail(EGV_ERR_ARG, "dropout_bwd: bad arguments");
  const unsigned grid = (unsigned)std::min<long long>(cdiv(cdiv(n, 4), 256), 148 * 8);
  if (dy_is_bf16)
    launch_k(dr::dropout_bwd_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, dy, p_drop, (const unsigned long long*)seed_dev, site, out_f32, (bf16*)out_bf16, n);
  else
    launch_k(dr::dropout_bwd_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, dy, p_drop, (const unsigned long long*)seed_dev, site, out_f32, (bf16*)out_bf16, n);
  return check_launch("dropout_bwd_kernel");
}

static int text_attention(const egv_text_attn_args* t, bool bwd, cudaStream_t stream) {
  if (!t || !t->q || !t->k || !t->v || !t->lse || t->B <= 0 || t->H <= 0 || t->S <= 0) return fail(EGV_ERR_ARG, "text_attention: bad arguments");
  if (t->S > dr::SMAX) return fail(EGV_ERR_UNSUPPORTED, "text_attention: %d tokens > %d", t->S, dr::SMAX);
  if (t->p_drop < 0.f || t->p_drop >= 1.f || (t->p_drop > 0.f && !t->seed_dev)) return fail(EGV_ERR_ARG, "text_attention: dropout arguments");
  if (bwd ? (!t->d_o || !t->dq || !t->dk || !t->dv) : !t->o) return fail(EGV_ERR_ARG, "text_attention: missing output");
  dr::AttP a;
  a.q = (const bf16*)t->q; a.k = (const bf16*)t->k; a.v = (const bf16*)t->v; a.ld = t->ld;
  a.key_bias = t->key_bias; a.scale = t->scale; a.p_drop = t->p_drop;
  a.seed_dev = (const unsigned long long*)t->seed_dev; a.site = t->site;
  a.B = t->B; a.H = t->H; a.S = t->S;
  a.o = (bf16*)t->o; a.ldo = t->ldo; a.lse = t->lse;
  a.d_o = (const bf16*)t->d_o; a.dq = (bf16*)t->dq; a.dk = (bf16*)t->dk; a.dv = (bf16*)t->dv; a.ldd = t->ldd;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(dr::text_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dr::ATT_SMEM);
    cudaFuncSetAttribute(dr::text_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dr::ATT_SMEM);
    configured = true;
  }
  if (bwd) launch_k(dr::text_attention_kernel<true>, dim3((unsigned)(t->B * t->H)), dim3(dr::ATT_THREADS), dr::ATT_SMEM, stream, a);
  else launch_k(dr::text_attention_kernel<false>, dim3((unsigned)(t->B * t->H)), dim3(dr::ATT_THREADS), dr::ATT_SMEM, stream, a);
  return check_launch("text_attention_kernel");
}

extern "C" int egv_text_attention_fwd(const egv_text_attn_args* t, egv_stream_t stream) { return text_attention(t, false, (cudaStream_t)stream); }
extern "C" int egv_text_attention_bwd(const egv_text_attn_args* t, egv_stream_t stream) { return text_attention(t, true, (cudaStream_t)stream); }
