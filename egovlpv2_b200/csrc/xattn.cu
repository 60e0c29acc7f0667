// Row-wise pieces of the re-associated gated cross-attention that are not GEMM-shaped (the GEMMs are xgemm.cu):
//   text -> video (roberta.py:470-486, 281-321): softmax over the N video tokens of every (clip, head, text query) row of
//     the score matrix Qp x^T, its backward, and (train mode) the Philox dropout of the probabilities (roberta.py:313);
//   video -> text (video_transformer.py:155-185): the query-bias term of the scores, c0[b,s,h] = d^-1/2 bq_h . k_h[b,s],
//     plus the additive key mask, and its backward.
#include <cuda.h>

#include "common.cuh"
#include "host_common.h"
#include "philox.cuh"

namespace egv {
namespace xa {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int ROW_THREADS = 256;
constexpr int MAX_PER_THREAD = 16;   // rows of up to 4096 columns live in registers

EGV_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool IS_MAX>
EGV_DEVINL float block_reduce(float v, float* red) {
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();   // red may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < ROW_THREADS / 32; ++i) r = IS_MAX ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// One CTA per row r of scores [rows, ld_s] (row r = batch r / rows_per_batch, local row r % rows_per_batch; output
// rows of batch b start at b * p_bstride elements).  P = softmax(scores) in bf16, lse = log sum exp (natural log).
// Dropout (p_drop > 0): P_out = keep ? P / (1 - p) : 0, with `rsum` = the row sum of the dropped probabilities.
// Column mapping: element (r, c) has the flat dropout index r * n + c and one Philox4x32 call decides four consecutive
// indices, so a thread owns the columns of whole Philox counters (pass i, thread t: counter (r * n) / 4 + i * 256 + t) --
// one generator call per four probabilities instead of one per probability (the first version spent its 54 us per launch
// at cfg 3 on 9.6 M ten-round Philox calls).
constexpr int QUADS = MAX_PER_THREAD / 4;

struct RowMap {
  long long ctr0;   // first Philox counter that overlaps the row
  int shift;        // (r * n) % 4: column of the counter's word w is 4 * (ctr - ctr0) + w - shift
};
EGV_DEVINL RowMap row_map(long long r, int n) {
  const unsigned long long first = (unsigned long long)r * (unsigned long long)n;
  RowMap m;
  m.ctr0 = (long long)(first >> 2);
  m.shift = (int)(first & 3ull);
  return m;
}

__global__ void __launch_bounds__(ROW_THREADS)
row_softmax_kernel(const float* __restrict__ scores, long long ld_s, int rows_per_batch, long long s_bstride, int n,
                   bf16* __restrict__ P, long long ld_p, long long p_bstride, float* __restrict__ lse, float p_drop,
                   const unsigned long long* __restrict__ seed_dev, unsigned long long site, float* __restrict__ rsum) {
  pdl_enter();
  __shared__ float red[ROW_THREADS / 32];
  const long long r = blockIdx.x;
  const long long b = r / rows_per_batch, lr = r % rows_per_batch;
  const float* src = scores + b * s_bstride + lr * ld_s;
  bf16* dst = P + b * p_bstride + lr * ld_p;
  const RowMap rm = row_map(r, n);
  float x[MAX_PER_THREAD];
  float mx = -3.0e38f;
#pragma unroll
  for (int i = 0; i < QUADS; ++i) {
    const int c0 = 4 * (threadIdx.x + i * ROW_THREADS) - rm.shift;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int c = c0 + w;
      x[4 * i + w] = (c >= 0 && c < n) ? __ldg(src + c) : -3.0e38f;
      mx = fmaxf(mx, x[4 * i + w]);
    }
  }
  mx = block_reduce<true>(mx, red);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_PER_THREAD; ++i) {
    x[i] = x[i] > -1.0e38f ? ex2((x[i] - mx) * LOG2E) : 0.f;
    sum += x[i];
  }
  sum = block_reduce<false>(sum, red);
  const float inv = 1.0f / sum;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  const unsigned long long seed = p_drop > 0.f ? philox_key(seed_dev, site) : 0ull;
  float dsum = 0.f;
#pragma unroll
  for (int i = 0; i < QUADS; ++i) {
    const int q = threadIdx.x + i * ROW_THREADS;
    const int c0 = 4 * q - rm.shift;
    if (c0 >= n) continue;
    uint32_t rw[4] = {0u, 0u, 0u, 0u};
    if (p_drop > 0.f) {
      const unsigned long long ctr = (unsigned long long)(rm.ctr0 + q);
      const uint4 rr = philox4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
      rw[0] = rr.x; rw[1] = rr.y; rw[2] = rr.z; rw[3] = rr.w;
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int c = c0 + w;
      if (c >= 0 && c < n) {
        float pv = x[4 * i + w] * inv;
        if (p_drop > 0.f) {
          pv = philox_keep_word(rw[w], p_drop) ? pv * keep_scale : 0.f;
          dsum += pv;
        }
        dst[c] = __float2bfloat16(pv);
      }
    }
  }
  if (threadIdx.x == 0 && lse) lse[r] = mx + logf(sum);
  if (rsum) {
    dsum = block_reduce<false>(dsum, red);
    if (threadIdx.x == 0) rsum[r] = p_drop > 0.f ? dsum : 1.0f;
  }
}

// Backward of row_softmax_kernel: dS = P * (dP~ - sum_c P dP~), where P = exp(scores - lse) is recomputed from the saved
// scores and dP~ = dP * keep / (1 - p) undoes the dropout (the mask is regenerated from the Philox stream; same column
// mapping as the forward).
__global__ void __launch_bounds__(ROW_THREADS)
row_dsoftmax_kernel(const float* __restrict__ scores, long long ld_s, int rows_per_batch, long long s_bstride, int n,
                    const float* __restrict__ lse, const float* __restrict__ dP, long long ld_dp, long long dp_bstride,
                    bf16* __restrict__ dS, long long ld_ds, long long ds_bstride, float p_drop,
                    const unsigned long long* __restrict__ seed_dev, unsigned long long site, const float* __restrict__ row_const) {
  pdl_enter();
  __shared__ float red[ROW_THREADS / 32];
  const long long r = blockIdx.x;
  const long long b = r / rows_per_batch, lr = r % rows_per_batch;
  const float* src = scores + b * s_bstride + lr * ld_s;
  const float* dp = dP + b * dp_bstride + lr * ld_dp;
  bf16* dst = dS + b * ds_bstride + lr * ld_ds;
  const float l2 = lse[r] * LOG2E;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  const float rc = row_const ? row_const[r] : 0.f;   // value-bias term d_ox_h . bv_h (cancels unless dropout is on)
  const unsigned long long seed = p_drop > 0.f ? philox_key(seed_dev, site) : 0ull;
  const RowMap rm = row_map(r, n);
  float pr[MAX_PER_THREAD], g[MAX_PER_THREAD];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < QUADS; ++i) {
    const int q = threadIdx.x + i * ROW_THREADS;
    const int c0 = 4 * q - rm.shift;
    uint32_t rw[4] = {0u, 0u, 0u, 0u};
    if (p_drop > 0.f && c0 < n) {
      const unsigned long long ctr = (unsigned long long)(rm.ctr0 + q);
      const uint4 rr = philox4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
      rw[0] = rr.x; rw[1] = rr.y; rw[2] = rr.z; rw[3] = rr.w;
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int c = c0 + w;
      pr[4 * i + w] = 0.f;
      g[4 * i + w] = 0.f;
      if (c >= 0 && c < n) {
        pr[4 * i + w] = ex2(fmaf(__ldg(src + c), LOG2E, -l2));
        float gv = __ldg(dp + c) + rc;
        if (p_drop > 0.f) gv = philox_keep_word(rw[w], p_drop) ? gv * keep_scale : 0.f;
        g[4 * i + w] = gv;
        dot = fmaf(pr[4 * i + w], gv, dot);
      }
    }
  }
  dot = block_reduce<false>(dot, red);
#pragma unroll
  for (int i = 0; i < QUADS; ++i) {
    const int c0 = 4 * (threadIdx.x + i * ROW_THREADS) - rm.shift;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int c = c0 + w;
      if (c >= 0 && c < n) dst[c] = __float2bfloat16(pr[4 * i + w] * (g[4 * i + w] - dot));
    }
  }
}

// Rows longer than the 4093 columns the register-resident kernels hold (TimeSformer-L/14 at 32 frames: 18 433 video tokens
// per clip): the same arithmetic and the same Philox column mapping, streaming the row from global memory (it stays in L2)
// in three passes (max, sum, write) / two passes (dot, write).
__global__ void __launch_bounds__(ROW_THREADS)
row_softmax_long_kernel(const float* __restrict__ scores, long long ld_s, int rows_per_batch, long long s_bstride, int n,
                        bf16* __restrict__ P, long long ld_p, long long p_bstride, float* __restrict__ lse, float p_drop,
                        const unsigned long long* __restrict__ seed_dev, unsigned long long site, float* __restrict__ rsum) {
  pdl_enter();
  __shared__ float red[ROW_THREADS / 32];
  const long long r = blockIdx.x;
  const long long b = r / rows_per_batch, lr = r % rows_per_batch;
  const float* src = scores + b * s_bstride + lr * ld_s;
  bf16* dst = P + b * p_bstride + lr * ld_p;
  const RowMap rm = row_map(r, n);
  float mx = -3.0e38f;
  for (int c = threadIdx.x; c < n; c += ROW_THREADS) mx = fmaxf(mx, __ldg(src + c));
  mx = block_reduce<true>(mx, red);
  float sum = 0.f;
  for (int c = threadIdx.x; c < n; c += ROW_THREADS) sum += ex2((__ldg(src + c) - mx) * LOG2E);
  sum = block_reduce<false>(sum, red);
  const float inv = 1.0f / sum;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  const unsigned long long seed = p_drop > 0.f ? philox_key(seed_dev, site) : 0ull;
  float dsum = 0.f;
  for (int q = threadIdx.x; 4 * q - rm.shift < n; q += ROW_THREADS) {
    const int c0 = 4 * q - rm.shift;
    uint32_t rw[4] = {0u, 0u, 0u, 0u};
    if (p_drop > 0.f) {
      const unsigned long long ctr = (unsigned long long)(rm.ctr0 + q);
      const uint4 rr = philox4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
      rw[0] = rr.x; rw[1] = rr.y; rw[2] = rr.z; rw[3] = rr.w;
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int c = c0 + w;
      if (c >= 0 && c < n) {
        float pv = ex2((__ldg(src + c) - mx) * LOG2E) * inv;
        if (p_drop > 0.f) {
          pv = philox_keep_word(rw[w], p_drop) ? pv * keep_scale : 0.f;
          dsum += pv;
        }
        dst[c] = __float2bfloat16(pv);
      }
    }
  }
  if (threadIdx.x == 0 && lse) lse[r] = mx + logf(sum);
  if (rsum) {
    dsum = block_reduce<false>(dsum, red);
    if (threadIdx.x == 0) rsum[r] = p_drop > 0.f ? dsum : 1.0f;
  }
}

__global__ void __launch_bounds__(ROW_THREADS)
row_dsoftmax_long_kernel(const float* __restrict__ scores, long long ld_s, int rows_per_batch, long long s_bstride, int n,
                         const float* __restrict__ lse, const float* __restrict__ dP, long long ld_dp, long long dp_bstride,
                         bf16* __restrict__ dS, long long ld_ds, long long ds_bstride, float p_drop,
                         const unsigned long long* __restrict__ seed_dev, unsigned long long site,
                         const float* __restrict__ row_const) {
  pdl_enter();
  __shared__ float red[ROW_THREADS / 32];
  const long long r = blockIdx.x;
  const long long b = r / rows_per_batch, lr = r % rows_per_batch;
  const float* src = scores + b * s_bstride + lr * ld_s;
  const float* dp = dP + b * dp_bstride + lr * ld_dp;
  bf16* dst = dS + b * ds_bstride + lr * ld_ds;
  const float l2 = lse[r] * LOG2E;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  const float rc = row_const ? row_const[r] : 0.f;
  const unsigned long long seed = p_drop > 0.f ? philox_key(seed_dev, site) : 0ull;
  const RowMap rm = row_map(r, n);
  float dot = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) dot = block_reduce<false>(dot, red);
    for (int q = threadIdx.x; 4 * q - rm.shift < n; q += ROW_THREADS) {
      const int c0 = 4 * q - rm.shift;
      uint32_t rw[4] = {0u, 0u, 0u, 0u};
      if (p_drop > 0.f) {
        const unsigned long long ctr = (unsigned long long)(rm.ctr0 + q);
        const uint4 rr = philox4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
        rw[0] = rr.x; rw[1] = rr.y; rw[2] = rr.z; rw[3] = rr.w;
      }
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int c = c0 + w;
        if (c >= 0 && c < n) {
          const float pr = ex2(fmaf(__ldg(src + c), LOG2E, -l2));
          float gv = __ldg(dp + c) + rc;
          if (p_drop > 0.f) gv = philox_keep_word(rw[w], p_drop) ? gv * keep_scale : 0.f;
          if (pass == 0) dot = fmaf(pr, gv, dot);
          else dst[c] = __float2bfloat16(pr * (gv - dot));
        }
      }
    }
  }
}

// bias1[b, h*S + s] = scale * sum_j k[b*S+s, h*64+j] * bq[h*64+j] + mask[b, s]      one warp per (b, s, h)
__global__ void qbias_fwd_kernel(const bf16* __restrict__ k, long long ldk, const float* __restrict__ bq,
                                 const float* __restrict__ mask, float scale, int B, int S, int H, float* __restrict__ out) {
  pdl_enter();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= B * S * H) return;
  const int h = gw % H, bs = gw / H;
  const int b = bs / S, s = bs % S;
  const bf16* kr = k + (long long)bs * ldk + h * 64;
  const float2 kv = unpack_bf16(*reinterpret_cast<const uint32_t*>(kr + 2 * lane));
  const float2 q = *reinterpret_cast<const float2*>(bq + h * 64 + 2 * lane);
  float v = warp_sum(kv.x * q.x + kv.y * q.y);
  if (lane == 0) out[(long long)b * H * S + h * S + s] = scale * v + (mask ? mask[bs] : 0.f);
}

// dk[b*S+s, h*64+j] += scale * dbias[b, h*S+s] * bq[h*64+j];   dbq[h*64+j] += scale * sum_{b,s} k[..] * dbias[..]
// grid (H, ceil(B*S / 4)), 64 threads (j): a CTA owns 4 rows of one head and adds its partial dbq with one atomic per j
__global__ void qbias_bwd_kernel(const bf16* __restrict__ k, long long ldk, const float* __restrict__ bq,
                                 const float* __restrict__ dbias, float scale, int B, int S, int H, float* __restrict__ dk,
                                 long long lddk, float* __restrict__ dbq) {
  pdl_enter();
  const int h = blockIdx.x, j = threadIdx.x;
  const float q = bq[h * 64 + j];
  float acc = 0.f;
  const int r0 = blockIdx.y * 4, r1 = min(r0 + 4, B * S);   // 4 rows per CTA: the loop is a chain of dependent global round trips (32 rows: 29 us)
#pragma unroll 4
  for (int bs = r0; bs < r1; ++bs) {
    const int b = bs / S, s = bs % S;
    const float g = scale * __ldg(dbias + (long long)b * H * S + h * S + s);
    acc = fmaf(g, __bfloat162float(k[(long long)bs * ldk + h * 64 + j]), acc);
    if (dk) dk[(long long)bs * lddk + h * 64 + j] += g * q;
  }
  if (dbq) atomicAdd(dbq + h * 64 + j, acc);
}

// ox[b*S+s, h*64+j] += rsum[b, h*S+s] * bv[h*64+j]   (value bias under dropout: the dropped probabilities do not sum to 1)
__global__ void rowscale_bias_kernel(bf16* __restrict__ ox, long long ld, const float* __restrict__ rsum,
                                     const float* __restrict__ bv, int B, int S, int H) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int Cw = H * 64;
  if (i >= B * S * Cw) return;
  const int col = i % Cw, bs = i / Cw, h = col / 64, b = bs / S, s = bs % S;
  bf16* o = ox + (long long)bs * ld + col;
  *o = __float2bfloat16(__bfloat162float(*o) + rsum[(long long)b * H * S + h * S + s] * bv[col]);
}
// dbv[h*64+j] += sum_{b,s} rsum[b, h*S+s] * d_ox[b*S+s, h*64+j]     grid (H*64 / 64, ceil(B*S / 16)): 16 rows per CTA + one atomic
__global__ void rowscale_bias_bwd_kernel(const bf16* __restrict__ d_ox, long long ld, const float* __restrict__ rsum,
                                         float* __restrict__ dbv, int B, int S, int H) {
  pdl_enter();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= H * 64) return;
  const int h = col / 64;
  const int r0 = blockIdx.y * 16, r1 = min(r0 + 16, B * S);
  float acc = 0.f;
#pragma unroll 4
  for (int bs = r0; bs < r1; ++bs)
    acc = fmaf(rsum[(long long)(bs / S) * H * S + h * S + (bs % S)], __bfloat162float(d_ox[(long long)bs * ld + col]), acc);
  atomicAdd(dbv + col, acc);
}

}  // namespace xa
}  // namespace egv

using namespace egv;

extern "C" int egv_xattn_row_softmax(const float* scores, int64_t ld_s, int64_t rows, int rows_per_batch, int64_t s_bstride,
                                     int n, void* P, int64_t ld_p, int64_t p_bstride, float* lse, float p_drop,
                                     const uint64_t* seed_dev, uint64_t site, float* rsum, egv_stream_t stream) {
  if (!scores || !P || rows <= 0 || n <= 0 || rows_per_batch <= 0) return fail(EGV_ERR_ARG, "row_softmax: bad arguments");
  if (p_drop < 0.f || p_drop >= 1.f) return fail(EGV_ERR_ARG, "row_softmax: dropout probability %f", p_drop);
  if (n > xa::ROW_THREADS * xa::MAX_PER_THREAD - 3) {   // long rows: streaming variant
    launch_k(xa::row_softmax_long_kernel, dim3((unsigned)rows), dim3(xa::ROW_THREADS), 0, (cudaStream_t)stream, scores, ld_s, rows_per_batch, s_bstride, n, (bf16*)P, ld_p, p_bstride, lse, p_drop, (const unsigned long long*)seed_dev, site, rsum);
    return check_launch("row_softmax_long_kernel");
  }
  launch_k(xa::row_softmax_kernel, dim3((unsigned)rows), dim3(xa::ROW_THREADS), 0, (cudaStream_t)stream,  scores, ld_s, rows_per_batch, s_bstride, n, (bf16*)P, ld_p, p_bstride, lse, p_drop, (const unsigned long long*)seed_dev, site, rsum);
  return check_launch("row_softmax_kernel");
}

extern "C" int egv_xattn_row_dsoftmax(const float* scores, int64_t ld_s, int64_t rows, int rows_per_batch, int64_t s_bstride,
                                      int n, const float* lse, const float* dP, int64_t ld_dp, int64_t dp_bstride, void* dS,
                                      int64_t ld_ds, int64_t ds_bstride, float p_drop, const uint64_t* seed_dev, uint64_t site,
                                      const float* row_const, egv_stream_t stream) {
  if (!scores || !lse || !dP || !dS || rows <= 0 || n <= 0 || rows_per_batch <= 0) return fail(EGV_ERR_ARG, "row_dsoftmax: bad arguments");
  if (n > xa::ROW_THREADS * xa::MAX_PER_THREAD - 3) {   // long rows: streaming variant
    launch_k(xa::row_dsoftmax_long_kernel, dim3((unsigned)rows), dim3(xa::ROW_THREADS), 0, (cudaStream_t)stream, scores, ld_s, rows_per_batch, s_bstride, n, lse, dP, ld_dp, dp_bstride, (bf16*)dS, ld_ds, ds_bstride, p_drop, (const unsigned long long*)seed_dev, site, row_const);
    return check_launch("row_dsoftmax_long_kernel");
  }
  launch_k(xa::row_dsoftmax_kernel, dim3((unsigned)rows), dim3(xa::ROW_THREADS), 0, (cudaStream_t)stream,  scores, ld_s, rows_per_batch, s_bstride, n, lse, dP, ld_dp, dp_bstride, (bf16*)dS, ld_ds, ds_bstride, p_drop, (const unsigned long long*)seed_dev, site, row_const);
  return check_launch("row_dsoftmax_kernel");
}

extern "C" int egv_xattn_qbias_fwd(const void* k, int64_t ldk, const float* bq, const float* mask, float scale, int B, int S,
                                   int H, float* out, egv_stream_t stream) {
  if (!k || !bq || !out || B <= 0 || S <= 0 || H <= 0) return fail(EGV_ERR_ARG, "qbias_fwd: bad arguments");
  const long long warps = (long long)B * S * H;
  launch_k(xa::qbias_fwd_kernel, dim3((unsigned)cdiv(warps, 8)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)k, ldk, bq, mask, scale, B, S, H, out);
  return check_launch("qbias_fwd_kernel");
}

extern "C" int egv_xattn_qbias_bwd(const void* k, int64_t ldk, const float* bq, const float* dbias, float scale, int B, int S,
                                   int H, float* dk, int64_t lddk, float* dbq, egv_stream_t stream) {
  if (!k || !bq || !dbias || B <= 0 || S <= 0 || H <= 0) return fail(EGV_ERR_ARG, "qbias_bwd: bad arguments");
  launch_k(xa::qbias_bwd_kernel, dim3(dim3((unsigned)H, (unsigned)cdiv((long long)B * S, 4))), dim3(64), 0, (cudaStream_t)stream, (const bf16*)k, ldk, bq, dbias, scale, B, S, H, dk, lddk, dbq);
  return check_launch("qbias_bwd_kernel");
}

extern "C" int egv_xattn_rowscale_bias(void* ox, int64_t ld, const float* rsum, const float* bv, int B, int S, int H,
                                       egv_stream_t stream) {
  if (!ox || !rsum || !bv || B <= 0 || S <= 0 || H <= 0) return fail(EGV_ERR_ARG, "rowscale_bias: bad arguments");
  const long long n = (long long)B * S * H * 64;
  launch_k(xa::rowscale_bias_kernel, dim3((unsigned)cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, (bf16*)ox, ld, rsum, bv, B, S, H);
  return check_launch("rowscale_bias_kernel");
}

extern "C" int egv_xattn_rowscale_bias_bwd(const void* d_ox, int64_t ld, const float* rsum, float* dbv, int B, int S, int H,
                                           egv_stream_t stream) {
  if (!d_ox || !rsum || !dbv || B <= 0 || S <= 0 || H <= 0) return fail(EGV_ERR_ARG, "rowscale_bias_bwd: bad arguments");
  launch_k(xa::rowscale_bias_bwd_kernel, dim3((unsigned)cdiv(H * 64, 64), (unsigned)cdiv(B * S, 16)), dim3(64), 0, (cudaStream_t)stream, (const bf16*)d_ox, ld, rsum, dbv, B, S, H);
  return check_launch("rowscale_bias_bwd_kernel");
}
