// egv_layernorm_fwd / egv_layernorm_bwd: nn.LayerNorm over the last dimension, one warp per row, fp32 statistics.
// HBM-bound: each row is read once (kept in registers for C <= 1024) and written once.
#include <type_traits>

#include "common.cuh"
#define EGV_PDL_CLASS 8
#include "host_common.h"

namespace egv {

constexpr int LN_MAXV = 8;  // float4 per lane -> C <= 1024

template <bool XBF>
EGV_DEVINL float4 ld4(const void* base, long long idx) {
  if (XBF) {
    uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(base) + idx);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
}
EGV_DEVINL void st4_bf16(bf16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16(v.x, v.y);
  u.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

template <bool XBF>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, long long rows, int C,
                                                     bf16* __restrict__ y_bf16, float* __restrict__ y_f32,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nv = C >> 2;  // float4 per row
  for (long long row = warp0; row < rows; row += nwarps) {
    const long long base = row * C;
    float4 v[LN_MAXV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c4 = lane + i * 32;
      if (c4 < nv) {
        v[i] = ld4<XBF>(x, base + 4 * c4);
        sum += v[i].x + v[i].y + v[i].z + v[i].w;
      }
    }
    const float mean = warp_sum(sum) / C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c4 = lane + i * 32;
      if (c4 < nv) {
        float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += a * a + b * b + c * c + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / C + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c4 = lane + i * 32;
      if (c4 < nv) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c4);
        const float4 bb = *reinterpret_cast<const float4*>(beta + 4 * c4);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + bb.x;
        o.y = (v[i].y - mean) * rstd * g.y + bb.y;
        o.z = (v[i].z - mean) * rstd * g.z + bb.z;
        o.w = (v[i].w - mean) * rstd * g.w + bb.w;
        if (y_f32) *reinterpret_cast<float4*>(y_f32 + base + 4 * c4) = o;
        if (y_bf16) st4_bf16(y_bf16 + base + 4 * c4, o);
      }
    }
  }
}

// val = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dx = (add ? add : 0) + val (add may alias dx);
// dx_bf16 = bf16(bf16_total ? dx : val);  dgamma += sum dy*xhat;  dbeta += sum dy;  out_colsum += sum of the value
// written to dx_bf16 (bias gradient of the Linear that produced this LayerNorm's input branch).
// One warp per row.  The three column accumulators live in per-warp shared memory (lane-private float4 slots, no
// conflicts, no atomics) so that registers only hold one row: high occupancy for an HBM-bound kernel.
EGV_DEVINL float4 unpack_dy(const float4& v) { return v; }
EGV_DEVINL float4 unpack_dy(const uint2& u) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

template <bool DYBF, bool XBF>
__global__ void __launch_bounds__(256, 2) ln_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, long long rows, int C,
                                                     const float* add, float* dx, int bf16_total, bf16* __restrict__ dx_bf16,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     float* __restrict__ out_colsum) {
  pdl_enter();
  extern __shared__ __align__(16) float ln_smem[];   // [8 warps][3][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp0 = (long long)blockIdx.x * 8 + warp;
  const long long nwarps = (long long)gridDim.x * 8;
  const int nv = C >> 2;
  float4* acc_g = reinterpret_cast<float4*>(ln_smem + (size_t)warp * 3 * C);
  float4* acc_b = acc_g + nv;
  float4* acc_s = acc_b + nv;
  const bool want_params = dgamma != nullptr || dbeta != nullptr;
  const bool want_cs = out_colsum != nullptr;
  for (int c4 = lane; c4 < nv; c4 += 32) acc_g[c4] = acc_b[c4] = acc_s[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = warp0; row < rows; row += nwarps) {
    const long long base = row * C;
    // Every global load of the row (x, dy and the residual gradient `add`) is issued before the first use: one memory
    // round trip per row instead of two -- the kernel is bound by bytes in flight per SM, not by issue slots.
    // dy stays in its storage format (bf16 pairs when DYBF) and x stays raw: xhat and g * dy are recomputed in the
    // second pass, which keeps the kernel at 128 registers = two resident 256-thread blocks per SM.
    typedef typename std::conditional<DYBF, uint2, float4>::type dy_t;
    float4 xv[LN_MAXV], av[LN_MAXV];
    dy_t dv[LN_MAXV];
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c4 = lane + i * 32;
      if (c4 < nv) {
        xv[i] = ld4<XBF>(x, base + 4 * c4);
        if (DYBF) dv[i] = *reinterpret_cast<const dy_t*>(reinterpret_cast<const bf16*>(dy) + base + 4 * c4);
        else dv[i] = *reinterpret_cast<const dy_t*>(reinterpret_cast<const float*>(dy) + base + 4 * c4);
        if (add) av[i] = *reinterpret_cast<const float4*>(add + base + 4 * c4);
      }
    }
    const float rs = rstd[row];
    const float nmu = -mean[row] * rs;   // xhat = x * rs + nmu
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c4 = lane + i * 32;
      if (c4 < nv) {
        const float4 d = unpack_dy(dv[i]);
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c4);
        const float4 xh = make_float4(fmaf(xv[i].x, rs, nmu), fmaf(xv[i].y, rs, nmu), fmaf(xv[i].z, rs, nmu), fmaf(xv[i].w, rs, nmu));
        const float4 gy = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
        s1 += gy.x + gy.y + gy.z + gy.w;
        s2 += gy.x * xh.x + gy.y * xh.y + gy.z * xh.z + gy.w * xh.w;
        if (want_params) {
          float4 a = acc_g[c4], b = acc_b[c4];
          a.x += d.x * xh.x; a.y += d.y * xh.y; a.z += d.z * xh.z; a.w += d.w * xh.w;
          b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
          acc_g[c4] = a;
          acc_b[c4] = b;
        }
      }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c4 = lane + i * 32;
      if (c4 < nv) {
        const float4 d = unpack_dy(dv[i]);
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c4);
        float4 o;
        o.x = rs * (d.x * g.x - s1 - fmaf(xv[i].x, rs, nmu) * s2);
        o.y = rs * (d.y * g.y - s1 - fmaf(xv[i].y, rs, nmu) * s2);
        o.z = rs * (d.z * g.z - s1 - fmaf(xv[i].z, rs, nmu) * s2);
        o.w = rs * (d.w * g.w - s1 - fmaf(xv[i].w, rs, nmu) * s2);
        float4 tot = o;
        if (add) {
          tot.x += av[i].x; tot.y += av[i].y; tot.z += av[i].z; tot.w += av[i].w;
        }
        if (dx) *reinterpret_cast<float4*>(dx + base + 4 * c4) = tot;
        const float4 ob = bf16_total ? tot : o;
        if (dx_bf16) st4_bf16(dx_bf16 + base + 4 * c4, ob);
        if (want_cs) {
          float4 a = acc_s[c4];
          a.x += ob.x; a.y += ob.y; a.z += ob.z; a.w += ob.w;
          acc_s[c4] = a;
        }
      }
    }
  }
  __syncthreads();
  // block reduction over the 8 warps, then one atomic per column per block
  for (int c = threadIdx.x; c < C; c += 256) {
    float g = 0.f, b = 0.f, sc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float* base = ln_smem + (size_t)w * 3 * C;
      g += base[c];
      b += base[C + c];
      sc += base[2 * C + c];
    }
    if (dgamma) atomicAdd(dgamma + c, g);
    if (dbeta) atomicAdd(dbeta + c, b);
    if (out_colsum) atomicAdd(out_colsum + c, sc);
  }
}

}  // namespace egv

using namespace egv;

extern "C" int egv_layernorm_fwd(const void* x, int x_is_bf16, const float* gamma, const float* beta, float eps,
                                 int64_t rows, int C, void* y_bf16, float* y_f32, float* mean, float* rstd,
                                 egv_stream_t stream) {
  if (!x || !gamma || !beta || (!y_bf16 && !y_f32)) return fail(EGV_ERR_ARG, "layernorm_fwd: null pointer");
  if (C % 4 || C > 128 * LN_MAXV || C <= 0) return fail(EGV_ERR_UNSUPPORTED, "layernorm: C=%d must be a multiple of 4 and <= 1024", C);
  if (rows <= 0) return EGV_OK;
  long long blocks = cdiv(rows, 8);
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (x_is_bf16)
    launch_k(ln_fwd_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, gamma, beta, eps, rows, C, (bf16*)y_bf16, y_f32, mean, rstd);
  else
    launch_k(ln_fwd_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, gamma, beta, eps, rows, C, (bf16*)y_bf16, y_f32, mean, rstd);
  return check_launch("ln_fwd_kernel");
}

extern "C" int egv_layernorm_bwd(const void* dy, int dy_is_bf16, const void* x, int x_is_bf16, const float* gamma,
                                 const float* mean, const float* rstd, int64_t rows, int C, const float* add, float* dx,
                                 void* dx_bf16, int bf16_total, float* dgamma, float* dbeta, float* out_colsum,
                                 egv_stream_t stream) {
  if (!dy || !x || !gamma || !mean || !rstd) return fail(EGV_ERR_ARG, "layernorm_bwd: null pointer");
  if (C % 4 || C > 128 * LN_MAXV || C <= 0) return fail(EGV_ERR_UNSUPPORTED, "layernorm: C=%d must be a multiple of 4 and <= 1024", C);
  if (rows <= 0) return EGV_OK;
  // >= 8 rows per warp amortises the column-sum atomics of the big (video) calls; the 256-row text-tower calls are latency
  // bound instead: 4 CTAs walking 8 rows per warp one after the other took 30 us -- one row per warp there
  long long blocks = rows <= 4096 ? cdiv(rows, 8) : cdiv(rows, 8 * 8);
  const long long cap = (long long)sm_count() * 2;   // measured: 3 resident blocks per SM are slower (129 vs 93 us at cfg 3)
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaStream_t s = (cudaStream_t)stream;
  const int smem = 8 * 3 * C * (int)sizeof(float);
#define EGV_LN_BWD(DYB, XB)                                                                                          \
  do {                                                                                                               \
    static bool cfg = false;                                                                                         \
    if (!cfg) {                                                                                                      \
      cudaFuncSetAttribute(ln_bwd_kernel<DYB, XB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 1024 * 4);   \
      cfg = true;                                                                                                    \
    }                                                                                                                \
    launch_k(ln_bwd_kernel<DYB, XB>, dim3((unsigned)blocks), dim3(256), smem, s, dy, x, gamma, mean, rstd, rows, C, add, dx, bf16_total, (bf16*)dx_bf16, dgamma, dbeta, out_colsum); \
  } while (0)
  if (dy_is_bf16 && x_is_bf16) EGV_LN_BWD(true, true);
  else if (dy_is_bf16) EGV_LN_BWD(true, false);
  else if (x_is_bf16) EGV_LN_BWD(false, true);
  else EGV_LN_BWD(false, false);
#undef EGV_LN_BWD
  return check_launch("ln_bwd_kernel");
}
