// Losses of the pre-training step: softmax cross-entropy with ignore_index (MLM / ITM) and EgoNCE
// (sim_matrix + positives mask + two-direction InfoNCE) with their gradients.  All fp32.
#include "common.cuh"
#include "host_common.h"

namespace egv {

template <int NT>
EGV_DEVINL float block_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < NT / 32; ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}
template <int NT>
EGV_DEVINL float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) r += red[i];
  __syncthreads();
  return r;
}

// one block per row: loss_sum += lse - logit[label], count += 1, dlogits = softmax - onehot (0 for ignored rows)
EGV_DEVINL void st_grad(bf16* p, float v) { *p = __float2bfloat16(v); }
EGV_DEVINL void st_grad(float* p, float v) { *p = v; }

template <typename DT>
__global__ void __launch_bounds__(256) xent_kernel(const float* __restrict__ logits, long long ld,
                                                   const long long* __restrict__ labels, int V, int ignore_index,
                                                   float* __restrict__ loss_sum, float* __restrict__ count,
                                                   DT* __restrict__ dlogits, long long ld_d) {
  pdl_enter();
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const long long label = labels[row];
  const float* x = logits + row * ld;
  DT* d = dlogits ? dlogits + row * ld_d : nullptr;
  if (label == ignore_index) {
    if (d) for (int c = threadIdx.x; c < V; c += 256) st_grad(d + c, 0.f);
    return;
  }
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < V; c += 256) mx = fmaxf(mx, x[c]);
  mx = block_max<256>(mx, red);
  float s = 0.f;
  for (int c = threadIdx.x; c < V; c += 256) s += __expf(x[c] - mx);
  s = block_sum<256>(s, red);
  const float lse = mx + logf(s);
  if (threadIdx.x == 0) {
    atomicAdd(loss_sum, lse - x[label]);
    atomicAdd(count, 1.0f);
  }
  if (d) {
    const float inv = 1.0f / s;
    for (int c = threadIdx.x; c < V; c += 256) {
      float p = __expf(x[c] - mx) * inv;
      if (c == label) p -= 1.0f;
      st_grad(d + c, p);
    }
  }
}

// loss = loss_sum / max(count,1); inv_count = 1 / max(count,1)
__global__ void xent_finalize_kernel(const float* loss_sum, const float* count, float* loss, float* inv_count) {
  pdl_enter();
  const float c = fmaxf(count[0], 1.0f);
  if (loss) loss[0] = loss_sum[0] / c;
  if (inv_count) inv_count[0] = 1.0f / c;
}

// ------------------------------------------------------------------------------------------- EgoNCE
// xn = x / max(||x||, eps); inv[i] = 1 / max(||x_i||, eps)
__global__ void __launch_bounds__(256) rownorm_kernel(const float* __restrict__ x, int P, float eps, float* __restrict__ xn,
                                                      float* __restrict__ inv) {
  pdl_enter();
  __shared__ float red[8];
  const long long row = blockIdx.x;
  float s = 0.f;
  for (int p = threadIdx.x; p < P; p += 256) {
    const float v = x[row * P + p];
    s = fmaf(v, v, s);
  }
  s = block_sum<256>(s, red);
  const float r = 1.0f / fmaxf(sqrtf(s), eps);
  if (threadIdx.x == 0 && inv) inv[row] = r;
  for (int p = threadIdx.x; p < P; p += 256) xn[row * P + p] = x[row * P + p] * r;
}

// out[i, j] = sum_p a[i,p] * b[j,p]; one warp per (i,j)
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, int Ga, int Gb,
                                                     int P, float* __restrict__ out) {
  pdl_enter();
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= Ga * Gb) return;
  const int i = w / Gb, j = w % Gb;
  float s = 0.f;
  for (int p = threadIdx.x & 31; p < P; p += 32) s = fmaf(a[(long long)i * P + p], b[(long long)j * P + p], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) out[w] = s;
}

// single block. sim [G,G] rows=text. mask_ij = (simv_ij*simn_ij + delta_ij) > 0.
// loss = -mean_i log sum_j M_ij softmax_j(sim_i./tau) - mean_i log sum_j M_ij softmax_j(sim_.i/tau)
// dsim = dloss/dsim.
__global__ void __launch_bounds__(256) egonce_loss_kernel(const float* __restrict__ sim, const float* __restrict__ simv,
                                                          const float* __restrict__ simn, int G, float tau,
                                                          uint8_t* __restrict__ mask, float* __restrict__ loss,
                                                          float* __restrict__ dsim) {
  pdl_enter();
  __shared__ float red[8];
  const float it = 1.0f / tau;
  for (int e = threadIdx.x; e < G * G; e += 256) {
    const int i = e / G, j = e % G;
    const float m = simv[e] * simn[e] + (i == j ? 1.0f : 0.0f);
    mask[e] = m > 0.0f ? 1 : 0;
    dsim[e] = 0.f;
  }
  __syncthreads();
  float part = 0.f;
  // row direction: thread i owns row i
  for (int i = threadIdx.x; i < G; i += 256) {
    float mx = -INFINITY;
    for (int j = 0; j < G; ++j) mx = fmaxf(mx, sim[i * G + j] * it);
    float z = 0.f, a = 0.f;
    for (int j = 0; j < G; ++j) {
      const float ex = __expf(sim[i * G + j] * it - mx);
      z += ex;
      if (mask[i * G + j]) a += ex;
    }
    part -= logf(a / z);
    for (int j = 0; j < G; ++j) {
      const float ex = __expf(sim[i * G + j] * it - mx);
      const float p = ex / z;
      const float gr = p - (mask[i * G + j] ? ex / a : 0.f);
      dsim[i * G + j] += gr * it / G;  // row i only touched by this thread in this phase
    }
  }
  __syncthreads();
  // column direction: thread i owns column i of sim (row i of sim^T); mask indexed [i][j] as in the reference
  for (int i = threadIdx.x; i < G; i += 256) {
    float mx = -INFINITY;
    for (int j = 0; j < G; ++j) mx = fmaxf(mx, sim[j * G + i] * it);
    float z = 0.f, a = 0.f;
    for (int j = 0; j < G; ++j) {
      const float ex = __expf(sim[j * G + i] * it - mx);
      z += ex;
      if (mask[i * G + j]) a += ex;
    }
    part -= logf(a / z);
    for (int j = 0; j < G; ++j) {
      const float ex = __expf(sim[j * G + i] * it - mx);
      const float p = ex / z;
      const float gr = p - (mask[i * G + j] ? ex / a : 0.f);
      dsim[j * G + i] += gr * it / G;  // column i only touched by this thread in this phase
    }
  }
  part = block_sum<256>(part, red);
  if (threadIdx.x == 0) loss[0] = part / G;
}

// Fine-tuning losses on the text x video cosine matrix (model_epic_charades.py:419-431), single block like the above.
//   kind 0  NormSoftmaxLoss (loss.py:13-31):  -mean_i log softmax(sim_i. / tau)_i - mean_i log softmax(sim_.i / tau)_i
//   kind 1  MaxMarginRankingLoss (loss.py:65-100):          mean relu(m       - (sim_ii - sim_ij)), relu(m       - (sim_ii - sim_ji))
//   kind 2  AdaptiveMaxMarginRankingLoss (loss.py:102-143): mean relu(w_i * m - (sim_ii - sim_ij)), relu(w_i * m - (sim_ii - sim_ji))
// (fix_norm: the mean runs over the 2G(G-1) off-diagonal terms; otherwise over all 2G^2, the diagonal adding w_i * m each).
__global__ void __launch_bounds__(256) dual_loss_kernel(const float* __restrict__ sim, int G, int kind, float param,
                                                        const float* __restrict__ weight, int fix_norm,
                                                        float* __restrict__ loss, float* __restrict__ dsim) {
  pdl_enter();
  __shared__ float red[8];
  for (int e = threadIdx.x; e < G * G; e += 256) dsim[e] = 0.f;
  __syncthreads();
  float part = 0.f;
  if (kind == 0) {
    const float it = 1.0f / param;
    for (int dir = 0; dir < 2; ++dir) {   // dir 0: thread i owns row i; dir 1: column i
      for (int i = threadIdx.x; i < G; i += 256) {
        const int si = dir == 0 ? G : 1, sj = dir == 0 ? 1 : G;
        float mx = -INFINITY;
        for (int j = 0; j < G; ++j) mx = fmaxf(mx, sim[i * si + j * sj] * it);
        float z = 0.f;
        for (int j = 0; j < G; ++j) z += __expf(sim[i * si + j * sj] * it - mx);
        part -= sim[i * G + i] * it - mx - logf(z);
        for (int j = 0; j < G; ++j) {
          const float pj = __expf(sim[i * si + j * sj] * it - mx) / z;
          dsim[i * si + j * sj] += (pj - (i == j ? 1.f : 0.f)) * it / G;
        }
      }
      __syncthreads();
    }
    part = block_sum<256>(part, red);
    if (threadIdx.x == 0) loss[0] = part / G;
    return;
  }
  const float cnt = fix_norm ? 2.f * G * (G - 1) : 2.f * G * G;
  const float gq = 1.0f / cnt;
  for (int dir = 0; dir < 2; ++dir) {
    for (int i = threadIdx.x; i < G; i += 256) {
      const int si = dir == 0 ? G : 1, sj = dir == 0 ? 1 : G;
      const float m = (kind == 2 ? weight[i] : 1.0f) * param;
      const float xii = sim[i * G + i];
      float dii = 0.f;
      for (int j = 0; j < G; ++j) {
        if (j == i) {
          if (!fix_norm) part += fmaxf(m, 0.f);
          continue;
        }
        const float a = m - (xii - sim[i * si + j * sj]);
        if (a > 0.f) {
          part += a;
          dii -= gq;
          dsim[i * si + j * sj] += gq;
        }
      }
      dsim[i * G + i] += dii;
    }
    __syncthreads();
  }
  part = block_sum<256>(part, red);
  if (threadIdx.x == 0) loss[0] = part / cnt;
}

// block (r, which): which=0 -> d t[row0+r] ; which=1 -> d v[row0+r].  Back-propagates through the L2 normalisation.
__global__ void __launch_bounds__(256) egonce_grad_kernel(const float* __restrict__ dsim, const float* __restrict__ tn,
                                                          const float* __restrict__ vn, const float* __restrict__ inv_t,
                                                          const float* __restrict__ inv_v, int G, int P, int row0,
                                                          float* __restrict__ dt, float* __restrict__ dv) {
  pdl_enter();
  __shared__ float red[8];
  extern __shared__ float coef[];  // G coefficients
  const int r = blockIdx.x, which = blockIdx.y;
  const int i = row0 + r;
  for (int j = threadIdx.x; j < G; j += 256) coef[j] = which == 0 ? dsim[i * G + j] : dsim[j * G + i];
  __syncthreads();
  const float* other = which == 0 ? vn : tn;
  const float* self = which == 0 ? tn : vn;
  float* out = which == 0 ? dt : dv;
  if (!out) return;
  const float inv = which == 0 ? inv_t[i] : inv_v[i];
  float dot = 0.f;
  for (int p = threadIdx.x; p < P; p += 256) {
    float g = 0.f;
    for (int j = 0; j < G; ++j) g = fmaf(coef[j], other[(long long)j * P + p], g);
    out[(long long)r * P + p] = g;
    dot = fmaf(g, self[(long long)i * P + p], dot);
  }
  dot = block_sum<256>(dot, red);
  for (int p = threadIdx.x; p < P; p += 256) {
    const float g = out[(long long)r * P + p];
    out[(long long)r * P + p] = inv * (g - self[(long long)i * P + p] * dot);
  }
}

}  // namespace egv

using namespace egv;

extern "C" int egv_softmax_xent(const float* logits, int64_t ld, const int64_t* labels, int64_t rows, int V, int ignore_index,
                                float* loss_sum, float* count, void* dlogits, int dlogits_is_f32, int64_t ld_d,
                                egv_stream_t stream) {
  if (!logits || !labels || !loss_sum || !count) return fail(EGV_ERR_ARG, "softmax_xent: null pointer");
  if (rows <= 0) return EGV_OK;
  if (dlogits_is_f32)
    launch_k(xent_kernel<float>, dim3((unsigned)rows), dim3(256), 0, (cudaStream_t)stream, logits, ld, (const long long*)labels, V, ignore_index, loss_sum, count, (float*)dlogits, ld_d);
  else
    launch_k(xent_kernel<bf16>, dim3((unsigned)rows), dim3(256), 0, (cudaStream_t)stream, logits, ld, (const long long*)labels, V, ignore_index, loss_sum, count, (bf16*)dlogits, ld_d);
  return check_launch("xent_kernel");
}

extern "C" int egv_xent_finalize(const float* loss_sum, const float* count, float* loss, float* inv_count, egv_stream_t stream) {
  if (!loss_sum || !count) return fail(EGV_ERR_ARG, "xent_finalize: null pointer");
  launch_k(xent_finalize_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, loss_sum, count, loss, inv_count);
  return check_launch("xent_finalize_kernel");
}

extern "C" int64_t egv_egonce_scratch_floats(int G, int P, int Dn, int Dv) {
  return 2ll * G * P + 2ll * G + (long long)G * Dn + (long long)G * Dv + 3ll * G * G;
}

extern "C" int egv_egonce(const float* t, const float* v, int G, int P, const float* noun, int Dn, const float* verb, int Dv,
                          float temperature, float* sim, uint8_t* mask, float* loss, int grad_row0, int grad_rows, float* dt,
                          float* dv, float* scratch, egv_stream_t stream) {
  if (!t || !v || !noun || !verb || !sim || !mask || !loss || !scratch) return fail(EGV_ERR_ARG, "egonce: null pointer");
  if (G <= 0 || P <= 0 || G > 4096) return fail(EGV_ERR_ARG, "egonce: bad sizes");
  if (grad_rows < 0 || grad_row0 < 0 || grad_row0 + grad_rows > G) return fail(EGV_ERR_ARG, "egonce: bad gradient row range");
  cudaStream_t s = (cudaStream_t)stream;
  float* tn = scratch;
  float* vn = tn + (long long)G * P;
  float* inv_t = vn + (long long)G * P;
  float* inv_v = inv_t + G;
  float* nn = inv_v + G;
  float* vb = nn + (long long)G * Dn;
  float* simv = vb + (long long)G * Dv;
  float* simn = simv + (long long)G * G;
  float* dsim = simn + (long long)G * G;
  int rc;
  launch_k(rownorm_kernel, dim3(G), dim3(256), 0, s, t, P, 1e-8f, tn, inv_t);
  if ((rc = check_launch("rownorm_kernel"))) return rc;
  launch_k(rownorm_kernel, dim3(G), dim3(256), 0, s, v, P, 1e-8f, vn, inv_v);
  if ((rc = check_launch("rownorm_kernel"))) return rc;
  launch_k(rownorm_kernel, dim3(G), dim3(256), 0, s, noun, Dn, 1e-8f, nn, nullptr);
  if ((rc = check_launch("rownorm_kernel"))) return rc;
  launch_k(rownorm_kernel, dim3(G), dim3(256), 0, s, verb, Dv, 1e-8f, vb, nullptr);
  if ((rc = check_launch("rownorm_kernel"))) return rc;
  const unsigned gb = (unsigned)cdiv((long long)G * G, 8);
  launch_k(rowdot_kernel, dim3(gb), dim3(256), 0, s, tn, vn, G, G, P, sim);
  if ((rc = check_launch("rowdot_kernel"))) return rc;
  launch_k(rowdot_kernel, dim3(gb), dim3(256), 0, s, nn, nn, G, G, Dn, simn);
  if ((rc = check_launch("rowdot_kernel"))) return rc;
  launch_k(rowdot_kernel, dim3(gb), dim3(256), 0, s, vb, vb, G, G, Dv, simv);
  if ((rc = check_launch("rowdot_kernel"))) return rc;
  launch_k(egonce_loss_kernel, dim3(1), dim3(256), 0, s, sim, simv, simn, G, temperature, mask, loss, dsim);
  if ((rc = check_launch("egonce_loss_kernel"))) return rc;
  if (grad_rows > 0 && (dt || dv)) {
    dim3 grid((unsigned)grad_rows, 2);
    launch_k(egonce_grad_kernel, dim3(grid), dim3(256), G * sizeof(float), s, dsim, tn, vn, inv_t, inv_v, G, P, grad_row0, dt, dv);
    if ((rc = check_launch("egonce_grad_kernel"))) return rc;
  }
  return EGV_OK;
}

extern "C" int64_t egv_dual_loss_scratch_floats(int G, int P) { return 2ll * G * P + 2ll * G + (long long)G * G; }

extern "C" int egv_dual_loss(const float* t, const float* v, int G, int P, int kind, float param, const float* weight,
                             int fix_norm, float* sim, float* loss, int grad_row0, int grad_rows, float* dt, float* dv,
                             float* scratch, egv_stream_t stream) {
  if (!t || !v || !sim || !loss || !scratch) return fail(EGV_ERR_ARG, "dual_loss: null pointer");
  if (G <= 0 || P <= 0 || G > 4096) return fail(EGV_ERR_ARG, "dual_loss: bad sizes");
  if (kind < EGV_DUAL_NORM_SOFTMAX || kind > EGV_DUAL_ADAPTIVE_MAX_MARGIN) return fail(EGV_ERR_ARG, "dual_loss: unknown kind");
  if (kind == EGV_DUAL_ADAPTIVE_MAX_MARGIN && !weight) return fail(EGV_ERR_ARG, "dual_loss: the adaptive margin needs weights");
  if (kind == EGV_DUAL_NORM_SOFTMAX && !(param > 0.f)) return fail(EGV_ERR_ARG, "dual_loss: temperature must be positive");
  if (grad_rows < 0 || grad_row0 < 0 || grad_row0 + grad_rows > G) return fail(EGV_ERR_ARG, "dual_loss: bad gradient row range");
  cudaStream_t s = (cudaStream_t)stream;
  float* tn = scratch;
  float* vn = tn + (long long)G * P;
  float* inv_t = vn + (long long)G * P;
  float* inv_v = inv_t + G;
  float* dsim = inv_v + G;
  int rc;
  launch_k(rownorm_kernel, dim3(G), dim3(256), 0, s, t, P, 1e-8f, tn, inv_t);
  if ((rc = check_launch("rownorm_kernel"))) return rc;
  launch_k(rownorm_kernel, dim3(G), dim3(256), 0, s, v, P, 1e-8f, vn, inv_v);
  if ((rc = check_launch("rownorm_kernel"))) return rc;
  launch_k(rowdot_kernel, dim3((unsigned)cdiv((long long)G * G, 8)), dim3(256), 0, s, tn, vn, G, G, P, sim);
  if ((rc = check_launch("rowdot_kernel"))) return rc;
  launch_k(dual_loss_kernel, dim3(1), dim3(256), 0, s, sim, G, kind, param, weight, fix_norm, loss, dsim);
  if ((rc = check_launch("dual_loss_kernel"))) return rc;
  if (grad_rows > 0 && (dt || dv)) {
    dim3 grid((unsigned)grad_rows, 2);
    launch_k(egonce_grad_kernel, dim3(grid), dim3(256), G * sizeof(float), s, dsim, tn, vn, inv_t, inv_v, G, P, grad_row0, dt, dv);
    if ((rc = check_launch("egonce_grad_kernel"))) return rc;
  }
  return EGV_OK;
}
