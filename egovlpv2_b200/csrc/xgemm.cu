// egv_bgemm_bf16: BATCHED tcgen05 GEMM with the softmax epilogues of the re-associated gated cross-attention
// (video_transformer.py:155-185, roberta.py:470-486; DESIGN.md "cross-attention by re-association").
//
// Same machine as gemm.cu's persistent kernel (warp 0 TMA producer, warp 1 tcgen05.mma issuer, warp 2 TMEM allocator,
// warps 4-11 epilogue), with
//   * a two-level batch grid z = (z1, z0): operands are addressed through rank-4 TMA tensor maps
//     {inner, outer, z0, z1}, so a batch is a coordinate, never a pointer table: per-clip operands ([B, N, C] streams,
//     per-clip text-side matrices [B, H*S, C]) and per-head column slices of a [rows, H*64] tensor are all views;
//     rows / reduction elements past a batch's own M / N / K are zero-filled by the TMA unit (N = 3137 tokens per clip is
//     not a multiple of anything);
//   * epilogue EGV_BGEMM_EPI_SOFTMAX32: every aligned group of 32 output columns (= the S = 32 text keys of one head)
//     of a row is soft-maxed in registers straight out of TMEM (thread = row, one tcgen05.ld = one group):
//         P[n, h, :] = softmax_s(acc[n, h*32+s] + bias[h*32+s])
//   * epilogue EGV_BGEMM_EPI_DSOFTMAX32: the backward of that softmax on dP = acc with the saved P as `aux`:
//         dS = scale * P * (dP - sum_s P dP),  colsum[j] += sum_n dS[n, j],  dot_out += sum_n,h sum_s P dP
//     (the last sum is the gradient of the scalar fusion gate: d alpha = sum(d_out * c) = sum(P * dP'));
//   * the plain epilogue: bias, scale, fp32 residual, fp32 store / atomic accumulate (split-K), bf16 store.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#define EGV_PDL_CLASS 2
#include "host_common.h"

namespace egv {

int get_tensor_map_nd(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMap* out);
int get_tensor_map_nd_ex(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                         int f32, CUtensorMap* out);

namespace xg {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 128 + EPI_WARPS * 32;
constexpr int EPI_STAGE_FLOATS = 32 * 32;
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
  int M, N, K;
  int nb0, nb1;
  int items_per_batch, total_items, num_n_tiles, split_k, k_blocks_total, k_blocks_per_split;
  int a_z0, a_z1, b_z0, b_z1;   // 1: the operand has that batch level (its coordinate is z), 0: shared (coordinate 0)
  const float* bias; long long sbias0, sbias1;
  float scale; const float* scale_dev;
  const float* residual; long long ld_res, sres0, sres1;
  float* out_f32; long long ld_o32, so32_0, so32_1;
  bf16* out_bf16; long long ld_o16, so16_0, so16_1;
  const bf16* aux; long long ld_aux, saux0, saux1;
  float* colsum; long long scol0, scol1;
  float* dot_out;
  int accumulate, epilogue;
  int prefetch_res, r_z0, r_z1;   // L2 prefetch of each tile's fp32 residual block by the producer (see gemm.cu)
};

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 192 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_STAGE_FLOATS * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SLACK = 1024;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + SLACK;
  static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);
};

EGV_DEVINL void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
EGV_DEVINL void tma_prefetch_l2_4d(const void* tmap, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
EGV_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
EGV_DEVINL int stg_off(int row, int chunk) { return row * 32 + ((chunk ^ (row & 7)) << 2); }

// per-item (batch-resolved) epilogue pointers
struct ItemPtrs {
  const float* bias;
  const float* residual;
  float* out_f32;
  bf16* out_bf16;
  const bf16* aux;
  float* colsum;
};

// One 32-row x 32-column chunk read back from the staging tile 4 rows per instruction: lane l owns columns
// 4*(l&7)..+3 of rows 4i + (l>>3).  rows_left / c_ok mask the row and column tails (N is handled in groups of 4 columns:
// the host guarantees every row stride covers round_up(N, 4)).
template <int EPI>
EGV_DEVINL void readback(const Params& p, const ItemPtrs& q, const float* stg, int lane, int row_base, int rows_left, int gcol,
                         bool c_ok, bool lead, float scale_total) {
  const int sub = lane >> 3, ch = lane & 7;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (EPI == EGV_BGEMM_EPI_NONE && lead && q.bias && c_ok) b = __ldg(reinterpret_cast<const float4*>(q.bias + gcol));
  const bool has_res = lead && q.residual != nullptr;
  const float* res_p = q.residual ? q.residual + (long long)row_base * p.ld_res + gcol : nullptr;
  const bf16* aux_p = q.aux ? q.aux + (long long)row_base * p.ld_aux + gcol : nullptr;
  float* o32 = q.out_f32 ? q.out_f32 + (long long)row_base * p.ld_o32 + gcol : nullptr;
  bf16* o16 = q.out_bf16 ? q.out_bf16 + (long long)row_base * p.ld_o16 + gcol : nullptr;
  float4 rs[8];
  uint2 ax[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = 4 * i + sub;
    const bool ok = c_ok && rl < rows_left;
    rs[i] = (has_res && ok) ? __ldg(reinterpret_cast<const float4*>(res_p + (long long)rl * p.ld_res)) : make_float4(0.f, 0.f, 0.f, 0.f);
    ax[i] = (EPI == EGV_BGEMM_EPI_DSOFTMAX32 && ok) ? __ldg(reinterpret_cast<const uint2*>(aux_p + (long long)rl * p.ld_aux))
                                                   : make_uint2(0u, 0u);
  }
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  float dsum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = 4 * i + sub;
    const bool ok = c_ok && rl < rows_left;
    const float4 a = *reinterpret_cast<const float4*>(stg + stg_off(rl, ch));
    float v0 = a.x + b.x, v1 = a.y + b.y, v2 = a.z + b.z, v3 = a.w + b.w;
    if (EPI == EGV_BGEMM_EPI_DSOFTMAX32) {
      // the 32 columns of this chunk are one softmax group; a row's group lives in the 8 lanes sharing `sub`
      const float2 p01 = unpack_bf16(ax[i].x), p23 = unpack_bf16(ax[i].y);
      float t = v0 * p01.x + v1 * p01.y + v2 * p23.x + v3 * p23.y;
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 4);
      if (ch == 0 && ok) dsum += t;
      v0 = p01.x * (v0 - t);
      v1 = p01.y * (v1 - t);
      v2 = p23.x * (v2 - t);
      v3 = p23.y * (v3 - t);
    }
    v0 = fmaf(v0, scale_total, rs[i].x);
    v1 = fmaf(v1, scale_total, rs[i].y);
    v2 = fmaf(v2, scale_total, rs[i].z);
    v3 = fmaf(v3, scale_total, rs[i].w);
    if (!ok) continue;
    if (o32) {
      float* o = o32 + (long long)rl * p.ld_o32;
      if (p.accumulate) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
      } else {
        *reinterpret_cast<float4*>(o) = make_float4(v0, v1, v2, v3);
      }
    }
    if (o16) {
      uint2 u;
      u.x = pack_bf16(v0, v1);
      u.y = pack_bf16(v2, v3);
      *reinterpret_cast<uint2*>(o16 + (long long)rl * p.ld_o16) = u;
    }
    cs.x += v0;
    cs.y += v1;
    cs.z += v2;
    cs.w += v3;
  }
  if (q.colsum) {
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o);
      cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o);
      cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
    }
    if (sub == 0 && c_ok) {
      // the column tail (N % 4 != 0) never carries a colsum: the host requires N % 4 == 0 with colsum
      atomicAdd(q.colsum + gcol, cs.x);
      atomicAdd(q.colsum + gcol + 1, cs.y);
      atomicAdd(q.colsum + gcol + 2, cs.z);
      atomicAdd(q.colsum + gcol + 3, cs.w);
    }
  }
  if (EPI == EGV_BGEMM_EPI_DSOFTMAX32 && p.dot_out) {
    dsum = warp_sum(dsum);
    if (lane == 0) atomicAdd(p.dot_out, dsum);
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 1)
bgemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ CUtensorMap tmap_res, const Params p) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  if (pad > (uint32_t)C::SLACK) {
    if (threadIdx.x == 0) printf("egv: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* smem = smem_raw + pad;
  float* epi_smem = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + C::EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // programmatic dependent launch (common.cuh): before any global access

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
        const int zb = w / p.items_per_batch;
        const int wi = w - zb * p.items_per_batch;
        const int z0 = zb % p.nb0, z1 = zb / p.nb0;
        const int ks = wi % p.split_k;
        const int tile = wi / p.split_k;
        const int m0 = (tile / p.num_n_tiles) * BM;
        const int n0 = (tile % p.num_n_tiles) * BN;
        const int kb0 = ks * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks_total);
        const int az0 = z0 * p.a_z0, az1 = z1 * p.a_z1, bz0 = z0 * p.b_z0, bz1 = z1 * p.b_z1;
        if (p.prefetch_res && ks == 0) tma_prefetch_l2_4d(&tmap_res, n0, m0, z0 * p.r_z0, z1 * p.r_z1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_sleep(&empty_bar[stage], phase ^ 1, 64);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          const int k0 = kb * BK;
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_4d(sa + j * 8192, &tmap_a, &full_bar[stage], m0 + 64 * j, k0, az0, az1);
          } else {
            tma_load_4d(sa, &tmap_a, &full_bar[stage], k0, m0, az0, az1);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_4d(sb + j * 8192, &tmap_b, &full_bar[stage], n0 + 64 * j, k0, bz0, bz1);
          } else {
            tma_load_4d(sb, &tmap_b, &full_bar[stage], k0, n0, bz0, bz1);
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
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t a_adv = A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      constexpr uint32_t b_adv = B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++it) {
        const int wi = w % p.items_per_batch;
        const int ks = wi % p.split_k;
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
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t adesc = A_MN ? umma_desc_sw128(sa, 8192, 1024) : umma_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = B_MN ? umma_desc_sw128(sb, 8192, 1024) : umma_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(d_tmem, adesc + (uint64_t)(k * a_adv), bdesc + (uint64_t)(k * b_adv), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int quad = warp & 3;
    const int chalf = ew >> 2;
    float* stg = epi_smem + ew * EPI_STAGE_FLOATS;
    const float scale_total = p.scale * (p.scale_dev ? __ldg(p.scale_dev) : 1.0f);
    int it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++it) {
      const int zb = w / p.items_per_batch;
      const int wi = w - zb * p.items_per_batch;
      const long long z0 = zb % p.nb0, z1 = zb / p.nb0;
      const int ks = wi % p.split_k;
      const int tile = wi / p.split_k;
      const int m0 = (tile / p.num_n_tiles) * BM;
      const int n0 = (tile % p.num_n_tiles) * BN;
      const int as = it & 1;
      const uint32_t use = (uint32_t)(it >> 1) & 1u;
      const bool lead = (ks == 0);
      ItemPtrs q;
      q.bias = p.bias ? p.bias + z0 * p.sbias0 + z1 * p.sbias1 : nullptr;
      q.residual = p.residual ? p.residual + z0 * p.sres0 + z1 * p.sres1 : nullptr;
      q.out_f32 = p.out_f32 ? p.out_f32 + z0 * p.so32_0 + z1 * p.so32_1 : nullptr;
      q.out_bf16 = p.out_bf16 ? p.out_bf16 + z0 * p.so16_0 + z1 * p.so16_1 : nullptr;
      q.aux = p.aux ? p.aux + z0 * p.saux0 + z1 * p.saux1 : nullptr;
      q.colsum = p.colsum ? p.colsum + z0 * p.scol0 + z1 * p.scol1 : nullptr;
      mbar_wait(&tmem_full[as], use);
      tc_fence_after();
      const int row_base = m0 + quad * 32;
      const int rows_left = p.M - row_base;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int col0 = chalf * (BN / 2) + c * 32;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + col0), v);
        tmem_ld_wait();
        if (p.epilogue == EGV_BGEMM_EPI_SOFTMAX32 && n0 + col0 < p.N) {
          // thread = row: the 32 registers are one softmax group (N % 32 == 0 on this path)
          const float4* bp = reinterpret_cast<const float4*>(q.bias + n0 + col0);
          float mx = -3.0e38f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = q.bias ? __ldg(bp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float x0 = __uint_as_float(v[4 * j]) + b4.x, x1 = __uint_as_float(v[4 * j + 1]) + b4.y;
            const float x2 = __uint_as_float(v[4 * j + 2]) + b4.z, x3 = __uint_as_float(v[4 * j + 3]) + b4.w;
            v[4 * j] = __float_as_uint(x0);
            v[4 * j + 1] = __float_as_uint(x1);
            v[4 * j + 2] = __float_as_uint(x2);
            v[4 * j + 3] = __float_as_uint(x3);
            mx = fmaxf(mx, fmaxf(fmaxf(x0, x1), fmaxf(x2, x3)));
          }
          const float ms = mx * LOG2E;
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e = ex2(fmaf(__uint_as_float(v[j]), LOG2E, -ms));
            sum += e;
            v[j] = __float_as_uint(e);
          }
          const float inv = 1.0f / sum;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * inv);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stg + stg_off(lane, j)) =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                          __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        const int gcol = n0 + col0 + 4 * (lane & 7);
        const bool c_ok = gcol < p.N;
        if (p.epilogue == EGV_BGEMM_EPI_DSOFTMAX32)
          readback<EGV_BGEMM_EPI_DSOFTMAX32>(p, q, stg, lane, row_base, rows_left, gcol, c_ok, lead, scale_total);
        else if (p.epilogue == EGV_BGEMM_EPI_SOFTMAX32)
          readback<EGV_BGEMM_EPI_SOFTMAX32>(p, q, stg, lane, row_base, rows_left, gcol, c_ok, lead, scale_total);
        else
          readback<EGV_BGEMM_EPI_NONE>(p, q, stg, lane, row_base, rows_left, gcol, c_ok, lead, scale_total);
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN, bool A_MN, bool B_MN>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const Params& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool configured = false;
  auto kern = bgemm_kernel<BN, A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "bgemm smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int grid = p.total_items < sm_count() ? p.total_items : sm_count();
  launch_k(kern, dim3(grid), dim3(THREADS), C::SMEM_BYTES, stream, ta, tb, tr, p);
  return check_launch("bgemm_kernel");
}

template <int BN>
static int dispatch(int layout, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const Params& p, cudaStream_t s) {
  if (layout == EGV_GEMM_NT) return launch<BN, false, false>(ta, tb, tr, p, s);
  if (layout == EGV_GEMM_NN) return launch<BN, false, true>(ta, tb, tr, p, s);
  return launch<BN, true, true>(ta, tb, tr, p, s);
}

}  // namespace xg
}  // namespace egv

using namespace egv;

extern "C" int egv_bgemm_bf16(const egv_bgemm_args* a, egv_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a || !a->A || !a->B) return fail(EGV_ERR_ARG, "bgemm: null operand");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0 || a->nb0 <= 0 || a->nb1 <= 0)
    return fail(EGV_ERR_ARG, "bgemm: bad shape %d %d %d x (%d, %d)", a->M, a->N, a->K, a->nb0, a->nb1);
  if (a->layout < 0 || a->layout > 2) return fail(EGV_ERR_ARG, "bgemm: bad layout %d", a->layout);
  if (!a->out_f32 && !a->out_bf16) return fail(EGV_ERR_ARG, "bgemm: no output");
  if (a->epilogue < 0 || a->epilogue > 2) return fail(EGV_ERR_ARG, "bgemm: bad epilogue %d", a->epilogue);
  int split_k = a->split_k < 1 ? 1 : a->split_k;
  if (split_k > 1 && (!a->accumulate || !a->out_f32 || a->out_bf16 || a->epilogue != EGV_BGEMM_EPI_NONE || a->colsum))
    return fail(EGV_ERR_ARG, "bgemm: split_k > 1 needs accumulate=1, f32 output only, plain epilogue");
  if (a->epilogue != EGV_BGEMM_EPI_NONE && a->N % 32) return fail(EGV_ERR_ARG, "bgemm: softmax epilogues need N %% 32 == 0 (N = %d)", a->N);
  if (a->epilogue == EGV_BGEMM_EPI_DSOFTMAX32 && !a->aux) return fail(EGV_ERR_ARG, "bgemm: the softmax-backward epilogue needs aux = P");
  if (a->epilogue == EGV_BGEMM_EPI_SOFTMAX32 && (a->residual || a->colsum))
    return fail(EGV_ERR_ARG, "bgemm: the softmax epilogue takes no residual / colsum");
  if ((a->bias || a->colsum) && a->N % 4) return fail(EGV_ERR_ARG, "bgemm: bias / colsum need N %% 4 == 0");
  // vector (4-column) epilogue I/O: 16-byte aligned fp32 rows, 8-byte aligned bf16 rows, row strides covering round_up(N, 4)
  const long long n4 = (a->N + 3) / 4 * 4;
  auto ok32 = [&](const void* ptr, long long ld, long long s0, long long s1) {
    return !ptr || (((uintptr_t)ptr) % 16 == 0 && ld % 4 == 0 && ld >= n4 && s0 % 4 == 0 && s1 % 4 == 0);
  };
  auto ok16 = [&](const void* ptr, long long ld, long long s0, long long s1) {
    return !ptr || (((uintptr_t)ptr) % 8 == 0 && ld % 4 == 0 && ld >= n4 && s0 % 4 == 0 && s1 % 4 == 0);
  };
  if (!ok32(a->out_f32, a->ld_out_f32, a->so32_0, a->so32_1) || !ok32(a->residual, a->ld_res, a->sres0, a->sres1) ||
      !ok16(a->out_bf16, a->ld_out_bf16, a->so16_0, a->so16_1) || !ok16(a->aux, a->ld_aux, a->saux0, a->saux1) ||
      (a->bias && (((uintptr_t)a->bias) % 16 || a->sbias0 % 4 || a->sbias1 % 4)) ||
      (a->colsum && (((uintptr_t)a->colsum) % 16 || a->scol0 % 4 || a->scol1 % 4)))
    return fail(EGV_ERR_UNSUPPORTED, "bgemm: epilogue tensors must be 16-byte aligned with row strides %% 4 == 0 covering round_up(N, 4)");
  const bool a_mn = a->layout == EGV_GEMM_TN;
  const bool b_mn = a->layout != EGV_GEMM_NT;
  auto tma_ok = [](const void* ptr, long long ld, long long s0, long long s1) {
    return (((uintptr_t)ptr) & 15) == 0 && ld % 8 == 0 && s0 % 8 == 0 && s1 % 8 == 0;
  };
  if (!tma_ok(a->A, a->lda, a->sa0, a->sa1) || !tma_ok(a->B, a->ldb, a->sb0, a->sb1))
    return fail(EGV_ERR_UNSUPPORTED, "bgemm: operands must be 16-byte aligned with row / batch strides %% 8 == 0");

  xg::Params p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.nb0 = a->nb0; p.nb1 = a->nb1;
  p.a_z0 = (a->nb0 > 1 && a->sa0 != 0) ? 1 : 0; p.a_z1 = (a->nb1 > 1 && a->sa1 != 0) ? 1 : 0;
  p.b_z0 = (a->nb0 > 1 && a->sb0 != 0) ? 1 : 0; p.b_z1 = (a->nb1 > 1 && a->sb1 != 0) ? 1 : 0;
  p.bias = a->bias; p.sbias0 = a->sbias0; p.sbias1 = a->sbias1;
  p.scale = a->scale; p.scale_dev = a->scale_dev;
  p.residual = a->residual; p.ld_res = a->ld_res; p.sres0 = a->sres0; p.sres1 = a->sres1;
  p.out_f32 = a->out_f32; p.ld_o32 = a->ld_out_f32; p.so32_0 = a->so32_0; p.so32_1 = a->so32_1;
  p.out_bf16 = (bf16*)a->out_bf16; p.ld_o16 = a->ld_out_bf16; p.so16_0 = a->so16_0; p.so16_1 = a->so16_1;
  p.aux = (const bf16*)a->aux; p.ld_aux = a->ld_aux; p.saux0 = a->saux0; p.saux1 = a->saux1;
  p.colsum = a->colsum; p.scol0 = a->scol0; p.scol1 = a->scol1;
  p.dot_out = a->dot_out;
  p.accumulate = a->accumulate; p.epilogue = a->epilogue;

  // tile width: the widest tile that does not waste more than a quarter of its columns; 192 divides H*S = 384
  int BN;
  if (a->N <= 64) BN = 64;
  else if (a->N <= 128) BN = 128;
  else if (a->N % 256 == 0 || a->N > 1024) BN = 256;
  else if (a->N % 192 == 0) BN = 192;
  else if (a->N <= 192) BN = 192;
  else BN = 256;
  const long long batches = (long long)a->nb0 * a->nb1;
  const int num_m_tiles = (int)cdiv(a->M, xg::BM);
  p.k_blocks_total = (int)cdiv(a->K, xg::BK);
  p.num_n_tiles = (int)cdiv(a->N, BN);
  // automatic split-K for accumulating fp32 outputs whose tile grid cannot fill the GPU (per-clip TN products with
  // K = tokens per clip): the caller accumulates anyway, so slices just add
  if (split_k == 1 && a->accumulate && a->out_f32 && !a->out_bf16 && a->epilogue == EGV_BGEMM_EPI_NONE && !a->colsum && !a->residual &&
      !a->bias) {
    const long long tiles = batches * num_m_tiles * p.num_n_tiles;
    if (tiles * 2 <= sm_count() * 3 / 2 && p.k_blocks_total >= 16) {
      long long want = cdiv(sm_count(), tiles);
      long long cap = p.k_blocks_total / 8;
      split_k = (int)std::max<long long>(1, std::min<long long>(want, cap));
    }
  }
  if (split_k > p.k_blocks_total) split_k = p.k_blocks_total;
  p.k_blocks_per_split = (int)cdiv(p.k_blocks_total, split_k);
  split_k = (int)cdiv(p.k_blocks_total, p.k_blocks_per_split);
  p.split_k = split_k;
  p.items_per_batch = num_m_tiles * p.num_n_tiles * split_k;
  const long long total = batches * p.items_per_batch;
  if (total > 0x7fffffffll) return fail(EGV_ERR_ARG, "bgemm: too many work items");
  p.total_items = (int)total;

  // rank-4 maps {inner, outer, z0, z1}; a shared level is a dimension of extent 1 (its coordinate is always 0)
  auto make_map = [&](const void* ptr, uint64_t inner, uint64_t outer, long long ld, long long s0, long long s1, int use0, int use1,
                      uint32_t box_outer, CUtensorMap* out) {
    uint64_t dims[4] = {inner, outer, (uint64_t)(use0 ? a->nb0 : 1), (uint64_t)(use1 ? a->nb1 : 1)};
    uint64_t st[3] = {(uint64_t)ld * 2, (uint64_t)(use0 ? s0 : ld) * 2, (uint64_t)(use1 ? s1 : ld) * 2};
    uint32_t box[4] = {64, box_outer, 1, 1};
    return get_tensor_map_nd(ptr, 4, dims, st, box, out);
  };
  CUtensorMap ta, tb;
  int rc;
  if (a_mn) rc = make_map(a->A, (uint64_t)a->M, (uint64_t)a->K, a->lda, a->sa0, a->sa1, p.a_z0, p.a_z1, xg::BK, &ta);
  else rc = make_map(a->A, (uint64_t)a->K, (uint64_t)a->M, a->lda, a->sa0, a->sa1, p.a_z0, p.a_z1, xg::BM, &ta);
  if (rc) return rc;
  if (b_mn) rc = make_map(a->B, (uint64_t)a->N, (uint64_t)a->K, a->ldb, a->sb0, a->sb1, p.b_z0, p.b_z1, xg::BK, &tb);
  else rc = make_map(a->B, (uint64_t)a->K, (uint64_t)a->N, a->ldb, a->sb0, a->sb1, p.b_z0, p.b_z1, (uint32_t)BN, &tb);
  if (rc) return rc;
  CUtensorMap tr = ta;
  p.prefetch_res = 0;
  p.r_z0 = (a->nb0 > 1 && a->sres0 != 0) ? 1 : 0;
  p.r_z1 = (a->nb1 > 1 && a->sres1 != 0) ? 1 : 0;
  static const bool res_prefetch = getenv("EGV_GEMM_RES_PREFETCH") != nullptr;   // off by default: see gemm.cu
  if (res_prefetch && a->residual && (long long)a->M * a->N * batches >= (1ll << 20)) {
    uint64_t dims[4] = {(uint64_t)a->N, (uint64_t)a->M, (uint64_t)(p.r_z0 ? a->nb0 : 1), (uint64_t)(p.r_z1 ? a->nb1 : 1)};
    uint64_t st[3] = {(uint64_t)a->ld_res * 4, (uint64_t)(p.r_z0 ? a->sres0 : a->ld_res) * 4, (uint64_t)(p.r_z1 ? a->sres1 : a->ld_res) * 4};
    uint32_t box[4] = {(uint32_t)(BN < a->N ? BN : a->N), (uint32_t)xg::BM, 1, 1};
    if (get_tensor_map_nd_ex(a->residual, 4, dims, st, box, 1, &tr) == EGV_OK) p.prefetch_res = 1;
  }
  switch (BN) {
    case 256: return xg::dispatch<256>(a->layout, ta, tb, tr, p, stream);
    case 192: return xg::dispatch<192>(a->layout, ta, tb, tr, p, stream);
    case 128: return xg::dispatch<128>(a->layout, ta, tb, tr, p, stream);
    default: return xg::dispatch<64>(a->layout, ta, tb, tr, p, stream);
  }
}
