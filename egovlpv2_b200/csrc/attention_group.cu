// Group-resident attention kernels (space attention of the divided space-time block: ~196 queries x 197 keys per
// (clip, frame, head) group; video_transformer.py:117-153).
//
// One persistent CTA per SM walks the (batch, group, head) problems.  A producer warp feeds shared memory with TMA
// (cp.async.bulk.tensor, 128B swizzle) one pass / one problem AHEAD of the 7 compute warps, which keep the "row side"
// of the problem (16 rows per warp, 112 rows per pass, two passes per group) in registers as mma.sync A fragments and
// walk the "stream side" -- the WHOLE group, resident in shared memory -- in 64-row chunks:
//     FWD : rows = queries        stream = keys (K, V)         O = softmax(QK^T) V, lse
//     DQ  : rows = queries        stream = keys (K, V)         dQ, delta = rowsum(dO * O)
//     DKV : rows = keys           stream = queries (Q, dO)     dK, dV (the shared CLS key goes to the fp32 accumulator)
// Compared with the generic kernels in attention.cu: the group's K/V are fetched once (not once per 64-row tile), the
// tiles are 112 x 208 instead of 256 x 256 for 196 x 197 real entries (1.2x instead of 1.7x padding), loads are issued
// by one thread instead of per-row cp.async address arithmetic in every thread, and they overlap the previous pass.
// Index convention inside the kernel: regular rows first, the shared CLS key LAST (index = number of regular rows).
#include <stdlib.h>

#include "attention.cuh"

namespace egv {

constexpr int GSCAP = 208;                     // stream-side capacity: 13 x 16 rows
constexpr int STR_TILE_BYTES = GSCAP * 128;    // 64 bf16 = 128 B per row
constexpr int STAT_BYTES = 1024;               // GSCAP floats, padded
constexpr int OUT_STAGE_BYTES = 16 * 128;

struct GroupMaps {
  CUtensorMap row0[3];   // row-side tensors, pass-0 box
  CUtensorMap row1[3];   // row-side tensors, pass-1 box
  CUtensorMap str[2];    // stream-side tensors (box = all regular stream rows)
};

struct GroupP {
  int rows_reg, row_cls;     // regular rows of the row side; 1 when the CLS key is a row (DKV with has_cls)
  int str_reg, str_cls;      // regular stream rows; 1 when the CLS key is a stream row (FWD / DQ with has_cls)
  int npass, r0, r1;         // passes per problem, TMA box rows of pass 0 / pass 1
  int cls_pass, cls_local;   // where the CLS row lives on the row side (DKV)
  long long total;           // problems = B * G * H
};

// GW compute warps (16 rows each) + 1 producer warp; KC = stream rows per chunk; RS = row-side stages.
// 13 warps cover a whole 196/197-row group in ONE pass; the register file (64 K) then allows ~146 registers per thread,
// which the backward modes reach with 32-row chunks.  The row side needs a single stage in that case: its tiles are
// dead as soon as every warp holds its fragments, i.e. the next problem's rows load during this problem's math.
template <int MODE, int GW_, int KC_, int RS_>
struct GCfg {
  static constexpr int GW = GW_, KC = KC_, RS = RS_;
  static constexpr int GROWS = 16 * GW;
  static constexpr int THREADS = (GW + 1) * 32;
  static constexpr int ROW_TILE_BYTES = ((GROWS * 128 + 1023) / 1024) * 1024;
  static constexpr int NR = MODE == MODE_FWD ? 1 : (MODE == MODE_DQ ? 3 : 2);
  static constexpr int NOUT = MODE == MODE_DKV ? 2 : 1;
  static constexpr int ROW_OFF = 0;
  static constexpr int STR_OFF = ROW_OFF + RS * NR * ROW_TILE_BYTES;
  static constexpr int STAT_OFF = STR_OFF + 4 * STR_TILE_BYTES;
  static constexpr int OUT_OFF = STAT_OFF + 4 * STAT_BYTES;
  static constexpr int BAR_OFF = OUT_OFF + GW * NOUT * OUT_STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 128 + 1024;   // barriers + alignment slack
};

// acc[nt] (16 x 8 tiles, nt < KC/8) = A (16 x 64 fragments) * T^T, T = KC rows of a swizzled tile starting at `tile`
template <int KC>
EGV_DEVINL void g_mma_nt(float (&acc)[KC / 8][4], const uint32_t (&af)[4][4], uint32_t tile, const LaneAddr& la) {
  // k-step outermost: consecutive HMMAs hit KC/8 different accumulators, so none waits for its predecessor's result
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int n2 = 0; n2 < KC / 16; ++n2) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4(b0, b1, b2, b3, tile + n2 * 2048 + la.nt_row + la.nt_x[kk]);
      mma_16816(acc[2 * n2], af[kk], b0, b1);
      mma_16816(acc[2 * n2 + 1], af[kk], b2, b3);
    }
  }
}
// out[dt] (16 x 8 tiles over the 64 columns) += P (16 x KC, C layout) * T, T = KC rows of a swizzled tile
template <int KC>
EGV_DEVINL void g_mma_p(float (&out)[8][4], const float (&pacc)[KC / 8][4], uint32_t tile, const LaneAddr& la) {
#pragma unroll
  for (int k2 = 0; k2 < KC / 16; ++k2) {
    uint32_t af[4];
    af[0] = pack_bf16(pacc[2 * k2][0], pacc[2 * k2][1]);
    af[1] = pack_bf16(pacc[2 * k2][2], pacc[2 * k2][3]);
    af[2] = pack_bf16(pacc[2 * k2 + 1][0], pacc[2 * k2 + 1][1]);
    af[3] = pack_bf16(pacc[2 * k2 + 1][2], pacc[2 * k2 + 1][3]);
#pragma unroll
    for (int d2 = 0; d2 < 4; ++d2) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4_trans(b0, b1, b2, b3, tile + k2 * 2048 + la.p_row + la.p_x[d2]);
      mma_16816(out[2 * d2], af, b0, b1);
      mma_16816(out[2 * d2 + 1], af, b2, b3);
    }
  }
}
EGV_DEVINL void g_load_a(uint32_t (&af)[4][4], uint32_t tile_rows, const LaneAddr& la) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldmatrix_x4(af[kk][0], af[kk][1], af[kk][2], af[kk][3], tile_rows + la.p_row + la.p_x[kk]);
}
// staging tile -> global rows (16-byte stores); row r of the tile is row idx0 + r of the tensor, valid below `limit`
EGV_DEVINL void g_store_rows(const uint8_t* stg, bf16* base, long long row_first, long long ld, int col0, int idx0, int limit,
                             int lane) {
  const int ch = lane & 7;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 3) + 4 * i;
    const int idx = idx0 + r;
    if (idx < limit) {
      const uint4 v = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch ^ (r & 7)) << 4));
      *reinterpret_cast<uint4*>(base + (row_first + idx) * ld + col0 + ch * 8) = v;
    }
  }
}

// One chunk of KC stream rows starting at c0.  FULL: no invalid stream rows inside the chunk.
template <int MODE, int KC, bool FULL>
EGV_DEVINL void g_chunk(int c0, int n_str, float scale2, const uint32_t (&fa)[4][4], const uint32_t (&fb)[4][4],
                        uint32_t strA, uint32_t strB, const float* st0, const float* st1, float (&acc0)[8][4],
                        float (&acc1)[8][4], float& m_lo, float& m_hi, float& l_lo, float& l_hi, float r0_lo, float r0_hi,
                        float r1_lo, float r1_hi, const LaneAddr& la, int tq) {
  const uint32_t tA = strA + c0 * 128, tB = strB + c0 * 128;
  const int lim = n_str - c0;
  float s[KC / 8][4];
#pragma unroll
  for (int i = 0; i < KC / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
  g_mma_nt<KC>(s, fa, tA, la);   // FWD/DQ: Q K^T    DKV: K Q^T
  if (MODE == MODE_FWD) {
    float cm_lo = -1e30f, cm_hi = -1e30f;
#pragma unroll
    for (int nt = 0; nt < KC / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = s[nt][e];
        if (!FULL) {
          const int col = nt * 8 + 2 * tq + (e & 1);
          v = col < lim ? v : -1e30f;
          s[nt][e] = v;
        }
        if (e < 2) cm_lo = fmaxf(cm_lo, v);
        else cm_hi = fmaxf(cm_hi, v);
      }
    }
    cm_lo = fmaxf(cm_lo, __shfl_xor_sync(0xffffffffu, cm_lo, 1));
    cm_lo = fmaxf(cm_lo, __shfl_xor_sync(0xffffffffu, cm_lo, 2));
    cm_hi = fmaxf(cm_hi, __shfl_xor_sync(0xffffffffu, cm_hi, 1));
    cm_hi = fmaxf(cm_hi, __shfl_xor_sync(0xffffffffu, cm_hi, 2));
    cm_lo *= scale2;   // to the scaled log2 domain (scale > 0); -1e30 stays hugely negative
    cm_hi *= scale2;
    const float mn_lo = fmaxf(m_lo, cm_lo), mn_hi = fmaxf(m_hi, cm_hi);
    const float al_lo = ex2(m_lo - mn_lo), al_hi = ex2(m_hi - mn_hi);
    m_lo = mn_lo;
    m_hi = mn_hi;
    l_lo *= al_lo;
    l_hi *= al_hi;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      acc0[nt][0] *= al_lo;
      acc0[nt][1] *= al_lo;
      acc0[nt][2] *= al_hi;
      acc0[nt][3] *= al_hi;
    }
#pragma unroll
    for (int nt = 0; nt < KC / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float pv = ex2(fmaf(s[nt][e], scale2, -(e < 2 ? m_lo : m_hi)));
        if (!FULL) {
          const int col = nt * 8 + 2 * tq + (e & 1);
          pv = col < lim ? pv : 0.f;
        }
        s[nt][e] = pv;
        if (e < 2) l_lo += pv;
        else l_hi += pv;
      }
    }
    g_mma_p<KC>(acc0, s, tB, la);   // O += P V
  } else {
#pragma unroll
    for (int nt = 0; nt < KC / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + 2 * tq + (e & 1);
        float pv;
        if (MODE == MODE_DQ) pv = ex2(fmaf(s[nt][e], scale2, -(e < 2 ? r0_lo : r0_hi)));   // row lse
        else pv = ex2(fmaf(s[nt][e], scale2, -st0[c0 + col]));                             // column (query) lse
        if (!FULL) pv = col < lim ? pv : 0.f;
        s[nt][e] = pv;
      }
    }
    if (MODE == MODE_DKV) g_mma_p<KC>(acc1, s, tB, la);   // dV += P^T dO
    float dp[KC / 8][4];
#pragma unroll
    for (int i = 0; i < KC / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dp[i][j] = 0.f;
    g_mma_nt<KC>(dp, fb, tB, la);   // DQ: dO V^T    DKV: V dO^T
#pragma unroll
    for (int nt = 0; nt < KC / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + 2 * tq + (e & 1);
        const float dl = (MODE == MODE_DQ) ? (e < 2 ? r1_lo : r1_hi) : st1[c0 + col];
        float ds = s[nt][e] * (dp[nt][e] - dl);
        if (!FULL) ds = col < lim ? ds : 0.f;
        s[nt][e] = ds;
      }
    }
    g_mma_p<KC>(acc0, s, tA, la);   // DQ: dQ += dS K    DKV: dK += dS^T Q
  }
}

template <int MODE, int GW, int KC, int RS>
// (the register file is 4 x 16 K registers, one slice per SM sub-partition: with 14 warps some sub-partition holds 4 of
// them, so the per-thread budget is 16384 / (4 * 32) = 128 registers, which is what __launch_bounds__(448) yields)
__global__ void __launch_bounds__((GW + 1) * 32, 1)
attn_group_kernel(const __grid_constant__ GroupMaps maps, const AttnP a, const GroupP gp) {
  using Cfg = GCfg<MODE, GW, KC, RS>;
  constexpr int NR = Cfg::NR;
  constexpr int G_THREADS = Cfg::THREADS, GROWS = Cfg::GROWS, ROW_TILE_BYTES = Cfg::ROW_TILE_BYTES;
  extern __shared__ __align__(1024) uint8_t gsm_raw[];
  const uint32_t pad = (1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u;
  uint8_t* smem = gsm_raw + pad;
  uint8_t* rowbuf = smem + Cfg::ROW_OFF;
  uint8_t* strbuf = smem + Cfg::STR_OFF;
  uint8_t* statbuf = smem + Cfg::STAT_OFF;
  uint8_t* outbuf = smem + Cfg::OUT_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* row_full = bars;        // [RS]
  uint64_t* row_empty = bars + 2;   // [RS]
  uint64_t* str_full = bars + 4;    // [2]
  uint64_t* str_empty = bars + 6;   // [2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool row_manual = (MODE == MODE_DKV) && a.has_cls;   // producer lanes write the CLS row of the row side
  const bool str_manual = (MODE != MODE_DKV) && a.has_cls;   // ... of the stream side

  // zero every tile once: rows the TMA boxes never touch (padding up to the 16-row granularity) must hold finite values
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = Cfg::OUT_OFF / 16;
    for (int i = threadIdx.x; i < n16; i += G_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();   // generic-proxy zero fill before async-proxy (TMA) writes to the same tiles
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < RS; ++s) {
      mbar_init(&row_full[s], row_manual ? 2 : 1);
      mbar_init(&row_empty[s], GW);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&str_full[s], str_manual ? 2 : 1);
      mbar_init(&str_empty[s], GW);
    }
    fence_barrier_init();
    for (int t = 0; t < NR; ++t) {
      tma_prefetch_desc(&maps.row0[t]);
      tma_prefetch_desc(&maps.row1[t]);
    }
    tma_prefetch_desc(&maps.str[0]);
    tma_prefetch_desc(&maps.str[1]);
  }
  __syncthreads();
  pdl_wait();    // programmatic dependent launch (common.cuh): shared memory zeroed, barriers up; global accesses from here

  const int HG = a.H * a.G;
  if (warp == GW) {
    // ------------------------------------------------------------------------------------------ producer
    int rs = 0, ks = 0;
    uint32_t rph = 0, kph = 0;
    for (long long p = blockIdx.x; p < gp.total; p += gridDim.x) {
      const int h = (int)(p % a.H), g = (int)((p / a.H) % a.G), b = (int)(p / HG);
      const long long q_first = (long long)b * a.q_bstride + a.q_row0 + (long long)g * a.q_gstride;
      const long long o_first = (long long)b * a.o_bstride + a.q_row0 + (long long)g * a.q_gstride;
      const long long k_first = (long long)b * a.kv_bstride + a.k_row0 + (long long)g * a.k_gstride;
      const long long cls_off = ((long long)b * a.kv_bstride + a.cls_row) * a.ldkv + h * HD + (lane & 7) * 8;
      const long long stat_base = (((long long)b * a.H + h) * a.G + g) * a.Lq;
      // ---- stream side of this problem
      mbar_wait_sleep(&str_empty[ks], kph ^ 1, 100);
      uint8_t* sA = strbuf + (ks * 2 + 0) * STR_TILE_BYTES;
      uint8_t* sB = strbuf + (ks * 2 + 1) * STR_TILE_BYTES;
      if (lane == 0) {
        uint32_t bytes = 2u * (uint32_t)gp.str_reg * 128u;
        if (MODE == MODE_DKV) bytes += 2u * (uint32_t)a.Lq * 4u;
        mbar_arrive_expect_tx(&str_full[ks], bytes);
        const int rA = (int)(MODE == MODE_DKV ? q_first : k_first);
        const int rB = (int)(MODE == MODE_DKV ? o_first : k_first);
        tma_load_2d(sA, &maps.str[0], &str_full[ks], h * HD, rA);
        tma_load_2d(sB, &maps.str[1], &str_full[ks], h * HD, rB);
        if (MODE == MODE_DKV) {
          bulk_copy_g2s(statbuf + (ks * 2 + 0) * STAT_BYTES, a.lse + stat_base, (uint32_t)a.Lq * 4u, &str_full[ks]);
          bulk_copy_g2s(statbuf + (ks * 2 + 1) * STAT_BYTES, a.delta + stat_base, (uint32_t)a.Lq * 4u, &str_full[ks]);
        }
      }
      if (str_manual) {
        if (lane < 16) {
          const uint4 val = *reinterpret_cast<const uint4*>((lane < 8 ? a.k : a.v) + cls_off);
          const int r = gp.str_reg;
          *reinterpret_cast<uint4*>((lane < 8 ? sA : sB) + r * 128 + (((lane & 7) ^ (r & 7)) << 4)) = val;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&str_full[ks]);
      }
      if (++ks == 2) {
        ks = 0;
        kph ^= 1;
      }
      // ---- row side, pass by pass
      for (int pass = 0; pass < gp.npass; ++pass) {
        mbar_wait_sleep(&row_empty[rs], rph ^ 1, 100);
        const int rows = pass == 0 ? gp.r0 : gp.r1;
        uint8_t* rb = rowbuf + rs * NR * ROW_TILE_BYTES;
        if (lane == 0) {
          if (rows > 0) {
            mbar_arrive_expect_tx(&row_full[rs], (uint32_t)(NR * rows) * 128u);
            const CUtensorMap* mp = pass == 0 ? maps.row0 : maps.row1;
            if (MODE == MODE_FWD) {
              tma_load_2d(rb, &mp[0], &row_full[rs], h * HD, (int)q_first + pass * GROWS);
            } else if (MODE == MODE_DQ) {
              tma_load_2d(rb, &mp[0], &row_full[rs], h * HD, (int)q_first + pass * GROWS);
              tma_load_2d(rb + ROW_TILE_BYTES, &mp[1], &row_full[rs], h * HD, (int)o_first + pass * GROWS);
              tma_load_2d(rb + 2 * ROW_TILE_BYTES, &mp[2], &row_full[rs], h * HD, (int)o_first + pass * GROWS);
            } else {
              tma_load_2d(rb, &mp[0], &row_full[rs], h * HD, (int)k_first + pass * GROWS);
              tma_load_2d(rb + ROW_TILE_BYTES, &mp[1], &row_full[rs], h * HD, (int)k_first + pass * GROWS);
            }
          } else {
            mbar_arrive(&row_full[rs]);
          }
        }
        if (row_manual) {
          if (pass == gp.cls_pass && lane < 16) {
            const uint4 val = *reinterpret_cast<const uint4*>((lane < 8 ? a.k : a.v) + cls_off);
            const int r = gp.cls_local;
            *reinterpret_cast<uint4*>(rb + (lane < 8 ? 0 : ROW_TILE_BYTES) + r * 128 + (((lane & 7) ^ (r & 7)) << 4)) = val;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&row_full[rs]);
        }
        if (++rs == RS) {
          rs = 0;
          rph ^= 1;
        }
      }
    }
    return;
  }

  // ---------------------------------------------------------------------------------------------- compute warps
  const int wq = warp;
  const int gq = lane >> 2, tq = lane & 3;
  const LaneAddr la = lane_addr(lane);
  const float scale2 = a.scale * LOG2E;
  const int n_rows = gp.rows_reg + gp.row_cls;
  const int n_str = gp.str_reg + gp.str_cls;
  uint8_t* stg0 = outbuf + (wq * Cfg::NOUT) * OUT_STAGE_BYTES;
  int rs = 0, ks = 0;
  uint32_t rph = 0, kph = 0;
  for (long long p = blockIdx.x; p < gp.total; p += gridDim.x) {
    const int h = (int)(p % a.H), g = (int)((p / a.H) % a.G), b = (int)(p / HG);
    const long long q_first = (long long)b * a.q_bstride + a.q_row0 + (long long)g * a.q_gstride;
    const long long o_first = (long long)b * a.o_bstride + a.q_row0 + (long long)g * a.q_gstride;
    const long long k_first = (long long)b * a.kv_bstride + a.k_row0 + (long long)g * a.k_gstride;
    const long long stat_base = (((long long)b * a.H + h) * a.G + g) * a.Lq;
    const uint32_t strA = smem_u32(strbuf + (ks * 2 + 0) * STR_TILE_BYTES);
    const uint32_t strB = smem_u32(strbuf + (ks * 2 + 1) * STR_TILE_BYTES);
    const float* st0 = reinterpret_cast<const float*>(statbuf + (ks * 2 + 0) * STAT_BYTES);
    const float* st1 = reinterpret_cast<const float*>(statbuf + (ks * 2 + 1) * STAT_BYTES);
    for (int pass = 0; pass < gp.npass; ++pass) {
      const int idx0 = pass * GROWS + wq * 16;   // first row (index on the row side) of this warp's tile
      mbar_wait(&row_full[rs], rph);
      const uint8_t* rb = rowbuf + rs * NR * ROW_TILE_BYTES;
      const uint32_t rows_u32 = smem_u32(rb) + wq * 2048;
      uint32_t fa[4][4], fb[4][4];
      g_load_a(fa, rows_u32, la);
      if (MODE != MODE_FWD) g_load_a(fb, rows_u32 + ROW_TILE_BYTES, la);
      const int r_lo = idx0 + gq, r_hi = r_lo + 8;
      float r0_lo = 0.f, r0_hi = 0.f, r1_lo = 0.f, r1_hi = 0.f;
      if (MODE == MODE_DQ) {
        // delta_i = sum_d dO[i, d] * O[i, d] from the staged tiles (one row per iteration, 2 columns per lane)
        const uint8_t* t_do = rb + ROW_TILE_BYTES + wq * 2048;
        const uint8_t* t_o = rb + 2 * ROW_TILE_BYTES + wq * 2048;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int off = r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + 4 * (lane & 3);
          const float2 x = unpack_bf16(*reinterpret_cast<const uint32_t*>(t_do + off));
          const float2 y = unpack_bf16(*reinterpret_cast<const uint32_t*>(t_o + off));
          const float part = warp_sum(x.x * y.x + x.y * y.y);
          if (idx0 + r < n_rows && lane == 0) a.delta[stat_base + idx0 + r] = part;
          if (r == gq) r1_lo = part;
          if (r == gq + 8) r1_hi = part;
        }
        r0_lo = r_lo < n_rows ? a.lse[stat_base + r_lo] : 0.f;
        r0_hi = r_hi < n_rows ? a.lse[stat_base + r_hi] : 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&row_empty[rs]);   // fragments are in registers: the producer may refill this stage
      if (++rs == RS) {
        rs = 0;
        rph ^= 1;
      }
      if (pass == 0) mbar_wait(&str_full[ks], kph);

      float acc0[8][4], acc1[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc0[i][j] = acc1[i][j] = 0.f;
      float m_lo = -1e30f, m_hi = -1e30f, l_lo = 0.f, l_hi = 0.f;
      if (idx0 < n_rows) {   // warp-uniform: tiles entirely past the last row have nothing to do
        int c0 = 0;
#pragma unroll 1
        for (; c0 + KC <= n_str; c0 += KC)
          g_chunk<MODE, KC, true>(c0, n_str, scale2, fa, fb, strA, strB, st0, st1, acc0, acc1, m_lo, m_hi, l_lo, l_hi, r0_lo,
                                  r0_hi, r1_lo, r1_hi, la, tq);
#pragma unroll 1
        for (; c0 < n_str; c0 += 16) {   // tail in 16-row pieces (197 keys = 3 x 64 + 5 or 6 x 32 + 5)
          if (n_str - c0 >= 16)
            g_chunk<MODE, 16, true>(c0, n_str, scale2, fa, fb, strA, strB, st0, st1, acc0, acc1, m_lo, m_hi, l_lo, l_hi,
                                    r0_lo, r0_hi, r1_lo, r1_hi, la, tq);
          else
            g_chunk<MODE, 16, false>(c0, n_str, scale2, fa, fb, strA, strB, st0, st1, acc0, acc1, m_lo, m_hi, l_lo, l_hi,
                                     r0_lo, r0_hi, r1_lo, r1_hi, la, tq);
        }
      }
      if (pass == gp.npass - 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&str_empty[ks]);   // this warp is done with the problem's stream tiles
        if (++ks == 2) {
          ks = 0;
          kph ^= 1;
        }
      }
      // ---- write-out
      if (idx0 < n_rows) {
        if (MODE == MODE_FWD) {
          l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
          l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
          l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
          l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
          g_stage(stg0, acc0, 1.0f / l_lo, 1.0f / l_hi, lane);
          if (tq == 0) {
            if (r_lo < n_rows) a.lse[stat_base + r_lo] = m_lo + log2f(l_lo);
            if (r_hi < n_rows) a.lse[stat_base + r_hi] = m_hi + log2f(l_hi);
          }
          __syncwarp();
          g_store_rows(stg0, a.o, o_first, a.ldo, h * HD, idx0, n_rows, lane);
        } else if (MODE == MODE_DQ) {
          g_stage(stg0, acc0, a.scale, a.scale, lane);
          __syncwarp();
          g_store_rows(stg0, a.dq, q_first, a.lddq, h * HD, idx0, n_rows, lane);
        } else {
          g_stage(stg0, acc0, a.scale, a.scale, lane);
          g_stage(stg0 + OUT_STAGE_BYTES, acc1, 1.0f, 1.0f, lane);
          __syncwarp();
          g_store_rows(stg0, a.dk, k_first, a.lddkv, h * HD, idx0, gp.rows_reg, lane);
          g_store_rows(stg0 + OUT_STAGE_BYTES, a.dv, k_first, a.lddkv, h * HD, idx0, gp.rows_reg, lane);
          if (gp.row_cls && a.dkv_cls) {
            const int cr = gp.rows_reg - idx0;   // row of the CLS key inside this warp's tile, if in [0, 16)
            if (cr >= 0 && cr < 16 && gq == (cr & 7)) {
              const bool lo = cr < 8;
              float* dst = a.dkv_cls + ((long long)b * a.H + h) * 2 * HD;
#pragma unroll
              for (int nt = 0; nt < 8; ++nt) {
                atomicAdd(dst + nt * 8 + 2 * tq, (lo ? acc0[nt][0] : acc0[nt][2]) * a.scale);
                atomicAdd(dst + nt * 8 + 2 * tq + 1, (lo ? acc0[nt][1] : acc0[nt][3]) * a.scale);
                atomicAdd(dst + HD + nt * 8 + 2 * tq, lo ? acc1[nt][0] : acc1[nt][2]);
                atomicAdd(dst + HD + nt * 8 + 2 * tq + 1, lo ? acc1[nt][1] : acc1[nt][3]);
              }
            }
          }
        }
        __syncwarp();   // staging tile is reused by the next pass
      }
    }
  }
}

static int g_group_mode = -1;   // env EGV_ATTN_GROUP: bit per MODE (default 7 = all)

template <int MODE, int GW, int KC, int RS>
static int launch_group(const GroupMaps& maps, const AttnP& a, const GroupP& gp, cudaStream_t stream) {
  using Cfg = GCfg<MODE, GW, KC, RS>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "group attention tile set exceeds the 227 KB shared memory of an SM");
  auto kern = attn_group_kernel<MODE, GW, KC, RS>;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "group attention smem attribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  const long long grid = gp.total < sm_count() ? gp.total : sm_count();
  launch_k(kern, dim3((unsigned)grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, maps, a, gp);
  int rc = check_launch("attn_group_kernel");
  return rc ? rc : 1;
}

static int g_group_wide = -1;   // env EGV_ATTN_GROUP_WIDE, bit per MODE: 1 = 13 compute warps / one pass per group, 0 = 7 warps / two passes
                                // (default 1: forward only -- measured: the 128-register budget of 14 warps costs the backward 20 %)

int launch_group_attention(int mode, const AttnP& a, cudaStream_t stream) {
  if (g_group_mode < 0) g_group_mode = getenv("EGV_ATTN_GROUP") ? atoi(getenv("EGV_ATTN_GROUP")) : 7;
  if (!((g_group_mode >> mode) & 1)) return 0;
  if (a.key_bias || a.q_istride != 1 || a.k_istride != 1) return 0;
  if (mode != MODE_FWD && a.dkv_accumulate) return 0;
  GroupP gp;
  const int lk_reg = a.LkT - (a.has_cls ? 1 : 0);
  if (mode == MODE_DKV) {
    gp.rows_reg = lk_reg; gp.row_cls = a.has_cls ? 1 : 0;
    gp.str_reg = a.Lq; gp.str_cls = 0;
    if (a.Lq % 4) return 0;   // 16-byte bulk copies of lse / delta
    if (a.has_cls && !a.dkv_cls) return 0;
  } else {
    gp.rows_reg = a.Lq; gp.row_cls = 0;
    gp.str_reg = lk_reg; gp.str_cls = a.has_cls ? 1 : 0;
  }
  const int n_rows = gp.rows_reg + gp.row_cls, n_str = gp.str_reg + gp.str_cls;
  if (g_group_wide < 0) g_group_wide = getenv("EGV_ATTN_GROUP_WIDE") ? atoi(getenv("EGV_ATTN_GROUP_WIDE")) : 1;
  const bool wide = (g_group_wide >> mode) & 1;
  const int GROWS = 16 * (wide ? 13 : 7);
  if (gp.rows_reg < 64 || n_rows > 2 * GROWS || gp.str_reg < 16 || n_str > GSCAP) return 0;
  gp.npass = (int)cdiv(n_rows, GROWS);
  gp.r0 = gp.rows_reg < GROWS ? gp.rows_reg : GROWS;
  gp.r1 = gp.rows_reg - gp.r0;
  gp.cls_pass = gp.rows_reg / GROWS;
  gp.cls_local = gp.rows_reg - gp.cls_pass * GROWS;
  gp.total = (long long)a.B * a.G * a.H;
  if (gp.total <= 0) return 0;

  GroupMaps maps;
  const uint64_t width = (uint64_t)a.H * HD;
  const uint64_t q_rows = (uint64_t)a.B * a.q_bstride, kv_rows = (uint64_t)a.B * a.kv_bstride, o_rows = (uint64_t)a.B * a.o_bstride;
  struct T { const void* p; uint64_t rows, ld; };
  T rowt[3], strt[2];
  int nr;
  if (mode == MODE_FWD) {
    nr = 1; rowt[0] = {a.q, q_rows, (uint64_t)a.ldq};
    strt[0] = {a.k, kv_rows, (uint64_t)a.ldkv}; strt[1] = {a.v, kv_rows, (uint64_t)a.ldkv};
  } else if (mode == MODE_DQ) {
    nr = 3; rowt[0] = {a.q, q_rows, (uint64_t)a.ldq}; rowt[1] = {a.d_o, o_rows, (uint64_t)a.ldo}; rowt[2] = {a.o, o_rows, (uint64_t)a.ldo};
    strt[0] = {a.k, kv_rows, (uint64_t)a.ldkv}; strt[1] = {a.v, kv_rows, (uint64_t)a.ldkv};
  } else {
    nr = 2; rowt[0] = {a.k, kv_rows, (uint64_t)a.ldkv}; rowt[1] = {a.v, kv_rows, (uint64_t)a.ldkv};
    strt[0] = {a.q, q_rows, (uint64_t)a.ldq}; strt[1] = {a.d_o, o_rows, (uint64_t)a.ldo};
  }
  int rc;
  for (int t = 0; t < 3; ++t) {
    const T& x = rowt[t < nr ? t : 0];
    if ((rc = get_tensor_map(x.p, width, x.rows, x.ld, 64, (uint32_t)gp.r0, &maps.row0[t]))) return rc;
    if ((rc = get_tensor_map(x.p, width, x.rows, x.ld, 64, (uint32_t)(gp.r1 > 0 ? gp.r1 : gp.r0), &maps.row1[t]))) return rc;
  }
  for (int t = 0; t < 2; ++t)
    if ((rc = get_tensor_map(strt[t].p, width, strt[t].rows, strt[t].ld, 64, (uint32_t)gp.str_reg, &maps.str[t]))) return rc;
  switch (mode) {
    // <mode, compute warps, chunk rows, row-side stages>
    case MODE_FWD: return wide ? launch_group<MODE_FWD, 13, 64, 2>(maps, a, gp, stream) : launch_group<MODE_FWD, 7, 64, 2>(maps, a, gp, stream);
    case MODE_DQ: return wide ? launch_group<MODE_DQ, 13, 32, 1>(maps, a, gp, stream) : launch_group<MODE_DQ, 7, 64, 2>(maps, a, gp, stream);
    default: return wide ? launch_group<MODE_DKV, 13, 32, 1>(maps, a, gp, stream) : launch_group<MODE_DKV, 7, 64, 2>(maps, a, gp, stream);
  }
}

}  // namespace egv
