// Fused attention kernels for MANY TINY groups -- the time attention of the divided space-time block
// (video_transformer.py:117-153, 'b (f n) d -> (b n) f d'): per (clip, patch position n, head) T <= 16 queries attend
// T <= 16 keys plus the shared CLS key; 18 816 such groups per call at cfg 3, each touching ~6 KB.  The work is pure
// HBM traffic, so the kernels are built around the memory system:
//   * a producer warp fetches, per (batch, head, tile of 8 ADJACENT groups), every operand as ONE 4-D TMA box
//     {64 head columns, 8 groups, 16 frames, 1 clip} (rows of one group are Nf rows apart, adjacent groups are
//     adjacent rows), 2-3 stages ahead of the math; out-of-range groups / frames are zero-filled by the TMA unit;
//   * 8 compute warps take one group each, entirely in registers (mma.sync m16n8k16: one 16 x 16 (+CLS) score tile);
//   * the BACKWARD is one launch instead of two: dQ from the query-row view, dK / dV from the transposed (key-row)
//     view recomputed from the same shared-memory tiles, the CLS key as an extra 1-row key tile whose gradient is
//     carried in registers across the groups of a (clip, head) and flushed with a handful of atomics.
//     Operands are read once (5 tensors) instead of 9 times by the generic DQ + DKV pair.
#include <stdlib.h>

#include "attention.cuh"

namespace egv {

constexpr int TG = 8;                        // groups per tile = compute warps
constexpr int TL = 16;                       // rows (frames) per group, padded
constexpr int T_THREADS = (TG + 1) * 32;
constexpr int T_TILE_BYTES = TL * TG * 128;  // 16 KB: [frame i][group g][64 bf16], row = i * 8 + g, 128B swizzle
constexpr int T_X_BYTES = 16 * 128;          // extra key tile: row 0 = CLS key (or value), rows 1..15 zero
constexpr int T_STAT_BYTES = 1024;           // lse of the tile's groups (backward): TG * TL floats, padded to the swizzle period

struct TinyMaps {
  CUtensorMap t[5];   // q, k, v, d_o, o
};
struct TinyP {
  int n_gt;           // group tiles per (batch, head)
  long long items;    // B * H * n_gt
  int istride;        // rows between consecutive frames of a group
  int lk_reg;         // regular keys per group
};

template <bool BWD>
struct TCfg {
  static constexpr int NT = BWD ? 5 : 3;
  static constexpr int STAGES = BWD ? 2 : 3;
  static constexpr int STAGE_BYTES = NT * T_TILE_BYTES + 2 * T_X_BYTES + (BWD ? T_STAT_BYTES : 0);
  static constexpr int OUT_OFF = STAGES * STAGE_BYTES;
  static constexpr int DELTA_OFF = OUT_OFF + TG * 2048;
  static constexpr int BAR_OFF = DELTA_OFF + TG * TL * 4;
  static constexpr int SMEM_BYTES = BAR_OFF + 128 + 1024;
};

EGV_DEVINL void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// per-lane byte offsets into a group tile for warp (= local group) w:
//   a[kk]: A fragments (rows = frames) and transposed B loads (rows = k index);  n[kk]: B loads with rows = n index
struct TinyAddr {
  uint32_t a[4], n[4];
};
EGV_DEVINL TinyAddr tiny_addr(int lane, int w) {
  TinyAddr t;
  const uint32_t ia = (lane & 7) + ((lane >> 3) & 1) * 8, jn = (lane & 7) + ((lane >> 4) << 3);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    t.a[kk] = ia * 1024 + w * 128 + ((((uint32_t)(kk * 2 + (lane >> 4))) ^ (uint32_t)w) << 4);
    t.n[kk] = jn * 1024 + w * 128 + ((((uint32_t)(kk * 2 + ((lane >> 3) & 1))) ^ (uint32_t)w) << 4);
  }
  return t;
}

EGV_DEVINL void t_load_a(uint32_t (&f)[4][4], uint32_t tile, const TinyAddr& ta) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(f[kk][0], f[kk][1], f[kk][2], f[kk][3], tile + ta.a[kk]);
}
// s[0..1] (16 x 16) = A * T^T for a group tile T;  s[2] (16 x 8, column 0 = CLS) = A * X^T for the extra tile X
EGV_DEVINL void t_mma_nt(float (&s)[3][4], const uint32_t (&f)[4][4], uint32_t tile, uint32_t xtile, bool with_x,
                         const TinyAddr& ta, const LaneAddr& la) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t b0, b1, b2, b3;
    ldmatrix_x4(b0, b1, b2, b3, tile + ta.n[kk]);
    mma_16816(s[0], f[kk], b0, b1);
    mma_16816(s[1], f[kk], b2, b3);
    if (with_x) {
      ldmatrix_x4(b0, b1, b2, b3, xtile + la.nt_row + la.nt_x[kk]);
      mma_16816(s[2], f[kk], b0, b1);
    }
  }
}
// out (16 x 64) += P (16 x 16 regular [+ CLS column]) * T (rows = k index) [+ X]
EGV_DEVINL void t_mma_p(float (&out)[8][4], const float (&p)[3][4], uint32_t tile, uint32_t xtile, bool with_x,
                        const TinyAddr& ta, const LaneAddr& la) {
  uint32_t af[4];
  af[0] = pack_bf16(p[0][0], p[0][1]);
  af[1] = pack_bf16(p[0][2], p[0][3]);
  af[2] = pack_bf16(p[1][0], p[1][1]);
  af[3] = pack_bf16(p[1][2], p[1][3]);
#pragma unroll
  for (int d2 = 0; d2 < 4; ++d2) {
    uint32_t b0, b1, b2, b3;
    ldmatrix_x4_trans(b0, b1, b2, b3, tile + ta.a[d2]);
    mma_16816(out[2 * d2], af, b0, b1);
    mma_16816(out[2 * d2 + 1], af, b2, b3);
  }
  if (with_x) {
    af[0] = pack_bf16(p[2][0], p[2][1]);
    af[1] = pack_bf16(p[2][2], p[2][3]);
    af[2] = 0u;
    af[3] = 0u;
#pragma unroll
    for (int d2 = 0; d2 < 4; ++d2) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4_trans(b0, b1, b2, b3, xtile + la.p_row + la.p_x[d2]);
      mma_16816(out[2 * d2], af, b0, b1);
      mma_16816(out[2 * d2 + 1], af, b2, b3);
    }
  }
}
// staging tile (16 rows) -> rows first + r * rstride of a global tensor, rows r < limit
EGV_DEVINL void t_store_rows(const uint8_t* stg, bf16* base, long long first, long long rstride, long long ld, int col0,
                             int limit, int lane) {
  const int ch = lane & 7;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 3) + 4 * i;
    if (r < limit) {
      const uint4 v = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch ^ (r & 7)) << 4));
      *reinterpret_cast<uint4*>(base + (first + (long long)r * rstride) * ld + col0 + ch * 8) = v;
    }
  }
}

// 9 warps: one SM sub-partition holds 3 of them, so the budget is 16384 / (3 * 32) = 170 registers per thread
template <bool BWD>
__global__ void __maxnreg__(168)
attn_tiny_kernel(const __grid_constant__ TinyMaps maps, const AttnP a, const TinyP tp) {
  using Cfg = TCfg<BWD>;
  constexpr int NT = Cfg::NT, STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t tsm_raw[];
  const uint32_t pad = (1024u - (smem_u32(tsm_raw) & 1023u)) & 1023u;
  uint8_t* smem = tsm_raw + pad;
  uint8_t* outbuf = smem + Cfg::OUT_OFF;
  float* sdelta = reinterpret_cast<float*>(smem + Cfg::DELTA_OFF);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_cls = a.has_cls != 0;
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < Cfg::OUT_OFF / 16; i += T_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], has_cls ? 2 : 1);
      mbar_init(&empty[s], TG);
    }
    fence_barrier_init();
    for (int t = 0; t < NT; ++t) tma_prefetch_desc(&maps.t[t]);
  }
  __syncthreads();

  // contiguous range of (batch, head, group tile) items per CTA: consecutive items share (batch, head)
  const long long per = (tp.items + gridDim.x - 1) / gridDim.x;
  const long long it0 = (long long)blockIdx.x * per;
  const long long it1 = it0 + per < tp.items ? it0 + per : tp.items;

  if (warp == TG) {
    // ------------------------------------------------------------------------------------------ producer
    int st = 0;
    uint32_t ph = 0;
    for (long long it = it0; it < it1; ++it) {
      const int gt = (int)(it % tp.n_gt);
      const int h = (int)((it / tp.n_gt) % a.H), b = (int)(it / ((long long)tp.n_gt * a.H));
      const int g0 = gt * TG;
      mbar_wait_sleep(&empty[st], ph ^ 1, 100);
      uint8_t* sb = smem + st * Cfg::STAGE_BYTES;
      if (lane == 0) {
        uint32_t bytes = (uint32_t)NT * T_TILE_BYTES;
        const int ng = a.G - g0 < TG ? a.G - g0 : TG;
        if (BWD) bytes += (uint32_t)(ng * a.Lq) * 4u;
        mbar_arrive_expect_tx(&full[st], bytes);
#pragma unroll
        for (int t = 0; t < NT; ++t) tma_load_4d(sb + t * T_TILE_BYTES, &maps.t[t], &full[st], h * HD, g0, 0, b);
        if (BWD) {
          const long long stat_base = (((long long)b * a.H + h) * a.G + g0) * a.Lq;
          bulk_copy_g2s(sb + NT * T_TILE_BYTES + 2 * T_X_BYTES, a.lse + stat_base, (uint32_t)(ng * a.Lq) * 4u, &full[st]);
        }
      }
      if (has_cls) {
        if (lane < 16) {
          const long long off = ((long long)b * a.kv_bstride + a.cls_row) * a.ldkv + h * HD + (lane & 7) * 8;
          const uint4 val = *reinterpret_cast<const uint4*>((lane < 8 ? a.k : a.v) + off);
          // row 0 of the extra tile: chunk c at c ^ 0
          *reinterpret_cast<uint4*>(sb + NT * T_TILE_BYTES + (lane < 8 ? 0 : T_X_BYTES) + ((lane & 7) << 4)) = val;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[st]);
      }
      if (++st == STAGES) {
        st = 0;
        ph ^= 1;
      }
    }
    return;
  }

  // ---------------------------------------------------------------------------------------------- compute warps
  const int w = warp;
  const int gq = lane >> 2, tq = lane & 3;
  const TinyAddr ta = tiny_addr(lane, w);
  const LaneAddr la = lane_addr(lane);
  const float scale2 = a.scale * LOG2E;
  uint8_t* stg = outbuf + w * 2048;
  float* wdelta = sdelta + w * TL;
  // running gradient of the shared CLS key / value (row 0 of the extra key tile) over the groups of one (batch, head)
  float clsk[8][4], clsv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) clsk[i][j] = clsv[i][j] = 0.f;
  long long cls_bh = -1;
  auto flush_cls = [&]() {
    if (BWD && has_cls && cls_bh >= 0 && gq == 0) {
      float* dst = a.dkv_cls + cls_bh * 2 * HD;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        atomicAdd(dst + nt * 8 + 2 * tq, clsk[nt][0] * a.scale);
        atomicAdd(dst + nt * 8 + 2 * tq + 1, clsk[nt][1] * a.scale);
        atomicAdd(dst + HD + nt * 8 + 2 * tq, clsv[nt][0]);
        atomicAdd(dst + HD + nt * 8 + 2 * tq + 1, clsv[nt][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) clsk[i][j] = clsv[i][j] = 0.f;
  };

  int st = 0;
  uint32_t ph = 0;
  for (long long it = it0; it < it1; ++it) {
    const int gt = (int)(it % tp.n_gt);
    const int h = (int)((it / tp.n_gt) % a.H), b = (int)(it / ((long long)tp.n_gt * a.H));
    const int g = gt * TG + w;
    const long long bh = (long long)b * a.H + h;
    if (BWD && bh != cls_bh) {
      flush_cls();
      cls_bh = bh;
    }
    mbar_wait(&full[st], ph);
    const uint8_t* sb = smem + st * Cfg::STAGE_BYTES;
    const uint32_t tQ = smem_u32(sb), tK = tQ + T_TILE_BYTES, tV = tQ + 2 * T_TILE_BYTES;
    const uint32_t tDO = tQ + 3 * T_TILE_BYTES;
    const uint32_t xK = tQ + NT * T_TILE_BYTES, xV = xK + T_X_BYTES;
    const float* slse = reinterpret_cast<const float*>(sb + NT * T_TILE_BYTES + 2 * T_X_BYTES) + w * a.Lq;
    if (g < a.G) {   // warp-uniform
      const long long q_first = (long long)b * a.q_bstride + a.q_row0 + g;
      const long long o_first = (long long)b * a.o_bstride + a.q_row0 + g;
      const long long k_first = (long long)b * a.kv_bstride + a.k_row0 + g;
      const long long stat_base = ((bh * a.G) + g) * a.Lq;
      uint32_t fq[4][4];
      t_load_a(fq, tQ, ta);
      float s[3][4];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
      t_mma_nt(s, fq, tK, xK, has_cls, ta, la);   // S = Q K^T  (16 x 16 regular, column 16 = CLS)
      if (!BWD) {
        float m_lo = -1e30f, m_hi = -1e30f;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = nt * 8 + 2 * tq + (e & 1);
            const bool ok = nt < 2 ? col < tp.lk_reg : (has_cls && col == 16);
            const float v = ok ? s[nt][e] * scale2 : -1e30f;
            s[nt][e] = v;
            if (e < 2) m_lo = fmaxf(m_lo, v);
            else m_hi = fmaxf(m_hi, v);
          }
        m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
        m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
        m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
        m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
        float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pv = s[nt][e] > -1e29f ? ex2(s[nt][e] - (e < 2 ? m_lo : m_hi)) : 0.f;
            s[nt][e] = pv;
            if (e < 2) l_lo += pv;
            else l_hi += pv;
          }
        l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
        l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
        l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
        l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        t_mma_p(acc, s, tV, xV, has_cls, ta, la);   // O = P V
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        g_stage(stg, acc, 1.0f / l_lo, 1.0f / l_hi, lane);
        if (tq == 0) {
          if (gq < a.Lq) a.lse[stat_base + gq] = m_lo + log2f(l_lo);
          if (gq + 8 < a.Lq) a.lse[stat_base + gq + 8] = m_hi + log2f(l_hi);
        }
        __syncwarp();
        t_store_rows(stg, a.o, o_first, tp.istride, a.ldo, h * HD, a.Lq, lane);
        __syncwarp();
      } else {
        const uint32_t tO = tQ + 4 * T_TILE_BYTES;
        uint32_t fdo[4][4];
        t_load_a(fdo, tDO, ta);
        // delta_i = rowsum(dO * O): one row per iteration, 2 columns per lane
        float d_lo = 0.f, d_hi = 0.f;
#pragma unroll
        for (int r = 0; r < TL; ++r) {
          const uint32_t off = r * 1024 + w * 128 + ((((uint32_t)(lane >> 2)) ^ (uint32_t)w) << 4) + 4 * (lane & 3);
          const float2 x = unpack_bf16(*reinterpret_cast<const uint32_t*>(sb + 3 * T_TILE_BYTES + off));
          const float2 y = unpack_bf16(*reinterpret_cast<const uint32_t*>(sb + 4 * T_TILE_BYTES + off));
          const float part = warp_sum(x.x * y.x + x.y * y.y);
          if (lane == 0) {
            wdelta[r] = part;
            if (r < a.Lq) a.delta[stat_base + r] = part;
          }
          if (r == gq) d_lo = part;
          if (r == gq + 8) d_hi = part;
        }
        (void)tO;
        const float lse_lo = gq < a.Lq ? slse[gq] : 0.f, lse_hi = gq + 8 < a.Lq ? slse[gq + 8] : 0.f;
        __syncwarp();
        // ---- query-row view: P, dP = dO V^T, dS;  dQ = dS K
        float dp[3][4];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dp[i][j] = 0.f;
        t_mma_nt(dp, fdo, tV, xV, has_cls, ta, la);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = nt * 8 + 2 * tq + (e & 1);
            const bool ok = nt < 2 ? col < tp.lk_reg : (has_cls && col == 16);
            const float pv = ok ? ex2(fmaf(s[nt][e], scale2, -(e < 2 ? lse_lo : lse_hi))) : 0.f;
            s[nt][e] = pv * (dp[nt][e] - (e < 2 ? d_lo : d_hi));
          }
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        t_mma_p(acc, s, tK, xK, has_cls, ta, la);   // dQ = dS K
        g_stage(stg, acc, a.scale, a.scale, lane);
        __syncwarp();
        t_store_rows(stg, a.dq, q_first, tp.istride, a.lddq, h * HD, a.Lq, lane);
        __syncwarp();
        // ---- key-row view (regular keys, then the CLS key tile): P^T, dP^T = V dO^T, dS^T;  dV = P^T dO, dK = dS^T Q
#pragma unroll 1
        for (int part = 0; part < (has_cls ? 2 : 1); ++part) {
          uint32_t fk[4][4], fv[4][4];
          if (part == 0) {
            t_load_a(fk, tK, ta);
            t_load_a(fv, tV, ta);
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              ldmatrix_x4(fk[kk][0], fk[kk][1], fk[kk][2], fk[kk][3], xK + la.p_row + la.p_x[kk]);
              ldmatrix_x4(fv[kk][0], fv[kk][1], fv[kk][2], fv[kk][3], xV + la.p_row + la.p_x[kk]);
            }
          }
          float stt[3][4], dpt[3][4];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) stt[i][j] = dpt[i][j] = 0.f;
          t_mma_nt(stt, fk, tQ, 0u, false, ta, la);    // S^T = K Q^T   (columns = queries)
          t_mma_nt(dpt, fv, tDO, 0u, false, ta, la);   // dP^T = V dO^T
          float pt[3][4];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int col = nt * 8 + 2 * tq + (e & 1);
              const bool ok = col < a.Lq;
              const float pv = ok ? ex2(fmaf(stt[nt][e], scale2, -slse[ok ? col : 0])) : 0.f;
              pt[nt][e] = pv;
              stt[nt][e] = pv * (dpt[nt][e] - wdelta[col]);
            }
#pragma unroll
          for (int e = 0; e < 4; ++e) pt[2][e] = stt[2][e] = 0.f;
          if (part == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            t_mma_p(acc, pt, tDO, 0u, false, ta, la);   // dV = P^T dO
            g_stage(stg, acc, 1.0f, 1.0f, lane);
            __syncwarp();
            t_store_rows(stg, a.dv, k_first, tp.istride, a.lddkv, h * HD, tp.lk_reg, lane);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            t_mma_p(acc, stt, tQ, 0u, false, ta, la);   // dK = dS^T Q
            g_stage(stg, acc, a.scale, a.scale, lane);
            __syncwarp();
            t_store_rows(stg, a.dk, k_first, tp.istride, a.lddkv, h * HD, tp.lk_reg, lane);
            __syncwarp();
          } else {
            t_mma_p(clsv, pt, tDO, 0u, false, ta, la);
            t_mma_p(clsk, stt, tQ, 0u, false, ta, la);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
      }
    } else {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (++st == STAGES) {
      st = 0;
      ph ^= 1;
    }
  }
  if (BWD) flush_cls();
}

static int g_tiny_mode = -1;   // env EGV_ATTN_TINY: bit 0 forward, bit 1 backward.  Default 0: correct (tests run it with
                               // EGV_ATTN_TINY=3) but measured 60 / 225 us vs 49 / 200 us for the generic kernels at cfg 3 -- the 8 warps
                               // of a CTA move in lock-step behind one barrier, so instruction latency is exposed; see DESIGN.md

template <bool BWD>
static int launch_tiny(const TinyMaps& maps, const AttnP& a, const TinyP& tp, cudaStream_t stream) {
  using Cfg = TCfg<BWD>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "tiny attention tile set exceeds the shared memory of an SM");
  auto kern = attn_tiny_kernel<BWD>;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "tiny attention smem attribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  const long long grid = tp.items < sm_count() ? tp.items : sm_count();
  kern<<<(unsigned)grid, T_THREADS, Cfg::SMEM_BYTES, stream>>>(maps, a, tp);
  int rc = check_launch("attn_tiny_kernel");
  return rc ? rc : 1;
}

void set_tiny_mode(int mode) { g_tiny_mode = mode; }

int launch_tiny_attention(int bwd, const AttnP& a, cudaStream_t stream) {
  if (g_tiny_mode < 0) g_tiny_mode = getenv("EGV_ATTN_TINY") ? atoi(getenv("EGV_ATTN_TINY")) : 0;
  if (!((g_tiny_mode >> bwd) & 1)) return 0;
  const int lk_reg = a.LkT - (a.has_cls ? 1 : 0);
  if (a.key_bias || a.Lq > TL || lk_reg > TL || lk_reg < 1 || a.Lq < 1) return 0;
  if (a.q_gstride != 1 || a.k_gstride != 1 || a.q_istride != a.k_istride || a.q_istride < a.G) return 0;
  if (a.G < TG || (long long)a.B * a.H * a.G < 1024) return 0;
  if (bwd && (a.dkv_accumulate || (a.Lq % 4) || (a.has_cls && !a.dkv_cls))) return 0;
  TinyP tp;
  tp.n_gt = (int)cdiv(a.G, TG);
  tp.items = (long long)a.B * a.H * tp.n_gt;
  tp.istride = a.q_istride;
  tp.lk_reg = lk_reg;
  TinyMaps maps;
  const uint32_t box[4] = {64, TG, TL, 1};
  struct T { const bf16* p; long long ld, bstride; int row0, L; };
  const T ts[5] = {{a.q, a.ldq, a.q_bstride, a.q_row0, a.Lq}, {a.k, a.ldkv, a.kv_bstride, a.k_row0, lk_reg},
                   {a.v, a.ldkv, a.kv_bstride, a.k_row0, lk_reg}, {a.d_o, a.ldo, a.o_bstride, a.q_row0, a.Lq},
                   {a.o, a.ldo, a.o_bstride, a.q_row0, a.Lq}};
  const int nt = bwd ? 5 : 3;
  for (int t = 0; t < 5; ++t) {
    const T& x = ts[t < nt ? t : 0];
    const uint64_t dims[4] = {(uint64_t)a.H * HD, (uint64_t)a.G, (uint64_t)x.L, (uint64_t)a.B};
    const uint64_t strides[3] = {(uint64_t)x.ld * 2, (uint64_t)a.q_istride * x.ld * 2, (uint64_t)x.bstride * x.ld * 2};
    int rc = get_tensor_map_nd(x.p + (long long)x.row0 * x.ld, 4, dims, strides, box, &maps.t[t]);
    if (rc) return getenv("EGV_ATTN_TINY_STRICT") ? rc : 0;   // not describable as a 4-D box: generic kernels
  }
  return bwd ? launch_tiny<true>(maps, a, tp, stream) : launch_tiny<false>(maps, a, tp, stream);
}

}  // namespace egv
