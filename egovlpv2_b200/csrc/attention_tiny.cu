// Fused attention kernels for MANY TINY groups -- the time attention of the divided space-time block
// (video_transformer.py:117-153, 'b (f n) d -> (b n) f d'): per (clip, patch position n, head) T <= 16 queries attend
// T <= 16 keys plus the shared CLS key; 18 816 such groups per call at cfg 3, each touching ~6 KB.  The work is pure
// HBM traffic and instruction latency, so:
//   * every WARP is its own pipeline: it walks a contiguous range of (clip, head, group) items, fetches each operand of
//     the next group as ONE 4-D TMA box {64 head columns, 1 group, 16 frames, 1 clip} (the frames of a group are Nf rows
//     apart) into its private 2-stage ring while it computes the current one -- no block-level barrier anywhere, so the
//     LDSM -> HMMA -> shuffle chains of different warps interleave (the first version moved 8 warps in lock-step
//     behind one stage barrier and lost to the generic kernels); out-of-range frames are zero-filled by the TMA unit;
//   * a group lives entirely in registers (mma.sync m16n8k16: one 16 x 16 (+CLS) score tile);
//   * the BACKWARD is one launch instead of two: dQ from the query-row view, dK / dV from the transposed (key-row)
//     view recomputed from the same shared-memory tiles, the CLS key as an extra 1-row key tile whose gradient is
//     carried in registers across the groups of a (clip, head) and flushed with a handful of atomics.
//     Operands are read once (5 tensors) instead of 9 times by the generic DQ + DKV pair.
#include <stdlib.h>

#include "attention.cuh"

namespace egv {

constexpr int TL = 16;                       // rows (frames) per group, padded
constexpr int T_TILE_BYTES = TL * 128;       // 2 KB: [frame i][64 bf16], 128B swizzle
constexpr int T_X_BYTES = 16 * 128;          // extra key tile: row 0 = CLS key (or value), rows 1..15 zero

struct TinyMaps {
  CUtensorMap t[5];   // q, k, v, d_o, o
};
struct TinyP {
  long long items;    // B * H * G
  int istride;        // rows between consecutive frames of a group
  int lk_reg;         // regular keys per group
};

template <bool BWD>
struct TCfg {
  static constexpr int WARPS = BWD ? 8 : 12;
  static constexpr int NT = BWD ? 5 : 3;
  static constexpr int STAGE_BYTES = NT * T_TILE_BYTES;
  // per warp: 2 stages, the CLS key / value tiles, one output staging tile, lse (2 stages) + delta
  static constexpr int X_OFF = 2 * STAGE_BYTES;
  static constexpr int OUT_OFF = X_OFF + 2 * T_X_BYTES;
  static constexpr int STAT_OFF = OUT_OFF + 2048;
  // lse[2][16] + delta[16] floats (backward only); regions stay multiples of 1024 B (128B-swizzle period)
  static constexpr int WARP_BYTES = ((STAT_OFF + (BWD ? 256 : 0) + 1023) / 1024) * 1024;
  static constexpr int BAR_OFF = WARPS * WARP_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + WARPS * 16 + 1024;
};

EGV_DEVINL void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// per-lane byte offsets into a [16 rows][128 B] swizzled tile:
//   a[kk]: A fragments (rows = frames) and transposed B loads (rows = k index);  n[kk]: B loads with rows = n index
struct TinyAddr {
  uint32_t a[4], n[4];
};
EGV_DEVINL TinyAddr tiny_addr(const LaneAddr& la) {
  TinyAddr t;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    t.a[kk] = la.p_row + la.p_x[kk];
    t.n[kk] = la.nt_row + la.nt_x[kk];
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

// 12 warps x 153 registers (forward) / 8 warps x 168 (backward): 3 / 2 warps per SM sub-partition, within 16384 / 32 each
template <bool BWD>
__global__ void __maxnreg__(168)
attn_tiny_kernel(const __grid_constant__ TinyMaps maps, const AttnP a, const TinyP tp) {
  using Cfg = TCfg<BWD>;
  constexpr int NT = Cfg::NT, WARPS = Cfg::WARPS;
  extern __shared__ __align__(1024) uint8_t tsm_raw[];
  const uint32_t pad = (1024u - (smem_u32(tsm_raw) & 1023u)) & 1023u;
  uint8_t* smem = tsm_raw + pad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wsm = smem + warp * Cfg::WARP_BYTES;                 // this warp's private region (multiple of 1024 B)
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF) + warp * 2;
  const bool has_cls = a.has_cls != 0;
  {
    uint4* z = reinterpret_cast<uint4*>(wsm);
    for (int i = lane; i < Cfg::WARP_BYTES / 16; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    if (lane == 0) {
      mbar_init(&full[0], 1);
      mbar_init(&full[1], 1);
      fence_barrier_init();
      if (warp == 0)
        for (int t = 0; t < NT; ++t) tma_prefetch_desc(&maps.t[t]);
    }
    __syncwarp();
  }
  pdl_wait();    // programmatic dependent launch (common.cuh): global accesses from here on; dependents start at exit

  // contiguous range of (batch, head, group) items per warp: consecutive items share (batch, head)
  const long long nwarps = (long long)gridDim.x * WARPS;
  const long long per = (tp.items + nwarps - 1) / nwarps;
  const long long it0 = ((long long)blockIdx.x * WARPS + warp) * per;
  const long long it1 = it0 + per < tp.items ? it0 + per : tp.items;
  if (it0 >= it1) return;

  const int gq = lane >> 2, tq = lane & 3;
  const LaneAddr la = lane_addr(lane);
  const TinyAddr ta = tiny_addr(la);
  const float scale2 = a.scale * LOG2E;
  uint8_t* stg = wsm + Cfg::OUT_OFF;
  float* wdelta = reinterpret_cast<float*>(wsm + Cfg::STAT_OFF) + 2 * TL;
  const uint32_t xK = smem_u32(wsm + Cfg::X_OFF), xV = xK + T_X_BYTES;

  auto issue = [&](long long it, int st) {   // lane 0: TMA boxes (+ lse) of item `it` into stage `st`
    const int g = (int)(it % a.G), h = (int)((it / a.G) % a.H), b = (int)(it / ((long long)a.G * a.H));
    uint8_t* sb = wsm + st * Cfg::STAGE_BYTES;
    uint32_t bytes = (uint32_t)NT * T_TILE_BYTES;
    if (BWD) bytes += (uint32_t)a.Lq * 4u;
    fence_proxy_async();   // this warp's generic-proxy reads of the stage (two items ago) precede the async writes
    mbar_arrive_expect_tx(&full[st], bytes);
#pragma unroll
    for (int t = 0; t < NT; ++t) tma_load_4d(sb + t * T_TILE_BYTES, &maps.t[t], &full[st], h * HD, g, 0, b);
    if (BWD) {
      const long long stat_base = (((long long)b * a.H + h) * a.G + g) * a.Lq;
      bulk_copy_g2s(wsm + Cfg::STAT_OFF + st * TL * 4, a.lse + stat_base, (uint32_t)a.Lq * 4u, &full[st]);
    }
  };

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

  if (lane == 0) issue(it0, 0);
  uint32_t phbits = 0u;   // bit s = parity the next wait on full[s] expects
  for (long long it = it0; it < it1; ++it) {
    const int st = (int)((it - it0) & 1);
    const int g = (int)(it % a.G), h = (int)((it / a.G) % a.H), b = (int)(it / ((long long)a.G * a.H));
    const long long bh = (long long)b * a.H + h;
    __syncwarp();                                     // every lane is done with the other stage (previous item)
    if (lane == 0 && it + 1 < it1) issue(it + 1, st ^ 1);
    if (bh != cls_bh) {
      if (BWD) flush_cls();
      cls_bh = bh;
      if (has_cls) {   // CLS key / value of this (batch, head) -> row 0 of the extra tiles (chunk c at c ^ 0)
        if (lane < 16) {
          const long long off = ((long long)b * a.kv_bstride + a.cls_row) * a.ldkv + h * HD + (lane & 7) * 8;
          const uint4 val = *reinterpret_cast<const uint4*>((lane < 8 ? a.k : a.v) + off);
          *reinterpret_cast<uint4*>(wsm + Cfg::X_OFF + (lane < 8 ? 0 : T_X_BYTES) + ((lane & 7) << 4)) = val;
        }
        __syncwarp();
      }
    }
    mbar_wait(&full[st], (phbits >> st) & 1u);
    phbits ^= 1u << st;
    const uint8_t* sb = wsm + st * Cfg::STAGE_BYTES;
    const uint32_t tQ = smem_u32(sb), tK = tQ + T_TILE_BYTES, tV = tQ + 2 * T_TILE_BYTES;
    const uint32_t tDO = tQ + 3 * T_TILE_BYTES;
    const float* slse = reinterpret_cast<const float*>(wsm + Cfg::STAT_OFF) + st * TL;
    {
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
          const uint32_t off = r * 128 + ((((uint32_t)(lane >> 2)) ^ (uint32_t)(r & 7)) << 4) + 4 * (lane & 3);
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
      }
    }
  }
  if (BWD) flush_cls();
}

static int g_tiny_mode = -1;   // env EGV_ATTN_TINY: bit 0 forward, bit 1 backward (default 3).  Measured at cfg 3 (18 816 groups):
                               // 39 / 115 us vs 49 / 200 us for the generic warp-per-group kernels (forward / backward).

template <bool BWD>
static int launch_tiny(const TinyMaps& maps, const AttnP& a, const TinyP& tp, cudaStream_t stream) {
  using Cfg = TCfg<BWD>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "tiny attention tile set exceeds the shared memory of an SM");
  static_assert(Cfg::WARP_BYTES % 1024 == 0, "per-warp regions must keep the 128B-swizzle period");
  auto kern = attn_tiny_kernel<BWD>;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "tiny attention smem attribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  long long grid = cdiv(tp.items, (long long)Cfg::WARPS * 4);   // >= 4 items per warp
  if (grid > sm_count()) grid = sm_count();
  launch_k(kern, dim3((unsigned)grid), dim3(Cfg::WARPS * 32), Cfg::SMEM_BYTES, stream, maps, a, tp);
  int rc = check_launch("attn_tiny_kernel");
  return rc ? rc : 1;
}

void set_tiny_mode(int mode) { g_tiny_mode = mode; }

int launch_tiny_attention(int bwd, const AttnP& a, cudaStream_t stream) {
  if (g_tiny_mode < 0) g_tiny_mode = getenv("EGV_ATTN_TINY") ? atoi(getenv("EGV_ATTN_TINY")) : 3;
  if (!((g_tiny_mode >> bwd) & 1)) return 0;
  const int lk_reg = a.LkT - (a.has_cls ? 1 : 0);
  if (a.key_bias || a.Lq > TL || lk_reg > TL || lk_reg < 1 || a.Lq < 1) return 0;
  if (a.q_gstride != 1 || a.k_gstride != 1 || a.q_istride != a.k_istride || a.q_istride < a.G) return 0;
  if ((long long)a.B * a.H * a.G < 1024) return 0;
  if (bwd && (a.dkv_accumulate || (a.Lq % 4) || (a.has_cls && !a.dkv_cls))) return 0;
  TinyP tp;
  tp.items = (long long)a.B * a.H * a.G;
  tp.istride = a.q_istride;
  tp.lk_reg = lk_reg;
  TinyMaps maps;
  const uint32_t box[4] = {64, 1, TL, 1};
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
