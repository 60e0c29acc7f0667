// egv_attention_fwd / egv_attention_bwd: strided multi-head attention, head_dim 64, bf16 in / fp32 softmax.
//
// One flash-style kernel family covers every attention on the EgoVLPv2 path (see include/egovlp_b200.h):
// divided time / space attention with the shared CLS key, the CLS query, the gated video->text and
// text->video cross-attention cores and RoBERTa self-attention.  A work item is
// (batch b, head h, group g, 16*NW-row tile); the "row side" lives in registers as mma.sync A fragments,
// the "stream side" is staged through shared memory in KC-row chunks:
//     FWD : rows = queries, stream = keys      O = softmax(QK^T) V, lse
//     DQ  : rows = queries, stream = keys      dQ (and delta = rowsum(dO*O))
//     DKV : rows = keys,    stream = queries   dK, dV
// Scores never touch HBM.  Small groups (time attention: T queries x T+1 keys) run one group per warp.
#include <stdlib.h>

#include <type_traits>

#include "attention.cuh"

namespace egv {

EGV_DEVINL long long q_row(const AttnP& a, int b, int g, int i) {
  return (long long)b * a.q_bstride + a.q_row0 + (long long)g * a.q_gstride + (long long)i * a.q_istride;
}
EGV_DEVINL long long o_row(const AttnP& a, int b, int g, int i) {
  return (long long)b * a.o_bstride + a.q_row0 + (long long)g * a.q_gstride + (long long)i * a.q_istride;
}
EGV_DEVINL long long k_row(const AttnP& a, int b, int g, int j) {
  long long r;
  if (a.has_cls) r = (j == 0) ? a.cls_row : a.k_row0 + (long long)g * a.k_gstride + (long long)(j - 1) * a.k_istride;
  else r = a.k_row0 + (long long)g * a.k_gstride + (long long)j * a.k_istride;
  return (long long)b * a.kv_bstride + r;
}

// which logical tensor a smem tile is filled from
enum { T_Q = 0, T_K = 1, T_V = 2, T_DO = 3, T_O = 4 };

template <int WHAT>
EGV_DEVINL const bf16* row_ptr(const AttnP& a, int b, int h, int g, int idx) {
  if (WHAT == T_Q) return a.q + q_row(a, b, g, idx) * a.ldq + h * HD;
  if (WHAT == T_DO) return a.d_o + o_row(a, b, g, idx) * a.ldo + h * HD;
  if (WHAT == T_O) return a.o + o_row(a, b, g, idx) * a.ldo + h * HD;
  if (WHAT == T_K) return a.k + k_row(a, b, g, idx) * a.ldkv + h * HD;
  return a.v + k_row(a, b, g, idx) * a.ldkv + h * HD;
}

// cp.async a [ROWS x 64] bf16 tile (rows idx0 .. idx0+ROWS-1 of the logical tensor, zero-filled past `count`)
template <int WHAT, int ROWS, int NT>
EGV_DEVINL void load_tile(bf16* s, const AttnP& a, int b, int h, int g, int idx0, int count, int tid) {
  for (int c = tid; c < ROWS * 8; c += NT) {
    const int r = c >> 3, ch = c & 7;
    const int idx = idx0 + r;
    const bool ok = idx < count;
    const bf16* src = ok ? row_ptr<WHAT>(a, b, h, g, idx) + ch * 8 : a.q;
    cp_async_16(s + r * LDS + ch * 8, src, ok);
  }
}

// Row addressing of one logical tensor inside a work item, linear in the row index:
// element offset(idx) = off + idx * step, except the shared CLS key at idx == 0 (has_cls) which lives at `cls`.
struct RowAddr {
  const bf16* base;
  long long off, step, cls;
  bool has_cls;
  EGV_DEVINL const bf16* at(int idx) const { return base + ((has_cls && idx == 0) ? cls : off + (long long)idx * step); }
};

// cp.async rows idx0 .. idx0+rows_cap-1 (zero-filled past `count`) into a [rows_cap x 64] tile; thread t moves the
// 16-byte chunk t & 7 of rows (t >> 3) + k * NT/8: one pointer increment per copy instead of re-deriving the address.
template <int NT>
EGV_DEVINL void load_rows(bf16* dst, const RowAddr& ra, int idx0, int count, int rows_cap, int tid) {
  constexpr int RSTEP = NT / 8;
  int r = tid >> 3;
  const int ch = tid & 7;
  const bf16* p = ra.base + ra.off + (long long)(idx0 + r) * ra.step + ch * 8;
  bf16* d = dst + r * LDS + ch * 8;
  const long long pstep = (long long)RSTEP * ra.step;
  for (; r < rows_cap; r += RSTEP, p += pstep, d += RSTEP * LDS) {
    const int idx = idx0 + r;
    const bool ok = idx < count;
    const bf16* sp = (ra.has_cls && idx == 0) ? ra.base + ra.cls + ch * 8 : p;
    cp_async_16(d, ok ? sp : ra.base, ok);
  }
}

template <int NW>
EGV_DEVINL void unit_sync() {
  if (NW == 1) __syncwarp();
  else __syncthreads();
}

// acc[nt] (16 x 8 fp32 tiles, nt < KC/8) = A(16 x 64, fragments af) * S^T, S = smem tile [KC rows][64] (row = n index)
template <int KC>
EGV_DEVINL void mma_a_stile_nt(float (&acc)[KC / 8][4], const uint32_t (&af)[4][4], const bf16* s, int lane) {
#pragma unroll
  for (int n2 = 0; n2 < KC / 16; ++n2) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t b0, b1, b2, b3;
      const bf16* p = s + (n2 * 16 + (lane & 7) + ((lane >> 4) << 3)) * LDS + kk * 16 + ((lane >> 3) & 1) * 8;
      ldmatrix_x4(b0, b1, b2, b3, smem_u32(p));
      mma_16816(acc[2 * n2], af[kk], b0, b1);
      mma_16816(acc[2 * n2 + 1], af[kk], b2, b3);
    }
  }
}

// out[dt] (16 x 8 fp32 tiles over the 64 columns, dt < 8) += P(16 x KC, taken from C-layout `pacc`) * S, S = smem [KC][64]
template <int KC>
EGV_DEVINL void mma_p_stile(float (&out)[8][4], const float (&pacc)[KC / 8][4], const bf16* s, int lane) {
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
      const bf16* p = s + (k2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + d2 * 16 + (lane >> 4) * 8;
      ldmatrix_x4_trans(b0, b1, b2, b3, smem_u32(p));
      mma_16816(out[2 * d2], af, b0, b1);
      mma_16816(out[2 * d2 + 1], af, b2, b3);
    }
  }
}

// load the 16 x 64 A fragments of this warp's rows from a smem row tile
EGV_DEVINL void load_a_frags(uint32_t (&af)[4][4], const bf16* s, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const bf16* p = s + ((lane & 7) + ((lane >> 3) & 1) * 8) * LDS + kk * 16 + (lane >> 4) * 8;
    ldmatrix_x4(af[kk][0], af[kk][1], af[kk][2], af[kk][3], smem_u32(p));
  }
}

// write a warp's 16 x 64 fp32 C-layout tile into smem as bf16 (rows r0.., stride LDS)
EGV_DEVINL void c_to_smem(bf16* s, const float (&c)[8][4], float mul0, float mul1, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(s + g * LDS + nt * 8 + 2 * t) = pack_bf16(c[nt][0] * mul0, c[nt][1] * mul0);
    *reinterpret_cast<uint32_t*>(s + (g + 8) * LDS + nt * 8 + 2 * t) = pack_bf16(c[nt][2] * mul1, c[nt][3] * mul1);
  }
}

// NCH = chunks of the stream side resident in shared memory.  NCH == 1: chunks are (re)loaded inside the loop;
// NCH > 1: the whole stream side of the group (<= NCH*KC rows) is loaded once, the loop has no loads and no barriers.
template <int MODE, int NW, int KC, int NCH = 1>
struct AttnSmem {
  static constexpr int ROWS = 16 * NW;
  static constexpr int ROW_TILES = (MODE == MODE_FWD) ? 1 : 2;
  static constexpr int SROWS = KC * NCH;
  static constexpr int BYTES = (ROW_TILES * ROWS + 2 * SROWS) * LDS * 2 + 2 * SROWS * 4;
};

// SPLIT: the kernel may be launched with a.n_split > 1 (stream-side split, fp32 partials); compiled out elsewhere so the
// warp-per-group kernels keep their register budget.
template <int MODE, int NW, int KC, int NCH = 1, bool SPLIT = false>
EGV_DEVINL void attn_unit(const AttnP& a, long long item, uint8_t* smem_unit, int tid) {
  constexpr int NT = NW * 32;
  constexpr int ROWS = 16 * NW;
  using SM = AttnSmem<MODE, NW, KC, NCH>;
  constexpr int SROWS = SM::SROWS;
  bf16* rowA = reinterpret_cast<bf16*>(smem_unit);
  bf16* rowB = rowA + (SM::ROW_TILES - 1) * ROWS * LDS;  // == rowA for FWD (unused)
  bf16* strA = rowA + SM::ROW_TILES * ROWS * LDS;
  bf16* strB = strA + SROWS * LDS;
  float* sf0 = reinterpret_cast<float*>(strB + SROWS * LDS);  // FWD/DQ: key bias (log2 units); DKV: lse of stream queries
  float* sf1 = sf0 + SROWS;                                   // DKV: delta of stream queries

  const int lane = tid & 31;
  const int wq = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;

  int split = 0;
  long long item_ns = item;   // item index without the split
  if (SPLIT && a.n_split > 1) {   // (64-bit division: keep it off the common path)
    split = (int)(item % a.n_split);
    item_ns = item / a.n_split;
  }
  const int tile = (int)(item_ns % a.row_tiles);
  long long rest = item_ns / a.row_tiles;
  const int h = (int)(rest % a.H);
  rest /= a.H;
  const int g = (int)(rest % a.G);
  const int b = (int)(rest / a.G);

  const int n_rows = (MODE == MODE_DKV) ? a.LkT : a.Lq;
  const int n_str = (MODE == MODE_DKV) ? a.Lq : a.LkT;
  const int row0 = tile * ROWS;
  const float scale2 = a.scale * LOG2E;
  const long long stat_base = (((long long)b * a.H + h) * a.G + g) * a.Lq;

  // linear row addressing of this item's tensors (element offsets)
  const long long q_row_first = (long long)b * a.q_bstride + a.q_row0 + (long long)g * a.q_gstride;
  const long long o_row_first = (long long)b * a.o_bstride + a.q_row0 + (long long)g * a.q_gstride;
  const long long k_row_first = (long long)b * a.kv_bstride + a.k_row0 + (long long)g * a.k_gstride;
  const long long k_cls_row = (long long)b * a.kv_bstride + a.cls_row;
  const int hc = a.has_cls ? 1 : 0;
  const RowAddr aQ{a.q, q_row_first * a.ldq + h * HD, (long long)a.q_istride * a.ldq, 0, false};
  const RowAddr aO{a.o, o_row_first * a.ldo + h * HD, (long long)a.q_istride * a.ldo, 0, false};
  const RowAddr aDO{a.d_o, o_row_first * a.ldo + h * HD, (long long)a.q_istride * a.ldo, 0, false};
  const RowAddr aK{a.k, (k_row_first - (long long)hc * a.k_istride) * a.ldkv + h * HD, (long long)a.k_istride * a.ldkv,
                   k_cls_row * a.ldkv + h * HD, a.has_cls != 0};
  const RowAddr aV{a.v, aK.off, aK.step, aK.cls, aK.has_cls};

  // ---------------------------------------------------------------- row-side operands -> registers
  if (MODE == MODE_FWD) {
    load_rows<NT>(rowA, aQ, row0, n_rows, ROWS, tid);
  } else if (MODE == MODE_DQ) {
    load_rows<NT>(rowA, aQ, row0, n_rows, ROWS, tid);
    load_rows<NT>(rowB, aDO, row0, n_rows, ROWS, tid);
    static_assert(2 * KC * NCH >= 16 * NW, "the O tile is staged in the (contiguous) stream buffers");
    load_rows<NT>(strA, aO, row0, n_rows, ROWS, tid);   // O rows, only for delta = rowsum(dO * O)
  } else {
    load_rows<NT>(rowA, aK, row0, n_rows, ROWS, tid);
    load_rows<NT>(rowB, aV, row0, n_rows, ROWS, tid);
  }
  cp_async_commit();
  cp_async_wait<0>();
  unit_sync<NW>();

  uint32_t fa[4][4];  // FWD/DQ: Q   DKV: K
  uint32_t fb[4][4];  // DQ: dO      DKV: V
  load_a_frags(fa, rowA + wq * 16 * LDS, lane);
  if (MODE != MODE_FWD) load_a_frags(fb, rowB + wq * 16 * LDS, lane);

  // per-row statistics for rows gq and gq+8 of this warp's tile
  const int r_lo = row0 + wq * 16 + gq, r_hi = r_lo + 8;
  float st0_lo = 0.f, st0_hi = 0.f, st1_lo = 0.f, st1_hi = 0.f;
  if (MODE == MODE_DQ) {
    // delta_i = sum_d dO[i,d] * O[i,d] from the staged tiles; one row per iteration, 2 columns per lane
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int idx = row0 + wq * 16 + r;
      const float2 x = unpack_bf16(*reinterpret_cast<const uint32_t*>(rowB + (wq * 16 + r) * LDS + 2 * lane));
      const float2 y = unpack_bf16(*reinterpret_cast<const uint32_t*>(strA + (wq * 16 + r) * LDS + 2 * lane));
      const float part = warp_sum(x.x * y.x + x.y * y.y);
      if (idx < n_rows && lane == 0) a.delta[stat_base + idx] = part;
      if (r == gq) st1_lo = part;
      if (r == gq + 8) st1_hi = part;
    }
    st0_lo = r_lo < n_rows ? a.lse[stat_base + r_lo] : 0.f;
    st0_hi = r_hi < n_rows ? a.lse[stat_base + r_hi] : 0.f;
  } else if (MODE == MODE_DKV) {
    // key bias of this thread's two key rows (log2 units)
    if (a.key_bias) {
      st0_lo = r_lo < n_rows ? fmaxf(a.key_bias[(long long)b * a.LkT + r_lo], -1e30f) * LOG2E : 0.f;
      st0_hi = r_hi < n_rows ? fmaxf(a.key_bias[(long long)b * a.LkT + r_hi], -1e30f) * LOG2E : 0.f;
    }
  }

  float acc0[8][4];  // FWD: O    DQ: dQ   DKV: dK
  float acc1[8][4];  // DKV: dV
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc0[i][j] = acc1[i][j] = 0.f;
  float m_lo = -1e30f, m_hi = -1e30f, l_lo = 0.f, l_hi = 0.f;  // FWD online softmax state

  // ---------------------------------------------------------------- stream loop
  auto load_stream = [&](int c0, int rows_cap, bf16* dA, bf16* dB, float* d0, float* d1) {
    // rows c0 .. c0+rows_cap-1 of the stream side -> dA / dB (+ per-row scalars); rows past n_str are zero-filled
    if (MODE == MODE_DKV) {
      load_rows<NT>(dA, aQ, c0, n_str, rows_cap, tid);
      load_rows<NT>(dB, aDO, c0, n_str, rows_cap, tid);
      for (int j = tid; j < rows_cap; j += NT) {
        const int idx = c0 + j;
        d0[j] = idx < n_str ? a.lse[stat_base + idx] : 0.f;
        d1[j] = idx < n_str ? a.delta[stat_base + idx] : 0.f;
      }
    } else {
      load_rows<NT>(dA, aK, c0, n_str, rows_cap, tid);
      load_rows<NT>(dB, aV, c0, n_str, rows_cap, tid);
      if (a.key_bias) {
        for (int j = tid; j < rows_cap; j += NT) {
          const int idx = c0 + j;
          d0[j] = idx < n_str ? fmaxf(a.key_bias[(long long)b * a.LkT + idx], -1e30f) * LOG2E : 0.f;
        }
      }
    }
  };
  if (NCH > 1) {
    unit_sync<NW>();  // row-side phase (delta from the staged O tile) finished with the stream buffers
    const int rows_needed = ((n_str + KC - 1) / KC) * KC;   // <= SROWS (checked by the host)
    load_stream(0, rows_needed, strA, strB, sf0, sf1);
    cp_async_commit();
    cp_async_wait<0>();
    unit_sync<NW>();
  }
  // stream range of this split (whole chunks)
  int c_begin = 0, c_end = n_str;
  if (SPLIT && a.n_split > 1) {
    const int chunks = (n_str + KC - 1) / KC;
    const int per = (chunks + a.n_split - 1) / a.n_split;
    c_begin = split * per * KC;
    c_end = min(n_str, c_begin + per * KC);
  }
  for (int c0 = c_begin; c0 < c_end; c0 += KC) {
    if (NCH == 1) {
      unit_sync<NW>();  // previous chunk fully consumed
      load_stream(c0, KC, strA, strB, sf0, sf1);
      cp_async_commit();
      cp_async_wait<0>();
      unit_sync<NW>();
    }
    const int coff = NCH > 1 ? c0 : 0;
    const bf16* cA = strA + coff * LDS;
    const bf16* cB = strB + coff * LDS;
    const float* cf0 = sf0 + coff;
    const float* cf1 = sf1 + coff;

    float s[KC / 8][4];
#pragma unroll
    for (int i = 0; i < KC / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    mma_a_stile_nt<KC>(s, fa, cA, lane);  // FWD/DQ: Q K^T    DKV: K Q^T

    // The scalar part below is what bounds these kernels (XU / ALU pipes), so it is specialised at compile time on
    // FULL (no stream-side tail in this chunk: no bounds predicates) and BIAS (additive key bias present).
    const int lim = n_str - c0;   // valid stream rows in this chunk (>= KC when full)
    auto chunk = [&](auto full_t, auto bias_t) {
      constexpr bool FULL = decltype(full_t)::value;
      constexpr bool BIAS = decltype(bias_t)::value;
      if (MODE == MODE_FWD) {
        // Without a key bias the softmax scale is folded into the exponent: max on the raw scores (scale > 0), then
        // p = 2^(s*scale2 - m) as one FFMA + EX2 per element.
        float cm_lo = -1e30f, cm_hi = -1e30f;
#pragma unroll
        for (int nt = 0; nt < KC / 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = nt * 8 + 2 * tq + (e & 1);
            float v = BIAS ? fmaf(s[nt][e], scale2, cf0[col]) : s[nt][e];
            if (!FULL) v = col < lim ? v : -1e30f;
            s[nt][e] = v;
            if (e < 2) cm_lo = fmaxf(cm_lo, v);
            else cm_hi = fmaxf(cm_hi, v);
          }
        }
        cm_lo = fmaxf(cm_lo, __shfl_xor_sync(0xffffffffu, cm_lo, 1));
        cm_lo = fmaxf(cm_lo, __shfl_xor_sync(0xffffffffu, cm_lo, 2));
        cm_hi = fmaxf(cm_hi, __shfl_xor_sync(0xffffffffu, cm_hi, 1));
        cm_hi = fmaxf(cm_hi, __shfl_xor_sync(0xffffffffu, cm_hi, 2));
        if (!BIAS) {   // to the scaled (log2) domain; -1e30 * scale2 stays a huge negative number
          cm_lo *= scale2;
          cm_hi *= scale2;
        }
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
            const float mm = e < 2 ? m_lo : m_hi;
            float pv = BIAS ? ex2(s[nt][e] - mm) : ex2(fmaf(s[nt][e], scale2, -mm));
            if (!FULL) {
              const int col = nt * 8 + 2 * tq + (e & 1);
              pv = col < lim ? pv : 0.f;
            }
            s[nt][e] = pv;
            if (e < 2) l_lo += pv;
            else l_hi += pv;
          }
        }
        mma_p_stile<KC>(acc0, s, cB, lane);  // O += P V
      } else {
        // P (or P^T) from the saved log-sum-exp
#pragma unroll
        for (int nt = 0; nt < KC / 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = nt * 8 + 2 * tq + (e & 1);
            float pv;
            if (MODE == MODE_DQ) {
              const float sc = BIAS ? fmaf(s[nt][e], scale2, cf0[col]) : s[nt][e] * scale2;
              pv = ex2(sc - (e < 2 ? st0_lo : st0_hi));
            } else {
              // DKV: the key bias is a per-row constant (0 when absent), the lse is per stream column
              pv = ex2(fmaf(s[nt][e], scale2, (e < 2 ? st0_lo : st0_hi)) - cf0[col]);
            }
            if (!FULL) pv = col < lim ? pv : 0.f;
            s[nt][e] = pv;
          }
        }
        if (MODE == MODE_DKV) mma_p_stile<KC>(acc1, s, cB, lane);  // dV += P^T dO
        float dp[KC / 8][4];
#pragma unroll
        for (int i = 0; i < KC / 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dp[i][j] = 0.f;
        mma_a_stile_nt<KC>(dp, fb, cB, lane);  // DQ: dO V^T    DKV: V dO^T
#pragma unroll
        for (int nt = 0; nt < KC / 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = nt * 8 + 2 * tq + (e & 1);
            const float dl = (MODE == MODE_DQ) ? (e < 2 ? st1_lo : st1_hi) : cf1[col];
            s[nt][e] = s[nt][e] * (dp[nt][e] - dl);  // dS (or dS^T)
          }
        }
        mma_p_stile<KC>(acc0, s, cA, lane);  // DQ: dQ += dS K    DKV: dK += dS^T Q
      }
    };
    const bool full = lim >= KC;
    const bool has_bias = (MODE != MODE_DKV) && a.key_bias != nullptr;
    if (full) {
      if (has_bias) chunk(std::true_type{}, std::true_type{});
      else chunk(std::true_type{}, std::false_type{});
    } else {
      if (has_bias) chunk(std::false_type{}, std::true_type{});
      else chunk(std::false_type{}, std::false_type{});
    }
  }

  // ---------------------------------------------------------------- split: fp32 partials, merged by attn_split_finalize_kernel
  if (SPLIT && a.n_split > 1) {
    // ws[((item_ns * n_split + split) * ROWS + row) * W + ...],  W = 66 (FWD: m, l, o[64]), 64 (DQ) or 128 (DKV: dk, dv)
    constexpr int W = MODE == MODE_FWD ? 66 : (MODE == MODE_DQ ? 64 : 128);
    float* base = a.ws + ((item_ns * a.n_split + split) * ROWS + wq * 16) * W;
    float* p_lo = base + (long long)gq * W, *p_hi = base + (long long)(gq + 8) * W;
    if (MODE == MODE_FWD) {
      l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
      l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
      l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
      l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
      if (tq == 0) {
        p_lo[0] = m_lo; p_lo[1] = l_lo;
        p_hi[0] = m_hi; p_hi[1] = l_hi;
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<float2*>(p_lo + 2 + nt * 8 + 2 * tq) = make_float2(acc0[nt][0], acc0[nt][1]);
        *reinterpret_cast<float2*>(p_hi + 2 + nt * 8 + 2 * tq) = make_float2(acc0[nt][2], acc0[nt][3]);
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<float2*>(p_lo + nt * 8 + 2 * tq) = make_float2(acc0[nt][0] * a.scale, acc0[nt][1] * a.scale);
        *reinterpret_cast<float2*>(p_hi + nt * 8 + 2 * tq) = make_float2(acc0[nt][2] * a.scale, acc0[nt][3] * a.scale);
        if (MODE == MODE_DKV) {
          *reinterpret_cast<float2*>(p_lo + 64 + nt * 8 + 2 * tq) = make_float2(acc1[nt][0], acc1[nt][1]);
          *reinterpret_cast<float2*>(p_hi + 64 + nt * 8 + 2 * tq) = make_float2(acc1[nt][2], acc1[nt][3]);
        }
      }
    }
    return;
  }
  // ---------------------------------------------------------------- write-out (through smem for 16-byte stores)
  unit_sync<NW>();
  bf16* outA = rowA + wq * 16 * LDS;
  if (MODE == MODE_FWD) {
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    c_to_smem(outA, acc0, 1.0f / l_lo, 1.0f / l_hi, lane);
    if (tq == 0) {
      if (r_lo < n_rows) a.lse[stat_base + r_lo] = m_lo + log2f(l_lo);
      if (r_hi < n_rows) a.lse[stat_base + r_hi] = m_hi + log2f(l_hi);
    }
  } else if (MODE == MODE_DQ) {
    c_to_smem(outA, acc0, a.scale, a.scale, lane);
  } else {
    c_to_smem(outA, acc0, a.scale, a.scale, lane);
    c_to_smem(rowB + wq * 16 * LDS, acc1, 1.0f, 1.0f, lane);
    if (a.has_cls && tile == 0 && wq == 0 && gq == 0 && a.dkv_cls) {
      float* dst = a.dkv_cls + ((long long)b * a.H + h) * 2 * HD;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        atomicAdd(dst + nt * 8 + 2 * tq, acc0[nt][0] * a.scale);
        atomicAdd(dst + nt * 8 + 2 * tq + 1, acc0[nt][1] * a.scale);
        atomicAdd(dst + HD + nt * 8 + 2 * tq, acc1[nt][0]);
        atomicAdd(dst + HD + nt * 8 + 2 * tq + 1, acc1[nt][1]);
      }
    }
  }
  unit_sync<NW>();
  {
    constexpr int RSTEP = NT / 8;
    const int ch = tid & 7;
    const RowAddr gQ{a.dq, q_row_first * a.lddq + h * HD, (long long)a.q_istride * a.lddq, 0, false};
    const RowAddr gK{a.dk, (k_row_first - (long long)hc * a.k_istride) * a.lddkv + h * HD, (long long)a.k_istride * a.lddkv, 0, false};
    for (int r = tid >> 3; r < ROWS; r += RSTEP) {
      const int idx = row0 + r;
      if (idx >= n_rows) break;
      const uint4 va = *reinterpret_cast<const uint4*>(rowA + r * LDS + ch * 8);
      if (MODE == MODE_FWD) {
        *reinterpret_cast<uint4*>(a.o + aO.off + (long long)idx * aO.step + ch * 8) = va;
      } else if (MODE == MODE_DQ) {
        *reinterpret_cast<uint4*>(a.dq + gQ.off + (long long)idx * gQ.step + ch * 8) = va;
      } else {
        if (a.has_cls && idx == 0) continue;  // shared CLS key: accumulated in dkv_cls above
        const long long off = gK.off + (long long)idx * gK.step + ch * 8;
        uint4 vk = va;
        uint4 vv = *reinterpret_cast<const uint4*>(rowB + r * LDS + ch * 8);
        if (a.dkv_accumulate) {
          uint4 ok = *reinterpret_cast<const uint4*>(a.dk + off);
          uint4 ov = *reinterpret_cast<const uint4*>(a.dv + off);
          uint32_t* pk = reinterpret_cast<uint32_t*>(&vk);
          uint32_t* pv = reinterpret_cast<uint32_t*>(&vv);
          const uint32_t* qk = reinterpret_cast<const uint32_t*>(&ok);
          const uint32_t* qv = reinterpret_cast<const uint32_t*>(&ov);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 x = unpack_bf16(pk[e]), y = unpack_bf16(qk[e]);
            pk[e] = pack_bf16(x.x + y.x, x.y + y.y);
            x = unpack_bf16(pv[e]);
            y = unpack_bf16(qv[e]);
            pv[e] = pack_bf16(x.x + y.x, x.y + y.y);
          }
        }
        *reinterpret_cast<uint4*>(a.dk + off) = vk;
        *reinterpret_cast<uint4*>(a.dv + off) = vv;
      }
    }
  }
}

// CTA-per-item variant (large groups): NW warps cooperate on one item.
template <int MODE, int NW, int KC, int NCH = 1, bool SPLIT = false>
__global__ void __launch_bounds__(NW * 32) attn_cta_kernel(const AttnP a) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t smem_attn[];
  for (long long item = blockIdx.x; item < a.items; item += gridDim.x) {
    attn_unit<MODE, NW, KC, NCH, SPLIT>(a, item, smem_attn, threadIdx.x);
    __syncthreads();
  }
}

// Warp-per-item variant (small groups): 4 independent warps per CTA, no block-level sync.
template <int MODE, int KC>
__global__ void __launch_bounds__(128) attn_warp_kernel(const AttnP a) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int warp = threadIdx.x >> 5;
  uint8_t* mine = smem_attn + warp * AttnSmem<MODE, 1, KC>::BYTES;
  for (long long item = (long long)blockIdx.x * 4 + warp; item < a.items; item += (long long)gridDim.x * 4) {
    attn_unit<MODE, 1, KC>(a, item, mine, threadIdx.x & 31);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------
// Single-query attention (the CLS query of VarAttention: one query row per (b, h), every token a key).
// One CTA per (b, h, g); every warp streams a strided subset of the keys with coalesced 128-byte row reads
// (2 head-dim elements per lane), keeps an online-softmax partial (m, l, o[2 per lane]) and the CTA merges the
// partials through shared memory.  Pure HBM/L2 streaming: 2 * Lk * 128 bytes per (b, h).
constexpr int SQ_WARPS = 16;

__global__ void __launch_bounds__(SQ_WARPS * 32) attn_single_fwd_kernel(const AttnP a) {
  pdl_enter();
  __shared__ float sm_m[SQ_WARPS], sm_l[SQ_WARPS];
  __shared__ float sm_o[SQ_WARPS][HD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.x % a.G;
  const int h = (blockIdx.x / a.G) % a.H;
  const int b = blockIdx.x / (a.G * a.H);
  const float scale2 = a.scale * LOG2E;
  const float2 q = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.q + q_row(a, b, g, 0) * a.ldq + h * HD + 2 * lane));
  float m = -1e30f, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int j0 = warp; j0 < a.LkT; j0 += SQ_WARPS * 4) {
    // 4 keys in flight per warp iteration
    float2 kk[4], vv[4];
    float sc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * SQ_WARPS;
      if (j < a.LkT) {
        const long long off = k_row(a, b, g, j) * a.ldkv + h * HD + 2 * lane;
        kk[u] = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.k + off));
        vv[u] = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.v + off));
      } else {
        kk[u] = vv[u] = make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * SQ_WARPS;
      float d = warp_sum(q.x * kk[u].x + q.y * kk[u].y);
      float bj = 0.f;
      if (a.key_bias && j < a.LkT) bj = fmaxf(a.key_bias[(long long)b * a.LkT + j], -1e30f) * LOG2E;
      sc[u] = j < a.LkT ? fmaf(d, scale2, bj) : -1e30f;
    }
    float mn = fmaxf(fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3])), m);
    const float al = ex2(m - mn);
    l *= al;
    o0 *= al;
    o1 *= al;
    m = mn;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * SQ_WARPS;
      const float pj = j < a.LkT ? ex2(sc[u] - m) : 0.f;
      l += pj;
      o0 = fmaf(pj, vv[u].x, o0);
      o1 = fmaf(pj, vv[u].y, o1);
    }
  }
  if (lane == 0) {
    sm_m[warp] = m;
    sm_l[warp] = l;
  }
  sm_o[warp][2 * lane] = o0;
  sm_o[warp][2 * lane + 1] = o1;
  __syncthreads();
  if (warp == 0) {
    float M = -1e30f;
#pragma unroll
    for (int w = 0; w < SQ_WARPS; ++w) M = fmaxf(M, sm_m[w]);
    float L = 0.f, r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int w = 0; w < SQ_WARPS; ++w) {
      const float f = ex2(sm_m[w] - M);
      L = fmaf(sm_l[w], f, L);
      r0 = fmaf(sm_o[w][2 * lane], f, r0);
      r1 = fmaf(sm_o[w][2 * lane + 1], f, r1);
    }
    const float inv = 1.0f / L;
    *reinterpret_cast<uint32_t*>(a.o + o_row(a, b, g, 0) * a.ldo + h * HD + 2 * lane) = pack_bf16(r0 * inv, r1 * inv);
    if (lane == 0) a.lse[((long long)b * a.H + h) * a.G + g] = M + log2f(L);
  }
}

// Backward of the single-query attention: P_j = exp2(s_j - lse), dP_j = dO . v_j, dS_j = P_j (dP_j - delta);
// dq = scale * sum_j dS_j k_j;  dk_j (+)= scale * dS_j q;  dv_j (+)= P_j dO.  The shared CLS key goes to dkv_cls.
__global__ void __launch_bounds__(SQ_WARPS * 32) attn_single_bwd_kernel(const AttnP a) {
  pdl_enter();
  __shared__ float sm_q[SQ_WARPS][HD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.x % a.G;
  const int h = (blockIdx.x / a.G) % a.H;
  const int b = blockIdx.x / (a.G * a.H);
  const float scale2 = a.scale * LOG2E;
  const long long stat = ((long long)b * a.H + h) * a.G + g;
  const long long qoff = q_row(a, b, g, 0) * a.ldq + h * HD + 2 * lane;
  const long long ooff = o_row(a, b, g, 0) * a.ldo + h * HD + 2 * lane;
  const float2 q = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.q + qoff));
  const float2 d_o = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.d_o + ooff));
  const float2 o = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.o + ooff));
  const float delta = warp_sum(d_o.x * o.x + d_o.y * o.y);
  const float lse = a.lse[stat];
  if (threadIdx.x == 0) a.delta[stat] = delta;
  float dq0 = 0.f, dq1 = 0.f;
  for (int j = warp; j < a.LkT; j += SQ_WARPS) {
    const long long row = k_row(a, b, g, j);
    const long long off = row * a.ldkv + h * HD + 2 * lane;
    const float2 kk = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.k + off));
    const float2 vv = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.v + off));
    float s = q.x * kk.x + q.y * kk.y;
    float dp = d_o.x * vv.x + d_o.y * vv.y;
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, sh);
      dp += __shfl_xor_sync(0xffffffffu, dp, sh);
    }
    float bj = 0.f;
    if (a.key_bias) bj = fmaxf(a.key_bias[(long long)b * a.LkT + j], -1e30f) * LOG2E;
    const float pj = ex2(fmaf(s, scale2, bj) - lse);
    const float ds = pj * (dp - delta);
    dq0 = fmaf(ds, kk.x, dq0);
    dq1 = fmaf(ds, kk.y, dq1);
    const float gk0 = a.scale * ds * q.x, gk1 = a.scale * ds * q.y;
    const float gv0 = pj * d_o.x, gv1 = pj * d_o.y;
    if (a.has_cls && j == 0) {
      float* dst = a.dkv_cls + ((long long)b * a.H + h) * 2 * HD;
      atomicAdd(dst + 2 * lane, gk0);
      atomicAdd(dst + 2 * lane + 1, gk1);
      atomicAdd(dst + HD + 2 * lane, gv0);
      atomicAdd(dst + HD + 2 * lane + 1, gv1);
    } else {
      const long long goff = row * a.lddkv + h * HD + 2 * lane;
      uint32_t* pk = reinterpret_cast<uint32_t*>(a.dk + goff);
      uint32_t* pv = reinterpret_cast<uint32_t*>(a.dv + goff);
      if (a.dkv_accumulate) {
        const float2 ok = unpack_bf16(*pk), ov = unpack_bf16(*pv);
        *pk = pack_bf16(ok.x + gk0, ok.y + gk1);
        *pv = pack_bf16(ov.x + gv0, ov.y + gv1);
      } else {
        *pk = pack_bf16(gk0, gk1);
        *pv = pack_bf16(gv0, gv1);
      }
    }
  }
  sm_q[warp][2 * lane] = dq0;
  sm_q[warp][2 * lane + 1] = dq1;
  __syncthreads();
  if (warp == 0) {
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int w = 0; w < SQ_WARPS; ++w) {
      r0 += sm_q[w][2 * lane];
      r1 += sm_q[w][2 * lane + 1];
    }
    *reinterpret_cast<uint32_t*>(a.dq + q_row(a, b, g, 0) * a.lddq + h * HD + 2 * lane) = pack_bf16(r0 * a.scale, r1 * a.scale);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Single-query attention, all heads per warp (H % 4 == 0, H <= 16).  A warp reads the full key / value row of one
// token (H*64 bf16 = 1.5 KB contiguous for H = 12) with 16-byte loads: lane l holds elements [256 i + 8 l, +8) of
// group i, i.e. head 4 i + l / 8, so a dot product is an 8-lane reduction.  CTA = (batch, key split): B * SQ_SPLITS
// CTAs fill the GPU; partial (m, l, o) per (b, split, h) are merged by a second tiny kernel.
constexpr int SQ_SPLITS_MAX = 64;   // key splits per batch element (runtime: a.n_split, ~2 CTAs per SM)
constexpr int SQH_WARPS = 8;

struct bf16x8 { float v[8]; };
EGV_DEVINL bf16x8 unpack8(const uint4& u) {
  bf16x8 r;
  float2 t = unpack_bf16(u.x); r.v[0] = t.x; r.v[1] = t.y;
  t = unpack_bf16(u.y); r.v[2] = t.x; r.v[3] = t.y;
  t = unpack_bf16(u.z); r.v[4] = t.x; r.v[5] = t.y;
  t = unpack_bf16(u.w); r.v[6] = t.x; r.v[7] = t.y;
  return r;
}
EGV_DEVINL bf16x8 ld8(const bf16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  bf16x8 r;
  float2 t = unpack_bf16(u.x); r.v[0] = t.x; r.v[1] = t.y;
  t = unpack_bf16(u.y); r.v[2] = t.x; r.v[3] = t.y;
  t = unpack_bf16(u.z); r.v[4] = t.x; r.v[5] = t.y;
  t = unpack_bf16(u.w); r.v[6] = t.x; r.v[7] = t.y;
  return r;
}
EGV_DEVINL float dot8(const bf16x8& a, const bf16x8& b) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s = fmaf(a.v[e], b.v[e], s);
  return s;
}
EGV_DEVINL float red8(float s) {   // sum over the 8 lanes that share a head
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  return s;
}

// partial layout: part[((b * n_split + split) * H + h) * 66 + {0: m, 1: l, 2..65: o}]
template <int NG>
__global__ void __launch_bounds__(SQH_WARPS * 32) attn_single_fwd_heads_kernel(const AttnP a, float* __restrict__ part) {
  pdl_enter();
  __shared__ float sm_m[SQH_WARPS][NG * 4], sm_l[SQH_WARPS][NG * 4];
  __shared__ float sm_o[SQH_WARPS][NG * 256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int split = blockIdx.x % a.n_split, b = blockIdx.x / a.n_split;
  const float scale2 = a.scale * LOG2E;
  const int per = (a.LkT + a.n_split - 1) / a.n_split;
  const int j_lo = split * per, j_hi = min(a.LkT, j_lo + per);
  bf16x8 q[NG];
  const bf16* qp = a.q + q_row(a, b, 0, 0) * a.ldq;
#pragma unroll
  for (int i = 0; i < NG; ++i) q[i] = ld8(qp + 256 * i + 8 * lane);
  float m[NG], l[NG], o[NG][8];
#pragma unroll
  for (int i = 0; i < NG; ++i) {
    m[i] = -1e30f;
    l[i] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[i][e] = 0.f;
  }
  // two keys per iteration: both rows' loads (2 x 3 KB per warp) are in flight before the first dot product
  for (int j = j_lo + warp; j < j_hi; j += 2 * SQH_WARPS) {
    const int j2 = j + SQH_WARPS;
    const bool has2 = j2 < j_hi;
    const long long off = k_row(a, b, 0, j) * a.ldkv;
    const long long off2 = k_row(a, b, 0, has2 ? j2 : j) * a.ldkv;
    uint4 rk[2][NG], rv[2][NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) {
      rk[0][i] = *reinterpret_cast<const uint4*>(a.k + off + 256 * i + 8 * lane);
      rv[0][i] = *reinterpret_cast<const uint4*>(a.v + off + 256 * i + 8 * lane);
      rk[1][i] = *reinterpret_cast<const uint4*>(a.k + off2 + 256 * i + 8 * lane);
      rv[1][i] = *reinterpret_cast<const uint4*>(a.v + off2 + 256 * i + 8 * lane);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !has2) break;
      const int ju = u ? j2 : j;
      float bj = 0.f;
      if (a.key_bias) bj = fmaxf(a.key_bias[(long long)b * a.LkT + ju], -1e30f) * LOG2E;
#pragma unroll
      for (int i = 0; i < NG; ++i) {
        const bf16x8 kk = unpack8(rk[u][i]), vv = unpack8(rv[u][i]);
        const float sc = fmaf(red8(dot8(q[i], kk)), scale2, bj);
        const float mn = fmaxf(m[i], sc);
        const float al = ex2(m[i] - mn), pj = ex2(sc - mn);
        m[i] = mn;
        l[i] = fmaf(l[i], al, pj);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[i][e] = fmaf(o[i][e], al, pj * vv.v[e]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NG; ++i) {
    if ((lane & 7) == 0) {
      sm_m[warp][4 * i + (lane >> 3)] = m[i];
      sm_l[warp][4 * i + (lane >> 3)] = l[i];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sm_o[warp][256 * i + 8 * lane + e] = o[i][e];
  }
  __syncthreads();
  // merge the warps: thread t < H*64 owns output element t (head t / 64)
  for (int t = threadIdx.x; t < a.H * HD; t += SQH_WARPS * 32) {
    const int h = t >> 6;
    float M = -1e30f;
#pragma unroll
    for (int w = 0; w < SQH_WARPS; ++w) M = fmaxf(M, sm_m[w][h]);
    float L = 0.f, r = 0.f;
#pragma unroll
    for (int w = 0; w < SQH_WARPS; ++w) {
      const float f = ex2(sm_m[w][h] - M);
      L = fmaf(sm_l[w][h], f, L);
      r = fmaf(sm_o[w][t], f, r);
    }
    float* dst = part + (((long long)b * a.n_split + split) * a.H + h) * 66;
    dst[2 + (t & 63)] = r;
    if ((t & 63) == 0) {
      dst[0] = M;
      dst[1] = L;
    }
  }
}

__global__ void __launch_bounds__(64) attn_single_combine_kernel(const AttnP a, const float* __restrict__ part) {
  pdl_enter();
  __shared__ float sm_f[SQ_SPLITS_MAX];   // 2^(m_s - M) per split
  __shared__ float sm_ml[2];
  const int h = blockIdx.x % a.H, b = blockIdx.x / a.H;
  const int d = threadIdx.x;
  const float* base = part + ((long long)b * a.n_split * a.H + h) * 66;
  const long long sstride = (long long)a.H * 66;
  // every split's (m, l) is fetched by its own thread: one memory round trip instead of n_split dependent ones
  float ms = -1e30f, ls = 0.f;
  if (d < a.n_split) {
    ms = base[d * sstride];
    ls = base[d * sstride + 1];
  }
  float M = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, 16));
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  if ((d & 31) == 0) sm_ml[d >> 5] = M;
  __syncthreads();
  M = fmaxf(sm_ml[0], sm_ml[1]);
  __syncthreads();
  const float f = d < a.n_split ? ex2(ms - M) : 0.f;
  if (d < a.n_split) sm_f[d] = f;
  float L = warp_sum(ls * f);
  if ((d & 31) == 0) sm_ml[d >> 5] = L;
  __syncthreads();
  L = sm_ml[0] + sm_ml[1];
  float r = 0.f;
#pragma unroll 8
  for (int s = 0; s < a.n_split; ++s) r = fmaf(base[s * sstride + 2 + d], sm_f[s], r);
  a.o[o_row(a, b, 0, 0) * a.ldo + h * HD + d] = __float2bfloat16(r / L);
  if (d == 0) a.lse[(long long)b * a.H + h] = M + log2f(L);
}

// backward, all heads per warp.  dq is accumulated in fp32 (dq_acc [B, H*64], zeroed by the host wrapper) and written
// out by attn_single_dq_finalize_kernel.
template <int NG>
__global__ void __launch_bounds__(SQH_WARPS * 32) attn_single_bwd_heads_kernel(const AttnP a, float* dq_acc, int* counters) {
  pdl_enter();
  __shared__ float sm_q[SQH_WARPS][NG * 256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int split = blockIdx.x % a.n_split, b = blockIdx.x / a.n_split;
  const float scale2 = a.scale * LOG2E;
  const int per = (a.LkT + a.n_split - 1) / a.n_split;
  const int j_lo = split * per, j_hi = min(a.LkT, j_lo + per);
  bf16x8 q[NG], d_o[NG];
  float delta[NG], lse[NG], dq[NG][8];
  const bf16* qp = a.q + q_row(a, b, 0, 0) * a.ldq;
  const long long orow = o_row(a, b, 0, 0) * a.ldo;
#pragma unroll
  for (int i = 0; i < NG; ++i) {
    q[i] = ld8(qp + 256 * i + 8 * lane);
    d_o[i] = ld8(a.d_o + orow + 256 * i + 8 * lane);
    const bf16x8 oo = ld8(a.o + orow + 256 * i + 8 * lane);
    delta[i] = red8(dot8(d_o[i], oo));
    const int h = 4 * i + (lane >> 3);
    lse[i] = a.lse[(long long)b * a.H + h];
    if (split == 0 && warp == 0 && (lane & 7) == 0) a.delta[(long long)b * a.H + h] = delta[i];
#pragma unroll
    for (int e = 0; e < 8; ++e) dq[i][e] = 0.f;
  }
  for (int j = j_lo + warp; j < j_hi; j += SQH_WARPS) {
    const long long row = k_row(a, b, 0, j);
    const long long off = row * a.ldkv, goff = row * a.lddkv;
    const bool to_cls = a.has_cls && j == 0;
    bf16x8 kk[NG], vv[NG];
    uint4 ok[NG], ov[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) {
      kk[i] = ld8(a.k + off + 256 * i + 8 * lane);
      vv[i] = ld8(a.v + off + 256 * i + 8 * lane);
      if (a.dkv_accumulate && !to_cls) {
        ok[i] = *reinterpret_cast<const uint4*>(a.dk + goff + 256 * i + 8 * lane);
        ov[i] = *reinterpret_cast<const uint4*>(a.dv + goff + 256 * i + 8 * lane);
      } else {
        ok[i] = ov[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    float bj = 0.f;
    if (a.key_bias) bj = fmaxf(a.key_bias[(long long)b * a.LkT + j], -1e30f) * LOG2E;
#pragma unroll
    for (int i = 0; i < NG; ++i) {
      const float s = red8(dot8(q[i], kk[i]));
      const float dp = red8(dot8(d_o[i], vv[i]));
      const float pj = ex2(fmaf(s, scale2, bj) - lse[i]);
      const float ds = pj * (dp - delta[i]);
      float gk[8], gv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dq[i][e] = fmaf(ds, kk[i].v[e], dq[i][e]);
        gk[e] = a.scale * ds * q[i].v[e];
        gv[e] = pj * d_o[i].v[e];
      }
      if (to_cls) {
        const int h = 4 * i + (lane >> 3);
        float* dst = a.dkv_cls + ((long long)b * a.H + h) * 2 * HD + 8 * (lane & 7);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          atomicAdd(dst + e, gk[e]);
          atomicAdd(dst + HD + e, gv[e]);
        }
      } else {
        const uint32_t* pk = reinterpret_cast<const uint32_t*>(&ok[i]);
        const uint32_t* pv = reinterpret_cast<const uint32_t*>(&ov[i]);
        uint4 wk, wv;
        uint32_t* qk = reinterpret_cast<uint32_t*>(&wk);
        uint32_t* qv = reinterpret_cast<uint32_t*>(&wv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x = unpack_bf16(pk[e]), y = unpack_bf16(pv[e]);
          qk[e] = pack_bf16(x.x + gk[2 * e], x.y + gk[2 * e + 1]);
          qv[e] = pack_bf16(y.x + gv[2 * e], y.y + gv[2 * e + 1]);
        }
        *reinterpret_cast<uint4*>(a.dk + goff + 256 * i + 8 * lane) = wk;
        *reinterpret_cast<uint4*>(a.dv + goff + 256 * i + 8 * lane) = wv;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NG; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) sm_q[warp][256 * i + 8 * lane + e] = dq[i][e];
  __syncthreads();
  for (int t = threadIdx.x; t < a.H * HD; t += SQH_WARPS * 32) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < SQH_WARPS; ++w) r += sm_q[w][t];
    atomicAdd(dq_acc + (long long)b * a.H * HD + t, r * a.scale);
  }
  // last CTA of this batch element: fp32 accumulator -> bf16 dq row, and leave the accumulator zeroed for the next call
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(&counters[b], 1);
    s_last = ticket == a.n_split - 1;
    if (s_last) counters[b] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int t = threadIdx.x; t < a.H * HD; t += SQH_WARPS * 32) {
    float* src = dq_acc + (long long)b * a.H * HD + t;
    a.dq[q_row(a, b, 0, 0) * a.lddq + t] = __float2bfloat16(__ldcg(src));
    *src = 0.f;
  }
}

// zero-initialised scratch that the kernels leave zeroed again (split tickets; the fp32 dq accumulator)
template <typename T, int TAG>
static T* sq_zeroed(size_t n) {
  static T* buf = nullptr;
  static size_t cap = 0;
  if (n > cap) {
    if (buf) cudaFree(buf);
    cap = n * 2;
    if (cudaMalloc(&buf, cap * sizeof(T)) != cudaSuccess || cudaMemset(buf, 0, cap * sizeof(T)) != cudaSuccess) {
      buf = nullptr;
      cap = 0;
    }
  }
  return buf;
}

static float* sq_workspace(size_t floats) {
  static float* buf = nullptr;
  static size_t cap = 0;
  if (floats > cap) {
    if (buf) cudaFree(buf);
    cap = floats * 2;
    if (cudaMalloc(&buf, cap * sizeof(float)) != cudaSuccess) {
      buf = nullptr;
      cap = 0;
    }
  }
  return buf;
}

// key splits of the all-heads single-query kernels: ~2 resident CTAs per SM, >= 32 keys per split
static int single_query_splits(const AttnP& a) {
  int n = (2 * sm_count() + a.B - 1) / a.B;
  const int cap = a.LkT / 32;
  if (n > cap) n = cap;
  if (n > SQ_SPLITS_MAX) n = SQ_SPLITS_MAX;
  return n < 1 ? 1 : n;
}

static bool single_heads_ok(const AttnP& a) {
  return a.Lq == 1 && a.G == 1 && a.H % 4 == 0 && a.H <= 16 && a.LkT >= 256;
}

// fp32 CLS accumulators -> bf16 rows of the dk / dv tensors
__global__ void attn_cls_finalize_kernel(const float* __restrict__ dkv_cls, bf16* dk, bf16* dv, long long lddkv,
                                         long long kv_bstride, int cls_row, int B, int H, int accumulate) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H * HD) return;
  const int d = idx % HD, h = (idx / HD) % H, b = idx / (HD * H);
  const float* src = dkv_cls + ((long long)b * H + h) * 2 * HD;
  const long long off = ((long long)b * kv_bstride + cls_row) * lddkv + h * HD + d;
  float vk = src[d], vv = src[HD + d];
  if (accumulate) {
    vk += __bfloat162float(dk[off]);
    vv += __bfloat162float(dv[off]);
  }
  dk[off] = __float2bfloat16(vk);
  dv[off] = __float2bfloat16(vv);
}

// Merge the per-split partials of attn_unit (n_split > 1): one 64-thread block per (item, row).
template <int MODE>
__global__ void __launch_bounds__(64) attn_split_finalize_kernel(const AttnP a, int rows_cap) {
  pdl_enter();
  constexpr int W = MODE == MODE_FWD ? 66 : (MODE == MODE_DQ ? 64 : 128);
  const int n_rows = (MODE == MODE_DKV) ? a.LkT : a.Lq;
  const long long item_ns = blockIdx.x / rows_cap;
  const int r = blockIdx.x % rows_cap;
  const int tile = (int)(item_ns % a.row_tiles);
  long long rest = item_ns / a.row_tiles;
  const int h = (int)(rest % a.H);
  rest /= a.H;
  const int g = (int)(rest % a.G);
  const int b = (int)(rest / a.G);
  const int idx = tile * rows_cap + r;
  if (idx >= n_rows) return;
  const int d = threadIdx.x;
  const float* p = a.ws + ((item_ns * a.n_split) * rows_cap + r) * W;
  const long long sstride = (long long)rows_cap * W;
  if (MODE == MODE_FWD) {
    float M = -1e30f;
    for (int s = 0; s < a.n_split; ++s) M = fmaxf(M, p[s * sstride]);
    float L = 0.f, o = 0.f;
    for (int s = 0; s < a.n_split; ++s) {
      const float f = ex2(p[s * sstride] - M);
      L = fmaf(p[s * sstride + 1], f, L);
      o = fmaf(p[s * sstride + 2 + d], f, o);
    }
    a.o[o_row(a, b, g, idx) * a.ldo + h * HD + d] = __float2bfloat16(o / L);
    if (d == 0) a.lse[(((long long)b * a.H + h) * a.G + g) * a.Lq + idx] = M + log2f(L);
  } else if (MODE == MODE_DQ) {
    float v = 0.f;
    for (int s = 0; s < a.n_split; ++s) v += p[s * sstride + d];
    a.dq[q_row(a, b, g, idx) * a.lddq + h * HD + d] = __float2bfloat16(v);
  } else {
    if (a.has_cls && idx == 0) return;   // the shared CLS key was accumulated into dkv_cls by the splits
    float vk = 0.f, vv = 0.f;
    for (int s = 0; s < a.n_split; ++s) {
      vk += p[s * sstride + d];
      vv += p[s * sstride + 64 + d];
    }
    const long long off = k_row(a, b, g, idx) * a.lddkv + h * HD + d;
    if (a.dkv_accumulate) {
      vk += __bfloat162float(a.dk[off]);
      vv += __bfloat162float(a.dv[off]);
    }
    a.dk[off] = __float2bfloat16(vk);
    a.dv[off] = __float2bfloat16(vv);
  }
}

// Stream splits for a (few rows, long stream) problem: enough CTAs to fill the GPU, >= 4 chunks of 64 rows per split.
static int pick_split(long long items, int n_str) {
  const int chunks = (n_str + 63) / 64;
  if (items >= 2 * sm_count() || chunks < 8) return 1;
  long long want = cdiv(4LL * sm_count(), items);
  long long cap = chunks / 4;
  long long n = want < cap ? want : cap;
  return (int)(n < 1 ? 1 : (n > 32 ? 32 : n));
}
template <int MODE>
static long long split_ws_floats(long long items_ns, int n_split, int rows_cap) {
  constexpr int W = MODE == MODE_FWD ? 66 : (MODE == MODE_DQ ? 64 : 128);
  return items_ns * n_split * rows_cap * W;
}

static int fill_params(const egv_attn_args* x, AttnP& a, bool bwd) {
  if (!x || !x->q || !x->k || !x->v) return fail(EGV_ERR_ARG, "attention: null q/k/v");
  if (x->B <= 0 || x->H <= 0 || x->G <= 0 || x->Lq <= 0 || x->Lk < 0) return fail(EGV_ERR_ARG, "attention: bad sizes");
  if (!x->o || !x->lse) return fail(EGV_ERR_ARG, "attention: o and lse are required");
  if ((x->ldq % 8) || (x->ldkv % 8) || (x->ldo % 8)) return fail(EGV_ERR_ARG, "attention: row strides must be multiples of 8");
  auto al = [](const void* p) { return (((uintptr_t)p) & 15) == 0; };
  if (!al(x->q) || !al(x->k) || !al(x->v) || !al(x->o)) return fail(EGV_ERR_ARG, "attention: pointers must be 16-byte aligned");
  a.B = x->B; a.H = x->H; a.G = x->G; a.Lq = x->Lq; a.LkT = x->Lk + (x->has_cls_key ? 1 : 0);
  if (a.LkT <= 0) return fail(EGV_ERR_ARG, "attention: no keys");
  a.q = (const bf16*)x->q; a.ldq = x->ldq; a.q_bstride = x->q_bstride;
  a.q_row0 = x->q_row0; a.q_gstride = x->q_gstride; a.q_istride = x->q_istride;
  a.k = (const bf16*)x->k; a.v = (const bf16*)x->v; a.ldkv = x->ldkv; a.kv_bstride = x->kv_bstride;
  a.k_row0 = x->k_row0; a.k_gstride = x->k_gstride; a.k_istride = x->k_istride;
  a.has_cls = x->has_cls_key ? 1 : 0; a.cls_row = x->cls_row;
  a.key_bias = x->key_bias; a.scale = x->scale;
  a.o = (bf16*)x->o; a.ldo = x->ldo; a.o_bstride = x->o_bstride;
  a.lse = x->lse;
  a.d_o = (const bf16*)x->d_o; a.dq = (bf16*)x->dq; a.lddq = x->lddq;
  a.dk = (bf16*)x->dk; a.dv = (bf16*)x->dv; a.lddkv = x->lddkv;
  a.delta = x->delta; a.dkv_cls = x->dkv_cls; a.dkv_accumulate = x->dkv_accumulate;
  a.lse_cls = bwd ? x->lse_cls : nullptr; a.dq_cls = bwd ? x->dq_cls : nullptr;
  a.cls_part = nullptr; a.fold_out = nullptr;
  a.n_split = 1; a.ws = (float*)x->workspace; a.ws_floats = x->workspace_bytes / 4;
  if (bwd) {
    if (!x->d_o || !x->dq || !x->dk || !x->dv || !x->delta) return fail(EGV_ERR_ARG, "attention bwd: null gradient buffer");
    if ((x->lddq % 8) || (x->lddkv % 8)) return fail(EGV_ERR_ARG, "attention bwd: gradient strides must be multiples of 8");
    if (!al(x->d_o) || !al(x->dq) || !al(x->dk) || !al(x->dv)) return fail(EGV_ERR_ARG, "attention bwd: unaligned gradient pointer");
    if (a.has_cls && !x->dkv_cls) return fail(EGV_ERR_ARG, "attention bwd: dkv_cls required with has_cls_key");
  }
  return EGV_OK;
}

template <int MODE>
static int launch_mode(AttnP a, cudaStream_t stream) {
  const int n_rows = (MODE == MODE_DKV) ? a.LkT : a.Lq;
  const int n_str = (MODE == MODE_DKV) ? a.Lq : a.LkT;
  const long long groups = (long long)a.B * a.H * a.G;
  const bool small = n_rows <= 32 && n_str <= 64 && groups >= 1024;
  cudaError_t e = cudaSuccess;
  {
    // large contiguous groups (space attention): tcgen05 / TMEM kernel (attention_tc.cu), else the group-resident
    // TMA-fed mma.sync kernels (attention_group.cu)
    int r = launch_tc_attention(MODE, a, stream);
    if (r != 0) return r < 0 ? r : EGV_OK;
    r = launch_group_attention(MODE, a, stream);
    if (r != 0) return r < 0 ? r : EGV_OK;
  }
  if (small) {
    constexpr int KC = 32;
    a.row_tiles = (int)cdiv(n_rows, 16);
    a.items = groups * a.row_tiles;
    auto kern = attn_warp_kernel<MODE, KC>;
    const int smem = 4 * AttnSmem<MODE, 1, KC>::BYTES;
    static bool cfg = false;
    if (!cfg) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cfg = true;
    }
    long long grid = cdiv(a.items, 4);
    const long long cap = (long long)sm_count() * 64;
    if (grid > cap) grid = cap;
    if (e == cudaSuccess) launch_k(kern, dim3((unsigned)grid), dim3(128), smem, stream, a);
  } else if (n_rows <= 16) {
    // few rows, long stream (CLS query; tiny text batches): one warp is all the row side can use
    constexpr int KC = 64;
    a.row_tiles = (int)cdiv(n_rows, 16);
    a.items = groups * a.row_tiles;
    {
      const int ns = pick_split(a.items, n_str);
      if (ns > 1 && a.ws && split_ws_floats<MODE>(a.items, ns, 16) <= a.ws_floats) {
        a.n_split = ns;
        a.items *= ns;
      }
    }
    auto kern = attn_cta_kernel<MODE, 1, KC, 1, true>;
    const int smem = AttnSmem<MODE, 1, KC>::BYTES;
    static bool cfg = false;
    if (!cfg) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cfg = true;
    }
    long long grid = std::min<long long>(a.items, (long long)sm_count() * 32);
    if (e == cudaSuccess) launch_k(kern, dim3((unsigned)grid), dim3(32), smem, stream, a);
    if (e == cudaSuccess && a.n_split > 1) {
      int rc = check_launch("attention kernel");
      if (rc) return rc;
      launch_k(attn_split_finalize_kernel<MODE>, dim3((unsigned)(a.items / a.n_split * 16)), dim3(64), 0, stream, a, 16);
    }
  } else if (n_rows <= 32) {
    constexpr int KC = 64;
    a.row_tiles = (int)cdiv(n_rows, 32);
    a.items = groups * a.row_tiles;
    {
      const int ns = pick_split(a.items, n_str);
      if (ns > 1 && a.ws && split_ws_floats<MODE>(a.items, ns, 32) <= a.ws_floats) {
        a.n_split = ns;
        a.items *= ns;
      }
    }
    auto kern = attn_cta_kernel<MODE, 2, KC, 1, true>;
    const int smem = AttnSmem<MODE, 2, KC>::BYTES;
    static bool cfg = false;
    if (!cfg) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cfg = true;
    }
    long long grid = std::min<long long>(a.items, (long long)sm_count() * 16);
    if (e == cudaSuccess) launch_k(kern, dim3((unsigned)grid), dim3(64), smem, stream, a);
    if (e == cudaSuccess && a.n_split > 1) {
      int rc = check_launch("attention kernel");
      if (rc) return rc;
      launch_k(attn_split_finalize_kernel<MODE>, dim3((unsigned)(a.items / a.n_split * 32)), dim3(64), 0, stream, a, 32);
    }
  } else {
    // tile shapes: 4 warps x 64-row chunks, or 7 warps x 112-row chunks when that wastes fewer MMA slots
    // (196-197 rows per space-attention group: 2 x 112 = 224 instead of 4 x 64 = 256); 32-row chunks for short streams.
    auto util = [](int n, int t) { return (double)n / (double)(cdiv(n, t) * t); };
    static const int wide_mode = getenv("EGV_ATTN_WIDE") ? atoi(getenv("EGV_ATTN_WIDE")) : 6;   // bit per MODE
    const bool wide = ((wide_mode >> MODE) & 1) && n_str > 32 &&
                      util(n_rows, 112) * util(n_str, 112) > 1.1 * util(n_rows, 64) * util(n_str, 64);
    static const int resident_mode = getenv("EGV_ATTN_RESIDENT") ? atoi(getenv("EGV_ATTN_RESIDENT")) : 6;
    if (((resident_mode >> MODE) & 1) && n_str > 64 && n_str <= 256 && n_rows > 64) {
      // the whole stream side of a group stays in shared memory (space attention: 196-197 rows)
      constexpr int KC = 64, NCH = 4;
      constexpr int NWR = (MODE == MODE_DKV) ? 4 : 7;
      a.row_tiles = (int)cdiv(n_rows, 16 * NWR);
      a.items = groups * a.row_tiles;
      auto kern = attn_cta_kernel<MODE, NWR, KC, NCH>;
      const int smem = AttnSmem<MODE, NWR, KC, NCH>::BYTES;
      static bool cfg = false;
      if (!cfg) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cfg = true;
      }
      long long grid = std::min<long long>(a.items, (long long)sm_count() * 4);
      if (e == cudaSuccess) launch_k(kern, dim3((unsigned)grid), dim3(NWR * 32), smem, stream, a);
    } else if (n_str <= 32) {
      constexpr int KC = 32;
      a.row_tiles = (int)cdiv(n_rows, 64);
      a.items = groups * a.row_tiles;
      auto kern = attn_cta_kernel<MODE, 4, KC>;
      const int smem = AttnSmem<MODE, 4, KC>::BYTES;
      static bool cfg = false;
      if (!cfg) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cfg = true;
      }
      long long grid = std::min<long long>(a.items, (long long)sm_count() * 8);
      if (e == cudaSuccess) launch_k(kern, dim3((unsigned)grid), dim3(128), smem, stream, a);
    } else if (wide) {
      constexpr int KC = 112;
      a.row_tiles = (int)cdiv(n_rows, 112);
      a.items = groups * a.row_tiles;
      auto kern = attn_cta_kernel<MODE, 7, KC>;
      const int smem = AttnSmem<MODE, 7, KC>::BYTES;
      static bool cfg = false;
      if (!cfg) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cfg = true;
      }
      long long grid = std::min<long long>(a.items, (long long)sm_count() * 4);
      if (e == cudaSuccess) launch_k(kern, dim3((unsigned)grid), dim3(224), smem, stream, a);
    } else {
      constexpr int KC = 64;
      a.row_tiles = (int)cdiv(n_rows, 64);
      a.items = groups * a.row_tiles;
      auto kern = attn_cta_kernel<MODE, 4, KC>;
      const int smem = AttnSmem<MODE, 4, KC>::BYTES;
      static bool cfg = false;
      if (!cfg) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cfg = true;
      }
      long long grid = std::min<long long>(a.items, (long long)sm_count() * 8);
      if (e == cudaSuccess) launch_k(kern, dim3((unsigned)grid), dim3(128), smem, stream, a);
    }
  }
  if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "attention smem attribute: %s", cudaGetErrorString(e));
  return check_launch("attention kernel");
}

}  // namespace egv

using namespace egv;

extern "C" void egv_attention_set_tc(int mode) { set_tc_attention_mode(mode); }
extern "C" void egv_attention_set_tiny(int mode) { set_tiny_mode(mode); }

extern "C" int64_t egv_attention_workspace_bytes(const egv_attn_args* x) {
  // upper bound over the forward and both backward modes of the split-stream path (0 when it cannot apply)
  if (!x || x->Lq <= 0) return 0;
  const int lkt = x->Lk + (x->has_cls_key ? 1 : 0);
  const long long groups = (long long)x->B * x->H * x->G;
  long long need = 0;
  auto consider = [&](int n_rows, int n_str, int w) {
    if (n_rows > 32 || (n_rows <= 32 && n_str <= 64 && groups >= 1024)) return;
    const int cap = n_rows <= 16 ? 16 : 32;
    const int ns = pick_split(groups, n_str);
    if (ns > 1) need = std::max<long long>(need, groups * ns * cap * w * 4);
  };
  consider(x->Lq, lkt, 66);    // FWD / DQ: rows = queries
  consider(lkt, x->Lq, 128);   // DKV: rows = keys
  return need;
}

extern "C" int egv_attention_fwd(const egv_attn_args* x, egv_stream_t stream) {
  AttnP a;
  int rc = fill_params(x, a, false);
  if (rc) return rc;
  if (single_heads_ok(a)) {
    a.n_split = single_query_splits(a);
    float* part = sq_workspace((size_t)a.B * SQ_SPLITS_MAX * a.H * 66);
    if (!part) return fail(EGV_ERR_CUDA, "attention: workspace allocation failed");
    const unsigned grid = (unsigned)(a.B * a.n_split);
    cudaStream_t s = (cudaStream_t)stream;
    switch (a.H / 4) {
      case 1: launch_k(attn_single_fwd_heads_kernel<1>, dim3(grid), dim3(SQH_WARPS * 32), 0, s, a, part); break;
      case 2: launch_k(attn_single_fwd_heads_kernel<2>, dim3(grid), dim3(SQH_WARPS * 32), 0, s, a, part); break;
      case 3: launch_k(attn_single_fwd_heads_kernel<3>, dim3(grid), dim3(SQH_WARPS * 32), 0, s, a, part); break;
      default: launch_k(attn_single_fwd_heads_kernel<4>, dim3(grid), dim3(SQH_WARPS * 32), 0, s, a, part); break;
    }
    int rc2 = check_launch("attn_single_fwd_heads_kernel");
    if (rc2) return rc2;
    // (merging the splits in the last CTA of each batch element instead was measured 5 us SLOWER: a serial tail)
    launch_k(attn_single_combine_kernel, dim3((unsigned)(a.B * a.H)), dim3(64), 0, s, a, part);
    return check_launch("attn_single_combine_kernel");
  }
  if (a.Lq == 1) {
    launch_k(attn_single_fwd_kernel, dim3((unsigned)(a.B * a.H * a.G)), dim3(SQ_WARPS * 32), 0, (cudaStream_t)stream, a);
    return check_launch("attn_single_fwd_kernel");
  }
  {
    const int r = launch_tiny_attention(0, a, (cudaStream_t)stream);   // many tiny groups (time attention)
    if (r != 0) return r < 0 ? r : EGV_OK;
  }
  // optional: the group kernel takes the clip's CLS query along (egv_attn_args::lse_cls is the OUTPUT here): per-frame
  // partials in the library's single-query scratch, merged by the single-query combine kernel
  int folded = 0;
  if (x->cls_query_folded) *x->cls_query_folded = 0;
  if (x->lse_cls && x->cls_query_folded && a.has_cls && a.G <= SQ_SPLITS_MAX && a.H % 4 == 0 && a.H <= 16) {
    a.cls_part = sq_workspace((size_t)a.B * SQ_SPLITS_MAX * a.H * 66);
    a.fold_out = &folded;
  }
  rc = launch_mode<MODE_FWD>(a, (cudaStream_t)stream);
  if (rc || !folded) return rc;
  AttnP c = a;     // the CLS query as a single-query problem whose G key splits are the frames
  c.n_split = a.G;
  c.G = 1;
  c.Lq = 1;
  c.q_row0 = a.cls_row;
  c.q_gstride = 0;
  c.lse = const_cast<float*>(x->lse_cls);
  launch_k(attn_single_combine_kernel, dim3((unsigned)(a.B * a.H)), dim3(64), 0, (cudaStream_t)stream, c, (const float*)a.cls_part);
  rc = check_launch("attn_single_combine_kernel");
  if (!rc) *x->cls_query_folded = 1;
  return rc;
}

extern "C" int egv_attention_bwd(const egv_attn_args* x, egv_stream_t stream) {
  AttnP a;
  int rc = fill_params(x, a, true);
  if (rc) return rc;
  if (x->cls_query_folded) *x->cls_query_folded = 0;
  if (single_heads_ok(a)) {
    const size_t n = (size_t)a.B * a.H * HD;
    float* acc = sq_zeroed<float, 1>(n);            // zero between calls: the last CTA per batch element re-zeroes it
    int* counters = sq_zeroed<int, 2>((size_t)a.B);
    if (!acc || !counters) return fail(EGV_ERR_CUDA, "attention: workspace allocation failed");
    cudaStream_t s = (cudaStream_t)stream;
    a.n_split = single_query_splits(a);
    const unsigned grid = (unsigned)(a.B * a.n_split);
    switch (a.H / 4) {
      case 1: launch_k(attn_single_bwd_heads_kernel<1>, dim3(grid), dim3(SQH_WARPS * 32), 0, s, a, acc, counters); break;
      case 2: launch_k(attn_single_bwd_heads_kernel<2>, dim3(grid), dim3(SQH_WARPS * 32), 0, s, a, acc, counters); break;
      case 3: launch_k(attn_single_bwd_heads_kernel<3>, dim3(grid), dim3(SQH_WARPS * 32), 0, s, a, acc, counters); break;
      default: launch_k(attn_single_bwd_heads_kernel<4>, dim3(grid), dim3(SQH_WARPS * 32), 0, s, a, acc, counters); break;
    }
    return check_launch("attn_single_bwd_heads_kernel");
  }
  if (a.Lq == 1) {
    launch_k(attn_single_bwd_kernel, dim3((unsigned)(a.B * a.H * a.G)), dim3(SQ_WARPS * 32), 0, (cudaStream_t)stream, a);
    return check_launch("attn_single_bwd_kernel");
  }
  {
    const int r = launch_tiny_attention(1, a, (cudaStream_t)stream);   // fused single-launch backward for tiny groups
    if (r != 0) return r < 0 ? r : EGV_OK;
  }
  {
    int folded = 0;
    const int r = launch_tc_attention_bwd(a, (cudaStream_t)stream, &folded);    // space attention: one tcgen05 launch
    if (r > 0 && x->cls_query_folded) *x->cls_query_folded = folded;
    if (r != 0) return r < 0 ? r : EGV_OK;
  }
  rc = launch_mode<MODE_DQ>(a, (cudaStream_t)stream);   // also produces delta
  if (rc) return rc;
  return launch_mode<MODE_DKV>(a, (cudaStream_t)stream);
}

__global__ void attn_cls_query_finalize_kernel(const float* __restrict__ dq_cls, bf16* dq, long long lddq, long long q_bstride,
                                               int cls_row, int B, int H) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H * HD) return;
  const int b = idx / (H * HD), c = idx % (H * HD);
  dq[((long long)b * q_bstride + cls_row) * lddq + c] = __float2bfloat16(dq_cls[idx]);
}

extern "C" int egv_attention_cls_query_finalize(const float* dq_cls, void* dq, int64_t lddq, int64_t q_bstride, int cls_row, int B,
                                                int H, egv_stream_t stream) {
  if (!dq_cls || !dq) return fail(EGV_ERR_ARG, "cls_query_finalize: null pointer");
  const int n = B * H * HD;
  launch_k(attn_cls_query_finalize_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, dq_cls, (bf16*)dq, lddq, q_bstride, cls_row, B, H);
  return check_launch("attn_cls_query_finalize_kernel");
}

extern "C" int egv_attention_cls_finalize(const float* dkv_cls, void* dk, void* dv, int64_t lddkv, int64_t kv_bstride,
                                          int cls_row, int B, int H, int accumulate, egv_stream_t stream) {
  if (!dkv_cls || !dk || !dv) return fail(EGV_ERR_ARG, "cls_finalize: null pointer");
  const int n = B * H * HD;
  launch_k(attn_cls_finalize_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, dkv_cls, (bf16*)dk, (bf16*)dv, lddkv, kv_bstride, cls_row, B, H, accumulate);
  return check_launch("attn_cls_finalize_kernel");
}
