// Divided SPACE attention (video_transformer.py:35-39, 117-153; the `(b f) n d` groups: 196 patch queries of one frame
// attend the frame's 196 keys + the clip's CLS key) on the 5th-generation tensor cores.
//
// One persistent CTA per SM walks the (clip, frame, head) problems.  Per problem:
//   warp 0        TMA producer: the group's Q rows (two 128-row m-tiles), K and V rows (one box each, 128B swizzle); the
//                 shared CLS key / value row is appended behind the frame's keys with 8 + 8 swizzled 16-byte copies
//   warp 1        tcgen05.mma issuer (one thread):  S_w = Q_w K^T  (M = 128, N = keys rounded up to 16, K = 64) into TMEM,
//                 then  O_w = P_w V  (M = 128, N = 64, K = keys) with P_w as a K-major bf16 A operand in shared memory and
//                 V as the MN-major B operand exactly as TMA delivered it
//   warp 2        TMEM allocator (512 columns: one 256-column region per m-tile; O_w aliases the consumed S_w columns)
//   warps 4-11 / 12-19   softmax warps of m-tile 0 / 1: thread = (query row, column half).  tcgen05.ld 32 score columns at a
//                 time, two passes over TMEM (row max, then exp2 / row sum) -- nothing but the running max and sum lives in
//                 registers; the two halves of a row meet through shared memory --, P written as swizzled bf16; after
//                 the PV product: O / rowsum -> bf16 -> swizzled staging -> coalesced 128-byte row stores; lse (log2
//                 domain, as the backward kernels expect it)
// Every buffer is single-buffered with early release: Q_w and K are free as soon as the S products retire (the producer
// refills them while the softmax of the same problem runs), V when the PV products retire.
// Eligibility (else the caller falls back to the mma.sync kernels of attention_group.cu): contiguous groups with the shared
// CLS key, 64 < queries <= 256, 16 <= keys + 1 <= 208, no key bias.
#include <cuda.h>
#include <stdlib.h>

#include "attention.cuh"

namespace egv {
namespace atc {

constexpr int THREADS = 128 + 16 * 32;   // producer, MMA issuer, TMEM allocator, (idle) + 16 softmax warps
constexpr int QT_BYTES = 128 * 128;          // one 128-row m-tile of Q, 128 B per row
constexpr int KV_ROWS = 208;
constexpr int KV_BYTES = 27 * 1024;          // 208 rows x 128 B rounded up to the 1024 B swizzle period
constexpr int P_BYTES = 4 * 16384;           // 128 x 256 bf16 as four K-major [128][64] sub-tiles
constexpr int OFF_Q = 0;
constexpr int OFF_K = 2 * QT_BYTES;
constexpr int OFF_V = OFF_K + KV_BYTES;
constexpr int OFF_P = OFF_V + KV_BYTES;
constexpr int OFF_XCH = OFF_P + 2 * P_BYTES;   // row max / row sum exchange between the two column halves: [2][2][2][128] f32
constexpr int OFF_BAR = OFF_XCH + 4096;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
static_assert(SMEM_BYTES <= 232448, "space attention tile set exceeds the 227 KB shared memory of an SM");

struct Maps {
  CUtensorMap q0, q1, k, v;
};
struct Prm {
  int n_mt;        // m-tiles per group (1 or 2)
  int rows1;       // query rows of m-tile 1
  int lk;          // regular keys (the CLS key is row lk of the K / V tiles)
  int npad;        // keys + 1 rounded up to a multiple of 16
  long long total;
  int fold;        // 1: the clip's CLS QUERY rides along as query slot `rows1` of m-tile 1 (video_transformer.py:134-150: it
                   // attends every token of the clip): per-frame partial (max, sum, unnormalised O) into `part`, merged by
                   // attn_single_combine_kernel; the (CLS query, CLS key) pair is counted in frame 0 only
  float* part;     // [B, G, H, 66] f32
};

EGV_DEVINL void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) attn_tc_fwd_kernel(const __grid_constant__ Maps maps, const AttnP a, const Prm pr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;          // [2]
  uint64_t* q_empty = bars + 2;     // [2]
  uint64_t* k_full = bars + 4;
  uint64_t* k_empty = bars + 5;
  uint64_t* v_full = bars + 6;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 8;      // [2]
  uint64_t* p_ready = bars + 10;    // [2]
  uint64_t* o_full = bars + 12;     // [2]
  uint64_t* r_empty = bars + 14;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows the TMA boxes never touch (the tail of m-tile 1, key rows behind the CLS row) must hold finite values
  for (int i = threadIdx.x; i < OFF_P / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q0);
    tma_prefetch_desc(&maps.q1);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
  }
  if (warp == 1 && lane == 0) {
    for (int w = 0; w < 2; ++w) {
      mbar_init(&q_full[w], w == 1 && pr.fold ? 2 : 1);   // m-tile 1: + the CLS query's row copy
      mbar_init(&q_empty[w], 1);
      mbar_init(&s_full[w], 1);
      mbar_init(&p_ready[w], 8);
      mbar_init(&o_full[w], 1);
      mbar_init(&r_empty[w], 8);
    }
    mbar_init(k_full, 2);     // TMA transaction + the CLS row copy
    mbar_init(k_empty, 1);
    mbar_init(v_full, 2);
    mbar_init(v_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int HG = a.H * a.G;
  pdl_wait();     // programmatic dependent launch (common.cuh): before any global access; dependents start when this CTA exits

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    uint32_t ph = 0;
    for (long long p = blockIdx.x; p < pr.total; p += gridDim.x, ph ^= 1) {
      const int h = (int)(p % a.H), g = (int)((p / a.H) % a.G), b = (int)(p / HG);
      const int q_first = (int)((long long)b * a.q_bstride + a.q_row0 + (long long)g * a.q_gstride);
      const int k_first = (int)((long long)b * a.kv_bstride + a.k_row0 + (long long)g * a.k_gstride);
      const long long cls_off = ((long long)b * a.kv_bstride + a.cls_row) * a.ldkv + h * HD + (lane & 7) * 8;
      for (int w = 0; w < pr.n_mt; ++w) {
        mbar_wait_sleep(&q_empty[w], ph ^ 1, 64);
        if (lane == 0) {
          const int rows = w == 0 ? (a.Lq < 128 ? a.Lq : 128) : pr.rows1;
          mbar_arrive_expect_tx(&q_full[w], (uint32_t)rows * 128u);
          tma_load_2d(smem + OFF_Q + w * QT_BYTES, w == 0 ? &maps.q0 : &maps.q1, &q_full[w], h * HD, q_first + 128 * w);
        }
        if (w == 1 && pr.fold) {   // the CLS query's row: slot rows1 of m-tile 1, 128B-swizzled like the TMA rows
          if (lane < 8) {
            const uint4 val = *reinterpret_cast<const uint4*>(a.q + ((long long)b * a.q_bstride + a.cls_row) * a.ldq + h * HD + lane * 8);
            *reinterpret_cast<uint4*>(smem + OFF_Q + QT_BYTES + pr.rows1 * 128 + ((lane ^ (pr.rows1 & 7)) << 4)) = val;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&q_full[1]);
        }
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        uint64_t* full = t == 0 ? k_full : v_full;
        uint64_t* empty = t == 0 ? k_empty : v_empty;
        uint8_t* dst = smem + (t == 0 ? OFF_K : OFF_V);
        mbar_wait_sleep(empty, ph ^ 1, 64);
        if (lane == 0) {
          mbar_arrive_expect_tx(full, (uint32_t)pr.lk * 128u);
          tma_load_2d(dst, t == 0 ? &maps.k : &maps.v, full, h * HD, k_first);
        }
        if (lane < 8) {   // the clip's CLS key / value: row lk of the tile, 128B-swizzled like the TMA rows
          const uint4 val = *reinterpret_cast<const uint4*>((t == 0 ? a.k : a.v) + cls_off);
          *reinterpret_cast<uint4*>(dst + pr.lk * 128 + (((lane & 7) ^ (pr.lk & 7)) << 4)) = val;
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's (async proxy) reads
        __syncwarp();
        if (lane == 0) mbar_arrive(full);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {   // elect.sync: bare UTCHMMA instead of a convergence loop per MMA
      const uint32_t idesc_s = umma_idesc_bf16(128, pr.npad, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(128, HD, 0, 1);
      const int ksteps = pr.npad / 16;
      const uint32_t sq = smem_u32(smem + OFF_Q), sk = smem_u32(smem + OFF_K), sv = smem_u32(smem + OFF_V), sp = smem_u32(smem + OFF_P);
      uint32_t ph = 0;
      for (long long p = blockIdx.x; p < pr.total; p += gridDim.x, ph ^= 1) {
        for (int w = 0; w < pr.n_mt; ++w) {
          mbar_wait_sleep(&r_empty[w], ph ^ 1, 32);
          mbar_wait_sleep(&q_full[w], ph, 20);
          if (w == 0) mbar_wait_sleep(k_full, ph, 20);
          tc_fence_after();
          const uint64_t qd = umma_desc_sw128(sq + w * QT_BYTES, 16, 1024), kd = umma_desc_sw128(sk, 16, 1024);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + (uint32_t)(w * 256), qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&q_empty[w]);
          umma_commit(&s_full[w]);
        }
        umma_commit(k_empty);
        for (int w = 0; w < pr.n_mt; ++w) {
          mbar_wait_sleep(&p_ready[w], ph, 20);
          if (w == 0) mbar_wait_sleep(v_full, ph, 20);
          tc_fence_after();
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t pd = umma_desc_sw128(sp + w * P_BYTES + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
            const uint64_t vd = umma_desc_sw128(sv + kk * 2048, 8192, 1024);
            umma_bf16(tmem_base + (uint32_t)(w * 256), pd, vd, idesc_pv, kk > 0 ? 1u : 0u);
          }
          umma_commit(&o_full[w]);
        }
        umma_commit(v_empty);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax / output warps
    // 8 warps per m-tile: TMEM lane quadrant `quad` (= warp % 4) x column half `half`.  A row's two halves exchange their
    // partial maximum / sum through shared memory behind a 64-thread named barrier.
    const int sw = warp - 4;
    const int w = sw >> 3;                    // m-tile
    const int half = (sw >> 2) & 1;
    const int quad = sw & 3;
    if (w < pr.n_mt) {
      const int row = quad * 32 + lane;       // row of the m-tile = TMEM lane
      const int rows_tile = w == 0 ? (a.Lq < 128 ? a.Lq : 128) : pr.rows1;
      const bool cls_q = pr.fold && w == 1 && row == pr.rows1;           // this thread's row is the CLS query
      const bool warp_live = quad * 32 < rows_tile + (w == 1 ? pr.fold : 0);    // warp-uniform: any valid row in this warp
      const float sl2 = a.scale * LOG2E;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(w * 256);
      uint8_t* pw = smem + OFF_P + w * P_BYTES;
      float* xch = reinterpret_cast<float*>(smem + OFF_XCH) + w * 256;   // [plane: max, sum][m-tile][half][128 rows]
      const int pair_bar = 1 + w * 4 + quad;   // named barrier of the two warps sharing these rows
      // column range of this half: the split is a multiple of 32 (TMEM load / P chunk alignment)
      const int c_mid = (pr.npad / 64) * 32;
      const int c_lo = half == 0 ? 0 : c_mid, c_hi = half == 0 ? c_mid : pr.npad;
      uint32_t ph = 0;
      for (long long p = blockIdx.x; p < pr.total; p += gridDim.x, ph ^= 1) {
        const int h = (int)(p % a.H), g = (int)((p / a.H) % a.G), b = (int)(p / HG);
        const long long o_first = (long long)b * a.o_bstride + a.q_row0 + (long long)g * a.q_gstride + 128 * w;
        const long long stat_base = (((long long)b * a.H + h) * a.G + g) * a.Lq + 128 * w;
        const int nkey = pr.lk + ((cls_q && g != 0) ? 0 : 1);   // the CLS query meets the CLS key in frame 0 only
        float* cpart = pr.part + (((long long)b * a.G + g) * a.H + h) * 66;
        mbar_wait_sleep(&s_full[w], ph, 32);   // sleeping polls: 16 tightly polling warps take issue slots from the MMA issuer and the producer
        tc_fence_after();
        float inv = 0.f;
        if (warp_live) {
          uint32_t v[32];
          // pass 1: row maximum over this half's columns (four independent chains)
          float mx0 = -3.0e38f, mx1 = -3.0e38f, mx2 = -3.0e38f, mx3 = -3.0e38f;
          for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
            const bool narrow = c_hi - c0 < 32;
            if (narrow) tmem_ld_32x16(taddr + (uint32_t)c0, v);
            else tmem_ld_32x32(taddr + (uint32_t)c0, v);
            tmem_ld_wait();
            const int nv = min(narrow ? 16 : 32, nkey - c0);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < nv) mx0 = fmaxf(mx0, __uint_as_float(v[j]));
              if (j + 1 < nv) mx1 = fmaxf(mx1, __uint_as_float(v[j + 1]));
              if (j + 2 < nv) mx2 = fmaxf(mx2, __uint_as_float(v[j + 2]));
              if (j + 3 < nv) mx3 = fmaxf(mx3, __uint_as_float(v[j + 3]));
            }
          }
          float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
          xch[half * 128 + row] = mx;
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          mx = fmaxf(mx, xch[(half ^ 1) * 128 + row]);
          const float m2 = mx * sl2;
          // pass 2: P = exp2(s * scale * log2e - m2) as bf16 (unnormalised; O is divided by the fp32 row sum), row sum
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
            const bool narrow = c_hi - c0 < 32;
            if (narrow) tmem_ld_32x16(taddr + (uint32_t)c0, v);
            else tmem_ld_32x32(taddr + (uint32_t)c0, v);
            tmem_ld_wait();
            const int nv = nkey - c0;
            uint8_t* dst = pw + (c0 >> 6) * 16384 + row * 128;
            const int ch0 = (c0 & 63) >> 3;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              if (narrow && q4 >= 2) break;
              float e[8];
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const int j = q4 * 8 + jj;
                e[jj] = j < nv ? ex2(fmaf(__uint_as_float(v[j]), sl2, -m2)) : 0.f;
              }
              s0 += e[0] + e[4];
              s1 += e[1] + e[5];
              s2 += e[2] + e[6];
              s3 += e[3] + e[7];
              *reinterpret_cast<uint4*>(dst + (((ch0 + q4) ^ (row & 7)) << 4)) =
                  make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
            }
          }
          float sum = (s0 + s1) + (s2 + s3);
          xch[512 + half * 128 + row] = sum;
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          sum += xch[512 + (half ^ 1) * 128 + row];
          inv = 1.0f / sum;
          if (half == 0 && row < rows_tile) a.lse[stat_base + row] = m2 + log2f(sum);
          if (half == 0 && cls_q) {
            cpart[0] = m2;
            cpart[1] = sum;
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[w]);
        // ---- output: O_w / rowsum -> bf16 -> the 32 rows of the staging tile these two warps share (= P_w sub-tile 0, free
        // once PV retired); this half converts 32 of the 64 columns and stores 16 of the 32 rows
        mbar_wait_sleep(&o_full[w], ph, 32);
        tc_fence_after();
        if (warp_live) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + (uint32_t)(half * 32), v);
          tmem_ld_wait();
          if (cls_q) {   // unnormalised partial output of the CLS query over this frame's keys
#pragma unroll
            for (int j = 0; j < 32; j += 2)   // (part rows are 66 floats: 8-byte alignment)
              *reinterpret_cast<float2*>(cpart + 2 + half * 32 + j) = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t pk[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              pk[jj] = pack_bf16(__uint_as_float(v[q4 * 8 + jj * 2]) * inv, __uint_as_float(v[q4 * 8 + jj * 2 + 1]) * inv);
            *reinterpret_cast<uint4*>(pw + row * 128 + (((half * 4 + q4) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          // 4 rows x 128 B per instruction
          const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = quad * 32 + 4 * (half * 4 + i) + sub;
            if (rl < rows_tile) {
              const uint4 val = *reinterpret_cast<const uint4*>(pw + rl * 128 + ((ch ^ (rl & 7)) << 4));
              *reinterpret_cast<uint4*>(a.o + (o_first + rl) * a.ldo + h * HD + ch * 8) = val;
            }
          }
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // partner's reads of my staging rows are done
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&r_empty[w]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int g_tc_mode = -1;   // env EGV_ATTN_TC: bit 0 = forward (default 1)

}  // namespace atc

void set_tc_attention_mode(int mode) { atc::g_tc_mode = mode; }

// returns 1 when it launched the tcgen05 kernel for this problem, 0 when the problem is not eligible, < 0 on error
int launch_tc_attention(int mode, const AttnP& a, cudaStream_t stream) {
  if (a.fold_out) *a.fold_out = 0;
  using namespace atc;
  if (g_tc_mode < 0) g_tc_mode = getenv("EGV_ATTN_TC") ? atoi(getenv("EGV_ATTN_TC")) : 1;
  if (mode != MODE_FWD || !(g_tc_mode & 1)) return 0;
  if (a.key_bias || a.q_istride != 1 || a.k_istride != 1 || !a.has_cls) return 0;
  const int lk = a.LkT - 1;
  if (a.Lq <= 64 || a.Lq > 256 || lk < 16 || a.LkT > KV_ROWS) return 0;
  if ((a.ldq % 8) || (a.ldkv % 8) || (a.ldo % 8)) return 0;
  Prm pr;
  pr.n_mt = a.Lq > 128 ? 2 : 1;
  pr.rows1 = a.Lq > 128 ? a.Lq - 128 : 0;
  pr.lk = lk;
  pr.npad = (a.LkT + 15) / 16 * 16;
  pr.total = (long long)a.B * a.G * a.H;
  if (pr.total <= 0) return 0;
  static int fold_mode = -1;   // env EGV_ATTN_FOLD_CLS=0 keeps the separate single-query kernels
  if (fold_mode < 0) fold_mode = getenv("EGV_ATTN_FOLD_CLS") ? atoi(getenv("EGV_ATTN_FOLD_CLS")) : 1;
  pr.part = a.cls_part;
  pr.fold = (fold_mode && a.cls_part && pr.n_mt == 2 && pr.rows1 < 128) ? 1 : 0;
  if (a.fold_out) *a.fold_out = pr.fold;
  Maps maps;
  const uint64_t width = (uint64_t)a.H * HD;
  const uint64_t q_rows = (uint64_t)a.B * a.q_bstride, kv_rows = (uint64_t)a.B * a.kv_bstride;
  int rc;
  if ((rc = get_tensor_map(a.q, width, q_rows, a.ldq, 64, (uint32_t)(a.Lq < 128 ? a.Lq : 128), &maps.q0))) return rc;
  if ((rc = get_tensor_map(a.q, width, q_rows, a.ldq, 64, (uint32_t)(pr.rows1 > 0 ? pr.rows1 : 8), &maps.q1))) return rc;
  if ((rc = get_tensor_map(a.k, width, kv_rows, a.ldkv, 64, (uint32_t)lk, &maps.k))) return rc;
  if ((rc = get_tensor_map(a.v, width, kv_rows, a.ldkv, 64, (uint32_t)lk, &maps.v))) return rc;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "tc attention smem attribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  const long long grid = pr.total < sm_count() ? pr.total : sm_count();
  launch_k(attn_tc_fwd_kernel, dim3((unsigned)grid), dim3(THREADS), SMEM_BYTES, stream, maps, a, pr);
  rc = check_launch("attn_tc_fwd_kernel");
  return rc ? rc : 1;
}

}  // namespace egv
