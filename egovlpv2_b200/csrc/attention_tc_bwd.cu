// Backward of the divided SPACE attention (video_transformer.py:35-39, 117-153; the `(b f) n d` groups: 196 patch queries
// of one frame x the frame's 196 keys + the clip's CLS key) on the 5th-generation tensor cores: dQ, dK, dV of one
// (clip, frame, head) problem in ONE pass, all five products on tcgen05.mma with TMEM accumulators.
//
// Everything is computed in the TRANSPOSED (key-row) form, so that the 512 TMEM columns hold a whole problem:
//
//   step (j, w), key tile j in {0, 1} (128 keys each), query tile w in {0, 1} (128 queries each), order (0,0) (0,1) (1,0) (1,1):
//     S^T  = K_j Q_w^T            M = 128 keys, N = queries of the tile, K = 64      -> TMEM columns [  0, 128)
//     dP^T = V_j dO_w^T                                                               -> TMEM columns [128, 256)
//     P^T  = exp2(S^T * scale * log2e - lse[q]),  dS^T = P^T * (dP^T - delta[q])      compute warps: thread = key row; bf16 -> smem
//     dV_j += P^T  dO_w           A = P^T  K-major,  B = dO_w MN-major (the TMA tile as loaded)   -> columns [448, 512)
//     dK_j += dS^T Q_w            A = dS^T K-major,  B = Q_w  MN-major                            -> columns [384, 448)
//     dQ_w += dS   K_j            A = dS^T read as an MN-major operand, B = K_j MN-major          -> columns [256 + 64 w, +64)
//
// i.e. the P^T / dS^T tiles are written once and consumed three times through different descriptor views; the key / query
// operand tiles are consumed both K-major (scores) and MN-major (gradients) as TMA delivered them.  dK_j / dV_j are drained
// one step after they complete (behind the next step's exponentials); `scale` is applied in the drain.
// delta = rowsum(dO * O) comes from the staged dO tiles and the O rows, which TMA fetches a problem ahead into their own buffer.
//
//   warp 0        TMA producer: O rows, Q_w / dO_w, K_j / V_j boxes (128B swizzle), the clip's CLS key / value row appended to
//                 tile 1; each buffer is refilled for the next problem as soon as its last product has retired
//   warp 1        tcgen05.mma issuer (one elected thread): the score products S^T / dP^T (issued one step ahead, as soon as
//                 the previous scores are read out of TMEM) and dQ_w
//   warp 3        second tcgen05.mma issuer: dV_j / dK_j
//   warp 2        TMEM allocator (512 columns)
//   warps 4-19    compute: TMEM lane quadrant (warp % 4) x 32-column quarter; tcgen05.ld -> exp2 -> swizzled bf16 stores;
//                 accumulator drains (bf16 rows of dq / dk / dv; the shared CLS key row goes to the fp32 dkv_cls accumulators)
// CLS query (video_transformer.py:134-150: the CLS token attends every token of the clip).  In the transposed form its
// backward is almost free: its q / dO / O rows are copied into the first unused query slot of tile 1 (196 queries = 128 + 68:
// slot 68 of the 80-wide tile), its log-sum-exp (over ALL keys of the clip, from the forward) into the statistics; the score
// products then deliver S^T[key, 68] = k . q_cls and dP^T[key, 68] = v . dO_cls, the exponentials its global probabilities,
// and the dV / dK products add its contribution to every key row -- no separate pass over K and V (attn_single_bwd_heads:
// 49 us per call).  Its dQ partial (this frame's keys) is accumulated with atomics; the (CLS key, CLS query) pair, which
// every frame's problem sees, is counted in frame 0 only.
// Eligibility (else the mma.sync kernels of attention_group.cu run): contiguous groups with the shared CLS key,
// 128 < queries <= 224, 129 < keys + 1 <= 256, no key bias.
#include <cuda.h>
#include <stdlib.h>

#include "attention.cuh"

namespace egv {
namespace atb {

constexpr int THREADS = 128 + 16 * 32;
constexpr int TILE = 128 * 128;              // one 128-row operand tile, 128 B per row
constexpr int OFF_Q = 0;                     // Q_0, Q_1
constexpr int OFF_DO = 2 * TILE;             // dO_0, dO_1
constexpr int OFF_K = 4 * TILE;              // K_0, K_1
constexpr int OFF_V = 6 * TILE;              // V_0, V_1
constexpr int OFF_P = 8 * TILE;              // P^T  [128 keys][128 queries] as two K-major [128][64] sub-tiles
constexpr int OFF_DS = OFF_P + 2 * TILE;     // dS^T, same layout
constexpr int OFF_STAT = OFF_DS + 2 * TILE;  // 2 x { lse (log2 domain) [256], delta [256] }
constexpr int OFF_O = OFF_STAT + 4096;       // O rows of the next problem (delta = rowsum(dO * O)): 128 + up to 96 rows
constexpr int O_BYTES = TILE + 96 * 128;
constexpr int OFF_BAR = OFF_O + O_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
static_assert(SMEM_BYTES <= 232448, "space attention backward tile set exceeds the 227 KB shared memory of an SM");

constexpr uint32_t COL_ST = 0, COL_DPT = 128, COL_DQ = 256, COL_DK = 384, COL_DV = 448;

struct Maps {
  CUtensorMap q0, q1, do0, do1, k0, k1, v0, v1, o0, o1;
};
struct Prm {
  int rows1;       // query rows of tile 1
  int keys1;       // regular key rows of tile 1 (the CLS key is row keys1 of tile 1)
  int nq0, nq1;    // query counts rounded up to 16 (N of the score products, K of the dK / dV products)
  int nk1;         // keys1 + 1 rounded up to 16 (K of the dQ product over tile 1)
  long long total;
  uint32_t smem_base;   // shared-window address of the 1024-aligned tile area, probed once by the host (see launch): the MMA
                        // issuer builds every operand descriptor from this kernel PARAMETER, i.e. in uniform registers
  uint32_t* probe;      // non-null: write that address here and exit
  int fold;             // 1: the clip's CLS QUERY rides along as query slot `rows1` of tile 1 (see "CLS query" below)
  int trace;       // debugging: EGV_ATB_TRACE=1 records SM clock stamps of the third problem of CTA 0 (printed by the host)
};

__device__ long long g_trace[64];
#define ATB_TR(cond, i)                                                   \
  do {                                                                    \
    if (pr.trace && blockIdx.x == 0 && (cond)) g_trace[i] = clock64();    \
  } while (0)

EGV_DEVINL void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// non-blocking phase test (the MMA issuer polls two independent conditions)
EGV_DEVINL bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// global load that the compiler may not sink below later code (the lse values are fetched a phase ahead of their use; an
// invariant __ldg load gets moved down to its first use)
// 32-byte global store (STG.256): the accumulator drains write one key / query row per lane, i.e. every store instruction
// costs 32 memory transactions whatever its width -- half as many instructions as with 16-byte stores
EGV_DEVINL void stg256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
EGV_DEVINL float ldg_now_f32(const float* p) {
  float v;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// tcgen05.mma with the 64-bit descriptors given as (lo, hi) words: lo = (address >> 4) | (LBO >> 4) << 16 varies, hi is a
// constant -- everything stays 32-bit uniform arithmetic.  Called from a region guarded by elect.sync (elect_one()):
// ptxas then knows a single thread is active and emits bare UTCHMMA instructions; guarded by `lane == 0` instead, every
// MMA is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY convergence loop (~15 instructions, ~120 clocks per MMA).
EGV_DEVINL void umma_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);     // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
EGV_DEVINL constexpr uint32_t desc_lo(uint32_t addr, uint32_t lbo_bytes) { return ((addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); }

// The MMA issue sequences of one step, straight-line: t (hence j, w and every operand offset) is a template constant and
// `sb` (the tile area's shared address) comes from a kernel parameter, so each tcgen05.mma costs a few uniform-datapath
// instructions.  Every product runs over the full 128-deep K: rows / columns behind the group are zero operands.
template <int T>
EGV_DEVINL void issue_scores_t(uint32_t tmem_base, uint32_t sb) {
  constexpr int j = T >> 1, w = T & 1;
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  const uint32_t kd = desc_lo(sb + OFF_K + j * TILE, 16), qd = desc_lo(sb + OFF_Q + w * TILE, 16);
  const uint32_t vd = desc_lo(sb + OFF_V + j * TILE, 16), dod = desc_lo(sb + OFF_DO + w * TILE, 16);
#pragma unroll
  for (int k = 0; k < HD / 16; ++k) umma_bf16_w(tmem_base + COL_ST, kd + k * 2, DESC_HI, qd + k * 2, DESC_HI, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < HD / 16; ++k) umma_bf16_w(tmem_base + COL_DPT, vd + k * 2, DESC_HI, dod + k * 2, DESC_HI, idesc_s, k > 0 ? 1u : 0u);
}
template <int T>
EGV_DEVINL void issue_dvk_t(uint32_t tmem_base, uint32_t sb) {      // dV_j, dK_j: K = the tile's queries
  constexpr int w = T & 1;
  constexpr uint32_t idesc_g = umma_idesc_bf16(128, HD, 0, 1);    // A K-major, B MN-major
  const uint32_t d_p = desc_lo(sb + OFF_P, 16), d_ds = desc_lo(sb + OFF_DS, 16);
  const uint32_t dob = desc_lo(sb + OFF_DO + w * TILE, 8192), qb = desc_lo(sb + OFF_Q + w * TILE, 8192);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk)
    umma_bf16_w(tmem_base + COL_DV, d_p + (((kk >> 2) * TILE + (kk & 3) * 32) >> 4), DESC_HI, dob + kk * 128, DESC_HI, idesc_g,
                (w > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk)
    umma_bf16_w(tmem_base + COL_DK, d_ds + (((kk >> 2) * TILE + (kk & 3) * 32) >> 4), DESC_HI, qb + kk * 128, DESC_HI, idesc_g,
                (w > 0 || kk > 0) ? 1u : 0u);
}
template <int T>
EGV_DEVINL void issue_dq_t(uint32_t tmem_base, uint32_t sb) {       // dQ_w: K = the tile's keys
  constexpr int j = T >> 1, w = T & 1;
  constexpr uint32_t idesc_q = umma_idesc_bf16(128, HD, 1, 1);    // both MN-major
  const uint32_t kb = desc_lo(sb + OFF_K + j * TILE, 8192), m_ds = desc_lo(sb + OFF_DS, TILE);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk)
    umma_bf16_w(tmem_base + COL_DQ + (uint32_t)(w * HD), m_ds + kk * 128, DESC_HI, kb + kk * 128, DESC_HI, idesc_q,
                (j > 0 || kk > 0) ? 1u : 0u);
}

EGV_DEVINL void compute_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__global__ void __launch_bounds__(THREADS, 1) attn_tc_bwd_kernel(const __grid_constant__ Maps maps, const AttnP a, const Prm pr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* qdo_full = bars;         // [2]  Q_w + dO_w landed
  uint64_t* qdo_empty = bars + 2;    // [2]
  uint64_t* kv_full = bars + 4;      // [2]  K_j + V_j landed (tile 1: + the CLS row copy)
  uint64_t* kv_empty = bars + 6;     // [2]
  uint64_t* s_full = bars + 8;       // S^T / dP^T of the step in TMEM
  uint64_t* s_free = bars + 9;       // ... read back by all 16 compute warps
  uint64_t* p_ready = bars + 10;     // P^T / dS^T of the step in shared memory
  uint64_t* p_free = bars + 11;      // ... consumed by the three gradient products
  uint64_t* dkv_full = bars + 12;    // dK_j / dV_j complete
  uint64_t* dkv_free = bars + 13;
  uint64_t* dq_full = bars + 14;     // [2]  dQ_w complete
  uint64_t* dq_free = bars + 16;     // [2]
  uint64_t* o_full = bars + 18;      // the O rows of the problem staged in their own buffer
  uint64_t* o_free = bars + 19;      // ... the statistics of the previous problem are computed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (pr.probe) {   // host probe of the shared-window address of the tile area (constant per kernel binary)
    if (threadIdx.x == 0) *pr.probe = smem_u32(smem);
    return;
  }
  if (smem_u32(smem) != pr.smem_base) {
    if (threadIdx.x == 0) printf("egv: attention backward: shared memory base %u differs from the probed %u\n", smem_u32(smem), pr.smem_base);
    __trap();
  }
  // rows the TMA boxes never touch (the tails of query tile 1 and key tile 1) must hold finite values
  for (int i = threadIdx.x; i < OFF_STAT / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q0);
    tma_prefetch_desc(&maps.q1);
    tma_prefetch_desc(&maps.do0);
    tma_prefetch_desc(&maps.do1);
    tma_prefetch_desc(&maps.k0);
    tma_prefetch_desc(&maps.k1);
    tma_prefetch_desc(&maps.v0);
    tma_prefetch_desc(&maps.v1);
    tma_prefetch_desc(&maps.o0);
    tma_prefetch_desc(&maps.o1);
  }
  if (warp == 1 && lane == 0) {
    for (int w = 0; w < 2; ++w) {
      mbar_init(&qdo_full[w], w == 1 && pr.fold ? 2 : 1);   // tile 1: + the CLS query's row copies
      mbar_init(&qdo_empty[w], 2);   // both MMA issuers read Q_w / dO_w
      mbar_init(&kv_empty[w], 1);
      mbar_init(&dq_full[w], 1);
      mbar_init(&dq_free[w], 16);
    }
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 2);   // TMA transaction + the CLS row copies
    mbar_init(s_full, 1);
    mbar_init(s_free, 16);
    mbar_init(p_ready, 16);
    mbar_init(p_free, 2);          // both MMA issuers read P^T / dS^T
    mbar_init(dkv_full, 1);
    mbar_init(dkv_free, 16);
    mbar_init(o_full, pr.fold ? 2 : 1);
    mbar_init(o_free, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int HG = a.H * a.G;
  pdl_wait();     // programmatic dependent launch (common.cuh): before any global access

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    uint32_t ph = 0;
    int pit = 0;
    for (long long p = blockIdx.x; p < pr.total; p += gridDim.x, ph ^= 1, ++pit) {
      const bool ptr = pit == 2 && lane == 0;
      const int h = (int)(p % a.H), g = (int)((p / a.H) % a.G), b = (int)(p / HG);
      const int q_first = (int)((long long)b * a.q_bstride + a.q_row0 + (long long)g * a.q_gstride);
      const int o_first = (int)((long long)b * a.o_bstride + a.q_row0 + (long long)g * a.q_gstride);
      const int k_first = (int)((long long)b * a.kv_bstride + a.k_row0 + (long long)g * a.k_gstride);
      const long long cls_off = ((long long)b * a.kv_bstride + a.cls_row) * a.ldkv + h * HD + (lane & 7) * 8;
      // O rows of this problem (for delta = rowsum(dO * O)): their buffer is free once the previous problem's statistics
      // are computed, i.e. a whole problem ahead
      mbar_wait_sleep(o_free, ph ^ 1, 64);
      ATB_TR(ptr, 53);
      if (lane == 0) {
        mbar_arrive_expect_tx(o_full, (uint32_t)a.Lq * 128u);
        tma_load_2d(smem + OFF_O, &maps.o0, o_full, h * HD, o_first);
        tma_load_2d(smem + OFF_O + TILE, &maps.o1, o_full, h * HD, o_first + 128);
      }
      const long long qcls_off = ((long long)b * a.q_bstride + a.cls_row) * a.ldq + h * HD + (lane & 7) * 8;
      const long long ocls_off = ((long long)b * a.o_bstride + a.cls_row) * a.ldo + h * HD + (lane & 7) * 8;
      const int cls_slot = pr.rows1 * 128 + (((lane & 7) ^ (pr.rows1 & 7)) << 4);   // query slot rows1 of tile 1, swizzled like a TMA row
      if (pr.fold) {
        if (lane < 8) *reinterpret_cast<uint4*>(smem + OFF_O + TILE + cls_slot) = *reinterpret_cast<const uint4*>(a.o + ocls_off);
        __syncwarp();
        if (lane == 0) mbar_arrive(o_full);
      }
      // order of first use: K_0 V_0 Q_0 dO_0 (step (0,0)), Q_1 dO_1 (step (0,1)), K_1 V_1 (step (1,0))
      mbar_wait_sleep(&kv_empty[0], ph ^ 1, 64);
      ATB_TR(ptr, 48);
      if (lane == 0) {
        mbar_arrive_expect_tx(&kv_full[0], 2u * 128u * 128u);
        tma_load_2d(smem + OFF_K, &maps.k0, &kv_full[0], h * HD, k_first);
        tma_load_2d(smem + OFF_V, &maps.v0, &kv_full[0], h * HD, k_first);
      }
      for (int w = 0; w < 2; ++w) {
        mbar_wait_sleep(&qdo_empty[w], ph ^ 1, 64);
        ATB_TR(ptr, 49 + w);
        if (lane == 0) {
          const int rows = w == 0 ? 128 : pr.rows1;
          mbar_arrive_expect_tx(&qdo_full[w], 2u * (uint32_t)rows * 128u);
          tma_load_2d(smem + OFF_Q + w * TILE, w == 0 ? &maps.q0 : &maps.q1, &qdo_full[w], h * HD, q_first + 128 * w);
          tma_load_2d(smem + OFF_DO + w * TILE, w == 0 ? &maps.do0 : &maps.do1, &qdo_full[w], h * HD, o_first + 128 * w);
        }
        if (w == 1 && pr.fold) {   // the CLS query's q and dO rows
          if (lane < 16) {
            const bool is_q = lane < 8;
            const uint4 val = *reinterpret_cast<const uint4*>(is_q ? a.q + qcls_off : a.d_o + ocls_off);
            *reinterpret_cast<uint4*>(smem + (is_q ? OFF_Q : OFF_DO) + TILE + cls_slot) = val;
          }
          fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's (async proxy) reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&qdo_full[1]);
        }
      }
      mbar_wait_sleep(&kv_empty[1], ph ^ 1, 64);
      ATB_TR(ptr, 51);
      if (lane == 0) {
        mbar_arrive_expect_tx(&kv_full[1], 2u * (uint32_t)pr.keys1 * 128u);
        tma_load_2d(smem + OFF_K + TILE, &maps.k1, &kv_full[1], h * HD, k_first + 128);
        tma_load_2d(smem + OFF_V + TILE, &maps.v1, &kv_full[1], h * HD, k_first + 128);
      }
      if (lane < 16) {   // the clip's CLS key / value: row keys1 of tile 1, 128B-swizzled like the TMA rows
        const int t = lane >> 3;
        const uint4 val = *reinterpret_cast<const uint4*>((t == 0 ? a.k : a.v) + cls_off);
        uint8_t* dst = smem + (t == 0 ? OFF_K : OFF_V) + TILE;
        *reinterpret_cast<uint4*>(dst + pr.keys1 * 128 + (((lane & 7) ^ (pr.keys1 & 7)) << 4)) = val;
      }
      fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's (async proxy) reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&kv_full[1]);
      ATB_TR(ptr, 52);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Flat pipeline over the steps n = 4 * problem + t: the score products of step n + 1 are issued as soon as the compute
    // warps have read step n's scores out of TMEM (s_free), i.e. they run under step n's exponentials; the gradient products
    // of step n follow when its P^T / dS^T tiles are in shared memory (p_ready).
    // The whole warp runs the control flow (warp votes keep every branch and every descriptor provably uniform, so the
    // compiler keeps them in uniform registers); only the elected lane issues tcgen05.mma / tcgen05.commit.  Issuing from
    // inside an `if (lane == 0)` region instead costs a 12-instruction ELECT / R2UR.BROADCAST loop per descriptor operand:
    // 130 clocks per MMA, more than the 32 clocks a 128 x 64 x 16 MMA takes.
    {
      const bool leader = elect_one();   // elect.sync: ptxas knows a single thread runs the regions it guards
      const long long nprob = pr.total > (long long)blockIdx.x ? (pr.total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const long long nsteps = 4 * nprob;
      auto scores_ready = [&](long long n) -> bool {
        const int t = (int)(n & 3), j = t >> 1, w = t & 1;
        const uint32_t ph = (uint32_t)((n >> 2) & 1);
        if (w == 0 && !mbar_test(&kv_full[j], ph)) return false;
        if (j == 0 && !mbar_test(&qdo_full[w], ph)) return false;
        return mbar_test(s_free, (uint32_t)(w ^ 1));
      };
      const uint32_t sb = pr.smem_base;
      auto issue_scores = [&](long long n) {
        tc_fence_after();
        switch ((int)(n & 3)) {
          case 0: if (leader) issue_scores_t<0>(tmem_base, sb); break;
          case 1: if (leader) issue_scores_t<1>(tmem_base, sb); break;
          case 2: if (leader) issue_scores_t<2>(tmem_base, sb); break;
          default: if (leader) issue_scores_t<3>(tmem_base, sb); break;
        }
        if (leader) umma_commit(s_full);
        __syncwarp();
        ATB_TR((n >> 2) == 2, (int)(n & 3));
      };
      auto grads_ready = [&](long long n) -> bool {
        const int t = (int)(n & 3), j = t >> 1, w = t & 1;
        const uint32_t ph = (uint32_t)((n >> 2) & 1);
        if (!mbar_test(p_ready, (uint32_t)w)) return false;
        if (j == 0 && !mbar_test(&dq_free[w], ph ^ 1)) return false;
        return true;
      };
      auto issue_grads = [&](long long n) {
        const int t = (int)(n & 3), j = t >> 1, w = t & 1;
        tc_fence_after();
        switch (t) {
          case 0: if (leader) issue_dq_t<0>(tmem_base, sb); break;
          case 1: if (leader) issue_dq_t<1>(tmem_base, sb); break;
          case 2: if (leader) issue_dq_t<2>(tmem_base, sb); break;
          default: if (leader) issue_dq_t<3>(tmem_base, sb); break;
        }
        if (leader) {
          umma_commit(p_free);
          if (j == 1) umma_commit(&dq_full[w]);
          if (t == 1) umma_commit(&kv_empty[0]);
          if (t == 2) umma_commit(&qdo_empty[0]);
          if (t == 3) {
            umma_commit(&qdo_empty[1]);
            umma_commit(&kv_empty[1]);
          }
        }
        __syncwarp();
        ATB_TR((n >> 2) == 2, 8 + (int)(n & 3));
      };
      // whichever is ready first: the score products of step n_s (operands landed, TMEM columns read back) or the gradient
      // products of step n_g (P^T / dS^T written, accumulators drained); scores run at most one step ahead
      long long n_s = 0, n_g = 0;
      uint32_t idle = 0;
      while (n_g < nsteps) {
        if (__all_sync(0xffffffffu, n_s < nsteps && n_s <= n_g + 1 && scores_ready(n_s))) {
          issue_scores(n_s++);
          idle = 0;
        } else if (__all_sync(0xffffffffu, n_g < n_s && grads_ready(n_g))) {
          issue_grads(n_g++);
          idle = 0;
        } else {
          __nanosleep(20);
          if (++idle > (1u << 24)) {
            if (leader) printf("egv: attention backward MMA issuer timeout block %d\n", blockIdx.x);
            __trap();
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ second MMA issuer: dV_j / dK_j of every step
    // (a single thread sustains one tcgen05.mma per ~100 clocks here: the 32 MMAs of a step are split over two issuers)
    if (elect_one()) {
      const uint32_t sb = pr.smem_base;
      const long long nprob = pr.total > (long long)blockIdx.x ? (pr.total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      for (long long n = 0; n < 4 * nprob; ++n) {
        const int t = (int)(n & 3), j = t >> 1, w = t & 1;
        mbar_wait_sleep(p_ready, (uint32_t)w, 20);
        if (w == 0) mbar_wait_sleep(dkv_free, (uint32_t)(j ^ 1), 20);
        tc_fence_after();
        switch (t) {
          case 0: issue_dvk_t<0>(tmem_base, sb); break;
          case 1: issue_dvk_t<1>(tmem_base, sb); break;
          case 2: issue_dvk_t<2>(tmem_base, sb); break;
          default: issue_dvk_t<3>(tmem_base, sb); break;
        }
        umma_commit(p_free);
        if (w == 1) umma_commit(dkv_full);
        if (t == 2) umma_commit(&qdo_empty[0]);
        if (t == 3) umma_commit(&qdo_empty[1]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ compute warps
    const int cw = warp - 4;
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read (hardware: warp id % 4)
    const int cq = cw >> 2;                 // 32-column quarter
    const int row = quad * 32 + lane;       // key row of the tile = TMEM lane
    const int ct = threadIdx.x - 128;       // 0..511
    const float sl2 = a.scale * LOG2E;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);

    // dK_j / dV_j: 128 key rows x 64 columns each; this warp: its 32 rows x 32 columns of dK (cq 0, 1) or dV (cq 2, 3)
    bool trd = false;
    auto drain_dkv = [&](int j, int b, int h, long long k_first) {
      mbar_wait_sleep(dkv_full, (uint32_t)j, 32);
      ATB_TR(trd, 56 + j);
      tc_fence_after();
      uint32_t acc[2][16];
      const uint32_t col = (cq < 2 ? COL_DK : COL_DV) + (uint32_t)((cq & 1) * 32);
      tmem_ld_32x16(lane_addr + col, acc[0]);
      tmem_ld_32x16(lane_addr + col + 16u, acc[1]);
      tmem_ld_wait();
      ATB_TR(trd, 58 + j);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dkv_free);
      const float mul = cq < 2 ? a.scale : 1.0f;
      const int nkeys = j == 0 ? 128 : pr.keys1 + 1;      // valid key rows of the tile (tile 1: + the CLS key)
      if (j == 1 && row == pr.keys1) {
        // the clip's CLS key is shared by every frame: fp32 accumulators, finalised by attn_cls_finalize_kernel
        float* dst = a.dkv_cls + ((long long)b * a.H + h) * 2 * HD + (cq < 2 ? 0 : HD) + (cq & 1) * 32;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(dst + 16 * hh + i, __uint_as_float(acc[hh][i]) * mul);
      } else if (row < nkeys) {
        bf16* dst = (cq < 2 ? a.dk : a.dv) + (k_first + 128 * j + row) * a.lddkv + h * HD + (cq & 1) * 32;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) pk[e] = pack_bf16(__uint_as_float(acc[hh][2 * e]) * mul, __uint_as_float(acc[hh][2 * e + 1]) * mul);
          stg256(dst + 16 * hh, pk);
        }
      }
    };
    // dQ_w: 128 query rows x 64 columns; this warp: its 32 rows x 16 columns
    auto drain_dq = [&](int w, uint32_t phase, int b, int h, long long q_first) {
      mbar_wait_sleep(&dq_full[w], phase, 32);
      tc_fence_after();
      uint32_t acc[16];
      tmem_ld_32x16(lane_addr + COL_DQ + (uint32_t)(w * HD + cq * 16), acc);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dq_free[w]);
      const int rows_tile = w == 0 ? 128 : pr.rows1;
      if (row < rows_tile) {
        bf16* dst = a.dq + (q_first + 128 * w + row) * a.lddq + h * HD + cq * 16;
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) pk[e] = pack_bf16(__uint_as_float(acc[2 * e]) * a.scale, __uint_as_float(acc[2 * e + 1]) * a.scale);
        stg256(dst, pk);
      } else if (pr.fold && w == 1 && row == pr.rows1) {   // the CLS query's dQ: this frame's keys' share
        float* dst = a.dq_cls + ((long long)b * a.H + h) * HD + cq * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(dst + i, __uint_as_float(acc[i]) * a.scale);
      }
    };

    // Accumulators are drained one step late, when their products have long retired: dK_0 / dV_0 (complete after step 1)
    // behind step 2's exponentials, dQ_0 (after step 2) behind step 3's, dK_1 / dV_1 / dQ_1 (after step 3) behind step 0
    // of the next problem.
    uint32_t ph = 0;
    int it = 0;
    int pb = 0, phd = 0;
    long long pq_first = 0, pk_first = 0;
    bool pending = false;
    for (long long p = blockIdx.x; p < pr.total; p += gridDim.x, ph ^= 1, ++it) {
      const bool tr = it == 2 && threadIdx.x == 128;
      trd = tr;
      ATB_TR(tr, 40);
      const int h = (int)(p % a.H), g = (int)((p / a.H) % a.G), b = (int)(p / HG);
      const long long q_first = (long long)b * a.q_bstride + a.q_row0 + (long long)g * a.q_gstride;
      const long long o_first = (long long)b * a.o_bstride + a.q_row0 + (long long)g * a.q_gstride;
      const long long k_first = (long long)b * a.kv_bstride + a.k_row0 + (long long)g * a.k_gstride;
      const long long stat_base = (((long long)b * a.H + h) * a.G + g) * a.Lq;
      float* s_lse = reinterpret_cast<float*>(smem + OFF_STAT) + (it & 1) * 512;   // double-buffered: one barrier per problem
      float* s_delta = s_lse + 256;
      // ---- lse and delta = rowsum(dO * O) of the 256 query slots (threads 0..255: one query row each) from the staged O and
      // dO rows.  Query tile 0 before step 0; tile 1 (whose dO rows land last) before step 1, the first step that reads it.
      ATB_TR(tr, 41);
      float lse_r = 1.0e30f;                // slots beyond the group: P = exp2(-huge) = 0
      const bool cls_slot_thread = pr.fold && ct == a.Lq;             // query slot rows1 of tile 1 carries the CLS query
      if (ct < a.Lq) lse_r = ldg_now_f32(a.lse + stat_base + ct);     // (Lq <= 224)
      else if (cls_slot_thread) lse_r = ldg_now_f32(a.lse_cls + (long long)b * a.H + h);
      auto stats_rows = [&](int tile) {
        if ((ct >> 7) == tile) {
          float dl = 0.f;
          if (ct < a.Lq || cls_slot_thread) {
            mbar_wait_sleep(&qdo_full[tile], ph, 32);
            mbar_wait_sleep(o_full, ph, 32);
            ATB_TR(tr, 43);
            const uint8_t* drow = smem + OFF_DO + tile * TILE + (ct & 127) * 128;
            const uint8_t* orow = smem + OFF_O + tile * TILE + (ct & 127) * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 dv = *reinterpret_cast<const uint4*>(drow + ((c ^ (ct & 7)) << 4));
              const uint4 ov = *reinterpret_cast<const uint4*>(orow + ((c ^ (ct & 7)) << 4));
              const uint32_t* po = reinterpret_cast<const uint32_t*>(&ov);
              const uint32_t* pd = reinterpret_cast<const uint32_t*>(&dv);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 x = unpack_bf16(po[e]), y = unpack_bf16(pd[e]);
                dl = fmaf(x.x, y.x, dl);
                dl = fmaf(x.y, y.y, dl);
              }
            }
            if (a.delta && ct < a.Lq) a.delta[stat_base + ct] = dl;
          }
          s_lse[ct] = lse_r;
          s_delta[ct] = dl;
        }
      };
      stats_rows(0);
      ATB_TR(tr, 44);
      compute_sync();
      ATB_TR(tr, 42);
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        const int w = t & 1;
        const int nq = w == 0 ? pr.nq0 : pr.nq1;
        if (t == 1) {
          stats_rows(1);
          compute_sync();
          if (threadIdx.x == 128) mbar_arrive(o_free);      // every O row of this problem is consumed
        }
        const int c0 = cq * 32;                             // this warp's query columns [c0, c0 + 32) of the tile
        const int nh = c0 >= nq ? 0 : (nq - c0 >= 32 ? 2 : 1);   // 16-column halves this warp owns in this step
        mbar_wait_sleep(s_full, (uint32_t)w, 32);
        ATB_TR(tr, 16 + 4 * t);
        tc_fence_after();
        // Rows of key tile 1 behind the CLS key need no masking: their K / V rows are zero, so S^T = dP^T = 0 there, P^T and
        // dS^T are finite, the dQ product multiplies them with zero K rows and the dK / dV rows are never stored.
        // Query slots behind the group carry lse = 1e30 (P = 0) and delta = 0.
        // The packed results wait in registers: the previous step's gradient products may still be reading the tiles.
        uint32_t pk[2][16];
        if (nh == 0) {                                      // no columns of this step: release the score columns right away
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (hh < nh) {
            const int cb = c0 + 16 * hh;                    // 16 columns = two 16-byte chunks of the row
            uint32_t sv[16], dv[16];
            tmem_ld_32x16(lane_addr + COL_ST + (uint32_t)cb, sv);
            tmem_ld_32x16(lane_addr + COL_DPT + (uint32_t)cb, dv);
            tmem_ld_wait();
            ATB_TR(tr && hh == 0, 18 + 4 * t);
            if (hh + 1 == nh) {                             // last read of this step's scores: the columns may be overwritten
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(s_free);
            }
            const float4* lp = reinterpret_cast<const float4*>(s_lse + 128 * w + cb);
            const float4* dp = reinterpret_cast<const float4*>(s_delta + 128 * w + cb);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 l4 = lp[e], d4 = dp[e];
              const int i0 = 4 * e;
              const float p0 = ex2(fmaf(__uint_as_float(sv[i0]), sl2, -l4.x));
              const float p1 = ex2(fmaf(__uint_as_float(sv[i0 + 1]), sl2, -l4.y));
              const float p2 = ex2(fmaf(__uint_as_float(sv[i0 + 2]), sl2, -l4.z));
              const float p3 = ex2(fmaf(__uint_as_float(sv[i0 + 3]), sl2, -l4.w));
              pk[hh][2 * e] = pack_bf16(p0, p1);
              pk[hh][2 * e + 1] = pack_bf16(p2, p3);
              pk[hh][8 + 2 * e] = pack_bf16(p0 * (__uint_as_float(dv[i0]) - d4.x), p1 * (__uint_as_float(dv[i0 + 1]) - d4.y));
              pk[hh][8 + 2 * e + 1] = pack_bf16(p2 * (__uint_as_float(dv[i0 + 2]) - d4.z), p3 * (__uint_as_float(dv[i0 + 3]) - d4.w));
            }
          }
        }
        mbar_wait_sleep(p_free, (uint32_t)(w ^ 1), 32);     // the previous step's gradient products have read P^T / dS^T
        ATB_TR(tr, 17 + 4 * t);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (hh < nh) {
            const int cb = c0 + 16 * hh;
            uint8_t* prow = smem + OFF_P + (cb >> 6) * TILE + row * 128;
            const int ch0 = (cb & 63) >> 3;
#pragma unroll
            for (int c8 = 0; c8 < 2; ++c8) {
              const int ch = ((ch0 + c8) ^ (row & 7)) << 4;
              *reinterpret_cast<uint4*>(prow + ch) = make_uint4(pk[hh][4 * c8], pk[hh][4 * c8 + 1], pk[hh][4 * c8 + 2], pk[hh][4 * c8 + 3]);
              *reinterpret_cast<uint4*>(prow + (OFF_DS - OFF_P) + ch) =
                  make_uint4(pk[hh][8 + 4 * c8], pk[hh][8 + 4 * c8 + 1], pk[hh][8 + 4 * c8 + 2], pk[hh][8 + 4 * c8 + 3]);
            }
          }
        }
        if (pr.fold && t == 3 && g != 0 && row == pr.keys1 && pr.rows1 >= c0 && pr.rows1 < c0 + 32) {
          // (CLS key, CLS query): every frame's problem sees the pair; it is counted in frame 0 only
          const int c = pr.rows1;
          uint8_t* el = smem + OFF_P + (c >> 6) * TILE + row * 128 + ((((c & 63) >> 3) ^ (row & 7)) << 4) + (c & 7) * 2;
          *reinterpret_cast<uint16_t*>(el) = 0;
          *reinterpret_cast<uint16_t*>(el + (OFF_DS - OFF_P)) = 0;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        ATB_TR(tr, 19 + 4 * t);
        // ---- deferred drains (see above)
        if (t == 0 && pending) {
          drain_dkv(1, pb, phd, pk_first);
          ATB_TR(tr, 60);
          drain_dq(1, ph ^ 1, pb, phd, pq_first);
          ATB_TR(tr, 61);
        }
        if (t == 2) {
          drain_dkv(0, b, h, k_first);
          ATB_TR(tr, 33);
        }
        if (t == 3) {
          drain_dq(0, ph, b, h, q_first);
          ATB_TR(tr, 38);
        }
      }
      pending = true;
      pb = b;
      phd = h;
      pq_first = q_first;
      pk_first = k_first;
    }
    if (pending) {
      drain_dkv(1, pb, phd, pk_first);
      drain_dq(1, ph ^ 1, pb, phd, pq_first);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace atb

// returns 1 when it launched the tcgen05 backward for this problem, 0 when the problem is not eligible, < 0 on error
int launch_tc_attention_bwd(const AttnP& a, cudaStream_t stream, int* cls_query_folded) {
  using namespace atb;
  static int mode = -1;   // env EGV_ATTN_TC: bit 1 = backward (default on)
  if (mode < 0) mode = getenv("EGV_ATTN_TC") ? atoi(getenv("EGV_ATTN_TC")) : 3;
  if (!(mode & 2)) return 0;
  if (a.key_bias || a.q_istride != 1 || a.k_istride != 1 || !a.has_cls || a.dkv_accumulate || !a.dkv_cls) return 0;
  const int lk = a.LkT - 1;
  if (a.Lq <= 128 || a.Lq > 224 || lk <= 128 || a.LkT > 256) return 0;   // (the O staging buffer holds 128 + 96 rows)
  if ((a.ldq % 8) || (a.ldkv % 8) || (a.ldo % 8) || (a.lddq % 16) || (a.lddkv % 16)) return 0;
  if (((uintptr_t)a.dq | (uintptr_t)a.dk | (uintptr_t)a.dv) & 31) return 0;   // 32-byte row stores
  Prm pr;
  pr.rows1 = a.Lq - 128;
  pr.keys1 = lk - 128;
  pr.nq0 = 128;
  pr.nq1 = (pr.rows1 + 15) / 16 * 16;
  pr.nk1 = (pr.keys1 + 1 + 15) / 16 * 16;
  pr.total = (long long)a.B * a.G * a.H;
  if (pr.total <= 0) return 0;
  // the CLS query takes the first unused query slot of tile 1, if there is one (and the O staging buffer has that row)
  static int fold_mode = -1;   // env EGV_ATTN_FOLD_CLS=0 keeps the separate single-query backward
  if (fold_mode < 0) fold_mode = getenv("EGV_ATTN_FOLD_CLS") ? atoi(getenv("EGV_ATTN_FOLD_CLS")) : 1;
  pr.fold = (fold_mode && a.lse_cls && a.dq_cls && pr.rows1 < pr.nq1 && pr.rows1 < 96) ? 1 : 0;
  if (cls_query_folded) *cls_query_folded = pr.fold;
  Maps maps;
  const uint64_t width = (uint64_t)a.H * HD;
  const uint64_t q_rows = (uint64_t)a.B * a.q_bstride, kv_rows = (uint64_t)a.B * a.kv_bstride, o_rows = (uint64_t)a.B * a.o_bstride;
  int rc;
  if ((rc = get_tensor_map(a.q, width, q_rows, a.ldq, 64, 128u, &maps.q0))) return rc;
  if ((rc = get_tensor_map(a.q, width, q_rows, a.ldq, 64, (uint32_t)pr.rows1, &maps.q1))) return rc;
  if ((rc = get_tensor_map(a.d_o, width, o_rows, a.ldo, 64, 128u, &maps.do0))) return rc;
  if ((rc = get_tensor_map(a.d_o, width, o_rows, a.ldo, 64, (uint32_t)pr.rows1, &maps.do1))) return rc;
  if ((rc = get_tensor_map(a.k, width, kv_rows, a.ldkv, 64, 128u, &maps.k0))) return rc;
  if ((rc = get_tensor_map(a.k, width, kv_rows, a.ldkv, 64, (uint32_t)pr.keys1, &maps.k1))) return rc;
  if ((rc = get_tensor_map(a.v, width, kv_rows, a.ldkv, 64, 128u, &maps.v0))) return rc;
  if ((rc = get_tensor_map(a.v, width, kv_rows, a.ldkv, 64, (uint32_t)pr.keys1, &maps.v1))) return rc;
  if ((rc = get_tensor_map(a.o, width, o_rows, a.ldo, 64, 128u, &maps.o0))) return rc;
  if ((rc = get_tensor_map(a.o, width, o_rows, a.ldo, 64, (uint32_t)pr.rows1, &maps.o1))) return rc;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "tc attention backward smem attribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  const long long grid = pr.total < sm_count() ? pr.total : sm_count();
  static uint32_t smem_base = 0;
  if (smem_base == 0) {   // one-time probe (first call; not under stream capture): the kernel reports where its tile area starts
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cs);
    if (cs != cudaStreamCaptureStatusNone) return 0;
    uint32_t* dev = nullptr;
    if (cudaMalloc(&dev, sizeof(uint32_t)) != cudaSuccess) return fail(EGV_ERR_CUDA, "tc attention backward: probe allocation failed");
    Prm pp = pr;
    pp.probe = dev;
    pp.smem_base = 0;
    pp.trace = 0;
    cudaError_t e = launch_k(attn_tc_bwd_kernel, dim3(1), dim3(THREADS), SMEM_BYTES, stream, maps, a, pp);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaMemcpy(&smem_base, dev, sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (e != cudaSuccess || smem_base == 0) return fail(EGV_ERR_CUDA, "tc attention backward: shared memory probe failed: %s", cudaGetErrorString(e));
  }
  pr.smem_base = smem_base;
  pr.probe = nullptr;
  static int trace = -1;
  if (trace < 0) trace = getenv("EGV_ATB_TRACE") ? atoi(getenv("EGV_ATB_TRACE")) : 0;
  pr.trace = trace;
  launch_k(attn_tc_bwd_kernel, dim3((unsigned)grid), dim3(THREADS), SMEM_BYTES, stream, maps, a, pr);
  rc = check_launch("attn_tc_bwd_kernel");
  if (trace) {   // debugging aid: SM-clock stamps of the third problem of CTA 0, relative to its start
    long long t[64];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(t, g_trace, sizeof(t));
    const char* names[64] = {};
    names[0] = "S(0) issued"; names[1] = "S(1) issued"; names[2] = "S(2) issued"; names[3] = "S(3) issued";
    names[8] = "G(0) issued"; names[9] = "G(1) issued"; names[10] = "G(2) issued"; names[11] = "G(3) issued";
    for (int s = 0; s < 4; ++s) {
      static char buf[16][32];
      snprintf(buf[4 * s], 32, "step %d s_full", s); names[16 + 4 * s] = buf[4 * s];
      snprintf(buf[4 * s + 1], 32, "step %d p_free", s); names[17 + 4 * s] = buf[4 * s + 1];
      snprintf(buf[4 * s + 2], 32, "step %d tmem ld", s); names[18 + 4 * s] = buf[4 * s + 2];
      snprintf(buf[4 * s + 3], 32, "step %d p_ready", s); names[19 + 4 * s] = buf[4 * s + 3];
    }
    names[33] = "dkv_full(0)"; names[35] = "dkv_full(1)"; names[38] = "dq_full(0)"; names[39] = "dq_full(1)";
    names[48] = "producer: K0 V0 issue"; names[49] = "producer: Q0 dO0 issue"; names[50] = "producer: Q1 dO1 issue";
    names[51] = "producer: K1 V1 issue"; names[52] = "producer: CLS rows copied"; names[53] = "producer: O issue";
    names[56] = "drain dkv(0): full"; names[57] = "drain dkv(1 prev): full"; names[58] = "drain dkv(0): tmem read"; names[59] = "drain dkv(1 prev): tmem read";
    names[60] = "drain dkv(1 prev) done"; names[61] = "drain dq(1 prev) done";
    names[40] = "problem start"; names[41] = "sync 1"; names[43] = "dO landed"; names[44] = "stats written"; names[42] = "sync 2";
    fprintf(stderr, "attn_tc_bwd trace (SM clocks since the problem's start):\n");
    for (int pass = 0; pass < 1; ++pass) {
      int order[64], n = 0;
      for (int i = 0; i < 64; ++i) if (names[i] && t[i]) order[n++] = i;
      for (int x = 0; x < n; ++x) for (int y = x + 1; y < n; ++y) if (t[order[y]] < t[order[x]]) { int z = order[x]; order[x] = order[y]; order[y] = z; }
      for (int x = 0; x < n; ++x) fprintf(stderr, "  %8lld  %s\n", t[order[x]] - t[40], names[order[x]]);
    }
  }
  return rc ? rc : 1;
}

}  // namespace egv
