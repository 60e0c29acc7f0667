// Shared declarations of the attention translation units (attention.cu: generic strided kernels;
// attention_group.cu: group-resident TMA-fed kernels for large contiguous groups, i.e. space attention).
#pragma once
#include <cuda.h>

#include "common.cuh"
#define EGV_PDL_CLASS 4
#include "host_common.h"

namespace egv {

enum { MODE_FWD = 0, MODE_DQ = 1, MODE_DKV = 2 };
constexpr int HD = 64;       // head dim
constexpr int LDS = 72;      // smem row stride in bf16 (144 B: conflict-free ldmatrix)
constexpr float LOG2E = 1.4426950408889634f;

// 2^x on the XU pipe in one instruction (inputs here are <= 0 or differences of bounded log-sum-exps)
EGV_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnP {
  int B, H, G, Lq, LkT;  // LkT = keys per group including the optional CLS key
  const bf16* q; long long ldq, q_bstride; int q_row0, q_gstride, q_istride;
  const bf16* k; const bf16* v; long long ldkv, kv_bstride; int k_row0, k_gstride, k_istride;
  int has_cls, cls_row;
  const float* key_bias;
  float scale;
  bf16* o; long long ldo, o_bstride;
  float* lse;
  const bf16* d_o;
  bf16* dq; long long lddq;
  bf16* dk; bf16* dv; long long lddkv;
  float* delta;
  float* dkv_cls;
  int dkv_accumulate;
  const float* lse_cls;   // optional CLS-query fold (egv_attn_args): that query's lse [B, H], its dq accumulator [B, H, 64]
  float* dq_cls;
  float* cls_part;        // forward fold: per-frame partials [B, G, H, 66] (library scratch), else NULL
  int* fold_out;          // HOST pointer: set to 1 by the kernel launcher that took the CLS query along
  int row_tiles;       // tiles of the row side per group
  long long items;     // B*G*H*row_tiles*n_split
  // stream-side split (few rows, long stream: text->video attention and the key side of video->text attention):
  // item = (group, tile, split); every split walks 1/n_split of the stream and leaves an fp32 partial in `ws`
  int n_split;
  float* ws;
  long long ws_floats;
};


// ---- device helpers shared by the TMA-fed kernels (attention_group.cu, attention_tiny.cu)
EGV_DEVINL void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// per-lane constants of the two ldmatrix address patterns over a 128B-swizzled [rows][64] bf16 tile
// (16-byte chunk c of row r lives at r * 128 + ((c ^ (r & 7)) << 4))
struct LaneAddr {
  uint32_t nt_row, nt_x[4];   // B operand of A * tile^T : rows = n index
  uint32_t p_row, p_x[4];     // B operand of P * tile (transposed load) and A fragments of a row tile
};
EGV_DEVINL LaneAddr lane_addr(int lane) {
  LaneAddr la;
  const uint32_t l7 = lane & 7;
  la.nt_row = ((lane & 7) + ((lane >> 4) << 3)) * 128;
  la.p_row = ((lane & 7) + ((lane >> 3) & 1) * 8) * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    la.nt_x[i] = (((uint32_t)(i * 2 + ((lane >> 3) & 1))) ^ l7) << 4;
    la.p_x[i] = (((uint32_t)(i * 2 + (lane >> 4))) ^ l7) << 4;
  }
  return la;
}

// a warp's 16 x 64 fp32 C-layout tile -> its swizzled bf16 staging tile
EGV_DEVINL void g_stage(uint8_t* stg, const float (&c)[8][4], float mul0, float mul1, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(stg + g * 128 + ((nt ^ (g & 7)) << 4) + 4 * t) = pack_bf16(c[nt][0] * mul0, c[nt][1] * mul0);
    *reinterpret_cast<uint32_t*>(stg + (g + 8) * 128 + ((nt ^ (g & 7)) << 4) + 4 * t) = pack_bf16(c[nt][2] * mul1, c[nt][3] * mul1);
  }
}

// attention_group.cu: returns 1 when it launched the group-resident kernel for this problem, 0 when the problem is not
// eligible (the caller falls back to the generic kernels), < 0 on error.
int launch_group_attention(int mode, const AttnP& a, cudaStream_t stream);

// attention_tc.cu: tcgen05 / TMEM kernel for the same problems (space attention, forward).  Same return convention; tried
// first, the mma.sync kernels of attention_group.cu remain the fallback for shapes it does not cover.
int launch_tc_attention(int mode, const AttnP& a, cudaStream_t stream);
void set_tc_attention_mode(int mode);
// attention_tc_bwd.cu: the whole backward (dQ, dK, dV) of the same problems in one tcgen05 launch; the caller zeroes
// dkv_cls (the shared CLS key's fp32 accumulators) and finalises it as for the mma.sync kernels.
int launch_tc_attention_bwd(const AttnP& a, cudaStream_t stream, int* cls_query_folded);

// gemm.cu: cached bf16 2-D tensor map (`inner` contiguous elements, `outer` rows `ld` elements apart, 128B swizzle)
int get_tensor_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                   CUtensorMap* out);

int get_tensor_map_nd(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMap* out);

// attention_tiny.cu: fused kernels for many tiny groups (time attention: <= 16 queries x <= 16 keys + the CLS key per
// group, adjacent groups on adjacent rows).  Same return convention as launch_group_attention; `bwd` = 0 forward,
// 1 = the whole backward (dQ, dK, dV, delta) in one launch.
int launch_tiny_attention(int bwd, const AttnP& a, cudaStream_t stream);
void set_tiny_mode(int mode);

}  // namespace egv
