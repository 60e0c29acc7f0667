// Shared declarations of the attention translation units (attention.cu: generic strided kernels;
// attention_group.cu: group-resident TMA-fed kernels for large contiguous groups, i.e. space attention).
#pragma once
#include <cuda.h>

#include "common.cuh"
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
  int row_tiles;       // tiles of the row side per group
  long long items;     // B*G*H*row_tiles
};


// attention_group.cu: returns 1 when it launched the group-resident kernel for this problem, 0 when the problem is not
// eligible (the caller falls back to the generic kernels), < 0 on error.
int launch_group_attention(int mode, const AttnP& a, cudaStream_t stream);

// gemm.cu: cached bf16 2-D tensor map (`inner` contiguous elements, `outer` rows `ld` elements apart, 128B swizzle)
int get_tensor_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                   CUtensorMap* out);

}  // namespace egv
