/* egovlp_b200.h -- C ABI of libegovlp_b200.so: the B200 (sm_100a) kernels behind the
 * EgoVLPv2 pre-training hot path (TimeSformer video tower + RoBERTa text tower + FIBER-style
 * gated cross-attention + EgoNCE/MLM/ITM heads).
 *
 * The reference (facebookresearch/EgoVLPv2 @ 550c0596) is pure PyTorch and has no FFI; its
 * operator surface is the nn.Module calls cited next to each entry point below (paths relative
 * to EgoVLPv2/).  Every pointer is a DEVICE pointer unless it says "host"; sizes are element
 * counts; `ld*` are row strides in ELEMENTS; `stream` is a cudaStream_t.  All functions are
 * asynchronous on `stream`, return 0 on success and a negative code on error
 * (egv_last_error() gives the message).  No torch types cross this boundary.
 *
 * dtypes: "bf16" = __nv_bfloat16 storage, "f32" = float.  GEMM operands are bf16 with fp32
 * accumulation in TMEM; the residual stream, LayerNorm statistics, softmax and losses are fp32.
 */
#ifndef EGOVLP_B200_H_
#define EGOVLP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* egv_stream_t;

#define EGV_OK 0
#define EGV_ERR_ARG (-1)
#define EGV_ERR_CUDA (-2)
#define EGV_ERR_UNSUPPORTED (-3)

/* library / device ------------------------------------------------------------------------ */
int egv_version(void);
const char* egv_last_error(void);
/* number of kernels this library has launched since load (bench.py's `gpu_launches`) */
long long egv_launch_count(void);
int egv_device_sm_count(void);

/* GEMM ---------------------------------------------------------------------------------------
 * D = epilogue(A x B) on tcgen05 tensor cores (TMA-fed, TMEM accumulators, persistent CTAs).
 *   EGV_GEMM_NT: A[M,K] row-major, B[N,K] row-major      y = x W^T      nn.Linear forward
 *                (video_transformer.py:53,56,120,152,160,166,183; roberta.py:268-278,339,407,420;
 *                 model.py:105-115,279-284,361; heads.py:22,33,41,49)
 *   EGV_GEMM_NN: A[M,K] row-major, B[K,N] row-major      dx = dy W      (autograd of the above)
 *   EGV_GEMM_TN: A[K,M] row-major, B[K,N] row-major      dW = dy^T x    (autograd of the above)
 * epilogue, in this order (each step optional):
 *   v = acc + bias[n];  out_pre = bf16(v);  v = act(v)  or  v *= act'(aux[m,n]);
 *   v *= scale * (scale_dev ? *scale_dev : 1);  v += residual[m,n];
 *   out_f32[m,n] = v (or atomically += v when accumulate != 0);  out_bf16[m,n] = bf16(v);  colsum[n] += v
 */
enum { EGV_GEMM_NT = 0, EGV_GEMM_NN = 1, EGV_GEMM_TN = 2 };
enum {
  EGV_ACT_NONE = 0,
  EGV_ACT_GELU = 1,     /* exact erf GELU (nn.GELU / ACT2FN["gelu"]) */
  EGV_ACT_RELU = 2,
  EGV_ACT_TANH = 3,
  EGV_ACT_GELU_BWD = 4, /* v *= gelu'(aux), aux = pre-activation   */
  EGV_ACT_RELU_BWD = 5, /* v *= (aux > 0),  aux = post-activation  */
  EGV_ACT_TANH_BWD = 6, /* v *= 1 - aux^2,  aux = tanh output      */
  EGV_ACT_GELU_DG = 7,  /* forward like EGV_ACT_GELU, but out_pre receives bf16(gelu'(v)) instead of bf16(v): the
                           backward of the MLP then is a plain multiply (the GELU' epilogue was its slowest GEMM) */
  EGV_ACT_MUL_AUX = 8   /* v *= aux,  aux = the derivative saved by EGV_ACT_GELU_DG */
};
typedef struct egv_gemm_args {
  int layout, M, N, K;
  const void* A; int64_t lda;        /* bf16 */
  const void* B; int64_t ldb;        /* bf16 */
  const float* bias;                 /* [N] or NULL */
  const void* aux; int64_t ld_aux;   /* bf16 [M,N] or NULL */
  const float* scale_dev; float scale;
  const float* residual; int64_t ld_res;
  float* out_f32; int64_t ld_out_f32;
  void* out_bf16; int64_t ld_out_bf16;
  void* out_pre_bf16; int64_t ld_out_pre;
  int act;
  int accumulate;  /* 1: out_f32 += v with fp32 atomics (required when split_k > 1) */
  int split_k;     /* >= 1; K is cut into split_k slices, one CTA pass each */
  float* colsum;   /* optional [N] f32: colsum[n] += sum_m of the final value v (bias gradients); zero it first */
} egv_gemm_args;
int egv_gemm_bf16(const egv_gemm_args* args, egv_stream_t stream);
/* test hook: route every GEMM through the SIMT fallback kernel (1) or restore normal dispatch (0) */
void egv_gemm_force_simt(int on);
/* tcgen05 CTA pairs (cta_group::2: 256 x BN tiles, each CTA stages its 128 rows of A and half of B): 0 = never,
 * 1 = wherever the tile grid allows, 2 (default) = for the TN layout (weight gradients: +7-11 %); bit-identical results. */
void egv_gemm_set_cluster(int mode);
/* tile-width / split-K planner for plain fp32-output GEMMs (weight gradients): 0 = round-1 heuristic, 1 = cost model
 * over {256,128}-wide tiles x split (default), 2 = also 192-wide tiles.  Env EGV_GEMM_PLAN sets the default. */
void egv_gemm_set_plan(int mode);

/* Batched GEMM with the softmax epilogues of the re-associated gated cross-attention -----------------------
 * (video_transformer.py:155-185 video->text, roberta.py:470-486 text->video).
 * D[z] = epilogue(A[z] x B[z]) for a two-level batch grid z = (z1, z0), z0 < nb0, z1 < nb1; operand / tensor X of batch z
 * starts at X + z0 * sX0 + z1 * sX1 elements (a stride of 0 = shared by that level).  Layouts as in egv_gemm_bf16.
 * Rows / reduction elements beyond a batch's own M, N, K read as zero (rank-4 TMA maps), so per-clip operands with
 * N = 3137 tokens need no padding.  Epilogues:
 *   EGV_BGEMM_EPI_NONE        v = acc + bias[n];  v = v * scale * (*scale_dev) + residual;  out_f32 (= or atomically +=),
 *                             out_bf16
 *   EGV_BGEMM_EPI_SOFTMAX32   every aligned group of 32 columns of a row (the 32 text keys of one head):
 *                             v = softmax(acc + bias) -- the attention probabilities P (video_transformer.py:176-181)
 *   EGV_BGEMM_EPI_DSOFTMAX32  backward of that softmax, aux = P (bf16): t = sum_group(P * acc);
 *                             v = scale * (*scale_dev) * P * (acc - t);  colsum[n] += sum_m v;  dot_out[0] += sum t
 * Row strides of every epilogue tensor must be multiples of 4 covering round_up(N, 4) (N itself may be odd: the tail
 * group of 4 columns is written into the row padding).  split_k > 1 (or 1 with accumulate: chosen automatically) adds
 * K slices with fp32 atomics. */
enum { EGV_BGEMM_EPI_NONE = 0, EGV_BGEMM_EPI_SOFTMAX32 = 1, EGV_BGEMM_EPI_DSOFTMAX32 = 2 };
typedef struct egv_bgemm_args {
  int layout, M, N, K;               /* per batch */
  int nb0, nb1;
  const void* A; int64_t lda, sa0, sa1;     /* bf16 */
  const void* B; int64_t ldb, sb0, sb1;     /* bf16 */
  const float* bias; int64_t sbias0, sbias1;
  float scale; const float* scale_dev;
  const float* residual; int64_t ld_res, sres0, sres1;
  float* out_f32; int64_t ld_out_f32, so32_0, so32_1;
  void* out_bf16; int64_t ld_out_bf16, so16_0, so16_1;
  const void* aux; int64_t ld_aux, saux0, saux1;   /* bf16 [M, N] per batch */
  float* colsum; int64_t scol0, scol1;             /* f32 [N] per batch, += */
  float* dot_out;                                  /* f32 [1], += */
  int accumulate, split_k, epilogue;
} egv_bgemm_args;
int egv_bgemm_bf16(const egv_bgemm_args* args, egv_stream_t stream);

/* Row-wise pieces of the re-associated cross-attention (csrc/xattn.cu) ------------------------------------------
 * text -> video (roberta.py:470-486 with 281-321): scores f32 [rows, ld_s] hold, per (clip, head, text query) row, the
 * n = N video-token scores; row r belongs to batch r / rows_per_batch (batch strides in elements).
 * P = softmax(scores) (bf16), lse = log-sum-exp.  p_drop > 0 applies roberta.py:313's dropout to P with a Philox4x32-10
 * stream keyed by (*seed_dev + site constant, r * n + column) -- see "Dropout" below: P = keep ? P / (1 - p) : 0,
 * rsum[r] = row sum of the result (1 without dropout). */
int egv_xattn_row_softmax(const float* scores, int64_t ld_s, int64_t rows, int rows_per_batch, int64_t s_bstride, int n,
                          void* P_bf16, int64_t ld_p, int64_t p_bstride, float* lse, float p_drop, const uint64_t* seed_dev,
                          uint64_t site, float* rsum, egv_stream_t stream);
/* dS = P * (dP~ - sum(P * dP~)), P recomputed from (scores, lse), dP~ = (dP + row_const[r]) * keep / (1 - p) (same Philox
 * stream; row_const, may be NULL, carries the value-bias term d_ctx_h . bv_h, which cancels unless dropout is on) */
int egv_xattn_row_dsoftmax(const float* scores, int64_t ld_s, int64_t rows, int rows_per_batch, int64_t s_bstride, int n,
                           const float* lse, const float* dP, int64_t ld_dp, int64_t dp_bstride, void* dS_bf16,
                           int64_t ld_ds, int64_t ds_bstride, float p_drop, const uint64_t* seed_dev, uint64_t site,
                           const float* row_const, egv_stream_t stream);
/* value bias under dropout: ox[b*S+s, h*64+j] (bf16, row stride ld) += rsum[b, h*S+s] * bv[h*64+j];  its backward:
 * dbv[h*64+j] += sum_{b,s} rsum[b, h*S+s] * d_ox[b*S+s, h*64+j] */
int egv_xattn_rowscale_bias(void* ox_bf16, int64_t ld, const float* rsum, const float* bv, int B, int S, int H,
                            egv_stream_t stream);
int egv_xattn_rowscale_bias_bwd(const void* d_ox_bf16, int64_t ld, const float* rsum, float* dbv, int B, int S, int H,
                                egv_stream_t stream);
/* video -> text (video_transformer.py:166-178): the query-bias term of the re-associated scores plus the key mask:
 * out[b, h*S+s] = scale * sum_j k[b*S+s, h*64+j] * bq[h*64+j] + mask[b*S+s]   (k bf16, row stride ldk; head dim 64) */
int egv_xattn_qbias_fwd(const void* k_bf16, int64_t ldk, const float* bq, const float* mask, float scale, int B, int S,
                        int H, float* out, egv_stream_t stream);
/* its backward: dk[b*S+s, h*64+j] += scale * dbias[b, h*S+s] * bq[h*64+j] (f32, row stride lddk; may be NULL);
 * dbq[h*64+j] += scale * sum_{b,s} k[b*S+s, h*64+j] * dbias[b, h*S+s] (may be NULL) */
int egv_xattn_qbias_bwd(const void* k_bf16, int64_t ldk, const float* bq, const float* dbias, float scale, int B, int S,
                        int H, float* dk, int64_t lddk, float* dbq, egv_stream_t stream);

/* Dropout of the RoBERTa tower in train mode (csrc/dropout.cu, csrc/philox.cuh) --------------------------------------
 * roberta.py:162,203 (embeddings), :244,313 (attention probabilities), :337,342 (attention output dense), :418,422
 * (feed-forward output dense); p = 0.1.  Masks come from Philox4x32-10: keep(element i of site s) is a pure function of
 * (key, i) with key = *seed_dev + s * 0x9E3779B97F4A7C15 (seed_dev NULL = 0).  `seed_dev` is a device-resident uint64
 * that egv_rng_advance bumps once per training step (inside the captured CUDA graph), `site` an immediate that names
 * the call site (pass, layer, kind); the backward regenerates the forward's mask from the same pair. */
int egv_rng_advance(uint64_t* state, egv_stream_t stream);
/* y = keep ? x / (1 - p) : 0 (x f32 or bf16, n elements); out_f32 (may be NULL) = (res ? res : 0) + scale * (*scale_dev) * y;
 * out_bf16 (may be NULL) = bf16(y)          -- dense -> dropout -> + residual (roberta.py:341-343, 421-425) */
int egv_dropout_add(const void* x, int x_is_bf16, const float* res, float scale, const float* scale_dev, float p_drop,
                    const uint64_t* seed_dev, uint64_t site, float* out_f32, void* out_bf16, int64_t n, egv_stream_t stream);
/* out = keep ? dy / (1 - p) : 0  (dy f32 or bf16; f32 and / or bf16 outputs) */
int egv_dropout_bwd(const void* dy, int dy_is_bf16, float p_drop, const uint64_t* seed_dev, uint64_t site, float* out_f32,
                    void* out_bf16, int64_t n, egv_stream_t stream);
/* RobertaSelfAttention core with dropout on the probabilities (roberta.py:281-321), S <= 64 tokens, head dim 64:
 * P = softmax(scale * q k^T + key_bias[b, :]);  o = dropout(P) v.  q / k / v / d* rows are b*S + s with row stride ld
 * (ldd for gradients), head h at columns h*64; lse [B, H, S] is written by fwd and read by bwd.  Dropout element index =
 * ((b*H + h)*S + query)*S + key. */
typedef struct egv_text_attn_args {
  const void* q; const void* k; const void* v; int64_t ld;     /* bf16 */
  const float* key_bias;                                      /* [B, S] additive or NULL */
  float scale, p_drop;
  const uint64_t* seed_dev; uint64_t site;
  int B, H, S;
  void* o; int64_t ldo;                                       /* bf16 */
  float* lse;
  const void* d_o;                                            /* bwd: bf16, addressed like o */
  void* dq; void* dk; void* dv; int64_t ldd;                  /* bwd: bf16 */
} egv_text_attn_args;
int egv_text_attention_fwd(const egv_text_attn_args* a, egv_stream_t stream);
int egv_text_attention_bwd(const egv_text_attn_args* a, egv_stream_t stream);

/* LayerNorm ------------------------------------------------------------------------------------
 * nn.LayerNorm over the last dim C (video_transformer.py:196,207,210,115,304; roberta.py:161,336,417;
 * model.py:155; heads.py:41).  x is f32 or bf16 (x_is_bf16), y is written as bf16 and/or f32.
 * mean/rstd [rows] are saved for the backward.  */
int egv_layernorm_fwd(const void* x, int x_is_bf16, const float* gamma, const float* beta, float eps,
                      int64_t rows, int C, void* y_bf16, float* y_f32, float* mean, float* rstd,
                      egv_stream_t stream);
/* val = LayerNorm'(dy);  dx (f32, may be NULL) = (add ? add : 0) + val, where add (f32, may be NULL) may alias dx;
 * dx_bf16 (may be NULL) = bf16(bf16_total ? dx : val);  dgamma/dbeta [C] (f32, may be NULL) are always accumulated
 * with atomics: zero them first for a plain result; out_colsum [C] (may be NULL) += column sums of the value written
 * to dx_bf16 (the bias gradient of the Linear feeding this branch).  dy is f32 or bf16. */
int egv_layernorm_bwd(const void* dy, int dy_is_bf16, const void* x, int x_is_bf16, const float* gamma,
                      const float* mean, const float* rstd, int64_t rows, int C, const float* add, float* dx,
                      void* dx_bf16, int bf16_total, float* dgamma, float* dbeta, float* out_colsum,
                      egv_stream_t stream);

/* Strided multi-head attention (head_dim 64) ------------------------------------------------------
 * One flash-style kernel family covers every attention on the path:
 *   divided time / space attention and the CLS query   video_transformer.py:35-39,117-153
 *   gated video->text cross-attention core              video_transformer.py:170-182
 *   RoBERTa self-attention and text->video cross core   roberta.py:281-321
 * Work unit = (batch b, head h, group g).  Query i of group g is row
 *   q_row0 + g*q_gstride + i*q_istride of batch b (row stride ldq elements, batch stride
 *   q_bstride rows); key/value j likewise with the k_* fields; when has_cls_key != 0 an extra
 *   key/value (row `cls_row` of the batch) is prepended to every group.
 * scores = scale * q.k + key_bias[b, j]  (key_bias: additive f32 [B, Lk], may be NULL).
 * lse [B, H, G, Lq] f32 is written by fwd and consumed by bwd.  */
typedef struct egv_attn_args {
  int B, H, G, Lq, Lk;
  const void* q; int64_t ldq, q_bstride; int q_row0, q_gstride, q_istride;
  const void* k; const void* v; int64_t ldkv, kv_bstride; int k_row0, k_gstride, k_istride;
  int has_cls_key, cls_row;
  const float* key_bias;
  float scale;
  void* o; int64_t ldo, o_bstride;      /* bf16, rows addressed like q */
  float* lse;
  /* backward only */
  const void* d_o;                      /* bf16, addressed like o */
  void* dq; int64_t lddq;               /* bf16, addressed like q */
  void* dk; void* dv; int64_t lddkv;    /* bf16, addressed like k/v */
  float* delta;                         /* scratch [B,H,G,Lq] f32 */
  float* dkv_cls;                       /* f32 [B, H, 2, 64] accumulator for the shared CLS key/value (or NULL) */
  int dkv_accumulate;                   /* 1: dk/dv rows are read-modify-written (+=) */
  /* optional scratch for the split-stream path (few rows x long stream, e.g. 32 text queries x 3137 video keys):
   * egv_attention_workspace_bytes() bytes of device memory, contents undefined; NULL = never split */
  void* workspace; int64_t workspace_bytes;
  /* optional, backward of the grouped problems only: fold the CLS QUERY of the divided attention into the group kernel
   * (video_transformer.py:134-150: the CLS token attends every token of the clip, i.e. every group).  lse_cls f32 [B, H] =
   * that query's log-sum-exp as egv_attention_fwd of the single-query problem wrote it; dq_cls f32 [B, H, 64] = accumulator
   * (zeroed by the caller) for its query gradient; its q / o / d_o rows are row `cls_row` of each batch.  When the kernel
   * takes the fold it adds the query's contribution to dk / dv (and to dkv_cls) itself, and *cls_query_folded (a HOST int)
   * is set to 1: the caller then skips the separate single-query backward and calls egv_attention_cls_query_finalize.
   * egv_attention_fwd takes the same fold with lse_cls as an OUTPUT (dq_cls unused): when *cls_query_folded comes back 1,
   * o row `cls_row` of each batch and lse_cls [B, H] hold the CLS query's attention over all keys of the clip and the
   * caller skips the separate single-query forward. */
  const float* lse_cls;
  float* dq_cls;
  int* cls_query_folded;
} egv_attn_args;
int64_t egv_attention_workspace_bytes(const egv_attn_args* a);
/* Threading: the entry points are stream-ordered and may be issued on several streams concurrently, with ONE exception:
 * single-query problems (Lq == 1 over all heads: the CLS query of the divided attention, video_transformer.py:134-150)
 * use library-owned scratch (key-split partials; a self-zeroing fp32 dq accumulator and per-clip tickets that the last
 * CTA of each clip resets) -- issue those from one stream at a time.  The scratch grows on first use: run every shape once
 * before capturing a CUDA graph (trainer.PretrainStep.capture does). */
int egv_attention_fwd(const egv_attn_args* a, egv_stream_t stream);
int egv_attention_bwd(const egv_attn_args* a, egv_stream_t stream);
/* fused tiny-group kernels (time attention; csrc/attention_tiny.cu): bit 0 = forward, bit 1 = backward.  Default 3
 * (env EGV_ATTN_TINY); 0 routes those shapes through the generic warp-per-group kernels. */
void egv_attention_set_tiny(int mode);
/* tcgen05 / TMEM space-attention kernel (csrc/attention_tc.cu): bit 0 = forward.  Default 1 (env EGV_ATTN_TC); 0 routes
 * those shapes through the mma.sync group kernels (csrc/attention_group.cu). */
void egv_attention_set_tc(int mode);
/* dk/dv row `cls_row` of every batch (+)= the fp32 accumulators dkv_cls [B,H,2,64] filled by egv_attention_bwd */
int egv_attention_cls_finalize(const float* dkv_cls, void* dk, void* dv, int64_t lddkv, int64_t kv_bstride, int cls_row,
                               int B, int H, int accumulate, egv_stream_t stream);
/* dq row `cls_row` of every batch = bf16 of the fp32 accumulator dq_cls [B, H, 64] of a folded CLS query (see egv_attn_args) */
int egv_attention_cls_query_finalize(const float* dq_cls, void* dq, int64_t lddq, int64_t q_bstride, int cls_row, int B, int H,
                                     egv_stream_t stream);

/* Elementwise / reductions --------------------------------------------------------------------- */
/* fp32 -> bf16 cast (weights, activations); n elements */
int egv_cast_f32_bf16(const float* x, void* y, int64_t n, egv_stream_t stream);
int egv_cast_bf16_f32(const void* x, float* y, int64_t n, egv_stream_t stream);
/* out[c] (+)= scale * (scale_dev ? *scale_dev : 1) * sum_r x[r, c]   -- bias gradients; x f32 or bf16 [rows, C], row stride ld */
int egv_colsum(const void* x, int x_is_bf16, int64_t rows, int C, int64_t ld, float* out, int accumulate, float scale,
               const float* scale_dev, egv_stream_t stream);
int egv_zero_f32(float* p, int64_t n, egv_stream_t stream);
/* out_bf16[i] = bf16( scale * (scale_dev ? *scale_dev : 1) * dy[i] * act'(aux[i]) )  -- gradient through a trailing
 * activation (act = EGV_ACT_NONE: scaled cast; *_BWD: aux as in egv_gemm_bf16).  dy is f32 or bf16. */
int egv_act_grad(const void* dy, int dy_is_bf16, const void* aux_bf16, int act, float scale, const float* scale_dev,
                 void* out_bf16, int64_t n, egv_stream_t stream);
/* out[0] (+)= sum_i a[i]*b[i]   (gradient of the scalar gates alpha_i2t / alpha_t2i) */
int egv_dot(const void* a, int a_is_bf16, const void* b, int b_is_bf16, int64_t n, float* out, int accumulate,
            egv_stream_t stream);
/* y = a + alpha * (alpha_dev ? *alpha_dev : 1) * b  (f32, in-place allowed; a may be NULL = 0);
 * writes y (f32) and/or y_bf16 */
int egv_axpy_f32(const float* a, const float* b, float alpha, const float* alpha_dev, float* y, void* y_bf16, int64_t n,
                 egv_stream_t stream);
/* y[r, 0:C] += x[r, 0:C] for r < rows; ldy / ldx = row strides (the residual add on the CLS rows of a [B, N, C]
 * stream when only the CLS output of a block is consumed: video_transformer.py:391, model.py:275) */
int egv_add_rows_f32(float* y, int64_t ldy, const float* x, int64_t ldx, int rows, int C, egv_stream_t stream);

/* Patch embedding (video_transformer.py:78-83, 354-372; model.py:211-232) -------------------------
 * im2col: video f32 [BT, 3, H, W] -> bf16 [BT*gh*gw, 3*p*p] (column order c,i,j = Conv2d weight order) */
int egv_patchify(const float* video, int BT, int Cin, int H, int W, int p, void* out_bf16, egv_stream_t stream);
/* any patch size (e.g. 14: TimeSformer-L/14, BASELINE cfg 5): rows of ld_out >= Cin*p*p elements (a multiple of 8 for the
 * GEMM's TMA maps), columns beyond Cin*p*p zero-filled; the GEMM then runs over the padded depth against a zero-padded weight */
int egv_patchify_padded(const float* video, int BT, int Cin, int H, int W, int p, int64_t ld_out, void* out_bf16,
                        egv_stream_t stream);
/* the same im2col from uint8 frames, fused with the loader's frames.float() / 255 (base/base_dataset.py:248,300) and
 * NormalizeVideo's (x - mean[c]) / std[c] (data_loader/transforms.py:49): video u8 [BT, Cin, H, W]; mean / std = HOST
 * arrays of Cin floats (Cin <= 4).  Output bit-identical to egv_patchify on the host-normalised fp32 frames. */
int egv_patchify_u8(const uint8_t* video, int BT, int Cin, int H, int W, int p, const float* mean, const float* stdv,
                    void* out_bf16, egv_stream_t stream);
/* tokens[b,0] = cls + pos[0]; tokens[b,1+f*Nf+n] = patch[b,f,n] + pos[1+n] + temporal[f]   (all f32) */
int egv_assemble_tokens(const float* patch, const float* cls, const float* pos, const float* temporal, int B,
                        int T, int Nf, int C, float* tokens, egv_stream_t stream);
/* backward of the above: d_patch (bf16 [B*T*Nf, C]), d_cls [C], d_pos [1+Nf, C], d_temporal [T, C] (f32, +=) */
int egv_assemble_tokens_bwd(const float* d_tokens, int B, int T, int Nf, int C, void* d_patch_bf16, float* d_cls,
                            float* d_pos, float* d_temporal, egv_stream_t stream);

/* RoBERTa embeddings (roberta.py:174-204, 881-892): word + type[0] + position(cumsum non-pad) (pre-LN sum) */
int egv_text_embed(const int64_t* ids, int B, int S, int C, int pad_id, const float* word, const float* pos,
                   const float* type0, float* out, egv_stream_t stream);
int egv_text_embed_bwd(const float* d_out, const int64_t* ids, int B, int S, int C, int pad_id, float* d_word,
                       float* d_pos, float* d_type0, egv_stream_t stream);

/* Losses ------------------------------------------------------------------------------------------
 * softmax cross-entropy with ignore_index (model.py:414-418, 478): per-row loss and dlogits.
 * logits f32 [rows, ld] (V valid columns); writes loss_sum[0] += sum of valid-row losses,
 * count[0] += valid rows; dlogits ([rows, ld_d] bf16, or f32 when dlogits_is_f32; may be NULL) = softmax - onehot
 * (unscaled; rows with label == ignore_index are zero). */
int egv_softmax_xent(const float* logits, int64_t ld, const int64_t* labels, int64_t rows, int V, int ignore_index,
                     float* loss_sum, float* count, void* dlogits, int dlogits_is_f32, int64_t ld_d, egv_stream_t stream);
/* loss[0] = loss_sum / max(count, 1); inv_count[0] = 1 / max(count, 1)  (device scalars: no host sync) */
int egv_xent_finalize(const float* loss_sum, const float* count, float* loss, float* inv_count, egv_stream_t stream);
/* EgoNCE (model.py:385-394, 576-584 sim_matrix; loss.py:40-61).
 * t, v: f32 [G, P] raw (gathered) embeddings; noun [G, Dn], verb [G, Dv] f32 multi-hot.
 * Outputs: sim [G,G] f32 (rows = text), mask [G,G] uint8, loss[0]; dt, dv f32 [grad_rows, P] = dloss/dt, dloss/dv
 * for rows grad_row0 .. grad_row0+grad_rows-1 (this rank's slice: trainer_egoclip.py:36-41).
 * scratch: egv_egonce_scratch_floats(G, P, Dn, Dv) floats. */
int64_t egv_egonce_scratch_floats(int G, int P, int Dn, int Dv);
int egv_egonce(const float* t, const float* v, int G, int P, const float* noun, int Dn, const float* verb, int Dv,
               float temperature, float* sim, uint8_t* mask, float* loss, int grad_row0, int grad_rows, float* dt,
               float* dv, float* scratch, egv_stream_t stream);

/* Fine-tuning losses of the dual-encoder ('Dual') path (model_epic_charades.py:408-445): sim_matrix (:542-550) of the
 * gathered embeddings + NormSoftmaxLoss (loss.py:13-31; param = temperature), MaxMarginRankingLoss (loss.py:65-100;
 * param = margin) or AdaptiveMaxMarginRankingLoss (loss.py:102-143; param = margin, weight f32 [G] = data['relation']).
 * Same layout and gradient-slice contract as egv_egonce.  scratch: egv_dual_loss_scratch_floats(G, P) floats. */
#define EGV_DUAL_NORM_SOFTMAX 0
#define EGV_DUAL_MAX_MARGIN 1
#define EGV_DUAL_ADAPTIVE_MAX_MARGIN 2
int64_t egv_dual_loss_scratch_floats(int G, int P);
int egv_dual_loss(const float* t, const float* v, int G, int P, int kind, float param, const float* weight, int fix_norm,
                  float* sim, float* loss, int grad_row0, int grad_rows, float* dt, float* dv, float* scratch,
                  egv_stream_t stream);

/* NVSwitch peer-to-peer all-gather (trainer_egoclip.py:25-41 AllGather_multi; model.py:385-388) ------------
 * Symmetric buffers: every rank egv_p2p_alloc()s `slots` (2 * W * slot_bytes: two sets, consecutive gathers alternate)
 * and `flags` (32 uint32), exchanges the 64-byte IPC handles out of band (torch.distributed) and maps the peers'
 * buffers with egv_p2p_open().  egv_p2p_allgather pushes `bytes` from src into slot `rank` of every peer, publishes a
 * device-resident, monotonically increasing sequence number, waits until all W local flags reached it and copies the
 * W gathered payloads (rank-major) into `out`.  No host-side state: the call can be captured in a CUDA graph. */
int egv_p2p_alloc(int64_t bytes, void** ptr, void* ipc_handle_64);
int egv_p2p_open(const void* ipc_handle_64, void** ptr);
int egv_p2p_close(void* ptr);
int egv_p2p_free(void* ptr);
int egv_p2p_allgather(const void* src, int64_t bytes, int64_t slot_bytes, void* const* slots, void* const* flags, int rank,
                      int world, void* out, egv_stream_t stream);
/* A rank that waits ~10 s for a peer's flag gives up WITHOUT killing the CUDA context: it leaves a sticky word in its own
 * flags buffer.  *err = 0: no gather of this rank ever timed out; else 0x100 | (rank it was waiting for).  Synchronous
 * 4-byte device-to-host read: call it at a point where the trainer synchronises anyway. */
int egv_p2p_error(const void* flags_local, int* err);

/* Optimiser (next row f-1): fused AdamW (transformers.AdamW semantics, set_optim_schedule.py:108) on one flat fp32
 * tensor; also refreshes the bf16 weight copy.  hyper_dev (optional, device float[3]) = {lr multiplier, 1-beta1^t,
 * 1-beta2^t}: when given it overrides bias_c1 / bias_c2 and scales lr, so a captured CUDA graph can be replayed
 * across optimiser steps. */
int egv_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
              float beta2, float eps, float weight_decay, float bias_c1, float bias_c2, float grad_scale,
              const float* hyper_dev, egv_stream_t stream);

/* step_dev[0] (device int64: optimiser steps completed) += 1 and hyper_dev = {LR multiplier of the HF cosine schedule with
 * warm-up (set_optim_schedule.py:115-119; 1 when max_steps <= 0), 1 - beta1^t, 1 - beta2^t} for the new step t -- computed
 * on the device so that the whole step, schedule included, replays from one CUDA graph. */
int egv_adamw_schedule(int64_t* step_dev, float* hyper_dev, int warmup_steps, int max_steps, float beta1, float beta2,
                       egv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EGOVLP_B200_H_ */
