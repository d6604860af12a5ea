// Host-side launchers of the non-GEMM kernels. All: no allocation, no synchronisation, return 0 on success.
#pragma once
#include "common.cuh"
#include "gemm.h"

// ---- front end (frontend.cu) ----------------------------------------------------------------------------
enum { FE_MODE_WARP = 0, FE_MODE_PASTE20 = 1, FE_MODE_FIX = 2, FE_MODE_NONE = 3 };
struct FrontendNorm {
  float mean[2][3];
  float std[2][3];
};
int patch_frontend_fwd(const uint8_t* obs, const float* patch, const int* xy, const float* theta, bf16* out, int B,
                       int H, int W, int ph, int pw, int mode, const FrontendNorm& nrm, cudaStream_t stream);
int patch_frontend_bwd(const bf16* dout, const float* patch, const int* xy, const float* theta, float* dpatch, int B,
                       int H, int W, int ph, int pw, int mode, const FrontendNorm& nrm, cudaStream_t stream);

// eval-time paste (simulation_random_patch): img / out uint8 [B,H,W,3]; patch f32 [3,ph,pw] in [0,1] (quantised in-kernel)
int patch_sim_paste(const uint8_t* img, const float* patch, const int* xy, const float* theta, uint8_t* out, int B, int H,
                    int W, int ph, int pw, int geometry, cudaStream_t stream);

// ---- layout / elementwise (elementwise.cu) --------------------------------------------------------------
// px [B,6,H,W] -> per-tower im2col rows [B*np, kpad] (k = c*P*P + ky*P + kx, zero padded to kpad)
int im2col_patches(const bf16* px, bf16* a_dino, bf16* a_sig, int B, int H, int W, int P, int kpad, cudaStream_t s);
// inverse scatter of the two im2col-layout gradients back to dpx [B,6,H,W]
int col2im_patches(const bf16* da_dino, const bf16* da_sig, bf16* dpx, int B, int H, int W, int P, int kpad,
                   cudaStream_t s);
// x[b, 0:npre] = prefix tokens (cls, reg...) ; rows npre.. are written by the patch-embed GEMM epilogue
int write_prefix_tokens(const bf16* cls, const bf16* reg, bf16* x, int B, int ntok, int npre, int d, cudaStream_t s);
// dst[b*rows_dst + dst_off + r, 0:cols] = src[b*rows_src + src_off + r, 0:cols]  for r < rows, b < B
int copy_rows(const bf16* src, int64_t lds, int rows_src, int src_off, bf16* dst, int64_t ldd, int rows_dst,
              int dst_off, int B, int rows, int cols, cudaStream_t s);
// x[b] = [table[ids[b,0]] | P rows left for the projector | table[ids[b,1..T-1]]]; ids row stride ld_ids (>= T)
int embed_tokens_splice(const int64_t* ids, int ld_ids, const bf16* table, bf16* x, int B, int T, int P, int d,
                        cudaStream_t s);
int gelu_bwd(const bf16* dy, const bf16* pre, bf16* dx, int64_t n, cudaStream_t s);
int scale_cols(const bf16* x, const bf16* gamma, bf16* y, int64_t rows, int cols, cudaStream_t s);
// gu [M, 2F] holds gate and up interleaved in groups of 64 features: columns [128k, 128k+64) = gate features
// [64k, 64k+64), columns [128k+64, 128k+128) = up features (the layout the packed gate|up GEMM emits; F % 64 == 0).
// act [M, F] = bf16(bf16(silu(gate)) * up)
int swiglu_fwd(const bf16* gu, bf16* act, int64_t M, int F, cudaStream_t s);
int swiglu_bwd(const bf16* dact, const bf16* gu, bf16* dgu, int64_t M, int F, cudaStream_t s);
// rotary embedding applied in place to the q and k thirds of qkv [M, 3*H*hd]; pos = row % L. dir=+1 fwd, -1 bwd
int rope_inplace(bf16* qkv, const float* cos_tab, const float* sin_tab, int64_t M, int L, int H, int hd, int dir,
                 cudaStream_t s);
// out[r] = a[r] + b[r] (bf16 rounding)
int add_bf16(const bf16* a, const bf16* b, bf16* out, int64_t n, cudaStream_t s);
// W [rows, cols] (row stride ldi) -> Wt [cols, rows] (row stride ldo)
int transpose_bf16(const bf16* w, int64_t ldi, bf16* wt, int64_t ldo, int rows, int cols, cudaStream_t s);
int gather_rows(const bf16* src, const int* rows, bf16* dst, int R, int d, cudaStream_t s);
int scatter_rows(const bf16* src, const int* rows, bf16* dst, int R, int d, cudaStream_t s);

// ---- norms (norm.cu) -------------------------------------------------------------------------------------
int layernorm_fwd(const bf16* x, const bf16* w, const bf16* b, bf16* y, float* mean, float* rstd, int64_t M, int d,
                  float eps, cudaStream_t s);
// dx_out = bf16(dres + bf16(layernorm_bwd(dy)))   (dres may be null)
// optional second output: scaled = bf16(dx * gamma2[col]) (timm LayerScale backward of the branch dx enters next)
int layernorm_bwd(const bf16* dy, const bf16* x, const bf16* w, const float* mean, const float* rstd, const bf16* dres,
                  bf16* dx, int64_t M, int d, cudaStream_t s, const bf16* gamma2 = nullptr, bf16* scaled = nullptr);
// two independent LayerNorm problems in one launch (the two vision towers at the same depth)
struct LnFwdProblem {
  const bf16 *x, *w, *b;
  bf16* y;
  float *mean, *rstd;
  int64_t M;
  int d;
};
struct LnBwdProblem {
  const bf16 *dy, *x, *w;
  const float *mean, *rstd;
  const bf16* dres;
  bf16* dx;
  int64_t M;
  int d;
  const bf16* gamma2;
  bf16* scaled;
};
int layernorm_fwd2(const LnFwdProblem& p0, const LnFwdProblem& p1, float eps, cudaStream_t s);
int layernorm_bwd2(const LnBwdProblem& p0, const LnBwdProblem& p1, cudaStream_t s);
int rmsnorm_fwd(const bf16* x, const bf16* w, bf16* y, float* rstd, int64_t M, int d, float eps, cudaStream_t s);
int rmsnorm_bwd(const bf16* dy, const bf16* x, const bf16* w, const float* rstd, const bf16* dres, bf16* dx, int64_t M,
                int d, cudaStream_t s);

// ---- attention (attention.cu) ----------------------------------------------------------------------------
// qkv [B*N, 3*H*hd] (q | k | v, heads contiguous inside each third); o [B*N, H*hd]; lse [B, H, N] fp32.
// causal != 0: key j visible to query i iff j <= i; kv_len (nullable) [B]: keys >= kv_len[b] are masked.
int attention_fwd(const bf16* qkv, bf16* o, float* lse, const int* kv_len, int B, int N, int H, int hd, int causal,
                  cudaStream_t s);
// persistent streaming tcgen05 forward (attention_fwd_tc2.cu): head dims 64 / 72 / 128, any N
bool attention_fwd_tc2_supported(int N, int hd);
int attention_fwd_tc2(const bf16* qkv, bf16* o, float* lse, const int* kv_len, int B, int N, int H, int hd, int causal,
                      cudaStream_t s);
// persistent tcgen05 forward with the whole score tile resident in TMEM (attention_fwd_tc3.cu): N <= 320
bool attention_fwd_tc3_supported(int N, int hd);
int attention_fwd_tc3(const bf16* qkv, bf16* o, float* lse, const int* kv_len, int B, int N, int H, int hd, int causal,
                      cudaStream_t s);
// tcgen05 backward (attention_bwd_tc.cu): head dims 64 / 72 / 128, any N; same contract as attention_bwd.
bool attention_bwd_tc_supported(int N, int hd);
bool attention_bwd_takes_delta(int N, int hd);
// rows by which the 128-row query tiles of the tcgen05 kernels are shifted towards the front under the causal mask (0 or 64;
// env VLA_ATTN_SHIFT=0 disables): the ragged tile then sits where the mask leaves the fewest keys
int attention_tile_shift(int N, int causal);   // attention_bwd(o == NULL): delta precomputed by the producer of dO
int attention_bwd_tc(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                     const int* kv_len, int B, int N, int H, int hd, int causal, const float* rope_cos, const float* rope_sin,
                     int rope_L, cudaStream_t s);
int attention_impl();      // resolved kernel selection (default / env VLA_ATTN_IMPL on first use)
extern int g_attn_impl;   // bit 0 = tcgen05 forward, bit 1 = tcgen05 backward where supported; -1 = unset (default 3 / env)
// dqkv [B*N, 3*H*hd]; delta scratch [B, H, N] fp32
// rope_cos / rope_sin (nullable; head dim 128 only): when given, d(q) and d(k) are returned with the rotary embedding's
// backward already applied (gradients wrt the PRE-RoPE projections), i.e. rope_inplace(dqkv, ..., dir = -1) is fused.
int attention_bwd(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                  const int* kv_len, int B, int N, int H, int hd, int causal, const float* rope_cos, const float* rope_sin,
                  int rope_L, cudaStream_t s);

// ---- loss head (loss_head.cu) ----------------------------------------------------------------------------
enum { LOSS_UADA = 0, LOSS_UADA_DDP = 1, LOSS_UPA = 2, LOSS_CE = 3, LOSS_NEG_CE = 4 };
struct LossParams {
  int kind;
  float mse_weight;   // UADA: 5 (UADA.py:396) / MSE_weights (UADA_ddp.py:114)
  float alpha, belta; // UPA (UPA.py:386)
  float ce_scale;     // TMA: 1/accumulate_steps (TMA.py:148)
};
// Scalars written by the loss head (device float[LOSS_NUM_SCALARS])
enum { LS_LOSS = 0, LS_CE = 1, LS_AUX0 = 2, LS_AUX1 = 3, LS_UAD = 4, LS_NTOK = 5, LS_NACT = 6, LS_GRAD_MEAN = 7,
       LOSS_NUM_SCALARS = 8 };
// logits fp32 [R, V] (bf16-rounded values) of the R supervised rows, sorted by (sample, position);
// meta int [R][3] = {label of the row, sample index, index of the row among its sample's supervised rows}.
// Writes dlogits bf16 [R, V], scalars, per-row argmax action id (pred_ids int [R]; -1 for non-action rows).
int loss_head_fwd_bwd(const float* logits, const int* meta, int R, int V, int B, const LossParams& lp, float* row_stats,
                      bf16* dlogits, float* scalars, int* pred_ids, cudaStream_t s);
size_t loss_head_row_stats_floats(int R);

// ---- patch update (patch_update.cu) ----------------------------------------------------------------------
enum { OPT_ADAMW = 0, OPT_PGD = 1 };
// grad is scaled by grad_scale (1/world after an all-reduce(sum)) first; clip_l1 > 0 applies
// clip_grad_norm_(max_norm=clip_l1, norm_type=1) (UPA.py:157); then the update and clamp(0,1).
int patch_update(float* patch, const float* grad, float* m, float* v, int n, int step, float lr, float beta1,
                 float beta2, float eps, int kind, float grad_scale, float clip_l1, float* scalars, cudaStream_t s);

// ---- device-resident step state of vla_attack_step (patch_update.cu) --------------------------------------
struct StepState {
  int place;     // index of the placement (and scalar-history row) the next attack step uses
  int adam_t;    // optimiser steps taken so far (transformers.AdamW state["step"])
  float lr;      // learning rate of the current outer iteration (LambdaLR value)
  int pad;
};
int patch_update_dev(float* patch, const float* grad, float* m, float* v, int n, const StepState* st, float beta1, float beta2,
                     float eps, int kind, float grad_scale, float clip_l1, float* scalars, float* zero_after, cudaStream_t s);
// copies placement st->place of the uploaded set into the "current placement" buffers the front-end kernels read
int step_begin(const StepState* st, const int* xy_all, const float* th_all, int* xy_cur, float* th_cur, int B, int n_place,
               float* scal_cur, cudaStream_t s);
// scal_hist[st->place] = scal_cur; st->place += 1; st->adam_t += adam_inc
int step_end(StepState* st, const float* scal_cur, float* scal_hist, int adam_inc, cudaStream_t s);
int accumulate_f32(float* acc, const float* g, int n, cudaStream_t s);

// ---- greedy decode with the KV cache (decode.cu) -----------------------------------------------------------
// dstate (nullable, device int[2] = {cache row of the token being processed, its column in the token buffer}): when given,
// the kernels read the position from it instead of the host argument, so that ONE recorded decode step serves every token
int embed_rows(const int* ids, const bf16* table, bf16* x, int B, int d, cudaStream_t s);
// qkv = a layer's cache [B*L, 3*H*hd]; row b*L + pos holds the new position's un-rotated q|k|v: q and k are rotated in place
// (tables [L, hd/2]) and the row attends to keys 0..pos; o [B, H*hd]
int attention_decode(bf16* qkv, bf16* o, const float* cos_tab, const float* sin_tab, int B, int L, int pos, int H, int hd, const int* dstate,
                     cudaStream_t s);
int argmax_rows(const float* logits, int R, int V, int* ids, int* out, int out_ld, int out_col, cudaStream_t s);
// closes a decode step (B <= 4): ids[b] = argmax, appended to out[b * out_ld + column]; x[b] = table[ids[b]]; dstate advances
int decode_select(const float* logits, int B, int V, int* ids, int* out, int out_ld, int out_col, int* dstate, const bf16* table, bf16* x,
                  int d, cudaStream_t s);
// skinny projection of the decode steps (decode.cu): M <= 4, HBM-bound weight streaming on the CUDA cores; optional fused
// RMSNorm of the activation rows (norm_w), residual, fp32 output, SwiGLU over interleaved gate|up weight rows, cache-row remap
bool gemv_supported(int M, int K, int64_t lda, int64_t ldw);
int gemv_bf16(const bf16* A, int64_t lda, const bf16* norm_w, float eps, const bf16* W, int64_t ldw, void* out, int64_t ldc, int M, int N,
              int K, const bf16* resid, int64_t ldr, int out_f32, int swiglu, int out_stride, int out_offset, const int* dstate,
              cudaStream_t s);
