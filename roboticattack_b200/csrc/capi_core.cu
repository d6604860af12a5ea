// Error state + the building-block entry points of include/vla_b200.h (thin casts onto kernels.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/vla_b200.h"
#include "kernels.h"

static thread_local char g_err[1024] = "";

static int init_pdl() {
  const char* s = getenv("VLA_PDL");
  return (s && atoi(s) == 0) ? 0 : 1;
}
int g_vla_pdl = init_pdl();

bool vla_ablated(const char* what) {
  static const char* env = getenv("VLA_ABLATE");
  if (!env || !*env) return false;
  const char* hit = strstr(env, what);
  if (!hit) return false;
  const char end = hit[strlen(what)];
  return (hit == env || hit[-1] == ',') && (end == '\0' || end == ',');
}

bool vla_doubled(const char* what) {
  static const char* env = getenv("VLA_DOUBLE");
  if (!env || !*env) return false;
  const char* hit = strstr(env, what);
  if (!hit) return false;
  const char end = hit[strlen(what)];
  return (hit == env || hit[-1] == ',') && (end == '\0' || end == ',');
}

void vla_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static_assert(VLA_FE_WARP == FE_MODE_WARP && VLA_FE_NONE == FE_MODE_NONE, "front-end mode enums diverged");
static_assert(VLA_LOSS_UADA == LOSS_UADA && VLA_LOSS_NEG_CE == LOSS_NEG_CE, "loss enums diverged");
static_assert(VLA_NUM_SCALARS == LOSS_NUM_SCALARS && VLA_S_GRAD_MEAN == LS_GRAD_MEAN, "scalar enums diverged");
static_assert(VLA_OPT_PGD == OPT_PGD, "optimiser enums diverged");

#define S(x) static_cast<cudaStream_t>(x)
#define BF(x) static_cast<bf16*>(x)
#define CBF(x) static_cast<const bf16*>(x)

static FrontendNorm make_norm(const float* n) {
  FrontendNorm r;
  for (int s = 0; s < 2; ++s)
    for (int c = 0; c < 3; ++c) {
      r.mean[s][c] = n[s * 3 + c];
      r.std[s][c] = n[6 + s * 3 + c];
    }
  return r;
}

extern "C" {

const char* vla_last_error(void) { return g_err; }
int vla_abi_version(void) { return VLA_B200_ABI_VERSION; }
long long vla_launch_count(void) { return g_vla_launch_count; }

int vla_patch_frontend_fwd(const uint8_t* obs, const float* patch, const int32_t* xy, const float* theta, void* out, int B,
                           int H, int W, int ph, int pw, int mode, const float* norm, void* stream) {
  VLA_REQUIRE(obs && out && norm, "vla_patch_frontend_fwd: null argument");
  return patch_frontend_fwd(obs, patch, xy, theta, BF(out), B, H, W, ph, pw, mode, make_norm(norm), S(stream));
}
int vla_patch_frontend_bwd(const void* dout, const float* patch, const int32_t* xy, const float* theta, float* dpatch,
                           int B, int H, int W, int ph, int pw, int mode, const float* norm, void* stream) {
  VLA_REQUIRE(dout && dpatch && norm, "vla_patch_frontend_bwd: null argument");
  return patch_frontend_bwd(CBF(dout), patch, xy, theta, dpatch, B, H, W, ph, pw, mode, make_norm(norm), S(stream));
}
int vla_patch_sim_paste(const uint8_t* img, const float* patch, const int32_t* xy, const float* theta, uint8_t* out, int B, int H,
                        int W, int ph, int pw, int geometry, void* stream) {
  VLA_REQUIRE(img && patch && xy && out && (theta || !geometry), "vla_patch_sim_paste: null argument");
  return patch_sim_paste(img, patch, xy, theta, out, B, H, W, ph, pw, geometry, S(stream));
}
int vla_loss_head(const float* logits, const int32_t* meta, int R, int V, int B, const vla_loss_params* lp, float* row_stats,
                  void* dlogits, float* scalars, int32_t* pred_ids, void* stream) {
  VLA_REQUIRE(logits && meta && lp && row_stats && dlogits && scalars && pred_ids, "vla_loss_head: null argument");
  LossParams p{lp->kind, lp->mse_weight, lp->alpha, lp->belta, lp->ce_scale};
  return loss_head_fwd_bwd(logits, meta, R, V, B, p, row_stats, BF(dlogits), scalars, pred_ids, S(stream));
}
int vla_patch_update(float* patch, const float* grad, float* m, float* v, int n, int step, float lr, float beta1, float beta2,
                     float eps, int kind, float grad_scale, float clip_l1, float* scalars, void* stream) {
  VLA_REQUIRE(patch && grad, "vla_patch_update: null argument");
  return patch_update(patch, grad, m, v, n, step, lr, beta1, beta2, eps, kind, grad_scale, clip_l1, scalars, S(stream));
}
int vla_gemm_bf16_tn(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldc, int M, int N, int K,
                     const void* bias, const void* gamma, const void* resid, int64_t ldr, int act, void* preact_out,
                     int out_f32, void* stream) {
  GemmEpilogue e;
  e.bias = CBF(bias);
  e.gamma = CBF(gamma);
  e.resid = CBF(resid);
  e.ldr = ldr;
  e.act = act;
  e.preact_out = BF(preact_out);
  e.out_f32 = out_f32;
  return gemm_bf16_tn(CBF(A), lda, CBF(W), ldw, out, ldc, M, N, K, e, S(stream));
}
int vla_gemv_bf16(const void* A, int64_t lda, const void* norm_w, float eps, const void* W, int64_t ldw, void* out, int64_t ldc, int M,
                  int N, int K, const void* resid, int64_t ldr, int out_f32, int swiglu, void* stream) {
  return gemv_bf16(CBF(A), lda, CBF(norm_w), eps, CBF(W), ldw, out, ldc, M, N, K, CBF(resid), ldr, out_f32, swiglu, 0, 0, nullptr, S(stream));
}
int vla_attention_decode(void* qkv, void* o, const float* cos_tab, const float* sin_tab, int B, int L, int pos, int H, int hd, void* stream) {
  return attention_decode(BF(qkv), BF(o), cos_tab, sin_tab, B, L, pos, H, hd, nullptr, S(stream));
}
int vla_gemm_bf16_tn_ex(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldc, int M, int N, int K,
                        const vla_gemm_epilogue* ep, void* stream) {
  VLA_REQUIRE(ep != nullptr, "vla_gemm_bf16_tn_ex: null epilogue");
  GemmEpilogue e;
  e.bias = CBF(ep->bias);
  e.gamma = CBF(ep->gamma);
  e.resid = CBF(ep->resid);
  e.ldr = ep->ldr;
  e.act = ep->act;
  e.preact_out = BF(ep->preact_out);
  e.out_f32 = ep->out_f32;
  e.out_group = ep->out_group;
  e.out_stride = ep->out_stride;
  e.out_offset = ep->out_offset;
  e.resid_mod = ep->resid_mod;
  e.aux_mode = ep->aux_mode;
  e.aux = CBF(ep->aux);
  e.ldaux = ep->ldaux;
  e.pair_mode = ep->pair_mode;
  e.rope_cos = ep->rope_cos;
  e.rope_sin = ep->rope_sin;
  e.rope_L = ep->rope_L;
  e.rope_cols = ep->rope_cols;
  e.act_out = BF(ep->act_out);
  e.ld_act = ep->ld_act;
  e.delta_out = ep->delta_out;
  e.delta_L = ep->delta_L;
  e.w_constant = ep->w_constant;
  return gemm_bf16_tn(CBF(A), lda, CBF(W), ldw, out, ldc, M, N, K, e, S(stream));
}
int vla_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t M, int d,
                      float eps, void* stream) {
  return layernorm_fwd(CBF(x), CBF(w), CBF(b), BF(y), mean, rstd, M, d, eps, S(stream));
}
int vla_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, const void* dres,
                      void* dx, int64_t M, int d, void* stream) {
  return layernorm_bwd(CBF(dy), CBF(x), CBF(w), mean, rstd, CBF(dres), BF(dx), M, d, S(stream));
}
int vla_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int64_t M, int d, float eps, void* stream) {
  return rmsnorm_fwd(CBF(x), CBF(w), BF(y), rstd, M, d, eps, S(stream));
}
int vla_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx, int64_t M,
                    int d, void* stream) {
  return rmsnorm_bwd(CBF(dy), CBF(x), CBF(w), rstd, CBF(dres), BF(dx), M, d, S(stream));
}
int vla_attention_fwd(const void* qkv, void* o, float* lse, const int32_t* kv_len, int B, int N, int H, int hd, int causal,
                      void* stream) {
  return attention_fwd(CBF(qkv), BF(o), lse, kv_len, B, N, H, hd, causal, S(stream));
}
int vla_attention_bwd(const void* qkv, const void* o, const void* dout, const float* lse, float* delta, void* dqkv,
                      const int32_t* kv_len, int B, int N, int H, int hd, int causal, void* stream) {
  return attention_bwd(CBF(qkv), CBF(o), CBF(dout), lse, delta, BF(dqkv), kv_len, B, N, H, hd, causal, nullptr, nullptr, 0,
                       S(stream));
}
int vla_attention_bwd_rope(const void* qkv, const void* o, const void* dout, const float* lse, float* delta, void* dqkv,
                           const int32_t* kv_len, int B, int N, int H, int hd, int causal, const float* cos_tab,
                           const float* sin_tab, int rope_L, void* stream) {
  return attention_bwd(CBF(qkv), CBF(o), CBF(dout), lse, delta, BF(dqkv), kv_len, B, N, H, hd, causal, cos_tab, sin_tab, rope_L,
                       S(stream));
}
int vla_attention_set_impl(int impl) {
  VLA_REQUIRE(impl >= 0 && impl <= 3, "vla_attention_set_impl: bit 0 = tcgen05 forward, bit 1 = tcgen05 backward (0 = legacy mma.sync only)");
  g_attn_impl = impl;
  return 0;
}
int vla_rope_inplace(void* qkv, const float* cos_tab, const float* sin_tab, int64_t M, int L, int H, int hd, int dir,
                     void* stream) {
  return rope_inplace(BF(qkv), cos_tab, sin_tab, M, L, H, hd, dir, S(stream));
}
int vla_swiglu_fwd(const void* gu, void* act, int64_t M, int F, void* stream) { return swiglu_fwd(CBF(gu), BF(act), M, F, S(stream)); }
int vla_swiglu_bwd(const void* dact, const void* gu, void* dgu, int64_t M, int F, void* stream) {
  return swiglu_bwd(CBF(dact), CBF(gu), BF(dgu), M, F, S(stream));
}
int vla_gelu_bwd(const void* dy, const void* pre, void* dx, int64_t n, void* stream) {
  return gelu_bwd(CBF(dy), CBF(pre), BF(dx), n, S(stream));
}

}  // extern "C"
