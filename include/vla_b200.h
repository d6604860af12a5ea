/* vla_b200.h -- C ABI of the B200 (sm_100a) adversarial-patch attack engine.
 *
 * Drop-in boundary for the inner loop of William-wAng618/roboticAttack's OpenVLAAttacker (file:line below are
 * relative to that repository @ a0bef502).  The reference is pure Python and has no FFI of its own; the entry
 * points here are what a binding for this path needs, one per reference call site.  Conventions:
 *   - every function returns 0 on success, non-zero on failure; vla_last_error() gives the message (thread local);
 *   - plain pointers and sizes only; device pointers unless a parameter says "host";
 *   - no function allocates device memory or synchronises the device, except the vla_engine_set_* staging calls,
 *     which synchronise the given stream once (they copy from pageable host staging);
 *   - all kernels are launched on the caller's stream (a cudaStream_t passed as void*);
 *   - one engine per process / GPU; an engine is not thread safe.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#define VLA_B200_ABI_VERSION 3

#ifdef __cplusplus
extern "C" {
#endif

const char* vla_last_error(void);
int vla_abi_version(void);
/* number of kernel launches issued by this library since process start */
long long vla_launch_count(void);

/* Measurement hook (bench.py roofline leg): while enabled, every tcgen05 GEMM launch is bracketed by CUDA events on
 * its launch stream; _end synchronises the device and returns the summed kernel time, the summed algorithmic FLOPs
 * (2*M*N*K with the true, unpadded dims) and the launch count. */
int vla_profile_gemm_begin(void);
/* pin one GEMM kernel variant: ctas in {1,2} (2 = tcgen05.mma.cta_group::2 CTA pairs), block_n in {128,256};
 * (0,0) restores the automatic per-shape choice.  For A/B measurements and the variant parity tests. */
int vla_gemm_set_mode(int ctas, int block_n);
/* autotune (default on): the FIRST call with a new (M,N,K) times every kernel variant on the real operands and keeps
 * the fastest for later calls; that first call synchronises the stream (warm-up only; never while capturing). */
int vla_gemm_set_autotune(int on);
int vla_profile_gemm_end(double* total_ms, double* total_flops, int* launches);

/* ------------------------------------------------------------------------------------------------------------
 * Front end.  Replaces RandomPatchTransform.apply_random_patch_batch (VLAAttacker/white_patch/
 * appply_random_transform.py:104-136), .paste_patch_fix (:160-188) / .random_paste_patch (:138-158), .im_process
 * (:190-197) plus the `.to(torch.bfloat16)` of UADA.py:142, and their autograd backward.
 *   obs    u8  [B,H,W,3]   clean observations (what torchvision ToTensor would read from the PIL images)
 *   patch  f32 [3,ph,pw]
 *   xy     i32 [B,2]       paste position (x, y) drawn by random.randint (:123-124)
 *   theta  f32 [B,2,3]     first two rows of S.R from combined_transform_matrix (:80-91); ignored unless WARP
 *   out    bf16 [B,6,H,W]  channels 0-2 DINOv2-normalised, 3-5 SigLIP-normalised (UADA.py:56-57)
 *   norm   f32 [12]        mean[2][3] then std[2][3]
 */
enum { VLA_FE_WARP = 0,    /* geometry=True : affine_grid + grid_sample(border), keep where canvas >= -20 (:127-131) */
       VLA_FE_PASTE20 = 1, /* geometry=False: no warp, same `canvas < -20` test */
       VLA_FE_FIX = 2,     /* paste_patch_fix / random_paste_patch: `canvas != -100` test (:153,:179) */
       VLA_FE_NONE = 3 };  /* im_process: no patch */
int vla_patch_frontend_fwd(const uint8_t* obs, const float* patch, const int32_t* xy, const float* theta, void* out_bf16,
                           int B, int H, int W, int ph, int pw, int mode, const float* norm, void* stream);
/* dpatch f32 [3,ph,pw] = d loss / d patch given dout bf16 [B,6,H,W]; overwrites dpatch */
int vla_patch_frontend_bwd(const void* dout_bf16, const float* patch, const int32_t* xy, const float* theta,
                           float* dpatch, int B, int H, int W, int ph, int pw, int mode, const float* norm, void* stream);
/* Eval-time paste: RandomPatchTransform.simulation_random_patch (appply_random_transform.py:43-78), the consumer-side
 * paste used by the closed-loop evaluation (experiments/robot/libero/run_libero_eval_args_geo_batch.py:207).
 * img / out uint8 [B,H,W,3] (device); patch f32 [3,ph,pw] in [0,1], quantised like ToPILImage (floor(p*255)); xy i32 [B,2]
 * = fixed (x, y); theta f32 [B,2,3] = (S.R)[:2] of the fixed (angle, shx, shy), read when geometry != 0.
 * out = canvas < 0 ? img : uint8(canvas). */
int vla_patch_sim_paste(const uint8_t* img, const float* patch, const int32_t* xy, const float* theta, uint8_t* out,
                        int B, int H, int W, int ph, int pw, int geometry, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Loss heads (UADA.py:145-147,381-418; UADA_ddp.py:99-136,203-206; UPA.py:145-150,367-387; TMA.py:148) */
enum { VLA_LOSS_UADA = 0,      /* mean((w e - w t)^2) + 1/CE,  w = mse_weight (5 in UADA.py:396) */
       VLA_LOSS_UADA_DDP = 1,  /* mean((w e - w t)^2),         w = MSE_weights (UADA_ddp.py:114) */
       VLA_LOSS_UPA = 2,       /* alpha*(cos+1).mean + belta/(mean||d|| + 1e-3) on the first 3 DoF */
       VLA_LOSS_CE = 3,        /* ce_scale * CE  (TMA: 1/accumulate_steps; UPA guide) */
       VLA_LOSS_NEG_CE = 4 };  /* -CE (UPA reverse_direction=False) */
typedef struct vla_loss_params {
  int kind;
  float mse_weight;
  float alpha, belta;
  float ce_scale;
} vla_loss_params;
/* indices into the per-step scalar record (device float[VLA_NUM_SCALARS]) */
enum { VLA_S_LOSS = 0, VLA_S_CE = 1, VLA_S_AUX0 = 2 /* UADA: MSE term; UPA: angle loss */, VLA_S_AUX1 = 3 /* UPA: distance loss */,
       VLA_S_UAD = 4, VLA_S_NTOK = 5, VLA_S_NACT = 6, VLA_S_GRAD_MEAN = 7 /* patch.grad.mean(), written by the update */,
       VLA_NUM_SCALARS = 8 };
/* logits f32 [R,V] of the R supervised rows; meta i32 [R,3] = {label, sample, index among the sample's supervised rows};
 * row_stats: scratch f32 [8*R]; dlogits bf16 [R,V]; pred_ids i32 [R] (argmax action id, -1 for non-action rows) */
int vla_loss_head(const float* logits, const int32_t* meta, int R, int V, int B, const vla_loss_params* lp,
                  float* row_stats, void* dlogits_bf16, float* scalars, int32_t* pred_ids, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Patch update.  transformers.AdamW(lr, betas=(0.9,0.999), eps=1e-6, wd=0) + clamp(0,1) (UADA.py:107-115,155-156),
 * with the optional clip_grad_norm_(max_norm=clip_l1, norm_type=1) of UPA.py:157; or sign-PGD (TMA.py:171-175).
 * grad is multiplied by grad_scale first (1/world_size after an all-reduce(sum): DDP's gradient mean). */
enum { VLA_OPT_ADAMW = 0, VLA_OPT_PGD = 1 };
int vla_patch_update(float* patch, const float* grad, float* exp_avg, float* exp_avg_sq, int n, int step, float lr,
                     float beta1, float beta2, float eps, int kind, float grad_scale, float clip_l1, float* scalars,
                     void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Building blocks, exported for the parity tests (each is checked against the oracle on its own). */
int vla_gemm_bf16_tn(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldc, int M, int N, int K,
                     const void* bias, const void* gamma, const void* resid, int64_t ldr, int act, void* preact_out,
                     int out_f32, void* stream);
/* Every fused epilogue of the tcgen05 GEMM (csrc/gemm.h: GemmEpilogue), for the per-variant x per-epilogue parity tests.
 * Zero-initialise and set what the mode needs; the kernel variant is pinned with vla_gemm_set_mode. */
typedef struct vla_gemm_epilogue {
  const void *bias, *gamma, *resid;   /* bf16 [N], [N], [M or resid_mod, ldr] */
  int64_t ldr;
  int act;                            /* 1 = erf GELU */
  void* preact_out;                   /* bf16 [M, ldc]: Linear output before the activation */
  int out_f32;
  int out_group, out_stride, out_offset, resid_mod;
  int aux_mode;                       /* 1 = GELU backward, 2 = SwiGLU backward (out = d(gate|up), ldc = 2N) */
  const void* aux;                    /* bf16: saved pre-activation / gate|up / attention output O */
  int64_t ldaux;
  int pair_mode;                      /* 1 = RoPE on columns < rope_cols, 2 = SwiGLU forward (act_out [M, N/2]) */
  const float *rope_cos, *rope_sin;   /* f32 [rope_L, 64] */
  int rope_L, rope_cols;
  void* act_out;
  int64_t ld_act;
  float* delta_out;                   /* f32 [M / delta_L, N / 128, delta_L]: rowsum(dO * O) per head; aux = O */
  int delta_L;
  int w_constant;                     /* scheduling hint: W is a constant of the stream (a weight matrix, not the output of a kernel
                                         launched just before): its first tiles may be requested before griddepcontrol.wait */
} vla_gemm_epilogue;
int vla_gemm_bf16_tn_ex(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldc, int M, int N, int K,
                        const vla_gemm_epilogue* ep, void* stream);
/* Skinny projection of the decode steps (M <= 4; HBM-bound weight streaming, csrc/decode.cu): out = epilogue(A' W^T) with
 * A' = A, or norm_w * bf16(A * rstd(A)) when norm_w is given (LlamaRMSNorm fused, transformers modeling_llama.py).
 * resid: out = bf16(resid + bf16(acc)); out_f32: fp32 output; swiglu: W rows are interleaved [gate 64 | up 64] groups and
 * out [M, N/2] = bf16(bf16(silu(g)) * u) (LlamaMLP).  Same rounding points as vla_gemm_bf16_tn_ex + vla_rmsnorm_fwd. */
int vla_gemv_bf16(const void* A, int64_t lda, const void* norm_w, float eps, const void* W, int64_t ldw, void* out, int64_t ldc, int M,
                  int N, int K, const void* resid, int64_t ldr, int out_f32, int swiglu, void* stream);
/* One decode position of causal attention over a KV cache, rotary embedding fused (csrc/decode.cu): qkv [B*L, 3*H*hd] bf16,
 * row b*L + pos holds the new position's un-rotated q|k|v; q and k of that row are rotated in place (cos/sin f32 [L, hd/2])
 * and o[b, h*hd..] = softmax(q k_j / sqrt(hd), j <= pos) v_j.  Replaces LlamaAttention.forward with past_key_values for one
 * new token (HF generate, modeling_prismatic.py:506-536).  The kernel is a programmatic dependent launch that requests the
 * cache rows BELOW pos before griddepcontrol.wait: those rows must be complete before the kernel launched immediately before
 * this one started (true for a cache: they were written by earlier decode steps / the prefill); row pos itself may come from
 * the preceding kernel. */
int vla_attention_decode(void* qkv, void* o, const float* cos_tab, const float* sin_tab, int B, int L, int pos, int H, int hd, void* stream);
int vla_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t M, int d,
                      float eps, void* stream);
int vla_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd,
                      const void* dres, void* dx, int64_t M, int d, void* stream);
int vla_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int64_t M, int d, float eps, void* stream);
int vla_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx,
                    int64_t M, int d, void* stream);
/* qkv bf16 [B*N, 3*H*hd]; o bf16 [B*N, H*hd]; lse f32 [B,H,N]; kv_len i32 [B] or NULL */
int vla_attention_fwd(const void* qkv, void* o, float* lse, const int32_t* kv_len, int B, int N, int H, int hd,
                      int causal, void* stream);
int vla_attention_bwd(const void* qkv, const void* o, const void* dout, const float* lse, float* delta_scratch,
                      void* dqkv, const int32_t* kv_len, int B, int N, int H, int hd, int causal, void* stream);
/* Same, with the backward of the rotary embedding fused (head dim 128): d(q), d(k) are returned wrt the PRE-RoPE
 * projections.  cos_tab / sin_tab f32 [rope_L, 64]; position of row n = n % rope_L
 * (replaces autograd through apply_rotary_pos_emb, transformers modeling_llama.py). */
int vla_attention_bwd_rope(const void* qkv, const void* o, const void* dout, const float* lse, float* delta_scratch,
                           void* dqkv, const int32_t* kv_len, int B, int N, int H, int hd, int causal,
                           const float* cos_tab, const float* sin_tab, int rope_L, void* stream);
/* Kernel selection, bit mask: bit 0 = persistent tcgen05 forward (hd 64/72/128), bit 1 = persistent tcgen05 backward
 * (hd 64/72/128).  Default 3; 0 = pipelined mma.sync kernels only (also the fallback for other head dims).
 * Env VLA_ATTN_IMPL sets the default. */
int vla_attention_set_impl(int impl);
int vla_rope_inplace(void* qkv, const float* cos_tab, const float* sin_tab, int64_t M, int L, int H, int hd, int dir,
                     void* stream);
int vla_swiglu_fwd(const void* gu, void* act, int64_t M, int F, void* stream);
int vla_swiglu_bwd(const void* dact, const void* gu, void* dgu, int64_t M, int F, void* stream);
int vla_gelu_bwd(const void* dy, const void* pre, void* dx, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Engine: forward + input-gradient of OpenVLAForActionPrediction for one attack iteration.
 * Replaces `self.vla(input_ids, attention_mask, pixel_values, labels)` + loss + `.backward()` of
 * UADA.py:139-148 / UADA_ddp.py:196-206 / UPA.py:139-152 / TMA.py:142-162 (prismatic/extern/hf/
 * modeling_prismatic.py:362-415 and the timm / HF-Llama bodies behind it). */
typedef struct vla_config {
  int img, patch;                                            /* 224, 14 */
  int dino_dim, dino_depth, dino_heads, dino_mlp, dino_prefix, dino_layerscale;   /* 1024,24,16,4096,5,1 */
  int sig_dim, sig_depth, sig_heads, sig_mlp, sig_prefix, sig_layerscale;         /* 1152,27,16,4304,0,0 */
  float vit_ln_eps;                                          /* 1e-6 */
  int llm_hidden, llm_layers, llm_heads, llm_ffn, vocab;     /* 4096,32,32,11008,32064 */
  float rms_eps;                                             /* 1e-6 */
  float norm_mean[2][3], norm_std[2][3];                     /* UADA.py:56-57 */
} vla_config;
typedef struct vla_engine vla_engine;

int vla_engine_create(const vla_config* cfg, vla_engine** out);
void vla_engine_destroy(vla_engine* e);
/* bytes of the weight arena (forward copies + the transposed copies the input-gradient GEMMs read) */
size_t vla_engine_weight_bytes(const vla_engine* e);
/* bytes of the activation arena for per-GPU batch B and text length T (sequence L = T + num_patches) */
size_t vla_engine_workspace_bytes(vla_engine* e, int B, int T);
int vla_engine_set_buffers(vla_engine* e, void* weight_arena, size_t weight_bytes, void* workspace,
                           size_t workspace_bytes, int B, int T);
/* one HF-checkpoint tensor (bf16, device), by its checkpoint name */
int vla_engine_load_weight(vla_engine* e, const char* name, const void* src_bf16, int64_t numel, void* stream);
int vla_engine_weights_ready(vla_engine* e);
/* cos/sin f32 [L, head_dim/2] (host), as HF LlamaRotaryEmbedding computes them (values rounded to bf16) */
int vla_engine_set_rope(vla_engine* e, const float* cos_host, const float* sin_host, int L, void* stream);
/* one outer iteration's batch, as PaddedCollatorForActionPrediction yields it (prismatic/util/data_utils.py:183-217):
 * obs u8 [B,H,W,3] (host unless obs_on_device), input_ids i64 [B,T] host, attention_mask u8 [B,T] host (right padded),
 * labels i64 [B,T] host (-100 = ignore; already masked by mask_labels / target substitution) */
int vla_engine_set_batch(vla_engine* e, const uint8_t* obs, int obs_on_device, const int64_t* input_ids,
                         const uint8_t* attention_mask, const int64_t* labels, int B, int T, void* stream);
/* placements of the next nsteps inner iterations: xy i32 [nsteps,B,2], theta f32 [nsteps,B,2,3] (host) */
int vla_engine_set_placements(vla_engine* e, const int32_t* xy_host, const float* theta_host, int nsteps, void* stream);
int vla_engine_num_supervised(const vla_engine* e);
/* ids i32 [num_supervised] (device): argmax over the FULL vocabulary of every supervised logits row of the last pass -- the
 * `action_preds = logits.argmax(dim=2)` the reference's metrics are built on (UADA.py:168,229; TMA.py:150,274); pred_ids of
 * vla_fwd_bwd is the argmax inside the 256 action classes (what weighted_loss / UAD use) */
int vla_engine_full_vocab_pred(vla_engine* e, int32_t* dst, void* stream);
/* The DINOv2 and SigLIP towers are independent until the feature concat; by default the SigLIP tower runs on an
 * engine-owned side stream (fork / join with events on the caller's stream).  on = 1 keeps everything on the caller's
 * stream (used for per-kernel timing). */
int vla_engine_set_single_stream(vla_engine* e, int on);

enum { VLA_FLAG_FORWARD_ONLY = 1 };   /* validation pass: loss heads + metrics, no backward */
/* patch f32 [3,ph,pw] -> dpatch f32 [3,ph,pw], scalars f32 [VLA_NUM_SCALARS], pred_ids i32 [num_supervised] */
int vla_fwd_bwd(vla_engine* e, const float* patch, int ph, int pw, int step_idx, int fe_mode,
                const vla_loss_params* lp, float* dpatch, float* scalars, int32_t* pred_ids, int flags, void* stream);
/* ------------------------------------------------------------------------------------------------------------
 * Patch-gradient exchange of the data-parallel attack (UADA_ddp.py:206: DDP's all-reduce of patch.grad on every backward).
 * A vla_comm wraps an NCCL communicator (libnccl.so.2 is resolved at run time; one communicator per process / GPU).
 * Rank 0 makes the 128-byte id, the host side distributes it (any channel), every rank calls vla_comm_create on its GPU. */
typedef struct vla_comm vla_comm;
#define VLA_COMM_ID_BYTES 128
int vla_comm_unique_id(void* id_host);
int vla_comm_create(const void* id_host, int rank, int world, vla_comm** out);
void vla_comm_destroy(vla_comm* c);
int vla_comm_world(const vla_comm* c);
/* in-place sum over ranks of dpatch f32 [n] on `stream` (ncclAllReduce); the 1/world of DDP's mean is folded into
 * vla_patch_update's grad_scale */
int vla_allreduce_patch_grad(vla_comm* c, float* dpatch, int n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * One whole attack iteration = the body of the reference's inner loop (UADA.py:134-158, UADA_ddp.py:192-209,
 * UPA.py:134-160, TMA.py:133-175) as ONE call: front end -> model forward -> loss head -> backward to the patch ->
 * [accumulate] -> [all-reduce over ranks] -> [clip] -> AdamW / sign-PGD -> clamp.
 * The placement index, the optimiser's step counter t and the learning rate live in device memory (so that the whole
 * iteration is one replayable CUDA graph): step k of an outer iteration reads placement k of vla_engine_set_placements,
 * writes its scalar record to scalars_hist[k] and advances the counters on the device.  vla_engine_set_step_state sets
 * them (start of an outer iteration: place = 0; resume: adam_step = restored value).
 * The first call with a new (plan, patch size, modes, buffers) runs eagerly (it also lets the GEMM autotuner see the
 * shapes), the second one records the same launch sequence into a CUDA graph, later ones replay it: one graph launch per
 * attack iteration.  VLA_STEP_NO_GRAPH keeps it eager (bit-identical results: same kernels, same order). */
enum { VLA_STEP_NO_GRAPH = 1, VLA_STEP_NO_UPDATE = 2 /* accumulate-only iteration (accumulate_steps > 1) */ };
typedef struct vla_step_params {
  int ph, pw, fe_mode;
  vla_loss_params loss;
  int opt_kind;                    /* VLA_OPT_* */
  float lr, beta1, beta2, eps;     /* transformers.AdamW defaults: 0.9, 0.999, 1e-6 */
  float clip_l1;                   /* > 0: clip_grad_norm_(max_norm, norm_type=1) of UPA.py:157 */
  int flags;                       /* VLA_STEP_* */
} vla_step_params;
int vla_engine_set_step_state(vla_engine* e, int placement_index, int adam_step, void* stream);
/* drops the recorded graphs (re-recorded on demand); call before vla_comm_destroy of a communicator they used: NCCL blocks the
 * destruction of a communicator while a graph still holds its collectives */
int vla_engine_drop_graphs(vla_engine* e);
/* host mirror of the device counters (no synchronisation) */
int vla_engine_get_step_state(const vla_engine* e, int* placement_index, int* adam_step);
/* patch / exp_avg / exp_avg_sq / dpatch f32 [3,ph,pw]; accumulate f32 [3,ph,pw] or NULL; comm NULL = single GPU;
 * scalars_hist f32 [>= placements, VLA_NUM_SCALARS]; pred_ids i32 [num_supervised] */
int vla_attack_step(vla_engine* e, float* patch, float* exp_avg, float* exp_avg_sq, float* dpatch, float* accumulate,
                    const vla_step_params* sp, vla_comm* comm, float* scalars_hist, int32_t* pred_ids, void* stream);
/* bookkeeping for bench.py: attack steps replayed from a graph so far / kernel nodes of the most recent graph */
long long vla_graph_replays(void);
int vla_graph_kernel_nodes(const vla_engine* e);

/* ------------------------------------------------------------------------------------------------------------
 * Greedy action decode with a KV cache.  Replaces `self.generate(input_ids, max_new_tokens=action_dim)` inside
 * OpenVLAForActionPrediction.predict_action (prismatic/extern/hf/modeling_prismatic.py:506-536; consumer:
 * experiments/robot/libero/run_libero_eval_args_geo_batch.py:105-228).
 * Prefill = one vla_fwd_bwd(VLA_FLAG_FORWARD_ONLY) over prompts of equal length `prompt_len` (BOS included) whose only
 * supervised row per sample is the last prompt position (set_batch: labels[b, prompt_len] != -100 on a plan with
 * T >= prompt_len + n_tokens); its post-RoPE k / v rows stay in the activation arena and are the cache.  Then n_tokens - 1
 * single-position steps: embed the previous argmax, project q|k|v into the cache row, attend to the cached rows, MLP, final norm,
 * lm_head, argmax over the full vocabulary -- all on the device, nothing read back in between.
 * tokens i32 [B, n_tokens] (device). */
int vla_engine_decode_greedy(vla_engine* e, int prompt_len, int n_tokens, int32_t* tokens, void* stream);

/* test tap: copies a named internal activation ("px", "dino_out", "llm_out", "logits", ...) to dst (device) */
int64_t vla_engine_debug_tap(vla_engine* e, const char* what, void* dst, int64_t max_bytes, void* stream);

#ifdef __cplusplus
}
#endif
