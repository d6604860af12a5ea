// Host-side interface of the tcgen05 GEMM (see gemm_sm100.cu).
#pragma once
#include "common.cuh"

// Y[M,N] = epilogue( A[M,K] . W[N,K]^T ), A and W bf16 row-major with the contraction dim contiguous
// ("TN" GEMM: exactly what nn.Linear computes). fp32 accumulation in TMEM.
//
// Epilogue, in this order, mirroring where eager PyTorch materialises bf16 tensors:
//   x = acc (+ bias[n])          -> rounded to bf16   (the nn.Linear output)
//   if preact_out: preact_out[m,n] = x                 (saved for the GELU backward)
//   if act == 1:   x = bf16(gelu_erf(x))
//   if gamma:      x = bf16(x * gamma[n])              (timm LayerScale)
//   if resid:      x = bf16(resid[m,n] + x)            (residual stream)
//   out[m,n] = x   (bf16, or fp32 when out_f32 != 0: the HF `.float()` of bf16 logits)
struct GemmEpilogue {
  const bf16* bias = nullptr;
  const bf16* gamma = nullptr;
  const bf16* resid = nullptr;
  int64_t ldr = 0;
  int act = 0;
  bf16* preact_out = nullptr;  // same ld as out
  int out_f32 = 0;
  // output row remap: row r -> (r / out_group) * out_stride + out_offset + r % out_group  (0 = identity); lets the
  // patch-embed / projector GEMMs write straight into the token buffers ([cls,reg | patches], [BOS | patches | text])
  int out_group = 0, out_stride = 0, out_offset = 0;
  // resid row = r % resid_mod when > 0 (position embedding broadcast over the batch), else the output row
  int resid_mod = 0;

  // ---- fused elementwise neighbours (each replaces a separate HBM-bound kernel) ----
  // aux_mode 1 (GELU backward): out = bf16( bf16(acc) * gelu'(aux[m, n]) ), aux = saved pre-activation (ld = ldaux)
  // aux_mode 2 (SwiGLU backward): the GEMM computes d(act) [M, N]; aux = saved gate|up pre-activations in the interleaved
  //   layout below ([M, 2N]); `out` receives d(gate|up) in the same layout (ldc = 2N); d(act) itself is not stored.
  int aux_mode = 0;
  const bf16* aux = nullptr;
  int64_t ldaux = 0;
  // pair_mode works on column pairs (n, n + 64) inside each 128-column group (one epilogue warp owns both):
  // pair_mode 1 (RoPE, HF rotate_half convention, head dim 128): columns < rope_cols are rotated with the cos / sin
  //   tables [rope_L, 64] at position row % rope_L; the rest (v) is stored as is.
  // pair_mode 2 (SwiGLU forward): columns are [gate 64 | up 64] per group (weights packed that way): `out` gets the raw
  //   gate|up (saved for the backward) and act_out [M, N/2] gets bf16(bf16(silu(gate)) * up).
  int pair_mode = 0;
  const float* rope_cos = nullptr;
  const float* rope_sin = nullptr;
  int rope_L = 0, rope_cols = 0;
  bf16* act_out = nullptr;
  int64_t ld_act = 0;
  // delta_out != nullptr: the output is dO of an attention block with head dim 128 ([M = B * delta_L rows, N = H * 128]); the
  //   epilogue also writes delta[b, h, n] = sum_c bf16(dO[b, n, h, c]) * O[b, n, h, c] (fp32, [B, H, delta_L], what the
  //   attention backward needs) with O = aux (ld = ldaux).  An epilogue thread owns whole heads of its row, so the sum is
  //   a plain serial reduction: no atomics, no extra pass over dO and O.  Plain epilogue otherwise; 256-wide tiles only.
  float* delta_out = nullptr;
  int delta_L = 0;

  // ---- scheduling hint (not part of the maths) ----
  // W does not depend on the preceding kernels of the stream (a weight matrix): every CTA may request the W tiles of its first
  // pipeline fill BEFORE griddepcontrol.wait, i.e. while the previous kernel is still draining; A (and every epilogue tensor) is
  // only touched after the wait.  Leave 0 when W was produced by a kernel launched just before.
  int w_constant = 0;
};

// One problem of a launch.
struct GemmProblem {
  const bf16* A;
  int64_t lda;
  const bf16* W;
  int64_t ldw;
  void* out;
  int64_t ldc;
  int M, N, K;
  GemmEpilogue epi;
};
// Two independent problems of the same epilogue kind in ONE persistent launch (the DINOv2 and the SigLIP GEMM of the same
// depth): one prologue / pipeline fill / drain for both.  Results are bit-identical to two separate launches (same tiles,
// same accumulation order).
int gemm_bf16_tn_dual(const GemmProblem& p0, const GemmProblem& p1, cudaStream_t stream);

// Returns 0 on success. No allocation, no synchronisation; launches on `stream`.
int gemm_bf16_tn(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, void* out, int64_t ldc, int M, int N, int K,
                 const GemmEpilogue& epi, cudaStream_t stream);

// Persistent kernels (GEMM, tcgen05 attention) size their grids to min(work, SMs).  While the two independent vision towers
// run on two streams the engine sets this limit to half the SMs, so that one tower's persistent kernel leaves the other
// half of the machine to the other tower (otherwise a 1-CTA-per-SM kernel of ~200 KB shared memory serialises them) and
// the 1.x-wave ViT shapes lose less to wave quantisation.  0 = no limit.
extern int g_vla_sm_limit;

// number of GEMM kernel launches since process start (for bench.py's gpu_launches accounting)
extern long long g_vla_launch_count;

// true while vla_profile_gemm_begin/_end bracket per-launch event timing (such steps must not be graph-captured)
bool vla_gemm_profiling();
