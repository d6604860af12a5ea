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
};

// Returns 0 on success. No allocation, no synchronisation; launches on `stream`.
int gemm_bf16_tn(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, void* out, int64_t ldc, int M, int N, int K,
                 const GemmEpilogue& epi, cudaStream_t stream);

// number of GEMM kernel launches since process start (for bench.py's gpu_launches accounting)
extern long long g_vla_launch_count;
