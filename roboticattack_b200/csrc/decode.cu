// Single-position kernels of the greedy action decode with a KV cache (predict_action -> HF generate(do_sample=False),
// prismatic/extern/hf/modeling_prismatic.py:506-536): after the prefill pass the post-RoPE q|k|v rows of every decoder layer
// are still in the activation arena ([B*L, 3*H*hd], row b*L + position): they ARE the cache.  A decode step embeds the B new
// token ids, and per layer projects them (tcgen05 GEMM with M = B, the output row remapped into the cache row of the new
// position), rotates q and k at that position, attends the single query row to the cached keys / values, and runs the
// row-wise rest of the layer.  All of it stays on the device: the argmax of one step feeds the embedding of the next.
#include <math.h>

#include "kernels.h"

namespace {

// x[b, :] = table[ids[b], :]
__global__ void embed_rows_kernel(const int* __restrict__ ids, const bf16* __restrict__ table, bf16* __restrict__ x, int B, int d8) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * d8) return;
  const int b = idx / d8, c = idx - b * d8;
  reinterpret_cast<uint4*>(x)[static_cast<int64_t>(b) * d8 + c] =
      __ldg(reinterpret_cast<const uint4*>(table) + static_cast<int64_t>(ids[b]) * d8 + c);
}

// One CTA per (sample, head): scores of the new position's query against keys 0..pos (fp32 dot products of bf16 values,
// scaled), softmax in fp32, probabilities rounded to bf16 (as the flash kernels of the prefill do), o = sum_j p_j v_j in fp32
// -> bf16.  hd <= 128, hd % 32 == 0; dynamic shared memory: (pos + 1) floats.
constexpr int DEC_THREADS = 128;
__global__ void __launch_bounds__(DEC_THREADS) attn_decode_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, int L, int pos,
                                                                 int H, int hd, float scale, const int* __restrict__ dstate) {
  extern __shared__ float sc[];
  if (dstate) pos = dstate[0];   // graph-replayed decode step: the position lives in device memory
  __shared__ float red[32];
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ld = 3LL * H * hd;
  const bf16* base = qkv + static_cast<int64_t>(b) * L * ld;
  const bf16* q = base + static_cast<int64_t>(pos) * ld + h * hd;
  const int per = hd / 32;                       // elements of the head per lane (<= 4)
  float qr[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) qr[i] = i < per ? b2f(q[lane * per + i]) : 0.f;
  const int nk = pos + 1;
  float mx = -INFINITY;
  for (int j = warp; j < nk; j += DEC_THREADS / 32) {
    const bf16* k = base + static_cast<int64_t>(j) * ld + H * hd + h * hd;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < per) s = fmaf(qr[i], b2f(k[lane * per + i]), s);
    s = warp_sum(s) * scale;
    if (lane == 0) sc[j] = s;
    mx = fmaxf(mx, s);
  }
  // block max
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < DEC_THREADS / 32; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < nk; j += DEC_THREADS) {
    const float p = __expf(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  __syncthreads();
  if (threadIdx.x < hd) {
    const bf16* v = base + 2LL * H * hd + h * hd + threadIdx.x;
    float acc = 0.f;
    for (int j = 0; j < nk; ++j) acc = fmaf(rbf(sc[j] * inv), b2f(v[static_cast<int64_t>(j) * ld]), acc);
    o[(static_cast<int64_t>(b) * H + h) * hd + threadIdx.x] = f2b(acc);
  }
}

// ids[b] = argmax_v logits[b, v] (first maximum, as torch.argmax); optionally also appended to out[b * out_ld + out_col]
__global__ void __launch_bounds__(1024) argmax_rows_kernel(const float* __restrict__ logits, int V, int* __restrict__ ids,
                                                          int* __restrict__ out, int out_ld, int out_col, const int* __restrict__ dstate) {
  __shared__ float bv[32];
  if (dstate) out_col = dstate[1];
  __shared__ int bi[32];
  const float* row = logits + static_cast<int64_t>(blockIdx.x) * V;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float x = row[v];
    if (x > best || (x == best && v < arg)) {
      best = x;
      arg = v;
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
    if (ob > best || (ob == best && oa < arg)) {
      best = ob;
      arg = oa;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    bv[warp] = best;
    bi[warp] = arg;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
      if (bv[w] > best || (bv[w] == best && bi[w] < arg)) {
        best = bv[w];
        arg = bi[w];
      }
    ids[blockIdx.x] = arg;
    if (out) out[blockIdx.x * out_ld + out_col] = arg;
  }
}

}  // namespace

__global__ void decode_advance_kernel(int* dstate) {
  dstate[0] += 1;   // cache row of the next token
  dstate[1] += 1;   // its column in the token buffer
}
int decode_advance(int* dstate, cudaStream_t s) {
  decode_advance_kernel<<<1, 1, 0, s>>>(dstate);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

int embed_rows(const int* ids, const bf16* table, bf16* x, int B, int d, cudaStream_t s) {
  VLA_REQUIRE(d % 8 == 0, "embed_rows: d %% 8 != 0");
  const int n = B * (d / 8);
  embed_rows_kernel<<<ceil_div(n, 256), 256, 0, s>>>(ids, table, x, B, d / 8);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

int attention_decode(const bf16* qkv, bf16* o, int B, int L, int pos, int H, int hd, const int* dstate, cudaStream_t s) {
  VLA_REQUIRE(hd % 32 == 0 && hd <= 128, "attention_decode: head dim %d not supported (multiple of 32, <= 128)", hd);
  VLA_REQUIRE(dstate || (pos >= 0 && pos < L), "attention_decode: position %d outside the cache (%d rows)", pos, L);
  VLA_REQUIRE(L * sizeof(float) <= 48 * 1024, "attention_decode: cache of %d rows exceeds the score buffer", L);
  attn_decode_kernel<<<B * H, DEC_THREADS, L * sizeof(float), s>>>(qkv, o, L, pos, H, hd, 1.f / sqrtf(static_cast<float>(hd)), dstate);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

int argmax_rows(const float* logits, int R, int V, int* ids, int* out, int out_ld, int out_col, const int* dstate, cudaStream_t s) {
  argmax_rows_kernel<<<R, 1024, 0, s>>>(logits, V, ids, out, out_ld, out_col, dstate);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Skinny GEMM for the decode steps: out[m, n] = epilogue(sum_k A[m, k] * W[n, k]) with M <= 4 rows (M = batch of the
// closed-loop evaluation, 1).  Such a product streams the whole weight matrix for a handful of FLOPs per byte: it is bound by
// HBM, and a tensor-core tile loop (one barrier round trip and four tcgen05.mma issues per 16 KB of weights) reaches ~2 TB/s
// on it.  Here one warp owns one output column: it streams the column's weight row with 16-byte loads (8 loads in flight per
// lane), multiplies with the activation rows (L1-resident) and reduces with shuffles; fp32 accumulation, bf16 rounding points
// of the general GEMM epilogue (bias -> round -> residual -> round; fp32 output holds bf16 values).
// ---------------------------------------------------------------------------------------------------------
namespace {
constexpr int GV_WARPS = 8;
template <int MR>
__global__ void __launch_bounds__(GV_WARPS * 32) gemv_bf16_kernel(const bf16* __restrict__ A, int64_t lda, const bf16* __restrict__ W,
                                                                 int64_t ldw, void* __restrict__ out, int64_t ldc, int N, int K,
                                                                 const bf16* __restrict__ bias, const bf16* __restrict__ resid,
                                                                 int64_t ldr, int out_f32, int out_stride, int out_offset,
                                                                 const int* __restrict__ dstate) {
  if (dstate) out_offset = dstate[0];
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * GV_WARPS + (threadIdx.x >> 5);
  if (n >= N) return;
  const bf16* w = W + static_cast<int64_t>(n) * ldw;
  float acc[MR];
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = 0.f;
  const int K8 = K / 8;
  constexpr int UN = 8;
  for (int c0 = lane; c0 < K8; c0 += 32 * UN) {
    uint4 wv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int c = c0 + u * 32;
      wv[u] = c < K8 ? __ldcs(reinterpret_cast<const uint4*>(w) + c) : make_uint4(0u, 0u, 0u, 0u);   // streamed once: evict first
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int c = c0 + u * 32;
      if (c >= K8) break;
      const uint32_t ww[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        const uint4 av = __ldg(reinterpret_cast<const uint4*>(A + m * lda) + c);
        const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 a = unpack_bf16x2(aw[t]), b = unpack_bf16x2(ww[t]);
          acc[m] = fmaf(a.x, b.x, acc[m]);
          acc[m] = fmaf(a.y, b.y, acc[m]);
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = warp_sum(acc[m]);
  if (lane == 0) {
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      float v = acc[m];
      if (bias) v += b2f(bias[n]);
      v = rbf(v);
      if (resid) v = rbf(b2f(resid[m * ldr + n]) + v);
      const int64_t row = out_stride ? static_cast<int64_t>(m) * out_stride + out_offset : m;
      if (out_f32) static_cast<float*>(out)[row * ldc + n] = v;
      else static_cast<bf16*>(out)[row * ldc + n] = f2b(v);
    }
  }
}
}  // namespace

bool gemv_supported(int M, int K, int64_t lda, int64_t ldw) { return M >= 1 && M <= 4 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0; }

// out row m -> m * out_stride + out_offset when out_stride > 0 (the decode's cache-row remap), else m
int gemv_bf16(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, void* out, int64_t ldc, int M, int N, int K, const bf16* bias,
              const bf16* resid, int64_t ldr, int out_f32, int out_stride, int out_offset, const int* dstate, cudaStream_t s) {
  VLA_REQUIRE(gemv_supported(M, K, lda, ldw), "gemv: needs 1 <= M <= 4 and K, lda, ldw multiples of 8 (M=%d K=%d)", M, K);
  VLA_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "gemv: operands must be 16-byte aligned");
  const dim3 grid(ceil_div(N, GV_WARPS)), block(GV_WARPS * 32);
  switch (M) {
    case 1: gemv_bf16_kernel<1><<<grid, block, 0, s>>>(A, lda, W, ldw, out, ldc, N, K, bias, resid, ldr, out_f32, out_stride, out_offset, dstate); break;
    case 2: gemv_bf16_kernel<2><<<grid, block, 0, s>>>(A, lda, W, ldw, out, ldc, N, K, bias, resid, ldr, out_f32, out_stride, out_offset, dstate); break;
    case 3: gemv_bf16_kernel<3><<<grid, block, 0, s>>>(A, lda, W, ldw, out, ldc, N, K, bias, resid, ldr, out_f32, out_stride, out_offset, dstate); break;
    default: gemv_bf16_kernel<4><<<grid, block, 0, s>>>(A, lda, W, ldw, out, ldc, N, K, bias, resid, ldr, out_f32, out_stride, out_offset, dstate); break;
  }
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
