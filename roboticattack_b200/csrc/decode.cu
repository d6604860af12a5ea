// Single-position kernels of the greedy action decode with a KV cache (predict_action -> HF generate(do_sample=False),
// prismatic/extern/hf/modeling_prismatic.py:506-536): after the prefill pass the post-RoPE q|k|v rows of every decoder layer
// are still in the activation arena ([B*L, 3*H*hd], row b*L + position): they ARE the cache.  A decode step streams every
// decoder weight once for B <= 4 rows of activations: it is bound by HBM (13.2 GB of bf16 weights per token at OpenVLA-7B), so
// the step is built to keep HBM busy from its first kernel to its last:
//   * five kernels per layer -- [RMSNorm + q|k|v projection into the cache row], [RoPE + attention over the cache],
//     [o projection + residual], [RMSNorm + gate|up projection + SwiGLU], [down projection + residual] -- then
//     [final norm + LM head] and [argmax + next embedding + position advance];
//   * every kernel is launched with programmatic dependent launch and issues its first weight loads BEFORE
//     griddepcontrol.wait: weights do not depend on the previous kernel, so the next projection's stream is already in flight
//     while the previous kernel drains (launch latency, ramp and tail of ~160 short kernels per token are hidden);
//   * the position lives in device memory, so one recorded step is replayed for every token of every action.
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

// ---------------------------------------------------------------------------------------------------------
// Attention of the new position over the cache, rotary embedding fused.  One CTA per (sample, head), 32 warps; a lane holds
// PER = hd / 32 consecutive dims of the head, so the rotate_half partner (dim +- hd/2) sits in lane ^ 16.
//   q, k of the new row arrive un-rotated from the projection: every warp rotates q in registers, the warp that owns key
//   `pos` rotates k; both are written back, so the cache row is post-RoPE like the prefill's rows.
//   scores: warp w takes keys w, w + 32, ... -- 10 rows of K and of V in flight per lane, i.e. a cache of <= 320 rows (256
//   patches + the prompt + the action) is requested in ONE round; fp32 dot products of bf16 values, scaled;
//   softmax in fp32, probabilities rounded to bf16 (as the flash kernels of the prefill do);
//   o = sum_j p_j v_j: the same key partition, fp32 partial sums per warp, reduced over the warps in a fixed order.
// Prologue before griddepcontrol.wait: rows below `pos` were written by earlier steps (earlier graph launches / the prefill),
// the position itself by the previous step's select kernel, the rotary tables at plan time -- all complete before this
// step's first kernel started, so their loads are issued while the q|k|v projection is still draining; only the new row
// is read after the wait.  Dynamic shared memory: L floats (the scores).
// ---------------------------------------------------------------------------------------------------------
constexpr int AD_WARPS = 32, AD_THREADS = AD_WARPS * 32, AD_UNR = 10;

template <int PER> struct RawDims;
template <> struct RawDims<4> { typedef uint2 type; };
template <> struct RawDims<2> { typedef uint32_t type; };
template <> struct RawDims<1> { typedef unsigned short type; };

template <int PER>
__device__ __forceinline__ typename RawDims<PER>::type load_raw(const bf16* p) {
  return __ldcg(reinterpret_cast<const typename RawDims<PER>::type*>(p));
}
__device__ __forceinline__ void unpack_raw(const uint2& v, float (&x)[4]) {
  x[0] = __uint_as_float(v.x << 16), x[1] = __uint_as_float(v.x & 0xFFFF0000u);
  x[2] = __uint_as_float(v.y << 16), x[3] = __uint_as_float(v.y & 0xFFFF0000u);
}
__device__ __forceinline__ void unpack_raw(const uint32_t& v, float (&x)[2]) {
  x[0] = __uint_as_float(v << 16), x[1] = __uint_as_float(v & 0xFFFF0000u);
}
__device__ __forceinline__ void unpack_raw(const unsigned short& v, float (&x)[1]) { x[0] = __uint_as_float(static_cast<uint32_t>(v) << 16); }
__device__ __forceinline__ uint2 pack_raw(const float (&x)[4]) { return make_uint2(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3])); }
__device__ __forceinline__ uint32_t pack_raw(const float (&x)[2]) { return pack_bf16x2(x[0], x[1]); }
__device__ __forceinline__ unsigned short pack_raw(const float (&x)[1]) { return __bfloat16_as_ushort(f2b(x[0])); }

// x (dims lane*PER .. +PER of one head) rotated with this lane's table entries cs / sn; the result holds bf16 values.
// Rounding points of the elementwise RoPE kernel: q*cos and rotate_half(q)*sin are bf16 tensors, so is their sum.
template <int PER>
__device__ __forceinline__ void rope_dims(float (&x)[PER], const float (&cs)[PER], const float (&sn)[PER], int lane) {
  const float sgn = lane < 16 ? -1.f : 1.f;   // rotate_half: (-x2, x1)
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const float partner = __shfl_xor_sync(0xffffffffu, x[i], 16);
    x[i] = rbf(rbf(x[i] * cs[i]) + rbf(sgn * partner * sn[i]));
  }
}

template <int PER>
__global__ void __launch_bounds__(AD_THREADS) attn_decode_kernel(bf16* __restrict__ qkv, bf16* __restrict__ o, const float* __restrict__ cos_tab,
                                                                 const float* __restrict__ sin_tab, int L, int pos, int H, float scale,
                                                                 const int* __restrict__ dstate) {
  typedef typename RawDims<PER>::type raw_t;
  extern __shared__ float sc[];
  __shared__ float red[32];
  __shared__ float pv[AD_WARPS][32 * PER];
  pdl_trigger();
  if (dstate) pos = __ldcg(dstate);   // graph-replayed decode step: the position lives in device memory
  constexpr int hd = 32 * PER;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ld = 3 * H * hd;   // L * ld < 2^31 (checked by the launcher): 32-bit row offsets
  bf16* base = qkv + static_cast<int64_t>(b) * L * ld + h * hd + lane * PER;   // this lane's dims of q in cache row 0
  const bf16* kbase = base + H * hd;
  const bf16* vbase = base + 2 * H * hd;
  const int nk = pos + 1;   // keys 0..pos
  raw_t kr[AD_UNR], vr[AD_UNR];
#pragma unroll
  for (int u = 0; u < AD_UNR; ++u) {   // first round, cached rows: before the wait
    const int j = warp + u * AD_WARPS;
    if (j < pos) {
      kr[u] = load_raw<PER>(kbase + j * ld);
      vr[u] = load_raw<PER>(vbase + j * ld);
    }
  }
  float cs[PER], sn[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    cs[i] = __ldg(cos_tab + static_cast<int64_t>(pos) * (hd / 2) + (lane & 15) * PER + i);
    sn[i] = __ldg(sin_tab + static_cast<int64_t>(pos) * (hd / 2) + (lane & 15) * PER + i);
  }
  pdl_wait();
  float q[PER];
  unpack_raw(load_raw<PER>(base + pos * ld), q);
  rope_dims<PER>(q, cs, sn, lane);
  if (warp == 0) *reinterpret_cast<raw_t*>(base + pos * ld) = pack_raw(q);
  float mx = -INFINITY;
  // the new key / value (row pos), by the warp its index falls to: rotate k, write it back to the cache, score it
  const bool owner = warp == (pos & (AD_WARPS - 1));
  raw_t vnew = raw_t();
  if (owner) {
    float kf[PER];
    unpack_raw(load_raw<PER>(kbase + pos * ld), kf);
    vnew = load_raw<PER>(vbase + pos * ld);
    rope_dims<PER>(kf, cs, sn, lane);
    *reinterpret_cast<raw_t*>(base + H * hd + pos * ld) = pack_raw(kf);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) s = fmaf(q[i], kf[i], s);
    s = warp_sum(s) * scale;
    if (lane == 0) sc[pos] = s;
    mx = s;
  }
  for (int j0 = warp; j0 < pos; j0 += AD_WARPS * AD_UNR) {   // cached keys
    if (j0 != warp) {
#pragma unroll
      for (int u = 0; u < AD_UNR; ++u) {
        const int j = j0 + u * AD_WARPS;
        if (j < pos) kr[u] = load_raw<PER>(kbase + j * ld);
      }
    }
#pragma unroll
    for (int u = 0; u < AD_UNR; ++u) {
      const int j = j0 + u * AD_WARPS;
      if (j >= pos) break;
      float kf[PER];
      unpack_raw(kr[u], kf);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) s = fmaf(q[i], kf[i], s);
      s = warp_sum(s) * scale;
      if (lane == 0) sc[j] = s;
      mx = fmaxf(mx, s);
    }
  }
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < AD_WARPS; ++w) mx = fmaxf(mx, red[w]);
  float sum = 0.f;
  for (int j = threadIdx.x; j < nk; j += AD_THREADS) {
    const float p = __expf(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  float acc[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) acc[i] = 0.f;
  for (int j0 = warp; j0 < pos; j0 += AD_WARPS * AD_UNR) {
    if (j0 != warp) {
#pragma unroll
      for (int u = 0; u < AD_UNR; ++u) {
        const int j = j0 + u * AD_WARPS;
        if (j < pos) vr[u] = load_raw<PER>(vbase + j * ld);
      }
    }
#pragma unroll
    for (int u = 0; u < AD_UNR; ++u) {
      const int j = j0 + u * AD_WARPS;
      if (j >= pos) break;
      const float p = rbf(sc[j] * inv);
      float vf[PER];
      unpack_raw(vr[u], vf);
#pragma unroll
      for (int i = 0; i < PER; ++i) acc[i] = fmaf(p, vf[i], acc[i]);
    }
  }
  if (owner) {
    const float p = rbf(sc[pos] * inv);
    float vf[PER];
    unpack_raw(vnew, vf);
#pragma unroll
    for (int i = 0; i < PER; ++i) acc[i] = fmaf(p, vf[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) pv[warp][lane * PER + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < hd) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < AD_WARPS; ++w) t += pv[w][threadIdx.x];
    o[(static_cast<int64_t>(b) * H + h) * hd + threadIdx.x] = f2b(t);
  }
}

// first maximum of a row (as torch.argmax) by the whole CTA; every thread returns it
__device__ __forceinline__ int block_argmax(const float* __restrict__ row, int V, float* bv, int* bi) {
  float best = -INFINITY;
  int arg = 0x7fffffff;
  auto take = [&](float x, int v) {
    if (x > best || (x == best && v < arg)) {
      best = x;
      arg = v;
    }
  };
  if (V % 4 == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    for (int v4 = threadIdx.x; v4 < V / 4; v4 += blockDim.x) {
      const float4 x = __ldcg(reinterpret_cast<const float4*>(row) + v4);
      take(x.x, 4 * v4), take(x.y, 4 * v4 + 1), take(x.z, 4 * v4 + 2), take(x.w, 4 * v4 + 3);
    }
  } else {
    for (int v = threadIdx.x; v < V; v += blockDim.x) take(__ldcg(row + v), v);
  }
  for (int off = 16; off > 0; off >>= 1) take(__shfl_xor_sync(0xffffffffu, best, off), __shfl_xor_sync(0xffffffffu, arg, off));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();   // bv / bi may still be read from a previous row
  if (lane == 0) {
    bv[warp] = best;
    bi[warp] = arg;
  }
  __syncthreads();
  best = bv[0], arg = bi[0];
  for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) take(bv[w], bi[w]);
  return arg == 0x7fffffff ? 0 : arg;   // a row of NaNs: token 0 rather than an out-of-range id
}

// ids[b] = argmax_v logits[b, v]; optionally also appended to out[b * out_ld + out_col]
__global__ void __launch_bounds__(1024) argmax_rows_kernel(const float* __restrict__ logits, int V, int* __restrict__ ids,
                                                          int* __restrict__ out, int out_ld, int out_col) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  const int arg = block_argmax(logits + static_cast<int64_t>(blockIdx.x) * V, V, bv, bi);
  if (threadIdx.x == 0) {
    ids[blockIdx.x] = arg;
    if (out) out[blockIdx.x * out_ld + out_col] = arg;
  }
}

// The token selection that closes a decode step, one CTA for the B <= 4 rows: id = argmax of the row, appended to the token
// buffer; x[b] = embedding of the id (the next step's input); then the position state advances (dstate non-null).
__global__ void __launch_bounds__(1024) decode_select_kernel(const float* __restrict__ logits, int B, int V, int* __restrict__ ids,
                                                            int* __restrict__ out, int out_ld, int out_col, int* __restrict__ dstate,
                                                            const bf16* __restrict__ table, bf16* __restrict__ x, int d8) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  pdl_trigger();
  pdl_wait();
  if (dstate) out_col = __ldcg(dstate + 1);
  for (int b = 0; b < B; ++b) {
    const int arg = block_argmax(logits + static_cast<int64_t>(b) * V, V, bv, bi);
    if (threadIdx.x == 0) {
      ids[b] = arg;
      out[b * out_ld + out_col] = arg;
    }
    for (int c = threadIdx.x; c < d8; c += blockDim.x)
      reinterpret_cast<uint4*>(x)[static_cast<int64_t>(b) * d8 + c] = __ldg(reinterpret_cast<const uint4*>(table) + static_cast<int64_t>(arg) * d8 + c);
  }
  if (threadIdx.x == 0 && dstate) {
    dstate[0] += 1;   // cache row of the next token
    dstate[1] += 1;   // its column in the token buffer
  }
}

}  // namespace

int embed_rows(const int* ids, const bf16* table, bf16* x, int B, int d, cudaStream_t s) {
  VLA_REQUIRE(d % 8 == 0, "embed_rows: d %% 8 != 0");
  const int n = B * (d / 8);
  embed_rows_kernel<<<ceil_div(n, 256), 256, 0, s>>>(ids, table, x, B, d / 8);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

int attention_decode(bf16* qkv, bf16* o, const float* cos_tab, const float* sin_tab, int B, int L, int pos, int H, int hd, const int* dstate,
                     cudaStream_t s) {
  VLA_REQUIRE(hd == 32 || hd == 64 || hd == 128, "attention_decode: head dim %d not supported (32, 64 or 128)", hd);
  VLA_REQUIRE(dstate || (pos >= 0 && pos < L), "attention_decode: position %d outside the cache (%d rows)", pos, L);
  VLA_REQUIRE(L * sizeof(float) <= 32 * 1024 && static_cast<int64_t>(L) * 3 * H * hd < (1LL << 31), "attention_decode: cache of %d rows too long", L);
  const float scale = 1.f / sqrtf(static_cast<float>(hd));
  const dim3 grid(B * H), block(AD_THREADS);
  const size_t smem = L * sizeof(float);
  if (hd == 128) VLA_CHECK_CUDA(vla_launch(attn_decode_kernel<4>, grid, block, smem, s, qkv, o, cos_tab, sin_tab, L, pos, H, scale, dstate));
  else if (hd == 64) VLA_CHECK_CUDA(vla_launch(attn_decode_kernel<2>, grid, block, smem, s, qkv, o, cos_tab, sin_tab, L, pos, H, scale, dstate));
  else VLA_CHECK_CUDA(vla_launch(attn_decode_kernel<1>, grid, block, smem, s, qkv, o, cos_tab, sin_tab, L, pos, H, scale, dstate));
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

int argmax_rows(const float* logits, int R, int V, int* ids, int* out, int out_ld, int out_col, cudaStream_t s) {
  argmax_rows_kernel<<<R, 1024, 0, s>>>(logits, V, ids, out, out_ld, out_col);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

int decode_select(const float* logits, int B, int V, int* ids, int* out, int out_ld, int out_col, int* dstate, const bf16* table, bf16* x,
                  int d, cudaStream_t s) {
  VLA_REQUIRE(d % 8 == 0 && out && ids, "decode_select: d %% 8 != 0 or null outputs");
  VLA_CHECK_CUDA(vla_launch(decode_select_kernel, dim3(1), dim3(1024), 0, s, logits, B, V, ids, out, out_ld, out_col, dstate, table, x, d / 8));
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Skinny projection of the decode steps: out[m, n] = epilogue(sum_k A[m, k] * W[n, k]) with M <= 4 rows (M = batch of the
// closed-loop evaluation, 1).  Such a product streams the whole weight matrix for a handful of FLOPs per byte: it is bound by
// HBM, and a tensor-core tile loop (one barrier round trip and four tcgen05.mma issues per 16 KB of weights) reaches ~2 TB/s.
// Here a CTA of 8 warps owns groups of 4 weight rows and splits K over its warps: per group a warp streams its K segment of
// the 4 rows with 16-byte loads (two chunks of 8 loads per lane in flight, 8 KB per warp, 128 KB per SM at 2 CTAs/SM),
// multiplies with the activations staged in shared memory, and the 8 partial sums per output meet in shared memory in a
// fixed order.  CTAs are persistent over the groups (grid ~ SMs x resident CTAs, every CTA the same number of groups).
//   prologue (before griddepcontrol.wait): the loads of the warp's first two chunks, the norm weights;
//   activations: plain bf16 rows, or RMSNorm fused -- A = w_norm * bf16(x * rstd) with the row's rstd computed by the CTA;
//   epilogue: fp32 sum -> bf16 (-> + residual -> bf16), or SwiGLU over interleaved [gate 64 | up 64] weight-row groups
//   (out[m, j] = bf16(bf16(silu(g)) * u), the PAIR epilogue of the tcgen05 GEMM), optionally fp32 output (holding bf16 values)
//   and the output row remapped to a cache row (m * out_stride + position).
// ---------------------------------------------------------------------------------------------------------
namespace {
constexpr int GV_WARPS = 8, GV_THREADS = GV_WARPS * 32, GV_UN = 4, GV_KC = 2, GV_NORM_VEC = 4;

struct GemvArgs {
  const bf16* A;
  int64_t lda;
  const bf16* norm_w;   // non-null: A rows are RMS-normalised and scaled by it while staged
  float eps;
  const bf16* W;
  int64_t ldw;
  void* out;
  int64_t ldc;
  const bf16* resid;
  int64_t ldr;
  int N, K;             // weight rows, reduction length
  int out_f32, swiglu, out_stride, out_offset;
  const int* dstate;    // non-null: out_offset = dstate[0]
};

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    f[2 * t] = __uint_as_float(w[t] << 16);
    f[2 * t + 1] = __uint_as_float(w[t] & 0xFFFF0000u);
  }
}

template <int MR>
__global__ void __launch_bounds__(GV_THREADS, 2) gemv_kernel(const GemvArgs a) {
  extern __shared__ __align__(16) unsigned char gv_smem[];
  uint4* sA = reinterpret_cast<uint4*>(gv_smem);   // [MR][K / 8] activations, 8 bf16 per element
  __shared__ float red[2][GV_WARPS][GV_UN * MR];
  __shared__ float nred[32];
  pdl_trigger();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K8 = a.K >> 3;
  const int seg = (K8 + GV_WARPS - 1) / GV_WARPS;
  const int kbeg = warp * seg, kend = min(K8, kbeg + seg);   // this warp's K segment, in 16-byte units
  const int ncols = a.swiglu ? a.N / 2 : a.N;                 // output columns
  const int per_group = a.swiglu ? GV_UN / 2 : GV_UN;         // output columns per group of GV_UN weight rows
  const int ngroups = (ncols + per_group - 1) / per_group;
  const uint4* W4 = reinterpret_cast<const uint4*>(a.W);
  const int64_t ldw8 = a.ldw >> 3;

  // weight rows of group g (-1: past the end).  SwiGLU: output column j pairs rows gate(j) = (j / 64) * 128 + j % 64 and
  // gate(j) + 64; rows 0, 1 of the group are the gates of its two columns, rows 2, 3 their ups.
  auto rows_of = [&](int g, int (&r)[GV_UN]) {
#pragma unroll
    for (int u = 0; u < GV_UN; ++u) {
      if (a.swiglu) {
        const int j = g * 2 + (u & 1);
        r[u] = j < ncols ? (j >> 6) * 128 + (j & 63) + (u >> 1) * 64 : -1;
      } else {
        r[u] = g * GV_UN + u < a.N ? g * GV_UN + u : -1;
      }
    }
  };
  auto load = [&](uint4 (&wv)[GV_UN][GV_KC], const int (&r)[GV_UN], int c0) {
#pragma unroll
    for (int u = 0; u < GV_UN; ++u)
#pragma unroll
      for (int i = 0; i < GV_KC; ++i) {
        const int c = c0 + i * 32 + lane;
        wv[u][i] = (r[u] >= 0 && c < kend) ? __ldcs(W4 + r[u] * ldw8 + c) : make_uint4(0u, 0u, 0u, 0u);   // streamed once: evict first
      }
  };

  // The warp's work is a sequence of chunks (group, K offset): GV_UN rows x GV_KC x 32 lanes x 16 bytes each.  Two chunks are
  // in flight per warp (two register buffers): the loads of chunk i + 2 are issued as soon as chunk i has been consumed,
  // so the stream does not pause while a group is reduced and stored.  Both buffers are filled before the wait.
  struct Chunk {
    int g, c0;
  };
  auto next_chunk = [&](Chunk c) {
    c.c0 += 32 * GV_KC;
    if (c.c0 >= kend) {
      c.g += gridDim.x;
      c.c0 = kbeg;
    }
    return c;
  };
  auto load_chunk = [&](uint4 (&wv)[GV_UN][GV_KC], const Chunk& c) {
    if (c.g >= ngroups) return;
    int r[GV_UN];
    rows_of(c.g, r);
    load(wv, r, c.c0);
  };
  Chunk cur = {static_cast<int>(blockIdx.x), kbeg};
  Chunk nxt = next_chunk(cur);
  uint4 w0[GV_UN][GV_KC], w1[GV_UN][GV_KC];
  load_chunk(w0, cur);
  load_chunk(w1, nxt);
  uint4 nw[GV_NORM_VEC];
  if (a.norm_w) {
#pragma unroll
    for (int it = 0; it < GV_NORM_VEC; ++it) {
      const int c = tid + it * GV_THREADS;
      if (c < K8) nw[it] = __ldg(reinterpret_cast<const uint4*>(a.norm_w) + c);
    }
  }
  pdl_wait();   // everything above reads constants only; the activations (and the position) come from the previous kernels
  const int out_offset = a.dstate ? __ldcg(a.dstate) : a.out_offset;
  const uint4* A4 = reinterpret_cast<const uint4*>(a.A);
  const int64_t lda8 = a.lda >> 3;
  if (!a.norm_w) {
#pragma unroll
    for (int m = 0; m < MR; ++m)
      for (int c = tid; c < K8; c += GV_THREADS) sA[m * K8 + c] = __ldcg(A4 + m * lda8 + c);
  } else {
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      uint4 xv[GV_NORM_VEC];
      float q = 0.f;
#pragma unroll
      for (int it = 0; it < GV_NORM_VEC; ++it) {
        const int c = tid + it * GV_THREADS;
        if (c < K8) {
          xv[it] = __ldcg(A4 + m * lda8 + c);
          float xf[8];
          unpack8(xv[it], xf);
#pragma unroll
          for (int k = 0; k < 8; ++k) q = fmaf(xf[k], xf[k], q);
        }
      }
      const float rstd = rsqrtf(block_sum(q, nred) / a.K + a.eps);
#pragma unroll
      for (int it = 0; it < GV_NORM_VEC; ++it) {
        const int c = tid + it * GV_THREADS;
        if (c < K8) {
          float xf[8], wf[8];
          unpack8(xv[it], xf);
          unpack8(nw[it], wf);
          uint32_t p[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) p[t] = pack_bf16x2(wf[2 * t] * rbf(xf[2 * t] * rstd), wf[2 * t + 1] * rbf(xf[2 * t + 1] * rstd));
          sA[m * K8 + c] = make_uint4(p[0], p[1], p[2], p[3]);   // weight * hidden.to(bf16), as rmsnorm_fwd
        }
      }
    }
  }
  __syncthreads();

  int par = 0;
  float acc[GV_UN][MR];
#pragma unroll
  for (int u = 0; u < GV_UN; ++u)
#pragma unroll
    for (int m = 0; m < MR; ++m) acc[u][m] = 0.f;
  auto consume = [&](const uint4 (&wv)[GV_UN][GV_KC], int c0) {
#pragma unroll
    for (int i = 0; i < GV_KC; ++i) {
      const int c = c0 + i * 32 + lane;
      if (c < kend) {
        float af[MR][8];
#pragma unroll
        for (int m = 0; m < MR; ++m) unpack8(sA[m * K8 + c], af[m]);
#pragma unroll
        for (int u = 0; u < GV_UN; ++u) {
          float wf[8];
          unpack8(wv[u][i], wf);
#pragma unroll
          for (int m = 0; m < MR; ++m)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[u][m] = fmaf(af[m][k], wf[k], acc[u][m]);
        }
      }
    }
  };
  // all chunks of group g consumed: reduce over lanes and warps, epilogue, reset the accumulators
  auto finish_group = [&](int g) {
#pragma unroll
    for (int u = 0; u < GV_UN; ++u)
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        const float v = warp_sum(acc[u][m]);
        if (lane == 0) red[par][warp][u * MR + m] = v;
        acc[u][m] = 0.f;
      }
    __syncthreads();
    if (!a.swiglu) {
      if (tid < GV_UN * MR) {
        const int u = tid / MR, m = tid - u * MR, n = g * GV_UN + u;
        if (n < a.N) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < GV_WARPS; ++w) v += red[par][w][tid];
          v = rbf(v);
          if (a.resid) v = rbf(b2f(__ldcg(a.resid + m * a.ldr + n)) + v);
          const int64_t row = a.out_stride ? static_cast<int64_t>(m) * a.out_stride + out_offset : m;
          if (a.out_f32) static_cast<float*>(a.out)[row * a.ldc + n] = v;
          else static_cast<bf16*>(a.out)[row * a.ldc + n] = f2b(v);
        }
      }
    } else if (tid < 2 * MR) {
      const int jj = tid / MR, m = tid - jj * MR, j = g * 2 + jj;
      if (j < ncols) {
        float gt = 0.f, up = 0.f;
#pragma unroll
        for (int w = 0; w < GV_WARPS; ++w) {
          gt += red[par][w][jj * MR + m];
          up += red[par][w][(2 + jj) * MR + m];
        }
        gt = rbf(gt), up = rbf(up);
        const float v = rbf(gt * sigmoid_fast(gt)) * up;
        static_cast<bf16*>(a.out)[m * a.ldc + j] = f2b(v);
      }
    }
    par ^= 1;   // two reduction buffers: one barrier per group is enough
  };
  while (cur.g < ngroups) {
    consume(w0, cur.c0);
    Chunk after = next_chunk(nxt);
    load_chunk(w0, after);
    if (nxt.g != cur.g) finish_group(cur.g);
    cur = nxt, nxt = after;
    if (cur.g >= ngroups) break;
    consume(w1, cur.c0);
    after = next_chunk(nxt);
    load_chunk(w1, after);
    if (nxt.g != cur.g) finish_group(cur.g);
    cur = nxt, nxt = after;
  }
}
}  // namespace

bool gemv_supported(int M, int K, int64_t lda, int64_t ldw) {
  return M >= 1 && M <= 4 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && static_cast<int64_t>(M) * K * 2 <= 96 * 1024;
}

// out row m -> m * out_stride + out_offset (or the position in dstate) when out_stride > 0 (the decode's cache-row remap), else m
int gemv_bf16(const bf16* A, int64_t lda, const bf16* norm_w, float eps, const bf16* W, int64_t ldw, void* out, int64_t ldc, int M, int N,
              int K, const bf16* resid, int64_t ldr, int out_f32, int swiglu, int out_stride, int out_offset, const int* dstate,
              cudaStream_t s) {
  VLA_REQUIRE(gemv_supported(M, K, lda, ldw), "gemv: needs 1 <= M <= 4, K, lda, ldw multiples of 8 and M * K <= 48K (M=%d K=%d)", M, K);
  VLA_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "gemv: operands must be 16-byte aligned");
  VLA_REQUIRE(!norm_w || (K <= 8 * GV_THREADS * GV_NORM_VEC && (reinterpret_cast<uintptr_t>(norm_w) & 15) == 0), "gemv: fused RMSNorm needs K <= %d",
              8 * GV_THREADS * GV_NORM_VEC);
  VLA_REQUIRE(!swiglu || (N % 128 == 0 && !resid && !out_f32 && !out_stride), "gemv: SwiGLU epilogue needs N %% 128 == 0 and a plain bf16 output");
  static int sms = 0;
  static bool attr_set[5] = {false, false, false, false, false};
  if (!sms) {
    int dev = 0;
    VLA_CHECK_CUDA(cudaGetDevice(&dev));
    VLA_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  GemvArgs a;
  a.A = A, a.lda = lda, a.norm_w = norm_w, a.eps = eps, a.W = W, a.ldw = ldw, a.out = out, a.ldc = ldc, a.resid = resid, a.ldr = ldr;
  a.N = N, a.K = K, a.out_f32 = out_f32, a.swiglu = swiglu, a.out_stride = out_stride, a.out_offset = out_offset, a.dstate = dstate;
  const int ngroups = swiglu ? ceil_div(N / 2, GV_UN / 2) : ceil_div(N, GV_UN);
  const int resident = 2;
  // every CTA takes the same number of groups (+-1): with ceil(ngroups / CTAs) rounds the grid is ngroups / rounds, not the
  // full SMs x resident -- CTAs progress in lock step on an HBM-bound stream, so a last partial round would run at the
  // bandwidth the few remaining CTAs can pull
  const int rounds = ceil_div(ngroups, sms * resident);
  const dim3 grid(ceil_div(ngroups, rounds)), block(GV_THREADS);
  const size_t smem = static_cast<size_t>(M) * K * sizeof(bf16);
  void (*kern)(const GemvArgs) = M == 1 ? gemv_kernel<1> : M == 2 ? gemv_kernel<2> : M == 3 ? gemv_kernel<3> : gemv_kernel<4>;
  if (!attr_set[M]) {
    VLA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set[M] = true;
  }
  VLA_CHECK_CUDA(vla_launch(kern, grid, block, smem, s, a));
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
