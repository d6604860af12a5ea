// Fused multi-head attention forward + backward (flash-style, no materialised score matrix or mask).
//
// Replaces F.scaled_dot_product_attention in timm Attention (DINOv2: 16 heads x 64, N=261; SigLIP: 16 heads x 72,
// N=256; non-causal) and HF LlamaSdpaAttention (32 heads x 128, L~289, causal AND key-padding: the reference builds
// an additive 4-D mask, which pushes SDPA onto its mem-efficient/math backend) and their autograd backward.
// Masking is done in-kernel from (causal, kv_len[b]).
//
// Round-1 implementation: legacy warp-level tensor-core path (ldmatrix + mma.sync m16n8k16, fp32 softmax
// statistics in registers).  Attention is ~1.3 % of the step's FLOPs; the tcgen05 GEMM carries the
// rest.  Head dim is padded to a multiple of 16 in shared memory (72 -> 80, zero filled).
//
// Layout: qkv [B*N, 3*H*hd] = [q | k | v] with heads contiguous inside each third (the natural output of the fused
// qkv GEMM); o, dout [B*N, H*hd]; lse, delta [B, H, N] fp32.
#include <math.h>
#include <stdlib.h>

#include "kernels.h"

namespace {

constexpr int BM = 64;   // rows of the "outer" tile owned by a CTA (4 warps x 16 rows)
constexpr int BN = 64;   // rows of the streamed tile
constexpr int ATT_THREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;

// ex2.approx: one MUFU op (inputs are <= 0 here: exponent of a softmax term; -inf -> 0)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- shared-memory tile helpers ------------------------------------------------------------------------
// Copy rows [r0, r0+64) x cols [0, hd) of a head slice (global row stride ldg) into smem [64][LD], zero filling
// rows >= nrows and the hd..HDP padding.
template <int HDP>
__device__ __forceinline__ void load_tile(bf16* __restrict__ s, const bf16* __restrict__ g, int64_t ldg, int r0, int nrows,
                                          int hd) {
  constexpr int LD = HDP + 8;
  constexpr int CH = HDP / 8;
  for (int idx = threadIdx.x; idx < 64 * CH; idx += ATT_THREADS) {
    const int r = idx / CH, c = idx % CH;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r0 + r < nrows && c * 8 < hd) v = *reinterpret_cast<const uint4*>(g + static_cast<int64_t>(r0 + r) * ldg + c * 8);
    *reinterpret_cast<uint4*>(s + r * LD + c * 8) = v;
  }
}

// Same tile copy with cp.async (16-byte, L2-only): issued for tile j+1 while tile j is being multiplied.
// Thread -> (row, 16-byte chunk) mapping is fixed; passes step the row by a constant, so the loop is adds only.
template <int HDP>
__device__ __forceinline__ void load_tile_async(bf16* __restrict__ s, const bf16* __restrict__ g, int64_t ldg, int r0,
                                                int nrows, int hd) {
  constexpr int LD = HDP + 8;
  constexpr int CH = HDP / 8;
  if constexpr (ATT_THREADS % CH == 0) {
    constexpr int ROWS_PER_PASS = ATT_THREADS / CH;
    const int c = threadIdx.x % CH, r = threadIdx.x / CH;
    const bool col_ok = c * 8 < hd;
    const bf16* src = g + static_cast<int64_t>(r0 + r) * ldg + c * 8;
    uint32_t dst = smem_u32(s + r * LD + c * 8);
    int row = r0 + r;
#pragma unroll
    for (int p = 0; p < 64 / ROWS_PER_PASS; ++p) {
      const bool ok = col_ok && row < nrows;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ok ? src : g), "r"(ok ? 16 : 0) : "memory");
      src += ROWS_PER_PASS * ldg;
      dst += ROWS_PER_PASS * LD * 2;
      row += ROWS_PER_PASS;
    }
  } else {
    for (int idx = threadIdx.x; idx < 64 * CH; idx += ATT_THREADS) {
      const int r = idx / CH, c = idx % CH;
      const bool ok = (r0 + r < nrows) && (c * 8 < hd);
      const bf16* src = ok ? g + static_cast<int64_t>(r0 + r) * ldg + c * 8 : g;
      const uint32_t dst = smem_u32(s + r * LD + c * 8);
      const int bytes = ok ? 16 : 0;   // src-size 0 -> the 16 destination bytes are zero filled
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// acc[NT][4] (16 x NT*8, fp32) += A(16 x K16*16) * B.
//   A: registers a[K16][4] (mma A-fragment order) when A_REG, else smem rows a_row0.. (row stride lda), cols k.
//   B_TRANS == false: Bs is [n][k] (contraction along smem columns), n range starts at b_row0.
//   B_TRANS == true : Bs is [k][n] (contraction along smem rows),    k range starts at b_row0.
// The lane-dependent part of every ldmatrix address is computed once; the unrolled loops only add constants.
template <int NT, int K16, bool B_TRANS, bool A_REG>
__device__ __forceinline__ void warp_gemm(float (&acc)[NT][4], const uint32_t (*a_reg)[4], const bf16* __restrict__ As,
                                          int lda, int a_row0, const bf16* __restrict__ Bs, int ldb, int b_row0) {
  const int lane = threadIdx.x & 31;
  const int mi = lane >> 3, r = lane & 7;
  uint32_t a_base = 0;
  if (!A_REG) a_base = smem_u32(As) + static_cast<uint32_t>(((a_row0 + (mi & 1) * 8 + r) * lda + (mi >> 1) * 8) * 2);
  const uint32_t b_base = smem_u32(Bs) + static_cast<uint32_t>(
      (B_TRANS ? ((b_row0 + (mi & 1) * 8 + r) * ldb + (mi >> 1) * 8) : ((b_row0 + (mi >> 1) * 8 + r) * ldb + (mi & 1) * 8)) * 2);
#pragma unroll
  for (int ks = 0; ks < K16; ++ks) {
    uint32_t a[4];
    if (A_REG) {
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = a_reg[ks][q];
    } else {
      ldmatrix_x4(a, a_base + ks * 32);
    }
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      if (!B_TRANS)
        ldmatrix_x4(b, b_base + static_cast<uint32_t>((np * 16 * ldb + ks * 16) * 2));
      else
        ldmatrix_x4_trans(b, b_base + static_cast<uint32_t>((ks * 16 * ldb + np * 16) * 2));
      mma_bf16_16816(acc[np * 2], a, b[0], b[1]);
      mma_bf16_16816(acc[np * 2 + 1], a, b[2], b[3]);
    }
  }
}

// C-fragment (16 x 64 fp32, 8 n-tiles) -> A-fragments (4 k-steps) in bf16
__device__ __forceinline__ void acc_to_afrag(const float (&c)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = pack_bf16x2(c[2 * ks][0], c[2 * ks][1]);
    a[ks][1] = pack_bf16x2(c[2 * ks][2], c[2 * ks][3]);
    a[ks][2] = pack_bf16x2(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    a[ks][3] = pack_bf16x2(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}

template <int HDP>
struct SmemLayout {
  static constexpr int LD = HDP + 8;
  static constexpr int TILE = 64 * LD;   // elements
};

// =========================================================================================================
// forward
// =========================================================================================================
template <int HDP>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o,
                                                               float* __restrict__ lse, const int* __restrict__ kv_len, int N,
                                                               int H, int hd, int causal, float scale) {
  using L = SmemLayout<HDP>;
  extern __shared__ __align__(16) uint8_t smem_att[];
  pdl_trigger();
  pdl_wait();
  bf16* sQ = reinterpret_cast<bf16*>(smem_att);
  bf16* sKb = sQ + L::TILE;            // K stage 0, K stage 1
  bf16* sVb = sKb + 2 * L::TILE;       // V stage 0, V stage 1

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int D = H * hd;
  const int64_t ld = 3 * static_cast<int64_t>(D);
  const bf16* base = qkv + static_cast<int64_t>(b) * N * ld + h * hd;
  const int klen = kv_len ? min(kv_len[b], N) : N;
  const float sl2 = scale * LOG2E;
  const int kv_end = causal ? min(klen, q0 + BM) : klen;

  load_tile_async<HDP>(sQ, base, ld, q0, N, hd);
  load_tile_async<HDP>(sKb, base + D, ld, 0, N, hd);
  load_tile_async<HDP>(sVb, base + 2 * D, ld, 0, N, hd);
  cp_async_commit();

  float acc_o[HDP / 8][4];
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc_o[i][k] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int j0 = 0, it = 0; j0 < kv_end; j0 += BN, ++it) {
    const bf16* sK = sKb + (it & 1) * L::TILE;
    const bf16* sV = sVb + (it & 1) * L::TILE;
    if (j0 + BN < kv_end) {   // prefetch the next K/V tile into the other stage (freed by the barrier ending iteration it-1)
      load_tile_async<HDP>(sKb + ((it + 1) & 1) * L::TILE, base + D, ld, j0 + BN, N, hd);
      load_tile_async<HDP>(sVb + ((it + 1) & 1) * L::TILE, base + 2 * D, ld, j0 + BN, N, hd);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) s[i][k] = 0.f;
    warp_gemm<8, HDP / 16, false, false>(s, nullptr, sQ, L::LD, warp * 16, sK, L::LD, 0);

    // mask + online softmax (rows g and g+8 of this warp's 16).  Tiles that lie entirely below the causal diagonal
    // and inside the key length need no masking (warp-uniform test): most of the work takes the short path.
    const int row0 = q0 + warp * 16 + g;
    const bool need_mask = (j0 + BN > klen) || (causal && j0 + BN - 1 > q0 + warp * 16);
    if (need_mask) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int col = j0 + nt * 8 + 2 * t + (k & 1);
          const int row = row0 + (k >> 1) * 8;
          const bool ok = (col < klen) && (!causal || col <= row);
          if (!ok) s[nt][k] = -INFINITY;
        }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float alpha[2], mref[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mnew = fmaxf(m_run[r], mx[r] * sl2);       // running max in the log2 domain (sl2 > 0)
      mref[r] = (mnew == -INFINITY) ? 0.f : mnew;
      alpha[r] = fast_exp2(m_run[r] - mref[r]);              // m_run = -inf -> 0
      m_run[r] = mnew;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float p = fast_exp2(fmaf(s[nt][k], sl2, -mref[k >> 1]));   // masked: -inf -> 0
        s[nt][k] = p;
        rs[k >> 1] += p;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * alpha[r] + rs[r];
    }
#pragma unroll
    for (int i = 0; i < HDP / 8; ++i) {
      acc_o[i][0] *= alpha[0];
      acc_o[i][1] *= alpha[0];
      acc_o[i][2] *= alpha[1];
      acc_o[i][3] *= alpha[1];
    }
    uint32_t pa[4][4];
    acc_to_afrag(s, pa);
    warp_gemm<HDP / 8, 4, true, true>(acc_o, pa, nullptr, 0, 0, sV, L::LD, 0);
    __syncthreads();   // every warp is done with this stage before the next prefetch overwrites it
  }

  // epilogue: O / l, LSE (natural log)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    if (row >= N) continue;
    const float inv = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
    bf16* orow = o + (static_cast<int64_t>(b) * N + row) * D + h * hd;
#pragma unroll
    for (int i = 0; i < HDP / 8; ++i) {
      const int col = i * 8 + 2 * t;
      if (col < hd)
        *reinterpret_cast<uint32_t*>(orow + col) = pack_bf16x2(acc_o[i][r * 2] * inv, acc_o[i][r * 2 + 1] * inv);
    }
    if (t == 0)
      lse[(static_cast<int64_t>(b) * H + h) * N + row] =
          (l_run[r] > 0.f) ? (m_run[r] + log2f(l_run[r])) / LOG2E : -INFINITY;
  }
}

// =========================================================================================================
// backward
// =========================================================================================================
// Writes one thread's share of a gradient row (columns i*8 + 2t, +1 of every 8-column tile) as bf16(acc * scale).
// With rope tables (head dim 128) the rotary embedding's backward is applied on the way out: the gradient wrt the
// pre-RoPE q / k is the inverse rotation of the gradient wrt the rotated ones (pairs (c, c + 64) live in tiles i, i + 8).
template <int HDP>
__device__ __forceinline__ void store_grad_row(bf16* __restrict__ drow, const float (&acc)[HDP / 8][4], int r, int t, int hd,
                                               float scale, const float* __restrict__ rope_cos,
                                               const float* __restrict__ rope_sin, int pos) {
  if constexpr (HDP == 128) {
    if (rope_cos != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int col = i * 8 + 2 * t;
        const float2 c = *reinterpret_cast<const float2*>(rope_cos + pos * 64 + col);
        const float2 sn = *reinterpret_cast<const float2*>(rope_sin + pos * 64 + col);
        const float l0 = rbf(acc[i][r * 2] * scale), l1 = rbf(acc[i][r * 2 + 1] * scale);
        const float h0 = rbf(acc[i + 8][r * 2] * scale), h1 = rbf(acc[i + 8][r * 2 + 1] * scale);
        *reinterpret_cast<uint32_t*>(drow + col) = pack_bf16x2(rbf(l0 * c.x) + rbf(h0 * sn.x), rbf(l1 * c.y) + rbf(h1 * sn.y));
        *reinterpret_cast<uint32_t*>(drow + 64 + col) = pack_bf16x2(rbf(h0 * c.x) + rbf(-l0 * sn.x), rbf(h1 * c.y) + rbf(-l1 * sn.y));
      }
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i) {
    const int col = i * 8 + 2 * t;
    if (col < hd) *reinterpret_cast<uint32_t*>(drow + col) = pack_bf16x2(acc[i][r * 2] * scale, acc[i][r * 2 + 1] * scale);
  }
}

// dQ: CTA owns 64 query rows, streams K/V tiles.
template <int HDP>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                                  const bf16* __restrict__ dout, const float* __restrict__ lse,
                                                                  float* __restrict__ delta, bf16* __restrict__ dqkv,
                                                                  const int* __restrict__ kv_len, int N, int H, int hd, int causal,
                                                                  float scale, const float* __restrict__ rope_cos,
                                                                  const float* __restrict__ rope_sin, int rope_L) {
  using L = SmemLayout<HDP>;
  extern __shared__ __align__(16) uint8_t smem_att[];
  pdl_trigger();
  pdl_wait();
  bf16* sQ = reinterpret_cast<bf16*>(smem_att);
  bf16* sdO = sQ + L::TILE;
  bf16* sKb = sdO + L::TILE;           // K stage 0, K stage 1
  bf16* sVb = sKb + 2 * L::TILE;       // V stage 0, V stage 1

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int D = H * hd;
  const int64_t ld = 3 * static_cast<int64_t>(D);
  const bf16* base = qkv + static_cast<int64_t>(b) * N * ld + h * hd;
  const int klen = kv_len ? min(kv_len[b], N) : N;
  const float sl2 = scale * LOG2E;
  const int kv_end = causal ? min(klen, q0 + BM) : klen;

  load_tile_async<HDP>(sQ, base, ld, q0, N, hd);
  load_tile_async<HDP>(sdO, dout + static_cast<int64_t>(b) * N * D + h * hd, D, q0, N, hd);
  load_tile_async<HDP>(sKb, base + D, ld, 0, N, hd);
  load_tile_async<HDP>(sVb, base + 2 * D, ld, 0, N, hd);
  cp_async_commit();

  float lse2[2], dl[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    const bool ok = row < N;
    const float l = ok ? lse[(static_cast<int64_t>(b) * H + h) * N + row] : INFINITY;
    lse2[r] = (l == -INFINITY) ? INFINITY : l * LOG2E;   // fully masked / padded rows -> p = 0
  }
  // delta = rowsum(dO * O) for this warp's 16 rows (was a separate kernel).  All 16 rows' loads are issued before any
  // reduction (independent, fully unrolled) so that their latency overlaps; written out for the dK/dV kernel.
  {
    const bf16* obase = o + static_cast<int64_t>(b) * N * D + h * hd;
    const bf16* dbase = dout + static_cast<int64_t>(b) * N * D + h * hd;
    float part[16];
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const int row = q0 + warp * 16 + rr;
      float sum = 0.f;
      if (row < N) {
#pragma unroll
        for (int c0 = 0; c0 < HDP; c0 += 64) {
          const int c = c0 + lane * 2;
          if (c < hd) {
            const float2 a = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(obase + static_cast<int64_t>(row) * D + c)));
            const float2 d = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(dbase + static_cast<int64_t>(row) * D + c)));
            sum += a.x * d.x + a.y * d.y;
          }
        }
      }
      part[rr] = sum;
    }
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const float sum = warp_sum(part[rr]);
      if (rr == g) dl[0] = sum;
      if (rr == g + 8) dl[1] = sum;
      const int row = q0 + warp * 16 + rr;
      if (lane == rr && row < N) delta[(static_cast<int64_t>(b) * H + h) * N + row] = sum;
    }
  }

  float acc_dq[HDP / 8][4];
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc_dq[i][k] = 0.f;

  for (int j0 = 0, it = 0; j0 < kv_end; j0 += BN, ++it) {
    const bf16* sK = sKb + (it & 1) * L::TILE;
    const bf16* sV = sVb + (it & 1) * L::TILE;
    if (j0 + BN < kv_end) {
      load_tile_async<HDP>(sKb + ((it + 1) & 1) * L::TILE, base + D, ld, j0 + BN, N, hd);
      load_tile_async<HDP>(sVb + ((it + 1) & 1) * L::TILE, base + 2 * D, ld, j0 + BN, N, hd);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) s[i][k] = dp[i][k] = 0.f;
    warp_gemm<8, HDP / 16, false, false>(s, nullptr, sQ, L::LD, warp * 16, sK, L::LD, 0);
    warp_gemm<8, HDP / 16, false, false>(dp, nullptr, sdO, L::LD, warp * 16, sV, L::LD, 0);
    const bool need_mask = (j0 + BN > klen) || (causal && j0 + BN - 1 > q0 + warp * 16);
    if (need_mask) {
      const int row0 = q0 + warp * 16 + g;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int col = j0 + nt * 8 + 2 * t + (k & 1);
          const int row = row0 + (k >> 1) * 8;
          const bool ok = (col < klen) && (!causal || col <= row);
          if (!ok) s[nt][k] = -INFINITY;
        }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float p = fast_exp2(fmaf(s[nt][k], sl2, -lse2[k >> 1]));   // masked / padded rows: -inf -> 0
        s[nt][k] = p * (dp[nt][k] - dl[k >> 1]);   // dS
      }
    uint32_t dsa[4][4];
    acc_to_afrag(s, dsa);
    warp_gemm<HDP / 8, 4, true, true>(acc_dq, dsa, nullptr, 0, 0, sK, L::LD, 0);
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    if (row >= N) continue;
    bf16* drow = dqkv + (static_cast<int64_t>(b) * N + row) * ld + h * hd;
    store_grad_row<HDP>(drow, acc_dq, r, t, hd, scale, rope_cos, rope_sin, row % (rope_L > 0 ? rope_L : 1));
  }
}

// dK, dV: CTA owns 64 key/value rows, streams Q/dO tiles; works on the transposed score tile S^T = K Q^T.
template <int HDP>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                   const float* __restrict__ lse, const float* __restrict__ delta,
                                                                   bf16* __restrict__ dqkv, const int* __restrict__ kv_len, int N,
                                                                   int H, int hd, int causal, float scale,
                                                                   const float* __restrict__ rope_cos,
                                                                   const float* __restrict__ rope_sin, int rope_L) {
  using L = SmemLayout<HDP>;
  extern __shared__ __align__(16) uint8_t smem_att[];
  pdl_trigger();
  pdl_wait();
  bf16* sK = reinterpret_cast<bf16*>(smem_att);
  bf16* sV = sK + L::TILE;
  bf16* sQb = sV + L::TILE;            // Q stage 0, Q stage 1
  bf16* sdOb = sQb + 2 * L::TILE;      // dO stage 0, dO stage 1
  float* sLseB = reinterpret_cast<float*>(sdOb + 2 * L::TILE);   // [2][64]
  float* sDeltaB = sLseB + 128;                                   // [2][64]

  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * BM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int D = H * hd;
  const int64_t ld = 3 * static_cast<int64_t>(D);
  const bf16* base = qkv + static_cast<int64_t>(b) * N * ld + h * hd;
  const bf16* dbase = dout + static_cast<int64_t>(b) * N * D + h * hd;
  const int klen = kv_len ? min(kv_len[b], N) : N;
  const float sl2 = scale * LOG2E;
  const bool tile_live = j0 < klen;   // a fully masked key tile gets zero gradients
  const int q_begin = causal ? (j0 / BN) * BN : 0;   // queries before this key tile never see it

  auto stage_stats = [&](int stage, int q0) {   // lse / delta of the 64 queries of a stage (plain loads, tiny)
    if (threadIdx.x < 64) {
      const int row = q0 + threadIdx.x;
      const float l = row < N ? lse[(static_cast<int64_t>(b) * H + h) * N + row] : INFINITY;
      sLseB[stage * 64 + threadIdx.x] = (l == -INFINITY) ? INFINITY : l * LOG2E;
      sDeltaB[stage * 64 + threadIdx.x] = row < N ? delta[(static_cast<int64_t>(b) * H + h) * N + row] : 0.f;
    }
  };

  load_tile_async<HDP>(sK, base + D, ld, j0, N, hd);
  load_tile_async<HDP>(sV, base + 2 * D, ld, j0, N, hd);
  if (tile_live && q_begin < N) {
    load_tile_async<HDP>(sQb, base, ld, q_begin, N, hd);
    load_tile_async<HDP>(sdOb, dbase, D, q_begin, N, hd);
    stage_stats(0, q_begin);
  }
  cp_async_commit();

  float acc_dk[HDP / 8][4], acc_dv[HDP / 8][4];
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc_dk[i][k] = acc_dv[i][k] = 0.f;

  for (int q0 = q_begin, it = 0; tile_live && q0 < N; q0 += BN, ++it) {
    const bf16* sQ = sQb + (it & 1) * L::TILE;
    const bf16* sdO = sdOb + (it & 1) * L::TILE;
    const float* sLse = sLseB + (it & 1) * 64;
    const float* sDelta = sDeltaB + (it & 1) * 64;
    if (q0 + BN < N) {
      load_tile_async<HDP>(sQb + ((it + 1) & 1) * L::TILE, base, ld, q0 + BN, N, hd);
      load_tile_async<HDP>(sdOb + ((it + 1) & 1) * L::TILE, dbase, D, q0 + BN, N, hd);
      stage_stats((it + 1) & 1, q0 + BN);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    // S^T (16 kv x 64 q) and dP^T = V dO^T
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) st[i][k] = dpt[i][k] = 0.f;
    warp_gemm<8, HDP / 16, false, false>(st, nullptr, sK, L::LD, warp * 16, sQ, L::LD, 0);
    warp_gemm<8, HDP / 16, false, false>(dpt, nullptr, sV, L::LD, warp * 16, sdO, L::LD, 0);
    // columns of S^T are queries, rows are keys.  Padded queries carry lse = +inf (p = 0) already; masking is only
    // needed on the causal diagonal and for key rows beyond the key length.
    const bool need_mask = (j0 + warp * 16 + 16 > klen) || (causal && j0 + warp * 16 + 15 > q0);
    if (need_mask) {
      const int krow0 = j0 + warp * 16 + g;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int qrow = q0 + nt * 8 + 2 * t + (k & 1);
          const int kcol = krow0 + (k >> 1) * 8;
          const bool ok = (kcol < klen) && (!causal || kcol <= qrow);
          if (!ok) st[nt][k] = -INFINITY;
        }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 l2 = *reinterpret_cast<const float2*>(sLse + nt * 8 + 2 * t);
      const float2 d2 = *reinterpret_cast<const float2*>(sDelta + nt * 8 + 2 * t);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float p = fast_exp2(fmaf(st[nt][k], sl2, -((k & 1) ? l2.y : l2.x)));
        st[nt][k] = p;                                                  // P^T
        dpt[nt][k] = p * (dpt[nt][k] - ((k & 1) ? d2.y : d2.x));        // dS^T
      }
    }
    uint32_t pa[4][4], dsa[4][4];
    acc_to_afrag(st, pa);
    acc_to_afrag(dpt, dsa);
    warp_gemm<HDP / 8, 4, true, true>(acc_dv, pa, nullptr, 0, 0, sdO, L::LD, 0);   // dV += P^T dO
    warp_gemm<HDP / 8, 4, true, true>(acc_dk, dsa, nullptr, 0, 0, sQ, L::LD, 0);   // dK += dS^T Q
    __syncthreads();
  }
  cp_async_wait<0>();

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = j0 + warp * 16 + g + r * 8;
    if (row >= N) continue;
    bf16* drow = dqkv + (static_cast<int64_t>(b) * N + row) * ld + h * hd;
    store_grad_row<HDP>(drow + D, acc_dk, r, t, hd, scale, rope_cos, rope_sin, row % (rope_L > 0 ? rope_L : 1));
    store_grad_row<HDP>(drow + 2 * D, acc_dv, r, t, hd, 1.f, nullptr, nullptr, 0);
  }
}

template <int HDP>
int launch_fwd(const bf16* qkv, bf16* o, float* lse, const int* kv_len, int B, int N, int H, int hd, int causal,
               cudaStream_t s) {
  constexpr int smem = 5 * SmemLayout<HDP>::TILE * 2;   // Q + 2 stages of (K, V)
  static bool configured = false;
  if (!configured) {
    VLA_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(ceil_div(N, BM), H, B);
  VLA_CHECK_CUDA(vla_launch(attn_fwd_kernel<HDP>, grid, dim3(ATT_THREADS), smem, s, qkv, o, lse, kv_len, N, H, hd, causal,
                            1.f / sqrtf(static_cast<float>(hd))));
  ++g_vla_launch_count;
  return 0;
}

template <int HDP>
int launch_bwd(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
               const int* kv_len, int B, int N, int H, int hd, int causal, const float* rope_cos, const float* rope_sin,
               int rope_L, cudaStream_t s) {
  constexpr int smem_dq = 6 * SmemLayout<HDP>::TILE * 2;    // Q, dO + 2 stages of (K, V)
  constexpr int smem_dkv = 6 * SmemLayout<HDP>::TILE * 2 + 4 * 64 * 4;   // K, V + 2 stages of (Q, dO, lse, delta)
  static bool configured = false;
  if (!configured) {
    VLA_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq));
    VLA_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv));
    configured = true;
  }
  const float scale = 1.f / sqrtf(static_cast<float>(hd));
  dim3 grid(ceil_div(N, BM), H, B);
  VLA_CHECK_CUDA(vla_launch(attn_bwd_dq_kernel<HDP>, grid, dim3(ATT_THREADS), smem_dq, s, qkv, o, dout, lse, delta, dqkv, kv_len, N, H,
                            hd, causal, scale, rope_cos, rope_sin, rope_L));
  VLA_CHECK_CUDA(vla_launch(attn_bwd_dkv_kernel<HDP>, grid, dim3(ATT_THREADS), smem_dkv, s, qkv, dout, lse,
                            static_cast<const float*>(delta), dqkv, kv_len, N, H, hd, causal, scale, rope_cos, rope_sin, rope_L));
  g_vla_launch_count += 2;
  return 0;
}

}  // namespace

// bit 0: tcgen05 forward (attention_fwd_tc3 for N <= 320, else the streaming attention_fwd_tc2); bit 1: tcgen05 backward.
// Default 3; VLA_ATTN_IMPL overrides it at first use.
int g_attn_impl = -1;
int attention_impl() {
  if (g_attn_impl < 0) {
    const char* e = getenv("VLA_ATTN_IMPL");
    g_attn_impl = e ? (atoi(e) & 3) : 3;
  }
  return g_attn_impl;
}

static int attention_fwd_once(const bf16* qkv, bf16* o, float* lse, const int* kv_len, int B, int N, int H, int hd, int causal,
                              cudaStream_t s);
int attention_fwd(const bf16* qkv, bf16* o, float* lse, const int* kv_len, int B, int N, int H, int hd, int causal,
                  cudaStream_t s) {
  if (vla_doubled("attn_fwd"))
    if (int rc = attention_fwd_once(qkv, o, lse, kv_len, B, N, H, hd, causal, s)) return rc;
  return attention_fwd_once(qkv, o, lse, kv_len, B, N, H, hd, causal, s);
}
static int attention_fwd_once(const bf16* qkv, bf16* o, float* lse, const int* kv_len, int B, int N, int H, int hd, int causal,
                              cudaStream_t s) {
  VLA_REQUIRE(hd % 8 == 0 && hd <= 128, "attention: unsupported head dim %d", hd);
  if (vla_ablated("attn_fwd")) return 0;
  if ((attention_impl() & 1) && attention_fwd_tc3_supported(N, hd)) return attention_fwd_tc3(qkv, o, lse, kv_len, B, N, H, hd, causal, s);
  if ((attention_impl() & 1) && attention_fwd_tc2_supported(N, hd)) return attention_fwd_tc2(qkv, o, lse, kv_len, B, N, H, hd, causal, s);
  if (hd <= 64) return launch_fwd<64>(qkv, o, lse, kv_len, B, N, H, hd, causal, s);
  if (hd <= 80) return launch_fwd<80>(qkv, o, lse, kv_len, B, N, H, hd, causal, s);
  return launch_fwd<128>(qkv, o, lse, kv_len, B, N, H, hd, causal, s);
}

int attention_tile_shift(int N, int causal) {
  const char* e = getenv("VLA_ATTN_SHIFT");   // read per call: A/B switch, tests
  if (!causal || (e != nullptr && atoi(e) == 0)) return 0;
  const int ntiles = (N + 127) / 128;
  return ((ntiles * 128 - N) / 64) * 64;
}

// true when attention_bwd() can be called with o == NULL and `delta` already holding rowsum(dO * O) per (b, h, n)
bool attention_bwd_takes_delta(int N, int hd) { return (attention_impl() & 2) && attention_bwd_tc_supported(N, hd); }

static int attention_bwd_once(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                              const int* kv_len, int B, int N, int H, int hd, int causal, const float* rope_cos, const float* rope_sin,
                              int rope_L, cudaStream_t s);
int attention_bwd(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                  const int* kv_len, int B, int N, int H, int hd, int causal, const float* rope_cos, const float* rope_sin,
                  int rope_L, cudaStream_t s) {
  if (vla_doubled("attn_bwd"))
    if (int rc = attention_bwd_once(qkv, o, dout, lse, delta, dqkv, kv_len, B, N, H, hd, causal, rope_cos, rope_sin, rope_L, s)) return rc;
  return attention_bwd_once(qkv, o, dout, lse, delta, dqkv, kv_len, B, N, H, hd, causal, rope_cos, rope_sin, rope_L, s);
}
static int attention_bwd_once(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                              const int* kv_len, int B, int N, int H, int hd, int causal, const float* rope_cos, const float* rope_sin,
                              int rope_L, cudaStream_t s) {
  VLA_REQUIRE(hd % 8 == 0 && hd <= 128, "attention: unsupported head dim %d", hd);
  if (vla_ablated("attn_bwd")) return 0;
  VLA_REQUIRE(rope_cos == nullptr || (hd == 128 && rope_sin != nullptr && rope_L > 0),
              "attention_bwd: the fused RoPE backward needs head dim 128 and both tables");
  VLA_REQUIRE(o != nullptr || attention_bwd_takes_delta(N, hd), "attention_bwd: a precomputed delta (o == NULL) needs the tcgen05 backward");
  if ((attention_impl() & 2) && attention_bwd_tc_supported(N, hd))
    return attention_bwd_tc(qkv, o, dout, lse, delta, dqkv, kv_len, B, N, H, hd, causal, rope_cos, rope_sin, rope_L, s);
  if (hd <= 64) return launch_bwd<64>(qkv, o, dout, lse, delta, dqkv, kv_len, B, N, H, hd, causal, nullptr, nullptr, 0, s);
  if (hd <= 80) return launch_bwd<80>(qkv, o, dout, lse, delta, dqkv, kv_len, B, N, H, hd, causal, nullptr, nullptr, 0, s);
  return launch_bwd<128>(qkv, o, dout, lse, delta, dqkv, kv_len, B, N, H, hd, causal, rope_cos, rope_sin, rope_L, s);
}
