// tcgen05 attention forward for sequences that fit one pass (N <= 320 keys): S = Q K^T and O = P V on the 5th-gen
// tensor cores with TMEM accumulators, TMA-staged operands, exact (non-online) softmax with one thread per query row.
//
// One CTA per (128-query tile, head, batch element), 160 threads:
//   warp 4 (one thread) : TMA loads of the Q tile and of the K / V rows this tile can see (3-D tensor map over
//                         [B][N][3*H*hd], 64x64 boxes, 128B swizzle, rows >= N zero filled), then the tcgen05.mma issue:
//                         S[128 x nkv] = Q . K^T  (both operands K-major), later O[128 x hd] = P . V with P (bf16, written
//                         by the softmax warps in the canonical K-major 128B-swizzled layout) and V as an MN-major B operand
//                         (V is [kv, hd] with hd contiguous: no transpose needed)
//   warps 0-3           : thread r owns query row r: tcgen05.ld its S row from TMEM, mask (causal, key length), max, exp2,
//                         sum, write P; after the second MMA read O from TMEM, scale by 1/l, store bf16 and the LSE.
// Replaces F.scaled_dot_product_attention forward for head dims 64 (DINOv2) and 128 (Llama); other head dims use the
// legacy kernels of attention.cu.
#include <math.h>

#include "kernels.h"
#include "tma_desc.h"

namespace {

constexpr int TC_THREADS = 288;      // warps 0-7: softmax / epilogue, warp 8: TMA + MMA issue
constexpr int MAX_KV = 320;          // keys per (batch, head) supported in one pass
constexpr int PANEL_ROWS_Q = 128;
constexpr float LOG2E_F = 1.4426950408889634f;

// shared-memory operand descriptor, 128B swizzle, version 1 (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// bf16 x bf16 -> fp32, M x N tile, A K-major, B K-major (b_mn = 0) or MN-major (b_mn = 1)
__device__ __forceinline__ uint32_t idesc_bf16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

template <int HD>
struct FwdSmem {
  static constexpr int PANELS = HD / 64;                       // 64-column (128-byte) panels of a head
  static constexpr int Q_BYTES = PANELS * PANEL_ROWS_Q * 128;
  static constexpr int KV_PANEL_BYTES = MAX_KV * 128;          // one panel of K or V: MAX_KV rows x 128 B
  static constexpr int KV_BYTES = PANELS * KV_PANEL_BYTES;
  static constexpr int P_PANEL_BYTES = PANEL_ROWS_Q * 128;     // one 64-key panel of P: 128 rows x 128 B
  static constexpr int P_BYTES = (MAX_KV / 64) * P_PANEL_BYTES;
  static constexpr bool P_ALIASES_K = (KV_BYTES >= P_BYTES);   // hd = 128: P reuses K's buffer once S is done
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_BYTES;
  static constexpr int OFF_P = P_ALIASES_K ? OFF_K : OFF_V + KV_BYTES;
  static constexpr int OFF_BAR = (P_ALIASES_K ? OFF_V + KV_BYTES : OFF_P + P_BYTES);
  static constexpr int TOTAL = OFF_BAR + 64 + 1024 /* s_red */ + 1024 /* alignment slack */;
};

// named barrier among the 256 softmax threads only (the control warp never joins it)
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD>
__global__ void __launch_bounds__(TC_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, bf16* __restrict__ o, float* __restrict__ lse,
                   const int* __restrict__ kv_len, int N, int H, int causal, float scale) {
  using SM = FwdSmem<HD>;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + SM::OFF_Q, sK = base + SM::OFF_K, sV = base + SM::OFF_V, sP = base + SM::OFF_P;
  const uint32_t bar_qk = base + SM::OFF_BAR, bar_v = bar_qk + 8, bar_s = bar_qk + 16, bar_p = bar_qk + 24, bar_o = bar_qk + 32;
  const uint32_t tmem_slot = bar_qk + 40;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + SM::OFF_BAR + 40);
  float* s_red = reinterpret_cast<float*>(gen + SM::OFF_BAR + 64);   // [2 halves][128 rows]: partial max, then partial sum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heavy tiles first: under the causal mask the last query tile sees the most keys
  const int b = blockIdx.z, h = blockIdx.y, q0 = (gridDim.x - 1 - blockIdx.x) * 128;
  const int D = H * HD;
  const int klen = kv_len ? min(kv_len[b], N) : N;
  const int kv_vis = causal ? min(klen, q0 + 128) : klen;        // keys any row of this tile can see
  const int nkv = ((kv_vis + 63) / 64) * 64;                     // MMA N extent (multiple of 64, <= MAX_KV)

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 256);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  const uint32_t tmem_S = tmem, tmem_O = tmem + 384;
  pdl_wait();

  if (warp == 8) {
    if (lane == 0) {
      // ---- loads ----
      const int kv_boxes = nkv / 64;
      mbar_arrive_expect_tx(bar_qk, (SM::PANELS * 2 + SM::PANELS * kv_boxes) * 8192);
      for (int p = 0; p < SM::PANELS; ++p) {
        for (int r = 0; r < 2; ++r)
          tma_load_3d(sQ + p * (PANEL_ROWS_Q * 128) + r * 8192, &map_qkv, bar_qk, h * HD + p * 64, q0 + r * 64, b);
        for (int r = 0; r < kv_boxes; ++r)
          tma_load_3d(sK + p * SM::KV_PANEL_BYTES + r * 8192, &map_qkv, bar_qk, D + h * HD + p * 64, r * 64, b);
      }
      mbar_arrive_expect_tx(bar_v, SM::PANELS * kv_boxes * 8192);
      for (int p = 0; p < SM::PANELS; ++p)
        for (int r = 0; r < kv_boxes; ++r)
          tma_load_3d(sV + p * SM::KV_PANEL_BYTES + r * 8192, &map_qkv, bar_v, 2 * D + h * HD + p * 64, r * 64, b);
    }
    __syncwarp();
    // The MMA issue runs warp-converged with the tcgen05 instructions under one elect.sync per batch (issued back to back).
    {
      // ---- S = Q K^T ----
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      if (elect_one_sync()) {
        for (int c0 = 0; c0 < nkv; c0 += 256) {
          const int n = min(256, nkv - c0);
          const uint32_t id = idesc_bf16(128, n, 0, 0);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint64_t da = smem_desc(sQ + (k / 4) * (PANEL_ROWS_Q * 128) + (k % 4) * 32, 16, 1024);
            const uint64_t db = smem_desc(sK + (k / 4) * SM::KV_PANEL_BYTES + c0 * 128 + (k % 4) * 32, 16, 1024);
            umma_bf16_ss<1>(tmem_S + c0, da, db, id, k > 0 ? 1u : 0u);
          }
        }
        umma_commit<1>(bar_s);
      }
      __syncwarp();
      // ---- O = P V  (P: K-major A operand written by the softmax warps; V: MN-major B operand) ----
      mbar_wait(bar_p, 0);
      mbar_wait(bar_v, 0);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t id = idesc_bf16(128, HD, 0, 1);
        for (int kk = 0; kk < nkv / 16; ++kk) {
          const uint64_t da = smem_desc(sP + (kk / 4) * SM::P_PANEL_BYTES + (kk % 4) * 32, 16, 1024);
          // 16 keys = 16 rows of 128 B; hd panels (64 columns each) are SM::KV_PANEL_BYTES apart (LBO)
          const uint64_t db = smem_desc(sV + kk * 16 * 128, SM::KV_PANEL_BYTES, 1024);
          umma_bf16_ss<1>(tmem_O, da, db, id, kk > 0 ? 1u : 0u);
        }
        umma_commit<1>(bar_o);
      }
    }
    __syncwarp();
  } else {
    // ---- softmax: 8 warps; warp w owns TMEM lanes 32*(w%4).. (query rows) and the 32-column chunks c with c%2 == w/4 ----
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const int q = q0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = scale * LOG2E_F;
    const int nchunks = nkv / 32;
    const int row_lo = q0 + quarter * 32;            // smallest query index of this warp
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // pass 1: row max over this warp's chunks (four independent chains), combined with the partner warp through smem
    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int c = half; c < nchunks; c += 2) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_S + lane_addr + c * 32, v);
      tmem_ld_wait();
      const bool need_mask = (c * 32 + 32 > klen) || (causal && c * 32 + 31 > row_lo);   // warp-uniform
      if (need_mask) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c * 32 + j;
          if (!((col < klen) && (!causal || col <= q))) v[j] = 0xff800000u;   // -inf
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
    }
    s_red[half * 128 + r] = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    softmax_bar();
    const float mx = fmaxf(s_red[r], s_red[128 + r]);
    const float mref = (mx == -INFINITY) ? 0.f : mx * sl2;
    softmax_bar();   // both halves have read the maxima before the slots are reused for the sums
    // pass 2: p = exp2(s*scale*log2e - m), row sum, P -> smem (bf16, canonical K-major 128B-swizzled panels)
    float l4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = half; c < nchunks; c += 2) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_S + lane_addr + c * 32, v);
      tmem_ld_wait();
      const bool need_mask = (c * 32 + 32 > klen) || (causal && c * 32 + 31 > row_lo);
      float p[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = fmaf(__uint_as_float(v[j]), sl2, -mref);
        if (need_mask) {
          const int col = c * 32 + j;
          if (!((col < klen) && (!causal || col <= q))) x = -INFINITY;
        }
        p[j] = fast_exp2(x);
        l4[j & 3] += p[j];
      }
      const uint32_t prow = sP + (c / 2) * SM::P_PANEL_BYTES + r * 128;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + t) ^ static_cast<uint32_t>(r & 7);
        const uint32_t a = prow + (chunk << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(p[t * 8], p[t * 8 + 1])),
                     "r"(pack_bf16x2(p[t * 8 + 2], p[t * 8 + 3])), "r"(pack_bf16x2(p[t * 8 + 4], p[t * 8 + 5])),
                     "r"(pack_bf16x2(p[t * 8 + 6], p[t * 8 + 7]))
                     : "memory");
      }
    }
    s_red[half * 128 + r] = (l4[0] + l4[1]) + (l4[2] + l4[3]);
    tc_fence_before();
    fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    mbar_arrive(bar_p);
    softmax_bar();
    const float l = s_red[r] + s_red[128 + r];
    // ---- epilogue: O / l, the two warps of a lane quarter take alternate 32-column chunks ----
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll 1
    for (int c = half; c < HD / 32; c += 2) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_O + lane_addr + c * 32, v);
      tmem_ld_wait();
      if (q < N) {
        uint4* op = reinterpret_cast<uint4*>(o + (static_cast<int64_t>(b) * N + q) * D + h * HD + c * 32);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          op[t] = make_uint4(pack_bf16x2(__uint_as_float(v[t * 8]) * inv, __uint_as_float(v[t * 8 + 1]) * inv),
                             pack_bf16x2(__uint_as_float(v[t * 8 + 2]) * inv, __uint_as_float(v[t * 8 + 3]) * inv),
                             pack_bf16x2(__uint_as_float(v[t * 8 + 4]) * inv, __uint_as_float(v[t * 8 + 5]) * inv),
                             pack_bf16x2(__uint_as_float(v[t * 8 + 6]) * inv, __uint_as_float(v[t * 8 + 7]) * inv));
      }
    }
    if (half == 0 && q < N) lse[(static_cast<int64_t>(b) * H + h) * N + q] = (l > 0.f) ? (mref + log2f(l)) / LOG2E_F : -INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, 512);
  }
}

template <int HD>
int launch_fwd_tc(const bf16* qkv, bf16* o, float* lse, const int* kv_len, int B, int N, int H, int causal, float scale,
                  cudaStream_t s) {
  using SM = FwdSmem<HD>;
  static bool configured = false;
  if (!configured) {
    VLA_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    configured = true;
  }
  const int D = H * HD;
  CUtensorMap map;
  if (int rc = make_tmap_bf16(qkv, 3 * D, N, B, 3 * static_cast<int64_t>(D), static_cast<int64_t>(N) * 3 * D, 64, 64, &map)) return rc;
  dim3 grid(ceil_div(N, 128), H, B);
  VLA_CHECK_CUDA(vla_launch(attn_fwd_tc_kernel<HD>, grid, dim3(TC_THREADS), static_cast<size_t>(SM::TOTAL), s, map, o, lse, kv_len, N, H,
                            causal, scale));
  ++g_vla_launch_count;
  return 0;
}

}  // namespace

bool attention_tc_supported(int N, int hd) { return (hd == 64 || hd == 128) && N >= 64 && N <= MAX_KV; }

int attention_fwd_tc(const bf16* qkv, bf16* o, float* lse, const int* kv_len, int B, int N, int H, int hd, int causal,
                     float scale, cudaStream_t s) {
  VLA_REQUIRE(attention_tc_supported(N, hd), "attention_fwd_tc: unsupported shape N=%d hd=%d", N, hd);
  if (hd == 64) return launch_fwd_tc<64>(qkv, o, lse, kv_len, B, N, H, causal, scale, s);
  return launch_fwd_tc<128>(qkv, o, lse, kv_len, B, N, H, causal, scale, s);
}
