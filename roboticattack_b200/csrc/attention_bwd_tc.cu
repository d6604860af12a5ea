// tcgen05 attention backward: dQ and dK/dV on the 5th-gen tensor cores with TMEM accumulators, TMA-staged operands and
// TMA-stored results.  Replaces the autograd backward of F.scaled_dot_product_attention (timm Attention; HF
// LlamaSdpaAttention with the causal + key-padding mask) for head dims 64 (DINOv2), 72 (SigLIP, run as 128 with
// zero-filled columns) and 128 (Llama).
//
// Three launches: a small HBM-bound kernel for delta = rowsum(dO o O), then two launches of one PERSISTENT kernel
// template (one CTA per SM looping over work items (batch, head, 128-row tile), heaviest items first).  An item owns
// a 128-row STATIONARY tile and streams 64-row steps of the other side through a 3-slot TMA ring:
//   MODE_DQ  : stationary = 128 queries (Q, dO), streamed = keys (K_s, V_s)
//                S  = Q K_s^T, dP  = dO V_s^T            (phase A, 128 x 64 fp32 each, TMEM stage g % 2)
//                dS = P o (dP - delta)  -> smem bf16      (softmax warps, one thread per query row)
//                dQ += dS K_s                              (phase B; K_s is read as an MN-major B operand)
//   MODE_DKV : stationary = 128 keys (K, V), streamed = queries (Q_s, dO_s); works on the transposed tiles
//                S^T = K Q_s^T, dP^T = V dO_s^T
//                P^T, dS^T -> smem bf16                    (one thread per key row; lse / delta per column from smem)
//                dV += P^T dO_s ; dK += dS^T Q_s           (Q_s / dO_s read as MN-major B operands)
// Roles (384 threads): warp 8 = TMA producer, warps 9 / 10 / 11 = MMA issuers (one elected thread each: S, dP, phase B),
// warps 0-7 = softmax / epilogue (warp w: TMEM lane quarter w % 4, column half w / 4).  Phase A of step g+1 and phase B of step g run on the tensor
// core while the softmax warps work on step g+1; the stationary tiles of the NEXT item are loaded as soon as the last
// phase A of the current item has read them, so the load latency hides behind the item's tail and epilogue.  The
// epilogue applies the scale (and the rotary embedding's backward for Llama), stages bf16 rows in shared memory and
// writes them with TMA stores (rows / columns outside the tensor are clipped by the TMA unit).
// No atomics: every output element is written exactly once, so results are bit-reproducible.
#include <math.h>

#include <type_traits>

#include "kernels.h"
#include "tma_desc.h"

namespace {

constexpr int BWD_THREADS = 384;
constexpr int TILE = 128;   // stationary rows per item (UMMA M)
constexpr int STEP = 64;    // streamed rows per step
constexpr int RING = 3;
constexpr float LOG2E_B = 1.4426950408889634f;
enum { MODE_DQ = 0, MODE_DKV = 1 };

// shared-memory operand descriptor, 128B swizzle, version 1 (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// bf16 x bf16 -> fp32, M x N, A K-major, B K-major (b_mn = 0) or MN-major (b_mn = 1)
__device__ __forceinline__ uint32_t idesc(int m, int n, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

#ifdef VLA_ATTN_TIMING
__device__ unsigned long long g_attn_dbg[2][64];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define STAMP(mode, who, i)                                                               \
  do {                                                                                    \
    if (blockIdx.x == 0 && (who) && (i) < 64) g_attn_dbg[mode][(i)] = gtime();            \
  } while (0)
#else
#define STAMP(mode, who, i) do {} while (0)
#endif

// ---------------------------------------------------------------------------------------------------------
// delta[b, h, n] = sum_c dO[b, n, h, c] * O[b, n, h, c]   (fp32).  One CTA per row (b, n): warp w takes heads w, w + 8, ...
// (32-bit index math only; all of a warp's loads are issued before the reductions), lane l takes 4 columns.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout,
                                                         float* __restrict__ delta, int N, int H, int hd) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x;                 // b * N + n
  const int b = row / N, n = row - b * N;
  const size_t base = static_cast<size_t>(row) * H * hd;
  constexpr int MAXH = 4;                     // heads per warp and pass
  for (int h0 = warp; h0 < H; h0 += 8 * MAXH) {
    uint2 av[MAXH][2], dv[MAXH][2];
#pragma unroll
    for (int i = 0; i < MAXH; ++i) {
      const int h = h0 + 8 * i;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = lane * 4 + j * 128;
        const bool ok = h < H && c < hd;
        av[i][j] = ok ? __ldg(reinterpret_cast<const uint2*>(o + base + h * hd + c)) : make_uint2(0u, 0u);
        dv[i][j] = ok ? __ldg(reinterpret_cast<const uint2*>(dout + base + h * hd + c)) : make_uint2(0u, 0u);
      }
    }
#pragma unroll
    for (int i = 0; i < MAXH; ++i) {
      const int h = h0 + 8 * i;
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float2 a0 = unpack_bf16x2(av[i][j].x), a1 = unpack_bf16x2(av[i][j].y);
        const float2 d0 = unpack_bf16x2(dv[i][j].x), d1 = unpack_bf16x2(dv[i][j].y);
        sum += (a0.x * d0.x + a0.y * d0.y) + (a1.x * d1.x + a1.y * d1.y);
      }
      sum = warp_sum(sum);
      if (lane == 0 && h < H) delta[(static_cast<size_t>(b) * H + h) * N + n] = sum;
    }
  }
}

template <int HD, int MODE>
struct BwdSmem {
  static constexpr int PANELS = HD / 64;                      // 64-column (128-byte) panels of a head
  static constexpr int STAT_BYTES = PANELS * TILE * 128;      // one stationary operand
  static constexpr int STREAM_PANEL = STEP * 128;             // one panel of one streamed operand
  static constexpr int STREAM_BYTES = PANELS * STREAM_PANEL;
  static constexpr int SLOT_BYTES = 2 * STREAM_BYTES;         // x_s then y_s
  static constexpr int PBUF_BYTES = TILE * 128;               // 128 rows x 64 columns bf16
  // P / dS buffers (single stage) double as the staging area of the TMA-stored results
  static constexpr int PD_MIN = (MODE == MODE_DKV) ? 2 * PBUF_BYTES : PBUF_BYTES;
  static constexpr int PD_BYTES = PD_MIN > STAT_BYTES ? PD_MIN : STAT_BYTES;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_Y = OFF_X + STAT_BYTES;
  static constexpr int OFF_RING = OFF_Y + STAT_BYTES;
  static constexpr int OFF_PD = OFF_RING + RING * SLOT_BYTES;
  static constexpr int OFF_AUX = OFF_PD + PD_BYTES;           // DKV: per ring slot [64 lse*log2e | 64 delta]
  static constexpr int AUX_BYTES = (MODE == MODE_DKV) ? RING * 2 * STEP * 4 : 0;
  static constexpr int OFF_BAR = OFF_AUX + AUX_BYTES;
  static constexpr int TOTAL = OFF_BAR + 192 + 1024 /* alignment slack */;
  static_assert(TOTAL <= 232448, "shared memory budget");
};

struct BwdArgs {
  const float* lse;
  const float* delta;
  const int* kv_len;
  const float* rope_cos;
  const float* rope_sin;
  int B, N, H, hd, causal, rope_L, ntiles, nitems;
  int shift;   // MODE_DQ under the causal mask: query tile i covers rows [128 i - shift, +128) (see Fwd3Args::shift)
  float scale;
};

// KS = phase-A contraction steps of 16 columns (HD / 16, or 5 for head dim 72: columns 72..79 are zero filled)
template <int HD, int KS, int MODE>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                   const __grid_constant__ CUtensorMap map_dqkv, const BwdArgs a) {
  using SM = BwdSmem<HD, MODE>;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  STAMP(MODE, threadIdx.x == 0, 0);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sX = base + SM::OFF_X, sY = base + SM::OFF_Y, sRing = base + SM::OFF_RING, sPD = base + SM::OFF_PD;
  const uint32_t bars = base + SM::OFF_BAR;
  const uint32_t bar_stat = bars;                                  // stationary tiles of the item landed
  const uint32_t bar_xfree = bars + 8;                             // last phase A of the item has read X, Y
  const uint32_t bar_out = bars + 16;                              // last phase B of the item completed
  const uint32_t bar_outfree = bars + 24;                          // epilogue has read the TMEM accumulators (8 warps)
  auto bar_full = [&](int i) { return bars + 32 + 8 * i; };        // ring slot i landed (TMA tx [+ stats])
  auto bar_empty = [&](int i) { return bars + 56 + 8 * i; };       // phase B that read ring slot i completed
  auto bar_sready = [&](int t) { return bars + 80 + 8 * t; };      // phase A into TMEM stage t completed
  auto bar_pdone = [&](int t) { return bars + 96 + 8 * t; };       // softmax warps wrote P/dS for a step of stage t (8 warps)
  auto bar_bdone = [&](int t) { return bars + 112 + 8 * t; };      // phase B of a step of stage t completed
  const uint32_t tmem_slot = bars + 128;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + SM::OFF_BAR + 128);
  float* aux = reinterpret_cast<float*>(gen + SM::OFF_AUX);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, H = a.H, causal = a.causal;
  const int BH = a.B * H;
  const int G = gridDim.x;

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_dqkv);
    mbar_init(bar_stat, 1);
    mbar_init(bar_xfree, 2);
    mbar_init(bar_out, 1);
    mbar_init(bar_outfree, 8);
    for (int i = 0; i < RING; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_sready(t), 2);
      mbar_init(bar_pdone(t), 8);
      mbar_init(bar_bdone(t), 1);
    }
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
  const uint32_t tmem_out0 = tmem + 256, tmem_out1 = tmem + 256 + HD;
  STAMP(MODE, threadIdx.x == 0, 1);
  pdl_wait();
  STAMP(MODE, threadIdx.x == 0, 2);

  // Work item i of this CTA: serpentine assignment over the heavy-first item order (round i, alternating direction).
  struct Item {
    int b, h, t0, klen, c_begin, nsteps, c_end;
  };
  auto get_item = [&](int i, Item& it) -> bool {
    const int k = i * G + ((i & 1) ? (G - 1 - static_cast<int>(blockIdx.x)) : static_cast<int>(blockIdx.x));
    if (k >= a.nitems) return false;
    const int y = k / BH, bh = k - y * BH;
    it.b = bh / H;
    it.h = bh - it.b * H;
    // heavy tiles first: under the causal mask the LAST query tile sees the most keys, the FIRST key tile the most queries
    const int ti = (MODE == MODE_DQ && causal) ? (a.ntiles - 1 - y) : y;
    it.t0 = ti * TILE - ((MODE == MODE_DQ) ? a.shift : 0);
    it.klen = a.kv_len ? min(a.kv_len[it.b], N) : N;
    if (MODE == MODE_DQ) {
      it.c_begin = 0;
      it.c_end = causal ? min(it.klen, it.t0 + TILE) : it.klen;
    } else {
      it.c_begin = causal ? it.t0 : 0;
      it.c_end = (it.t0 < it.klen) ? N : 0;   // a key tile entirely behind the padding boundary has zero gradient
    }
    it.nsteps = it.c_end > it.c_begin ? (it.c_end - it.c_begin + STEP - 1) / STEP : 0;
    return true;
  };
  auto step_cols = [&](const Item& it, int s) { return min(STEP, ((it.c_end - (it.c_begin + s * STEP) + 15) / 16) * 16); };

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    const CUtensorMap* mapY = (MODE == MODE_DQ) ? &map_do : &map_qkv;   // dO | V
    const CUtensorMap* mapy = (MODE == MODE_DQ) ? &map_qkv : &map_do;   // V_s | dO_s
    int g = 0, n_it = 0;
    Item it;
    for (int i = 0; get_item(i, it); ++i) {
      if (it.nsteps == 0) continue;
      const int headX = (MODE == MODE_DQ) ? it.h : H + it.h;          // Q | K
      const int headY = (MODE == MODE_DQ) ? it.h : 2 * H + it.h;      // dO | V
      const int headx = (MODE == MODE_DQ) ? H + it.h : it.h;          // K_s | Q_s
      const int heady = (MODE == MODE_DQ) ? 2 * H + it.h : it.h;      // V_s | dO_s
      if (n_it > 0) mbar_wait(bar_xfree, (n_it - 1) & 1);
      if (lane == 0) {
        // a 64-row box in front of the sequence (shifted first query tile) is not loaded: its rows belong to inactive warps
        mbar_arrive_expect_tx(bar_stat, it.t0 < 0 ? SM::STAT_BYTES : 2 * SM::STAT_BYTES);
#pragma unroll
        for (int p = 0; p < SM::PANELS; ++p)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (it.t0 + r * 64 < 0) continue;
            tma_load_4d(sX + p * (TILE * 128) + r * 8192, &map_qkv, bar_stat, p * 64, headX, it.t0 + r * 64, it.b);
            tma_load_4d(sY + p * (TILE * 128) + r * 8192, mapY, bar_stat, p * 64, headY, it.t0 + r * 64, it.b);
          }
      }
      for (int s = 0; s < it.nsteps; ++s, ++g) {
        const int slot = g % RING;
        const int c0 = it.c_begin + s * STEP;
        if (g >= RING) mbar_wait(bar_empty(slot), ((g / RING) - 1) & 1);
        if (MODE == MODE_DKV) {
          float* st = aux + slot * (2 * STEP);
          const int64_t sb = (static_cast<int64_t>(it.b) * H + it.h) * N;
#pragma unroll
          for (int i2 = 0; i2 < STEP / 32; ++i2) {
            const int q = c0 + lane + i2 * 32;
            float l2 = INFINITY, dl = 0.f;
            if (q < N) {
              const float l = a.lse[sb + q];
              l2 = (l == -INFINITY) ? INFINITY : l * LOG2E_B;
              dl = a.delta[sb + q];
            }
            st[lane + i2 * 32] = l2;
            st[STEP + lane + i2 * 32] = dl;
          }
          __syncwarp();
        }
        if (lane == 0) {
          const uint32_t xs = sRing + slot * SM::SLOT_BYTES, ys = xs + SM::STREAM_BYTES;
          mbar_arrive_expect_tx(bar_full(slot), SM::SLOT_BYTES);
#pragma unroll
          for (int p = 0; p < SM::PANELS; ++p) {
            tma_load_4d(xs + p * SM::STREAM_PANEL, &map_qkv, bar_full(slot), p * 64, headx, c0, it.b);
            tma_load_4d(ys + p * SM::STREAM_PANEL, mapy, bar_full(slot), p * 64, heady, c0, it.b);
          }
        }
      }
      ++n_it;
    }
  } else if (warp == 9 || warp == 10) {
    // ------------------------------------------------------------------ phase-A MMA issuers
    // Issuing a tcgen05.mma costs ~100 cycles of a single thread (descriptor moves to uniform registers + the
    // single-thread election sequence), more than these small MMAs take to execute, so the issue work is spread over
    // three warps: warp 9 -> S (X . x_s^T), warp 10 -> dP (Y . y_s^T), warp 11 -> phase B.
    {
      const bool second = (warp == 10);
      const uint64_t dA0 = sdesc(second ? sY : sX, 16, 1024);
      int g = 0, n_it = 0;
      Item it;
      for (int i = 0; get_item(i, it); ++i) {
        if (it.nsteps == 0) continue;
        mbar_wait(bar_stat, n_it & 1);
        for (int s = 0; s < it.nsteps; ++s, ++g) {
          const int slot = g % RING, t = g & 1;
          const uint32_t idA = idesc(TILE, step_cols(it, s), 0);
          const uint64_t dB = sdesc(sRing + slot * SM::SLOT_BYTES + (second ? SM::STREAM_BYTES : 0), 16, 1024);
          if (g >= 2) mbar_wait(bar_pdone(t), ((g - 2) >> 1) & 1);   // the softmax warps have read TMEM stage t (step g - 2)
          mbar_wait(bar_full(slot), (g / RING) & 1);
          tc_fence_after();
          const uint32_t tS = tmem + t * 128 + (second ? 64 : 0);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < KS; ++k)
              umma_bf16_ss<1>(tS, dA0 + (((k >> 2) * (TILE * 128) + (k & 3) * 32) >> 4),
                              dB + (((k >> 2) * SM::STREAM_PANEL + (k & 3) * 32) >> 4), idA, k > 0 ? 1u : 0u);
            umma_commit<1>(bar_sready(t));
            if (s == it.nsteps - 1) umma_commit<1>(bar_xfree);
          }
          __syncwarp();
        }
        ++n_it;
      }
    }
  } else if (warp == 11) {
    // ------------------------------------------------------------------ phase-B MMA issuer
    {
      const uint32_t idB = idesc(TILE, HD, 1);
      const uint64_t ddS0 = sdesc(sPD + ((MODE == MODE_DKV) ? SM::PBUF_BYTES : 0), 16, 1024);   // dS (K-major A)
      const uint64_t dP0 = sdesc(sPD, 16, 1024);                                                 // P (MODE_DKV)
      int g = 0, n_it = 0;
      Item it;
      for (int i = 0; get_item(i, it); ++i) {
        if (it.nsteps == 0) continue;
        for (int s = 0; s < it.nsteps; ++s, ++g) {
          const int slot = g % RING, t = g & 1;
          const int nk = step_cols(it, s) / 16;
          const uint32_t xs = sRing + slot * SM::SLOT_BYTES;
          // MN-major B operands: 16 streamed rows per K step (2048 B), 64-column blocks STREAM_PANEL apart
          const uint64_t dxm = sdesc(xs, SM::STREAM_PANEL, 1024), dym = sdesc(xs + SM::STREAM_BYTES, SM::STREAM_PANEL, 1024);
          STAMP(MODE, g == 2 && lane == 0, 50);
          mbar_wait(bar_pdone(t), (g >> 1) & 1);
          STAMP(MODE, g == 2 && lane == 0, 51);
          if (s == 0 && n_it > 0) mbar_wait(bar_outfree, (n_it - 1) & 1);
          tc_fence_after();
          if (elect_one_sync()) {
          if (MODE == MODE_DQ) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)   // dQ += dS K_s
              if (kk < nk) umma_bf16_ss<1>(tmem_out0, ddS0 + kk * 2, dxm + kk * 128, idB, (s | kk) ? 1u : 0u);
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)   // dV += P^T dO_s
              if (kk < nk) umma_bf16_ss<1>(tmem_out0, dP0 + kk * 2, dym + kk * 128, idB, (s | kk) ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)   // dK += dS^T Q_s
              if (kk < nk) umma_bf16_ss<1>(tmem_out1, ddS0 + kk * 2, dxm + kk * 128, idB, (s | kk) ? 1u : 0u);
          }
          umma_commit<1>(bar_empty(slot));
          umma_commit<1>(bar_bdone(t));
          if (s == it.nsteps - 1) umma_commit<1>(bar_out);
          }
          __syncwarp();
          STAMP(MODE, g == 2 && lane == 0, 52);
        }
        ++n_it;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / epilogue warps
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = a.scale * LOG2E_B;
    const uint32_t sdS = sPD + ((MODE == MODE_DKV) ? SM::PBUF_BYTES : 0) + r * 128;
    const uint32_t sPt = sPD + r * 128;   // MODE_DKV only
    const int D = H * a.hd;
    int g = 0, n_it = 0;
    bool store_pending = false;
    auto drain_store = [&]() {
      if (store_pending) {
        if (threadIdx.x == 0) tma_store_wait_read();
        softmax_bar();
        store_pending = false;
      }
    };
    Item it;
    for (int i = 0; get_item(i, it); ++i) {
      const int srow = it.t0 + r;                     // query (MODE_DQ) or key (MODE_DKV) index of this thread's row
      const int warp_row0 = it.t0 + quarter * 32;
      const int klen = it.klen;
      const bool warp_active = warp_row0 >= 0 && warp_row0 < ((MODE == MODE_DQ) ? N : klen);
      float lse2_r = INFINITY, delta_r = 0.f;
      if (MODE == MODE_DQ && srow >= 0 && srow < N) {
        const int64_t si = (static_cast<int64_t>(it.b) * H + it.h) * N + srow;
        const float l = __ldg(a.lse + si);
        delta_r = __ldg(a.delta + si);
        lse2_r = (l == -INFINITY) ? INFINITY : l * LOG2E_B;
      }
      // prefetch the rotary tables of this thread's row (consumed in the epilogue)
      for (int s = 0; s < it.nsteps; ++s, ++g) {
        const int slot = g % RING, t = g & 1;
        const int ncols = step_cols(it, s);
        const int col0 = it.c_begin + s * STEP + half * 32;   // first streamed index of this warp's 32-column chunk
        mbar_wait(bar_sready(t), (g >> 1) & 1);
        STAMP(MODE, threadIdx.x == 0, 4 + 2 * g);
        tc_fence_after();
        if (MODE == MODE_DKV) mbar_wait(bar_full(slot), (g / RING) & 1);   // acquire the producer's lse / delta stores
        if (s == 0) drain_store();
        if (half * 32 < ncols) {
          // a 32 x 32 block entirely above the causal diagonal (or behind the key padding): P = dS = 0 without touching TMEM
          bool chunk_masked;
          if (MODE == MODE_DQ) chunk_masked = (col0 >= klen) || (causal && col0 > warp_row0 + 31);
          else chunk_masked = causal && warp_row0 > col0 + 31;
#ifdef VLA_ATTN_EXPERIMENT_NOSOFTMAX   // diagnostics only (results are wrong): the softmax stage does no TMEM reads and no math
          if (false) {
#else
          if (warp_active && !chunk_masked) {
#endif
            uint32_t sv[32], dv[32];
            tmem_ld_32x32(tmem + lane_addr + t * 128 + half * 32, sv);
            tmem_ld_32x32(tmem + lane_addr + t * 128 + 64 + half * 32, dv);
            tmem_ld_wait();
            STAMP(MODE, threadIdx.x == 0 && g == 2, 44);
            bool need_mask;
            if (MODE == MODE_DQ) need_mask = (col0 + 32 > klen) || (causal && col0 + 31 > warp_row0);
            else need_mask = (warp_row0 + 31 >= klen) || (causal && warp_row0 + 31 > col0);
            const float* st = aux + slot * (2 * STEP) + half * 32;   // MODE_DKV: [lse2 | delta] of this chunk's queries
            uint32_t pk[16], dk[16];
            auto compute = [&](auto masked_tag) {
              constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                float l2[4], dl[4];
                if (MODE == MODE_DKV) {
                  const float4 la = *reinterpret_cast<const float4*>(st + j4 * 4);
                  const float4 lc = *reinterpret_cast<const float4*>(st + STEP + j4 * 4);
                  l2[0] = la.x, l2[1] = la.y, l2[2] = la.z, l2[3] = la.w;
                  dl[0] = lc.x, dl[1] = lc.y, dl[2] = lc.z, dl[3] = lc.w;
                } else {
#pragma unroll
                  for (int u = 0; u < 4; ++u) l2[u] = lse2_r, dl[u] = delta_r;
                }
                float p[4], ds[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const int j = j4 * 4 + u;
                  float x = fmaf(__uint_as_float(sv[j]), sl2, -l2[u]);
                  if (MASKED) {
                    const int key = (MODE == MODE_DQ) ? col0 + j : srow;
                    const int q = (MODE == MODE_DQ) ? srow : col0 + j;
                    if (!((key < klen) && (!causal || key <= q))) x = -INFINITY;
                  }
                  p[u] = ex2(x);
                  ds[u] = p[u] * (__uint_as_float(dv[j]) - dl[u]);
                }
                pk[j4 * 2] = pack_bf16x2(p[0], p[1]);
                pk[j4 * 2 + 1] = pack_bf16x2(p[2], p[3]);
                dk[j4 * 2] = pack_bf16x2(ds[0], ds[1]);
                dk[j4 * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
              }
            };
            if (need_mask) compute(std::true_type{}); else compute(std::false_type{});
            STAMP(MODE, threadIdx.x == 0 && g == 2, 45);
            // the single P / dS buffer is free once phase B of the previous step (and the previous item's stores) is done
            if (g >= 1) mbar_wait(bar_bdone((g - 1) & 1), ((g - 1) >> 1) & 1);
            STAMP(MODE, threadIdx.x == 0 && g == 2, 48);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint32_t chunk = (static_cast<uint32_t>(half * 4 + q4) ^ static_cast<uint32_t>(r & 7)) << 4;
              st_shared_v4(sdS + chunk, dk[q4 * 4], dk[q4 * 4 + 1], dk[q4 * 4 + 2], dk[q4 * 4 + 3]);
              if (MODE == MODE_DKV) st_shared_v4(sPt + chunk, pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
            }
            STAMP(MODE, threadIdx.x == 0 && g == 2, 46);
          } else if (warp_active || s == 0) {
            // masked block: zeros for this step; rows outside the sequence: zero once per item (the tensor core reads all
            // 128 rows of the A operand)
            if (g >= 1) mbar_wait(bar_bdone((g - 1) & 1), ((g - 1) >> 1) & 1);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              const uint32_t chunk = (static_cast<uint32_t>(q4) ^ static_cast<uint32_t>(r & 7)) << 4;
              if (half == (q4 >> 2)) {
                st_shared_v4(sdS + chunk, 0u, 0u, 0u, 0u);
                if (MODE == MODE_DKV) st_shared_v4(sPt + chunk, 0u, 0u, 0u, 0u);
              }
            }
          }
        }
        tc_fence_before();
        fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
        STAMP(MODE, threadIdx.x == 0 && g == 2, 47);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pdone(t));
        STAMP(MODE, threadIdx.x == 0, 5 + 2 * g);
      }

      // ---- epilogue: TMEM -> registers -> (scale, RoPE backward) -> bf16 rows staged in the P/dS area -> TMA store ----
      const bool have_acc = it.nsteps > 0;
      if (have_acc) {
        mbar_wait(bar_out, n_it & 1);
        tc_fence_after();
      }
      STAMP(MODE, threadIdx.x == 0 && i == 0, 60);
      drain_store();   // (items without steps reach the staging writes directly)
      constexpr int NOUT = (MODE == MODE_DKV) ? 2 : 1;
      constexpr bool SEQ = NOUT * SM::STAT_BYTES > SM::PD_BYTES;   // stage the two results one after the other
#pragma unroll
      for (int oi = 0; oi < NOUT; ++oi) {
        const uint32_t tacc = (oi == 0 ? tmem_out0 : tmem_out1) + lane_addr;
        const uint32_t stage = sPD + ((SEQ || oi == 0) ? 0 : SM::STAT_BYTES);
        const bool is_v = (MODE == MODE_DKV) && oi == 0;
        const float osc = is_v ? 1.f : a.scale;
        const bool rope = (HD == 128) && !is_v && a.rope_cos != nullptr;
        if (HD == 128 && rope) {
          uint32_t lo[32], hi[32];
          if (have_acc) {
            tmem_ld_32x32(tacc + half * 32, lo);
            tmem_ld_32x32(tacc + 64 + half * 32, hi);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) lo[j] = hi[j] = 0u;
          }
          const int pos = max(srow, 0) % a.rope_L;   // rows in front of the sequence are never stored
          const float* cp = a.rope_cos + pos * 64 + half * 32;
          const float* sp = a.rope_sin + pos * 64 + half * 32;
#pragma unroll
          for (int t4 = 0; t4 < 4; ++t4) {
            uint32_t wl[4], wh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = t4 * 8 + u * 2;
              const float2 c = *reinterpret_cast<const float2*>(cp + j);
              const float2 sn = *reinterpret_cast<const float2*>(sp + j);
              const float l0 = rbf(__uint_as_float(lo[j]) * osc), l1 = rbf(__uint_as_float(lo[j + 1]) * osc);
              const float h0 = rbf(__uint_as_float(hi[j]) * osc), h1 = rbf(__uint_as_float(hi[j + 1]) * osc);
              wl[u] = pack_bf16x2(rbf(l0 * c.x) + rbf(h0 * sn.x), rbf(l1 * c.y) + rbf(h1 * sn.y));
              wh[u] = pack_bf16x2(rbf(h0 * c.x) + rbf(-l0 * sn.x), rbf(h1 * c.y) + rbf(-l1 * sn.y));
            }
            // columns [32 half, +32) live in panel 0, columns 64 + [32 half, +32) in panel 1
            const uint32_t chunk = (static_cast<uint32_t>(half * 4 + t4) ^ static_cast<uint32_t>(r & 7)) << 4;
            st_shared_v4(stage + r * 128 + chunk, wl[0], wl[1], wl[2], wl[3]);
            st_shared_v4(stage + TILE * 128 + r * 128 + chunk, wh[0], wh[1], wh[2], wh[3]);
          }
        } else {
#pragma unroll
          for (int c = half; c < HD / 32; c += 2) {
            uint32_t v[32];
            if (have_acc) {
              tmem_ld_32x32(tacc + c * 32, v);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4) {
              const uint32_t chunk = (static_cast<uint32_t>((c & 1) * 4 + t4) ^ static_cast<uint32_t>(r & 7)) << 4;
              st_shared_v4(stage + (c >> 1) * (TILE * 128) + r * 128 + chunk,
                           pack_bf16x2(__uint_as_float(v[t4 * 8]) * osc, __uint_as_float(v[t4 * 8 + 1]) * osc),
                           pack_bf16x2(__uint_as_float(v[t4 * 8 + 2]) * osc, __uint_as_float(v[t4 * 8 + 3]) * osc),
                           pack_bf16x2(__uint_as_float(v[t4 * 8 + 4]) * osc, __uint_as_float(v[t4 * 8 + 5]) * osc),
                           pack_bf16x2(__uint_as_float(v[t4 * 8 + 6]) * osc, __uint_as_float(v[t4 * 8 + 7]) * osc));
            }
          }
        }
        if (oi == NOUT - 1 && have_acc) {   // every TMEM read of this item is done: the next item's phase B may overwrite
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_outfree);
        }
        if (SEQ || oi == NOUT - 1) {
          fence_proxy_async();
          softmax_bar();
          if (threadIdx.x == 0) {
            const int o_lo = SEQ ? oi : 0, o_hi = SEQ ? oi : NOUT - 1;
            for (int oo = o_lo; oo <= o_hi; ++oo) {
              const uint32_t stg = sPD + ((SEQ || oo == 0) ? 0 : SM::STAT_BYTES);
              // d(q) | d(k) | d(v) thirds of dqkv: MODE_DQ -> q; MODE_DKV: result 0 = dV, result 1 = dK
              const int head_out = (MODE == MODE_DQ) ? it.h : (oo == 0 ? 2 * H + it.h : H + it.h);
#pragma unroll
              for (int p = 0; p < SM::PANELS; ++p)
#pragma unroll
                for (int rb = 0; rb < 2; ++rb)
                  if (it.t0 + rb * 64 >= 0 && it.t0 + rb * 64 < N && p * 64 < a.hd)
                    tma_store_4d(&map_dqkv, stg + p * (TILE * 128) + rb * 8192, p * 64, head_out, it.t0 + rb * 64, it.b);
            }
            tma_store_commit();
            if (oi != NOUT - 1) tma_store_wait_read();
          }
          // The staging area (= P / dS buffers) is reusable once the store has read it.  Between two results of one item
          // that wait is taken here; after the last result it is deferred to the next item's first step (drain_store),
          // by which time the store has long finished, so the epilogue does not stall on it.
          if (oi != NOUT - 1) softmax_bar();
          else store_pending = true;
        }
      }
      STAMP(MODE, threadIdx.x == 0 && i == 0, 61);
      if (have_acc) ++n_it;
    }
    if (threadIdx.x == 0) tma_store_wait_all();
  }

  STAMP(MODE, threadIdx.x == 0, 62);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, 512);
  }
  STAMP(MODE, threadIdx.x == 0, 63);
}

int g_num_sms_attn = 0;

template <int HD, int KS>
int launch_bwd_tc(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                  const int* kv_len, int B, int N, int H, int hd, int causal, const float* rope_cos, const float* rope_sin,
                  int rope_L, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    VLA_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<HD, KS, MODE_DQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        BwdSmem<HD, MODE_DQ>::TOTAL));
    VLA_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<HD, KS, MODE_DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        BwdSmem<HD, MODE_DKV>::TOTAL));
    configured = true;
  }
  if (g_num_sms_attn == 0) {
    int dev = 0;
    VLA_CHECK_CUDA(cudaGetDevice(&dev));
    VLA_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms_attn, cudaDevAttrMultiProcessorCount, dev));
  }
  const int64_t D = static_cast<int64_t>(H) * hd;
  CUtensorMap map_qkv, map_do, map_dqkv;
  if (int rc = make_tmap_bf16_4d(qkv, hd, 3 * H, N, B, hd, 3 * D, N * 3 * D, 64, 64, &map_qkv)) return rc;
  if (int rc = make_tmap_bf16_4d(dout, hd, H, N, B, hd, D, N * D, 64, 64, &map_do)) return rc;
  if (int rc = make_tmap_bf16_4d(dqkv, hd, 3 * H, N, B, hd, 3 * D, N * 3 * D, 64, 64, &map_dqkv)) return rc;
  BwdArgs a;
  a.lse = lse;
  a.delta = delta;
  a.kv_len = kv_len;
  a.rope_cos = rope_cos;
  a.rope_sin = rope_sin;
  a.B = B, a.N = N, a.H = H, a.hd = hd, a.causal = causal, a.rope_L = rope_L > 0 ? rope_L : 1;
  a.ntiles = ceil_div(N, TILE);
  a.nitems = a.ntiles * B * H;
  a.shift = attention_tile_shift(N, causal);
  a.scale = 1.f / sqrtf(static_cast<float>(hd));
  VLA_REQUIRE(hd <= 256, "attention_bwd_tc: head dim %d too large for the delta kernel", hd);
  // o == nullptr: delta was already written by the producer of dO (the delta epilogue of the o_proj backward GEMM)
  if (o != nullptr) VLA_CHECK_CUDA(vla_launch(attn_delta_kernel, dim3(static_cast<unsigned>(B * N)), dim3(256), 0, s, o, dout, delta, N, H, hd));
  const int sms = (g_vla_sm_limit > 0 && g_vla_sm_limit < g_num_sms_attn) ? g_vla_sm_limit : g_num_sms_attn;
  const int grid = a.nitems < sms ? a.nitems : sms;
  VLA_CHECK_CUDA(vla_launch(attn_bwd_tc_kernel<HD, KS, MODE_DQ>, dim3(grid), dim3(BWD_THREADS), static_cast<size_t>(BwdSmem<HD, MODE_DQ>::TOTAL),
                            s, map_qkv, map_do, map_dqkv, a));
  VLA_CHECK_CUDA(vla_launch(attn_bwd_tc_kernel<HD, KS, MODE_DKV>, dim3(grid), dim3(BWD_THREADS),
                            static_cast<size_t>(BwdSmem<HD, MODE_DKV>::TOTAL), s, map_qkv, map_do, map_dqkv, a));
  g_vla_launch_count += o != nullptr ? 3 : 2;
  return 0;
}

}  // namespace

#ifdef VLA_ATTN_TIMING
extern "C" int vla_attn_dbg_read(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_attn_dbg, sizeof(unsigned long long) * 128) == cudaSuccess ? 0 : 1;
}
#endif

bool attention_bwd_tc_supported(int N, int hd) { return (hd == 64 || hd == 72 || hd == 128) && N >= 1; }

int attention_bwd_tc(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                     const int* kv_len, int B, int N, int H, int hd, int causal, const float* rope_cos, const float* rope_sin,
                     int rope_L, cudaStream_t s) {
  VLA_REQUIRE(attention_bwd_tc_supported(N, hd), "attention_bwd_tc: unsupported shape N=%d hd=%d", N, hd);
  if (hd == 64) return launch_bwd_tc<64, 4>(qkv, o, dout, lse, delta, dqkv, kv_len, B, N, H, hd, causal, rope_cos, rope_sin, rope_L, s);
  if (hd == 72) return launch_bwd_tc<128, 5>(qkv, o, dout, lse, delta, dqkv, kv_len, B, N, H, hd, causal, rope_cos, rope_sin, rope_L, s);
  return launch_bwd_tc<128, 8>(qkv, o, dout, lse, delta, dqkv, kv_len, B, N, H, hd, causal, rope_cos, rope_sin, rope_L, s);
}
