// tcgen05 attention backward: dQ and dK/dV on the 5th-gen tensor cores with TMEM accumulators and TMA-staged operands.
// Replaces the autograd backward of F.scaled_dot_product_attention (timm Attention; HF LlamaSdpaAttention with the
// causal + key-padding mask) for head dims 64 (DINOv2), 72 (SigLIP, run as 128 with zero-filled columns) and 128
// (Llama).
//
// Two launches of one kernel template, each CTA owning a 128-row STATIONARY tile and streaming 64-row steps of the other
// side through a 3-slot TMA ring:
//   MODE_DQ  : stationary = 128 queries (Q, dO), streamed = keys (K_s, V_s)
//                S  = Q K_s^T, dP  = dO V_s^T            (phase A, 128 x 64 fp32 each, TMEM stage s % 2)
//                dS = P o (dP - delta)  -> smem bf16      (softmax warps, one thread per query row)
//                dQ += dS K_s                              (phase B; K_s is read as an MN-major B operand)
//              also produces delta = rowsum(dO o O) for the second launch.
//   MODE_DKV : stationary = 128 keys (K, V), streamed = queries (Q_s, dO_s); works on the transposed tiles
//                S^T = K Q_s^T, dP^T = V dO_s^T
//                P^T, dS^T -> smem bf16                    (one thread per key row; lse / delta per column from smem)
//                dV += P^T dO_s ; dK += dS^T Q_s           (Q_s / dO_s read as MN-major B operands)
// Pipeline per CTA (320 threads): warp 8 = TMA producer, warp 9 = MMA issuer (one elected thread), warps 0-7 = softmax /
// epilogue (warp w: TMEM lane quarter w % 4, column half w / 4).  Phase A of step s+1 and phase B of step s run on the
// tensor core while the softmax warps work on step s+1 (two TMEM stages, two P/dS smem buffers).
// No atomics: every output element is written by exactly one thread, so results are bit-reproducible.
#include <math.h>

#include "kernels.h"
#include "tma_desc.h"

namespace {

constexpr int BWD_THREADS = 320;
constexpr int TILE = 128;   // stationary rows per CTA (UMMA M)
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
#define STAMP(mode, who, i)                                                                        \
  do {                                                                                             \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (who)) g_attn_dbg[mode][(i)] = gtime();              \
  } while (0)
#else
#define STAMP(mode, who, i) do {} while (0)
#endif

template <int HD, int MODE>
struct BwdSmem {
  static constexpr int PANELS = HD / 64;                      // 64-column (128-byte) panels of a head
  static constexpr int STAT_BYTES = PANELS * TILE * 128;      // one stationary operand
  static constexpr int STREAM_PANEL = STEP * 128;             // one panel of one streamed operand
  static constexpr int STREAM_BYTES = PANELS * STREAM_PANEL;
  static constexpr int SLOT_BYTES = 2 * STREAM_BYTES;         // x_s then y_s
  static constexpr int PBUF_BYTES = TILE * 128;               // 128 rows x 64 columns bf16
  static constexpr int NPBUF = (MODE == MODE_DKV) ? 4 : 2;    // per TMEM stage: dS (and P for MODE_DKV)
  static constexpr int OFF_X = 0;
  static constexpr int OFF_Y = OFF_X + STAT_BYTES;
  static constexpr int OFF_RING = OFF_Y + STAT_BYTES;
  static constexpr int OFF_P = OFF_RING + RING * SLOT_BYTES;
  static constexpr int OFF_AUX = OFF_P + NPBUF * PBUF_BYTES;  // DKV: per ring slot [64 lse*log2e | 64 delta]; DQ: delta partial sums [2][128]
  static constexpr int AUX_BYTES = (MODE == MODE_DKV) ? RING * 2 * STEP * 4 : 2 * TILE * 4;
  static constexpr int OFF_BAR = OFF_AUX + AUX_BYTES;
  static constexpr int TOTAL = OFF_BAR + 128 + 1024 /* alignment slack */;
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// Write one thread's row of an fp32 accumulator [128 lanes x HD columns] as bf16(acc * scale) to `dst` (row pointer, hd
// valid columns), optionally applying the rotary embedding's backward (pairs (c, c + 64); head dim 128 only).  The
// two warps of a lane quarter (half = 0 / 1) take alternate 32-column chunks.  All lanes execute the TMEM loads.
template <int HD>
__device__ __forceinline__ void write_out(uint32_t taddr, bf16* dst, bool row_ok, bool have_acc, int half, int hd, float scale,
                                          const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int pos) {
  if constexpr (HD == 128) {
    if (rope_cos != nullptr) {
      uint32_t lo[32], hi[32];
      if (have_acc) {
        tmem_ld_32x32(taddr + half * 32, lo);
        tmem_ld_32x32(taddr + 64 + half * 32, hi);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) lo[j] = hi[j] = 0u;
      }
      if (row_ok) {
        const float* cp = rope_cos + pos * 64 + half * 32;
        const float* sp = rope_sin + pos * 64 + half * 32;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint32_t wl[4], wh[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = t * 8 + u * 2;
            const float2 c = *reinterpret_cast<const float2*>(cp + j);
            const float2 sn = *reinterpret_cast<const float2*>(sp + j);
            const float l0 = rbf(__uint_as_float(lo[j]) * scale), l1 = rbf(__uint_as_float(lo[j + 1]) * scale);
            const float h0 = rbf(__uint_as_float(hi[j]) * scale), h1 = rbf(__uint_as_float(hi[j + 1]) * scale);
            wl[u] = pack_bf16x2(rbf(l0 * c.x) + rbf(h0 * sn.x), rbf(l1 * c.y) + rbf(h1 * sn.y));
            wh[u] = pack_bf16x2(rbf(h0 * c.x) + rbf(-l0 * sn.x), rbf(h1 * c.y) + rbf(-l1 * sn.y));
          }
          *reinterpret_cast<uint4*>(dst + half * 32 + t * 8) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
          *reinterpret_cast<uint4*>(dst + 64 + half * 32 + t * 8) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
        }
      }
      return;
    }
  }
#pragma unroll 1
  for (int c = half; c < HD / 32; c += 2) {
    if (c * 32 >= hd) break;
    uint32_t v[32];
    if (have_acc) {
      tmem_ld_32x32(taddr + c * 32, v);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0u;
    }
    if (row_ok) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (c * 32 + t * 8 + 8 <= hd)
          *reinterpret_cast<uint4*>(dst + c * 32 + t * 8) =
              make_uint4(pack_bf16x2(__uint_as_float(v[t * 8]) * scale, __uint_as_float(v[t * 8 + 1]) * scale),
                         pack_bf16x2(__uint_as_float(v[t * 8 + 2]) * scale, __uint_as_float(v[t * 8 + 3]) * scale),
                         pack_bf16x2(__uint_as_float(v[t * 8 + 4]) * scale, __uint_as_float(v[t * 8 + 5]) * scale),
                         pack_bf16x2(__uint_as_float(v[t * 8 + 6]) * scale, __uint_as_float(v[t * 8 + 7]) * scale));
      }
    }
  }
}

template <int HD, int MODE>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                   const bf16* __restrict__ o, const bf16* __restrict__ dout, const float* __restrict__ lse,
                   float* __restrict__ delta, bf16* __restrict__ dqkv, const int* __restrict__ kv_len, int N, int H, int hd,
                   int causal, float scale, const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int rope_L) {
  using SM = BwdSmem<HD, MODE>;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  STAMP(MODE, threadIdx.x == 0, 0);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sX = base + SM::OFF_X, sY = base + SM::OFF_Y, sRing = base + SM::OFF_RING, sP = base + SM::OFF_P;
  const uint32_t bars = base + SM::OFF_BAR;
  const uint32_t bar_stat = bars;                                  // stationary tiles landed
  auto bar_full = [&](int i) { return bars + 8 + 8 * i; };         // ring slot i landed (TMA tx [+ stats])
  auto bar_empty = [&](int i) { return bars + 32 + 8 * i; };       // phase B that read ring slot i completed
  auto bar_sready = [&](int t) { return bars + 56 + 8 * t; };      // phase A into TMEM stage t completed
  auto bar_pdone = [&](int t) { return bars + 72 + 8 * t; };       // softmax warps wrote P/dS buffer t (256 arrivals)
  auto bar_bdone = [&](int t) { return bars + 88 + 8 * t; };       // phase B that read P/dS buffer t completed
  const uint32_t bar_out = bars + 104;                             // last phase B completed
  const uint32_t tmem_slot = bars + 112;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + SM::OFF_BAR + 112);
  float* aux = reinterpret_cast<float*>(gen + SM::OFF_AUX);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int ntiles = gridDim.y;
  // heavy tiles first: under the causal mask the LAST query tile sees the most keys, the FIRST key tile the most queries
  const int ti = (MODE == MODE_DQ && causal) ? (ntiles - 1 - blockIdx.y) : blockIdx.y;
  const int t0 = ti * TILE;
  const int D = H * hd;
  const int klen = kv_len ? min(kv_len[b], N) : N;
  int c_begin, c_end;   // streamed range: keys (MODE_DQ) or queries (MODE_DKV)
  if (MODE == MODE_DQ) {
    c_begin = 0;
    c_end = causal ? min(klen, t0 + TILE) : klen;
  } else {
    c_begin = causal ? t0 : 0;
    c_end = (t0 < klen) ? N : 0;   // a key tile entirely behind the padding boundary has zero gradient
  }
  const int nsteps = c_end > c_begin ? (c_end - c_begin + STEP - 1) / STEP : 0;
  auto step_cols = [&](int s) { return min(STEP, ((c_end - (c_begin + s * STEP) + 15) / 16) * 16); };
  const int ksteps = (hd + 15) / 16;   // phase A contraction steps (columns >= hd are zero filled by TMA)

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    mbar_init(bar_stat, 1);
    for (int i = 0; i < RING; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_sready(t), 1);
      mbar_init(bar_pdone(t), 256);
      mbar_init(bar_bdone(t), 1);
    }
    mbar_init(bar_out, 1);
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

  // operand -> (tensor map, head index along the map's second dimension)
  const CUtensorMap* mapX = &map_qkv;
  const CUtensorMap* mapY = (MODE == MODE_DQ) ? &map_do : &map_qkv;
  const CUtensorMap* mapx = &map_qkv;
  const CUtensorMap* mapy = (MODE == MODE_DQ) ? &map_qkv : &map_do;
  const int headX = (MODE == MODE_DQ) ? h : H + h;          // Q | K
  const int headY = (MODE == MODE_DQ) ? h : 2 * H + h;      // dO | V
  const int headx = (MODE == MODE_DQ) ? H + h : h;          // K_s | Q_s
  const int heady = (MODE == MODE_DQ) ? 2 * H + h : h;      // V_s | dO_s

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (nsteps > 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_stat, 2 * SM::STAT_BYTES);
        for (int p = 0; p < SM::PANELS; ++p)
          for (int r = 0; r < 2; ++r) {
            tma_load_4d(sX + p * (TILE * 128) + r * 8192, mapX, bar_stat, p * 64, headX, t0 + r * 64, b);
            tma_load_4d(sY + p * (TILE * 128) + r * 8192, mapY, bar_stat, p * 64, headY, t0 + r * 64, b);
          }
      }
      for (int s = 0; s < nsteps; ++s) {
        const int slot = s % RING;
        const int c0 = c_begin + s * STEP;
        if (s >= RING) mbar_wait(bar_empty(slot), ((s - RING) / RING) & 1);
        if (MODE == MODE_DKV) {
          float* st = aux + slot * (2 * STEP);
          for (int i = lane; i < STEP; i += 32) {
            const int q = c0 + i;
            float l2 = INFINITY, dl = 0.f;
            if (q < N) {
              const float l = lse[(static_cast<int64_t>(b) * H + h) * N + q];
              l2 = (l == -INFINITY) ? INFINITY : l * LOG2E_B;
              dl = delta[(static_cast<int64_t>(b) * H + h) * N + q];
            }
            st[i] = l2;
            st[STEP + i] = dl;
          }
          __syncwarp();
        }
        if (lane == 0) {
          const uint32_t xs = sRing + slot * SM::SLOT_BYTES, ys = xs + SM::STREAM_BYTES;
          mbar_arrive_expect_tx(bar_full(slot), SM::SLOT_BYTES);
          for (int p = 0; p < SM::PANELS; ++p) {
            tma_load_4d(xs + p * SM::STREAM_PANEL, mapx, bar_full(slot), p * 64, headx, c0, b);
            tma_load_4d(ys + p * SM::STREAM_PANEL, mapy, bar_full(slot), p * 64, heady, c0, b);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && nsteps > 0) {
      const uint32_t idB = idesc(TILE, HD, 1);
      auto issueA = [&](int s) {
        const int slot = s % RING, t = s & 1;
        const uint32_t idA = idesc(TILE, step_cols(s), 0);
        const uint32_t xs = sRing + slot * SM::SLOT_BYTES, ys = xs + SM::STREAM_BYTES;
        mbar_wait(bar_full(slot), (s / RING) & 1);
        tc_fence_after();
        for (int k = 0; k < ksteps; ++k)
          umma_bf16_ss<1>(tmem + t * 128, sdesc(sX + (k >> 2) * (TILE * 128) + (k & 3) * 32, 16, 1024),
                          sdesc(xs + (k >> 2) * SM::STREAM_PANEL + (k & 3) * 32, 16, 1024), idA, k > 0 ? 1u : 0u);
        for (int k = 0; k < ksteps; ++k)
          umma_bf16_ss<1>(tmem + t * 128 + 64, sdesc(sY + (k >> 2) * (TILE * 128) + (k & 3) * 32, 16, 1024),
                          sdesc(ys + (k >> 2) * SM::STREAM_PANEL + (k & 3) * 32, 16, 1024), idA, k > 0 ? 1u : 0u);
        umma_commit<1>(bar_sready(t));
      };
      mbar_wait(bar_stat, 0);
      STAMP(MODE, true, 40);
      issueA(0);
      STAMP(MODE, true, 41);
      if (nsteps > 1) issueA(1);
      for (int s = 0; s < nsteps; ++s) {
        const int slot = s % RING, t = s & 1;
        const int nk = step_cols(s) / 16;
        const uint32_t xs = sRing + slot * SM::SLOT_BYTES, ys = xs + SM::STREAM_BYTES;
        mbar_wait(bar_pdone(t), (s >> 1) & 1);
        tc_fence_after();
        if (MODE == MODE_DQ) {
          const uint32_t sdS = sP + t * SM::PBUF_BYTES;
          for (int kk = 0; kk < nk; ++kk)   // dQ += dS K_s
            umma_bf16_ss<1>(tmem_out0, sdesc(sdS + kk * 32, 16, 1024), sdesc(xs + kk * 16 * 128, SM::STREAM_PANEL, 1024), idB,
                            (s | kk) ? 1u : 0u);
        } else {
          const uint32_t sPt = sP + (2 * t) * SM::PBUF_BYTES, sdS = sPt + SM::PBUF_BYTES;
          for (int kk = 0; kk < nk; ++kk)   // dV += P^T dO_s
            umma_bf16_ss<1>(tmem_out0, sdesc(sPt + kk * 32, 16, 1024), sdesc(ys + kk * 16 * 128, SM::STREAM_PANEL, 1024), idB,
                            (s | kk) ? 1u : 0u);
          for (int kk = 0; kk < nk; ++kk)   // dK += dS^T Q_s
            umma_bf16_ss<1>(tmem_out1, sdesc(sdS + kk * 32, 16, 1024), sdesc(xs + kk * 16 * 128, SM::STREAM_PANEL, 1024), idB,
                            (s | kk) ? 1u : 0u);
        }
        umma_commit<1>(bar_empty(slot));
        umma_commit<1>(bar_bdone(t));
        if (s + 2 < nsteps) issueA(s + 2);
      }
      umma_commit<1>(bar_out);
    }
  } else {
    // ------------------------------------------------------------------ softmax / epilogue warps
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const int srow = t0 + r;                        // query (MODE_DQ) or key (MODE_DKV) index of this thread's row
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = scale * LOG2E_B;
    const int warp_row0 = t0 + quarter * 32;
    const bool warp_active = warp_row0 < ((MODE == MODE_DQ) ? N : klen);
    float lse2_r = INFINITY, delta_r = 0.f;
    if (MODE == MODE_DQ) {
      // delta = rowsum(dO o O): the two warps of a lane quarter each take half of the head's columns
      float part = 0.f;
      if (srow < N) {
        const bf16* op = o + (static_cast<int64_t>(b) * N + srow) * D + h * hd;
        const bf16* dp = dout + (static_cast<int64_t>(b) * N + srow) * D + h * hd;
        const int nch = hd / 8, mid = (nch + 1) / 2;
        for (int c = half ? mid : 0; c < (half ? nch : mid); ++c) {
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(op + c * 8));
          const uint4 d = __ldg(reinterpret_cast<const uint4*>(dp + c * 8));
          const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
          const float2 d0 = unpack_bf16x2(d.x), d1 = unpack_bf16x2(d.y), d2 = unpack_bf16x2(d.z), d3 = unpack_bf16x2(d.w);
          part += (a0.x * d0.x + a0.y * d0.y) + (a1.x * d1.x + a1.y * d1.y) + (a2.x * d2.x + a2.y * d2.y) + (a3.x * d3.x + a3.y * d3.y);
        }
        const float l = lse[(static_cast<int64_t>(b) * H + h) * N + srow];
        lse2_r = (l == -INFINITY) ? INFINITY : l * LOG2E_B;
      }
      aux[half * TILE + r] = part;
      softmax_bar();
      delta_r = aux[r] + aux[TILE + r];
      if (half == 0 && srow < N) delta[(static_cast<int64_t>(b) * H + h) * N + srow] = delta_r;
    }
    STAMP(MODE, threadIdx.x == 0, 3);

    for (int s = 0; s < nsteps; ++s) {
      const int slot = s % RING, t = s & 1;
      const int ncols = step_cols(s);
      const int col0 = c_begin + s * STEP + half * 32;   // first streamed index of this warp's 32-column chunk
      mbar_wait(bar_sready(t), (s >> 1) & 1);
      STAMP(MODE, threadIdx.x == 0, 4 + 2 * s);
      tc_fence_after();
      if (s >= 2) mbar_wait(bar_bdone(t), ((s - 2) >> 1) & 1);
      if (MODE == MODE_DKV) mbar_wait(bar_full(slot), (s / RING) & 1);   // acquire the producer's lse / delta stores
      if (half * 32 < ncols) {
        const uint32_t sdS = sP + ((MODE == MODE_DKV) ? (2 * t + 1) : t) * SM::PBUF_BYTES + r * 128;
        const uint32_t sPt = sP + (2 * t) * SM::PBUF_BYTES + r * 128;   // MODE_DKV only
        if (warp_active) {
          uint32_t sv[32], dv[32];
          tmem_ld_32x32(tmem + lane_addr + t * 128 + half * 32, sv);
          tmem_ld_32x32(tmem + lane_addr + t * 128 + 64 + half * 32, dv);
          tmem_ld_wait();
          bool need_mask;
          if (MODE == MODE_DQ) need_mask = (col0 + 32 > klen) || (causal && col0 + 31 > warp_row0);
          else need_mask = (warp_row0 + 31 >= klen) || (causal && warp_row0 + 31 > col0);
          const float* st = aux + slot * (2 * STEP) + half * 32;   // MODE_DKV: [lse2 | delta] of this chunk's queries
          uint32_t pk[16], dk[16];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float l2[4], dl[4];
            if (MODE == MODE_DKV) {
              const float4 a = *reinterpret_cast<const float4*>(st + j4 * 4);
              const float4 c = *reinterpret_cast<const float4*>(st + STEP + j4 * 4);
              l2[0] = a.x, l2[1] = a.y, l2[2] = a.z, l2[3] = a.w;
              dl[0] = c.x, dl[1] = c.y, dl[2] = c.z, dl[3] = c.w;
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u) l2[u] = lse2_r, dl[u] = delta_r;
            }
            float p[4], ds[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = j4 * 4 + u;
              float x = fmaf(__uint_as_float(sv[j]), sl2, -l2[u]);
              if (need_mask) {
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
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t chunk = (static_cast<uint32_t>(half * 4 + q4) ^ static_cast<uint32_t>(r & 7)) << 4;
            st_shared_v4(sdS + chunk, dk[q4 * 4], dk[q4 * 4 + 1], dk[q4 * 4 + 2], dk[q4 * 4 + 3]);
            if (MODE == MODE_DKV) st_shared_v4(sPt + chunk, pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
          }
        } else if (s < 2) {
          // rows outside the sequence: zero once per buffer (the tensor core reads all 128 rows of the A operand)
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t chunk = (static_cast<uint32_t>(half * 4 + q4) ^ static_cast<uint32_t>(r & 7)) << 4;
            st_shared_v4(sdS + chunk, 0u, 0u, 0u, 0u);
            if (MODE == MODE_DKV) st_shared_v4(sPt + chunk, 0u, 0u, 0u, 0u);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(bar_pdone(t));
      STAMP(MODE, threadIdx.x == 0, 5 + 2 * s);
    }

    // ---- epilogue ----
    const bool have_acc = nsteps > 0;
    if (have_acc) {
      mbar_wait(bar_out, 0);
      tc_fence_after();
    }
    STAMP(MODE, threadIdx.x == 0, 30);
    const bool row_ok = srow < N;
    bf16* drow = dqkv + (static_cast<int64_t>(b) * N + (row_ok ? srow : 0)) * (3 * static_cast<int64_t>(D)) + h * hd;
    const int pos = rope_L > 0 ? srow % rope_L : 0;
    if (MODE == MODE_DQ) {
      write_out<HD>(tmem_out0 + lane_addr, drow, row_ok, have_acc, half, hd, scale, rope_cos, rope_sin, pos);
    } else {
      write_out<HD>(tmem_out0 + lane_addr, drow + 2 * D, row_ok, have_acc, half, hd, 1.f, nullptr, nullptr, 0);
      write_out<HD>(tmem_out1 + lane_addr, drow + D, row_ok, have_acc, half, hd, scale, rope_cos, rope_sin, pos);
    }
  }

  STAMP(MODE, threadIdx.x == 0, 31);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, 512);
  }
  STAMP(MODE, threadIdx.x == 0, 32);
}

template <int HD>
int launch_bwd_tc(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                  const int* kv_len, int B, int N, int H, int hd, int causal, const float* rope_cos, const float* rope_sin,
                  int rope_L, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    VLA_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<HD, MODE_DQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        BwdSmem<HD, MODE_DQ>::TOTAL));
    VLA_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<HD, MODE_DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        BwdSmem<HD, MODE_DKV>::TOTAL));
    configured = true;
  }
  const int64_t D = static_cast<int64_t>(H) * hd;
  CUtensorMap map_qkv, map_do;
  if (int rc = make_tmap_bf16_4d(qkv, hd, 3 * H, N, B, hd, 3 * D, N * 3 * D, 64, 64, &map_qkv)) return rc;
  if (int rc = make_tmap_bf16_4d(dout, hd, H, N, B, hd, D, N * D, 64, 64, &map_do)) return rc;
  const float scale = 1.f / sqrtf(static_cast<float>(hd));
  dim3 grid(B * H, ceil_div(N, TILE));
  VLA_CHECK_CUDA(vla_launch(attn_bwd_tc_kernel<HD, MODE_DQ>, grid, dim3(BWD_THREADS), static_cast<size_t>(BwdSmem<HD, MODE_DQ>::TOTAL), s,
                            map_qkv, map_do, o, dout, lse, delta, dqkv, kv_len, N, H, hd, causal, scale, rope_cos, rope_sin, rope_L));
  VLA_CHECK_CUDA(vla_launch(attn_bwd_tc_kernel<HD, MODE_DKV>, grid, dim3(BWD_THREADS), static_cast<size_t>(BwdSmem<HD, MODE_DKV>::TOTAL), s,
                            map_qkv, map_do, o, dout, lse, delta, dqkv, kv_len, N, H, hd, causal, scale, rope_cos, rope_sin, rope_L));
  g_vla_launch_count += 2;
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
  VLA_REQUIRE(B * H <= 65535 * 32, "attention_bwd_tc: batch x heads too large");
  if (hd == 64) return launch_bwd_tc<64>(qkv, o, dout, lse, delta, dqkv, kv_len, B, N, H, hd, causal, rope_cos, rope_sin, rope_L, s);
  return launch_bwd_tc<128>(qkv, o, dout, lse, delta, dqkv, kv_len, B, N, H, hd, causal, rope_cos, rope_sin, rope_L, s);
}
