// tcgen05 attention forward for sequences of up to 320 keys (every sequence of the attack path: 261 / 256 / <= 320): the
// default forward.  Replaces F.scaled_dot_product_attention forward (timm Attention; HF LlamaSdpaAttention with the causal +
// key-padding mask) for head dims 64, 72 (run as 128 with zero-filled columns) and 128.
//
// Persistent, one CTA per SM over (batch, head, 128-query tile) items, heaviest first.  The whole score tile of an item
// (128 x up to 320 fp32) stays in TMEM, so S = Q K^T is computed ONCE and the exact two-pass softmax needs no rescaling:
//   S = Q K^T   with the item's K tiles contiguous in shared memory, so one tcgen05.mma covers up to 256 keys (issuing a
//               tcgen05 instruction costs ~50 ns whatever its size: 16 wide MMAs per item instead of 40 narrow ones)
//   pass 0: row maxima straight from TMEM (no MUFU work)        pass 1: P = exp2(S * scale * log2e - m) -> bf16 in smem
//   O += P_s V_s  per step (V_s as an MN-major B operand), double-buffered P;   epilogue: O / l -> smem -> TMA store; LSE
// K has its own region (free again right after the S MMAs, so the next item's K loads behind the current item's softmax);
// the V_s tiles stream through a 4-slot TMA ring.  The S tiles of the NEXT item are issued
// as soon as the softmax warps have read the last score of the current one, i.e. behind its P V tail and epilogue.
// Roles (352 threads): warps 0-7 softmax / epilogue (warp w: TMEM lane quarter w % 4, column half w / 4), warp 8 TMA
// producer, warp 9 issues S, warp 10 issues O += P V (tcgen05 instructions under elect.sync).
// Longer sequences use the streaming two-pass kernel of attention_fwd_tc2.cu.
#include <math.h>

#include <type_traits>

#include "kernels.h"
#include "tma_desc.h"

namespace {

constexpr int FWD_THREADS = 352;
constexpr int MAXS = 5;    // key steps per item: N <= 320
constexpr int RING = 4;    // V_s ring slots
constexpr int TILE = 128;
constexpr int STEP = 64;
constexpr float LOG2E_F2 = 1.4426950408889634f;

__device__ __forceinline__ uint64_t sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
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
__device__ unsigned long long g_fwd3_dbg[64];
__device__ __forceinline__ unsigned long long gtime3() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define STAMP3(i)                                                                             \
  do {                                                                                        \
    if (blockIdx.x == 0 && threadIdx.x == 0 && (i) < 64) g_fwd3_dbg[(i)] = gtime3();          \
  } while (0)
#else
#define STAMP3(i) do {} while (0)
#endif

template <int HD>
struct FwdSmem3 {
  static constexpr int PANELS = HD / 64;
  static constexpr int Q_BYTES = PANELS * TILE * 128;
  static constexpr int STREAM_PANEL = STEP * 128;
  static constexpr int STREAM_BYTES = PANELS * STREAM_PANEL;   // one V_s tile = one ring slot
  static constexpr int K_PANEL = MAXS * STEP * 128;            // one 64-column panel of the item's keys: 320 rows x 128 B
  static constexpr int K_BYTES = PANELS * K_PANEL;
  static constexpr int PBUF_BYTES = TILE * 128;                // P: 128 rows x 64 keys bf16
  static constexpr int PD_BYTES = 2 * PBUF_BYTES > Q_BYTES ? 2 * PBUF_BYTES : Q_BYTES;   // two P buffers = the O staging area
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_RING = OFF_K + K_BYTES;
  static constexpr int OFF_PD = OFF_RING + RING * STREAM_BYTES;
  static constexpr int OFF_AUX = OFF_PD + PD_BYTES;            // [2][128] floats: max / sum exchange between the two column halves
  static constexpr int OFF_BAR = OFF_AUX + 2 * TILE * 4;
  static constexpr int TOTAL = OFF_BAR + 320 + 1024;
  static_assert(TOTAL <= 232448, "shared memory budget");
};

struct Fwd3Args {
  float* lse;
  const int* kv_len;
  int B, N, H, hd, causal, ntiles, nitems;
  // Causal mask: the ragged query tile goes to the FRONT of the sequence (tile i covers rows [128 i - shift, +128), shift =
  // 0 or 64 <= 128 ntiles - N), where a tile sees the fewest keys -- at N = 288 the tiles are [-64, 64), [64, 192), [192, 320)
  // with 1 + 3 + 5 key steps instead of [0, 128), [128, 256), [256, 384) with 2 + 4 + 5.  shift is a multiple of the 64-row
  // TMA box and of the 32 rows of a warp, so a box / a warp is either entirely in front of the sequence or inside it.
  int shift;
  float scale;
};

template <int HD, int KS>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc3_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_o, const Fwd3Args a) {
  using SM = FwdSmem3<HD>;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + SM::OFF_Q, sK = base + SM::OFF_K, sRing = base + SM::OFF_RING, sPD = base + SM::OFF_PD;
  const uint32_t bars = base + SM::OFF_BAR;
  const uint32_t bar_q = bars;                                     // Q tile of the item landed
  const uint32_t bar_qfree = bars + 8;                             // last S MMA of the item has read Q
  const uint32_t bar_out = bars + 16;                              // last P V MMA of the item completed
  const uint32_t bar_outfree = bars + 24;                          // epilogue has read the O accumulator (8 warps)
  const uint32_t bar_sfree = bars + 32;                            // softmax warps have read the item's last score (8 warps)
  auto bar_full = [&](int i) { return bars + 40 + 8 * i; };                          // ring slot i landed
  auto bar_empty = [&](int i) { return bars + 40 + 8 * (RING + i); };                // the MMA reading ring slot i completed
  auto bar_sready = [&](int s) { return bars + 40 + 8 * (2 * RING + s); };           // [0]: all S tiles of the item completed; [1]: K landed; [2]: K region free
  auto bar_pdone = [&](int t) { return bars + 40 + 8 * (2 * RING + MAXS + t); };     // softmax warps wrote P buffer t (8 warps)
  auto bar_bdone = [&](int t) { return bars + 40 + 8 * (2 * RING + MAXS + 2 + t); }; // P V MMA that read P buffer t completed
  constexpr int TMEM_SLOT_OFF = 40 + 8 * (2 * RING + MAXS + 4);
  static_assert(TMEM_SLOT_OFF + 8 <= 320, "barrier block");
  const uint32_t tmem_slot = bars + TMEM_SLOT_OFF;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + SM::OFF_BAR + TMEM_SLOT_OFF);
  float* aux = reinterpret_cast<float*>(gen + SM::OFF_AUX);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, H = a.H, causal = a.causal;
  const int BH = a.B * H;
  const int G = gridDim.x;

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_o);
    mbar_init(bar_q, 1);
    mbar_init(bar_qfree, 1);
    mbar_init(bar_out, 1);
    mbar_init(bar_outfree, 8);
    mbar_init(bar_sfree, 8);
    for (int i = 0; i < RING; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    for (int s = 0; s < 3; ++s) mbar_init(bar_sready(s), 1);
    for (int t = 0; t < 2; ++t) {
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
  const uint32_t tmem_O = tmem + MAXS * STEP;
  pdl_wait();

  struct Item {
    int b, h, t0, klen, nsteps, c_end;
  };
  auto get_item = [&](int i, Item& it) -> bool {
    const int k = i * G + ((i & 1) ? (G - 1 - static_cast<int>(blockIdx.x)) : static_cast<int>(blockIdx.x));
    if (k >= a.nitems) return false;
    const int y = k / BH, bh = k - y * BH;
    it.b = bh / H;
    it.h = bh - it.b * H;
    const int ti = causal ? (a.ntiles - 1 - y) : y;   // heavy first: under the causal mask the last query tile sees the most keys
    it.t0 = ti * TILE - a.shift;
    it.klen = a.kv_len ? min(a.kv_len[it.b], N) : N;
    it.c_end = causal ? min(it.klen, it.t0 + TILE) : it.klen;
    it.nsteps = it.c_end > 0 ? (it.c_end + STEP - 1) / STEP : 0;
    return true;
  };
  auto step_cols = [&](const Item& it, int s) { return min(STEP, ((it.c_end - s * STEP + 15) / 16) * 16); };

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer: Q + all K tiles, then the V_s tiles
    const uint32_t bar_k = bar_sready(1), bar_kfree = bar_sready(2);
    int gl = 0, n_it = 0;
    Item it;
    for (int i = 0; get_item(i, it); ++i) {
      if (it.nsteps == 0) continue;
      if (n_it > 0) {
        mbar_wait(bar_qfree, (n_it - 1) & 1);    // the S MMAs of the previous item have read Q and K
        mbar_wait(bar_kfree, (n_it - 1) & 1);
      }
      if (lane == 0) {
        // a 64-row box in front of the sequence (shifted first tile) is not loaded: its rows belong to inactive warps, whose
        // scores are never read and whose P rows are zeroed
        mbar_arrive_expect_tx(bar_q, it.t0 < 0 ? SM::Q_BYTES / 2 : SM::Q_BYTES);
#pragma unroll
        for (int p = 0; p < SM::PANELS; ++p)
#pragma unroll
          for (int r = 0; r < 2; ++r)
            if (it.t0 + r * 64 >= 0) tma_load_4d(sQ + p * (TILE * 128) + r * 8192, &map_qkv, bar_q, p * 64, it.h, it.t0 + r * 64, it.b);
        mbar_arrive_expect_tx(bar_k, it.nsteps * SM::STREAM_BYTES);
        for (int s = 0; s < it.nsteps; ++s)
#pragma unroll
          for (int p = 0; p < SM::PANELS; ++p)
            tma_load_4d(sK + p * SM::K_PANEL + s * SM::STREAM_PANEL, &map_qkv, bar_k, p * 64, H + it.h, s * STEP, it.b);
      }
      for (int s = 0; s < it.nsteps; ++s, ++gl) {
        const int slot = gl % RING;
        if (gl >= RING) mbar_wait(bar_empty(slot), ((gl / RING) - 1) & 1);
        if (lane == 0) {
          const uint32_t dst = sRing + slot * SM::STREAM_BYTES;
          mbar_arrive_expect_tx(bar_full(slot), SM::STREAM_BYTES);
#pragma unroll
          for (int p = 0; p < SM::PANELS; ++p) tma_load_4d(dst + p * SM::STREAM_PANEL, &map_qkv, bar_full(slot), p * 64, 2 * H + it.h, s * STEP, it.b);
        }
      }
      ++n_it;
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ S = Q K^T issuer: up to 256 keys per tcgen05.mma
    const uint32_t bar_k = bar_sready(1), bar_kfree = bar_sready(2);
    const uint64_t dQ0 = sdesc(sQ, 16, 1024), dK0 = sdesc(sK, 16, 1024);
    int n_it = 0;
    Item it;
    for (int i = 0; get_item(i, it); ++i) {
      if (it.nsteps == 0) continue;
      mbar_wait(bar_q, n_it & 1);
      mbar_wait(bar_k, n_it & 1);
      if (n_it > 0) mbar_wait(bar_sfree, (n_it - 1) & 1);   // the previous item's scores have been read
      tc_fence_after();
      const int nkv = (it.nsteps - 1) * STEP + step_cols(it, it.nsteps - 1);   // multiple of 16
      if (elect_one_sync()) {
        for (int c0 = 0; c0 < nkv; c0 += 256) {
          const uint32_t idA = idesc(TILE, min(256, nkv - c0), 0);
#pragma unroll
          for (int k = 0; k < KS; ++k)
            umma_bf16_ss<1>(tmem + c0, dQ0 + (((k >> 2) * (TILE * 128) + (k & 3) * 32) >> 4),
                            dK0 + (((k >> 2) * SM::K_PANEL + c0 * 128 + (k & 3) * 32) >> 4), idA, k > 0 ? 1u : 0u);
        }
        umma_commit<1>(bar_sready(0));
        umma_commit<1>(bar_qfree);
        umma_commit<1>(bar_kfree);
      }
      __syncwarp();
      ++n_it;
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------ O += P_s V_s issuer
    const uint32_t idB = idesc(TILE, HD, 1);
    int gl = 0, gp = 0, n_it = 0;
    Item it;
    for (int i = 0; get_item(i, it); ++i) {
      if (it.nsteps == 0) continue;
      for (int s = 0; s < it.nsteps; ++s, ++gp) {
        const int buf = gp & 1;
        const int gv = gl + s, slot = gv % RING;
        const int nk = step_cols(it, s) / 16;
        const uint64_t dP = sdesc(sPD + buf * SM::PBUF_BYTES, 16, 1024);
        const uint64_t dvm = sdesc(sRing + slot * SM::STREAM_BYTES, SM::STREAM_PANEL, 1024);   // MN-major V_s
        mbar_wait(bar_pdone(buf), (gp >> 1) & 1);
        mbar_wait(bar_full(slot), (gv / RING) & 1);
        if (s == 0 && n_it > 0) mbar_wait(bar_outfree, (n_it - 1) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (kk < nk) umma_bf16_ss<1>(tmem_O, dP + kk * 2, dvm + kk * 128, idB, (s | kk) ? 1u : 0u);
          umma_commit<1>(bar_empty(slot));
          umma_commit<1>(bar_bdone(buf));
          if (s == it.nsteps - 1) umma_commit<1>(bar_out);
        }
        __syncwarp();
      }
      gl += it.nsteps;
      ++n_it;
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ softmax / epilogue warps
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = a.scale * LOG2E_F2;
    int gp = 0, n_it = 0;
    Item it;
    for (int i = 0; get_item(i, it); ++i) {
      const int q = it.t0 + r;
      const int warp_row0 = it.t0 + quarter * 32;
      const int klen = it.klen;
      const bool warp_active = warp_row0 >= 0 && warp_row0 < N;
      STAMP3(i * 8 + 0);
      // ---- pass 0: row maximum over all key steps ----
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (it.nsteps > 0) {
        mbar_wait(bar_sready(0), n_it & 1);
        tc_fence_after();
      }
      for (int s = 0; s < it.nsteps; ++s) {
        const int col0 = s * STEP + half * 32;
        const bool chunk_masked = (col0 >= klen) || (causal && col0 > warp_row0 + 31);   // 32 x 32 block with no visible key
        if (half * 32 < step_cols(it, s) && warp_active && !chunk_masked) {
          uint32_t sv[32];
          tmem_ld_32x32(tmem + lane_addr + s * STEP + half * 32, sv);
          tmem_ld_wait();
          const bool need_mask = (col0 + 32 > klen) || (causal && col0 + 31 > warp_row0);
          if (need_mask) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              const bool ok = (col < klen) && (!causal || col <= q);
              m4[j & 3] = fmaxf(m4[j & 3], ok ? __uint_as_float(sv[j]) : -INFINITY);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(sv[j]));
          }
        }
      }
      STAMP3(i * 8 + 1);
      float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      float l = 0.f, mref = 0.f;
      if (it.nsteps > 0) {
        aux[half * TILE + r] = mx;
        softmax_bar();
        mx = fmaxf(aux[r], aux[TILE + r]);
        mref = (mx == -INFINITY) ? 0.f : mx * sl2;
        softmax_bar();
      }
      STAMP3(i * 8 + 2);
      // ---- pass 1: P = exp2(S * scale * log2e - m), row sums, P -> smem (double buffered) ----
      for (int s = 0; s < it.nsteps; ++s, ++gp) {
        const int buf = gp & 1;
        const int ncols = step_cols(it, s);
        const int col0 = s * STEP + half * 32;
        const uint32_t sPr = sPD + buf * SM::PBUF_BYTES + r * 128;
        // EVERY warp waits for the P V MMA that read this buffer two steps ago before it touches the buffer or arrives on
        // its barrier again: a warp without work in this step (rows outside the sequence, no columns) would otherwise run
        // ahead and arrive twice within one phase of bar_pdone, completing it before the working warps have written P.
        if (gp >= 2) mbar_wait(bar_bdone(buf), ((gp - 2) >> 1) & 1);
        const bool chunk_masked = (col0 >= klen) || (causal && col0 > warp_row0 + 31);
        if (half * 32 < ncols && warp_active && !chunk_masked) {
          uint32_t sv[32];
          tmem_ld_32x32(tmem + lane_addr + s * STEP + half * 32, sv);
          tmem_ld_wait();
          const bool need_mask = (col0 + 32 > klen) || (causal && col0 + 31 > warp_row0);
          uint32_t pk[16];
          float l4[4] = {0.f, 0.f, 0.f, 0.f};
          auto compute = [&](auto masked_tag) {
            constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
            for (int j2 = 0; j2 < 16; ++j2) {
              float p[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int j = j2 * 2 + e;
                float x = fmaf(__uint_as_float(sv[j]), sl2, -mref);
                if (MASKED) {
                  const int col = col0 + j;
                  if (!((col < klen) && (!causal || col <= q))) x = -INFINITY;
                }
                p[e] = ex2(x);
                l4[j & 3] += p[e];
              }
              pk[j2] = pack_bf16x2(p[0], p[1]);
            }
          };
          if (need_mask) compute(std::true_type{}); else compute(std::false_type{});
          l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t chunk = (static_cast<uint32_t>(half * 4 + q4) ^ static_cast<uint32_t>(r & 7)) << 4;
            st_shared_v4(sPr + chunk, pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
          }
        } else if ((s < 2 && !warp_active) || (warp_active && chunk_masked && half * 32 < ncols)) {
          // a block with no visible key: P = 0 for this step; query rows outside the sequence: zero both P buffers once per
          // item (the tensor core reads all 128 rows)
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t chunk = (static_cast<uint32_t>(half * 4 + q4) ^ static_cast<uint32_t>(r & 7)) << 4;
            st_shared_v4(sPr + chunk, 0u, 0u, 0u, 0u);
          }
        }
        if (s == it.nsteps - 1) {   // every score of this item has been read: the next item's S tiles may overwrite TMEM
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_sfree);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pdone(buf));
      }
      STAMP3(i * 8 + 3);

      // ---- epilogue: O / l -> bf16 rows staged in the P area -> TMA store; LSE ----
      const bool have_acc = it.nsteps > 0;
      if (have_acc) {
        aux[half * TILE + r] = l;
        softmax_bar();
        l = aux[r] + aux[TILE + r];
        mbar_wait(bar_out, n_it & 1);
        tc_fence_after();
      }
      STAMP3(i * 8 + 4);
      const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
      for (int c = half; c < HD / 32; c += 2) {
        uint32_t v[32];
        if (have_acc) {
          tmem_ld_32x32(tmem_O + lane_addr + c * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
          const uint32_t chunk = (static_cast<uint32_t>((c & 1) * 4 + t4) ^ static_cast<uint32_t>(r & 7)) << 4;
          st_shared_v4(sPD + (c >> 1) * (TILE * 128) + r * 128 + chunk,
                       pack_bf16x2(__uint_as_float(v[t4 * 8]) * inv, __uint_as_float(v[t4 * 8 + 1]) * inv),
                       pack_bf16x2(__uint_as_float(v[t4 * 8 + 2]) * inv, __uint_as_float(v[t4 * 8 + 3]) * inv),
                       pack_bf16x2(__uint_as_float(v[t4 * 8 + 4]) * inv, __uint_as_float(v[t4 * 8 + 5]) * inv),
                       pack_bf16x2(__uint_as_float(v[t4 * 8 + 6]) * inv, __uint_as_float(v[t4 * 8 + 7]) * inv));
        }
      }
      if (have_acc) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_outfree);
      }
      if (half == 0 && q >= 0 && q < N)
        a.lse[(static_cast<int64_t>(it.b) * H + it.h) * N + q] = (l > 0.f) ? (mref + log2f(l)) / LOG2E_F2 : -INFINITY;
      fence_proxy_async();
      softmax_bar();
      if (threadIdx.x == 0) {
#pragma unroll
        for (int p = 0; p < SM::PANELS; ++p)
#pragma unroll
          for (int rb = 0; rb < 2; ++rb)
            if (it.t0 + rb * 64 >= 0 && it.t0 + rb * 64 < N && p * 64 < a.hd)
              tma_store_4d(&map_o, sPD + p * (TILE * 128) + rb * 8192, p * 64, it.h, it.t0 + rb * 64, it.b);
        tma_store_commit();
        tma_store_wait_read();
      }
      softmax_bar();   // staging area (= the P buffers) reusable
      STAMP3(i * 8 + 5);
      if (have_acc) ++n_it;
    }
    if (threadIdx.x == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, 512);
  }
}

int g_num_sms_fwd3 = 0;

template <int HD, int KS>
int launch_fwd_tc3(const bf16* qkv, bf16* o, float* lse, const int* kv_len, int B, int N, int H, int hd, int causal, cudaStream_t s) {
  using SM = FwdSmem3<HD>;
  static bool configured = false;
  if (!configured) {
    VLA_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc3_kernel<HD, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    configured = true;
  }
  if (g_num_sms_fwd3 == 0) {
    int dev = 0;
    VLA_CHECK_CUDA(cudaGetDevice(&dev));
    VLA_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms_fwd3, cudaDevAttrMultiProcessorCount, dev));
  }
  const int64_t D = static_cast<int64_t>(H) * hd;
  CUtensorMap map_qkv, map_o;
  if (int rc = make_tmap_bf16_4d(qkv, hd, 3 * H, N, B, hd, 3 * D, N * 3 * D, 64, 64, &map_qkv)) return rc;
  if (int rc = make_tmap_bf16_4d(o, hd, H, N, B, hd, D, N * D, 64, 64, &map_o)) return rc;
  Fwd3Args a;
  a.lse = lse;
  a.kv_len = kv_len;
  a.B = B, a.N = N, a.H = H, a.hd = hd, a.causal = causal;
  a.ntiles = ceil_div(N, TILE);
  a.nitems = a.ntiles * B * H;
  a.shift = attention_tile_shift(N, causal);
  a.scale = 1.f / sqrtf(static_cast<float>(hd));
  const int sms = (g_vla_sm_limit > 0 && g_vla_sm_limit < g_num_sms_fwd3) ? g_vla_sm_limit : g_num_sms_fwd3;
  const int grid = a.nitems < sms ? a.nitems : sms;
  VLA_CHECK_CUDA(vla_launch(attn_fwd_tc3_kernel<HD, KS>, dim3(grid), dim3(FWD_THREADS), static_cast<size_t>(SM::TOTAL), s, map_qkv, map_o, a));
  ++g_vla_launch_count;
  return 0;
}

}  // namespace

#ifdef VLA_ATTN_TIMING
extern "C" int vla_attn_fwd3_dbg_read(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_fwd3_dbg, sizeof(unsigned long long) * 64) == cudaSuccess ? 0 : 1;
}
#endif

bool attention_fwd_tc3_supported(int N, int hd) { return (hd == 64 || hd == 72 || hd == 128) && N >= 1 && N <= MAXS * STEP; }

int attention_fwd_tc3(const bf16* qkv, bf16* o, float* lse, const int* kv_len, int B, int N, int H, int hd, int causal,
                      cudaStream_t s) {
  VLA_REQUIRE(attention_fwd_tc3_supported(N, hd), "attention_fwd_tc3: unsupported shape N=%d hd=%d", N, hd);
  if (hd == 64) return launch_fwd_tc3<64, 4>(qkv, o, lse, kv_len, B, N, H, hd, causal, s);
  if (hd == 72) return launch_fwd_tc3<128, 5>(qkv, o, lse, kv_len, B, N, H, hd, causal, s);
  return launch_fwd_tc3<128, 8>(qkv, o, lse, kv_len, B, N, H, hd, causal, s);
}
