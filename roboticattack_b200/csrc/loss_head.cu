// F3: action-logit loss heads, forward + gradient wrt the logits, on the supervised rows only.
//
// Replaces (a) HF's shifted cross-entropy over all L rows (the reference materialises fp32 logits [B, L, 32064]),
// (b) OpenVLAAttacker.weighted_loss -- UADA.py:381-406 / UADA_ddp.py:99-124 (softmax over the 256 action classes,
// expectation, MSE against a {0,1} target; `+ 1/CE` in UADA.py:147) and UPA.py:367-387 (cosine + inverse-L2 on the
// first three DoF), (c) cal_UAD (UADA.py:408-418) and the argmax decode, which cost the reference a GPU->CPU->numpy
// round trip per inner step, and (d) autograd's backward of all of the above.  Three tiny launches:
//   rows   : one CTA per supervised row: full-vocab log-sum-exp, CE term, 256-way softmax expectation + argmax
//   reduce : one CTA: batch scalars (loss, CE, UAD ...) and the per-row gradient coefficients
//   grad   : one CTA per row: dlogits (bf16, what `.float()`'s backward hands to lm_head) from the coefficients
#include <math.h>

#include "kernels.h"

namespace {

constexpr int LH_THREADS = 256;
constexpr int ACT_LO = 31744, ACT_N = 256, ACT_ZERO = 31872, VOCAB_TOK = 32000;
enum { RS_CE = 0, RS_E1 = 1, RS_ARGMAX = 2, RS_LSE256 = 3, RS_LSEFULL = 4, RS_CCE = 5, RS_CE_COEF = 6, RS_SCALE = 7, RS_N = 8 };

__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : -INFINITY;
  return warp_max(t);
}

// ActionTokenizer.decode_token_ids_to_actions (action_tokenizer.py:49-68): bin centres of linspace(-1, 1, 256)
__device__ __forceinline__ float decode_action(int id) {
  int d = VOCAB_TOK - id - 1;
  d = max(0, min(d, 254));
  return -1.f + (2.f * d + 1.f) / 255.f;
}

__global__ void __launch_bounds__(LH_THREADS) loss_rows_kernel(const float* __restrict__ logits, const int* __restrict__ meta, int V,
                                                               float* __restrict__ row_stats) {
  __shared__ float red[32];
  __shared__ int red_i[32];
  const int r = blockIdx.x;
  const float* z = logits + static_cast<int64_t>(r) * V;
  const int target = meta[r * 3 + 0];
  // full vocabulary log-sum-exp
  float mx = -INFINITY;
#pragma unroll 8
  for (int n = threadIdx.x; n < V; n += LH_THREADS) mx = fmaxf(mx, z[n]);   // unrolled: the loads of 8 iterations in flight
  mx = block_max(mx, red);
  float se = 0.f;
#pragma unroll 8
  for (int n = threadIdx.x; n < V; n += LH_THREADS) se += expf(z[n] - mx);
  se = block_sum(se, red);
  const float lse_full = mx + logf(se);
  // 256 action classes: thread k owns class k
  const int k = threadIdx.x;
  const float zk = z[ACT_LO + k];
  const float m256 = block_max(zk, red);
  const float ek = expf(zk - m256);
  const float s256 = block_sum(ek, red);
  const float e1 = block_sum(ek * static_cast<float>(k + 1), red) / s256;
  // argmax, lowest index on ties
  int cand = (zk == m256) ? k : ACT_N;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red_i[threadIdx.x >> 5] = cand;
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = ACT_N;
    for (int w = 0; w < LH_THREADS / 32; ++w) best = min(best, red_i[w]);
    float* rs = row_stats + r * RS_N;
    rs[RS_CE] = (target >= 0 && target < V) ? lse_full - z[target] : 0.f;
    rs[RS_E1] = e1;
    rs[RS_ARGMAX] = static_cast<float>(best);
    rs[RS_LSE256] = m256 + logf(s256);
    rs[RS_LSEFULL] = lse_full;
  }
}

__global__ void __launch_bounds__(LH_THREADS) loss_reduce_kernel(const int* __restrict__ meta, int R, int B, LossParams lp,
                                                                 float* __restrict__ row_stats, float* __restrict__ scalars,
                                                                 int* __restrict__ pred_ids) {
  __shared__ float red[32];
  // ---- pass 1: CE mean, action-row count, UADA MSE, UAD ----
  float ce = 0.f, nact = 0.f, mse = 0.f, uad = 0.f;
  for (int r = threadIdx.x; r < R; r += LH_THREADS) {
    const float* rs = row_stats + r * RS_N;
    const int target = meta[r * 3];
    ce += rs[RS_CE];
    const bool is_act = target > 2;   // temp_label > 2 (UADA.py:383)
    const int pred = ACT_LO + static_cast<int>(rs[RS_ARGMAX]);
    pred_ids[r] = is_act ? pred : -1;
    if (is_act) {
      nact += 1.f;
      const float es = rs[RS_E1] / 256.f;
      const float tt = (target <= ACT_ZERO) ? 1.f : 0.f;   // int-cast quirk reproduced (UADA.py:390-394)
      const float d = lp.mse_weight * es - lp.mse_weight * tt;
      mse += d * d;
      const float gt = decode_action(target), pr = decode_action(pred);
      const float maxd = gt > 0.f ? fabsf(gt + 1.f) : fabsf(gt - 1.f);
      uad += fabsf(pr - gt) / maxd;
    }
  }
  ce = block_sum(ce, red);
  nact = block_sum(nact, red);
  mse = block_sum(mse, red);
  uad = block_sum(uad, red);
  const float ce_mean = R > 0 ? ce / R : 0.f;
  const float mse_mean = nact > 0.f ? mse / nact : 0.f;

  // ---- UPA: per-sample cosine / distance over the first three supervised tokens ----
  float upa_angle = 0.f, upa_dist = 0.f, mean_norm = 0.f;
  if (lp.kind == LOSS_UPA) {
    // rows are sorted by (sample, position): thread b handles sample b by scanning for its first row
    float cs = 0.f, nr = 0.f;
    for (int b = threadIdx.x; b < B; b += LH_THREADS) {
      float xh[3] = {0, 0, 0}, x[3] = {0, 0, 0};
      for (int r = 0; r < R; ++r)
        if (meta[r * 3 + 1] == b && meta[r * 3 + 2] < 3) {
          const int j = meta[r * 3 + 2];
          xh[j] = (row_stats[r * RS_N + RS_E1] - 1.f) / 255.f;
          x[j] = (static_cast<float>(meta[r * 3] - 31743) - 1.f) / 255.f;
        }
      const float dot = xh[0] * x[0] + xh[1] * x[1] + xh[2] * x[2];
      const float nh = sqrtf(xh[0] * xh[0] + xh[1] * xh[1] + xh[2] * xh[2]);
      const float nx = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
      cs += dot / (fmaxf(nh, 1e-8f) * fmaxf(nx, 1e-8f)) + 1.f;
      const float d0 = xh[0] - x[0], d1 = xh[1] - x[1], d2 = xh[2] - x[2];
      nr += sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    }
    cs = block_sum(cs, red);
    nr = block_sum(nr, red);
    upa_angle = cs / B;
    mean_norm = nr / B;
    upa_dist = 1.f / (mean_norm + 1e-3f);
  }

  // ---- total loss + coefficients ----
  float loss = 0.f, c_ce = 0.f;
  switch (lp.kind) {
    case LOSS_UADA:
      loss = mse_mean + 1.f / ce_mean;               // UADA.py:147
      c_ce = -1.f / (ce_mean * ce_mean) / R;
      break;
    case LOSS_UADA_DDP:
      loss = mse_mean;                               // UADA_ddp.py:203-206
      break;
    case LOSS_UPA:
      loss = lp.alpha * upa_angle + lp.belta * upa_dist;
      break;
    case LOSS_CE:
      loss = lp.ce_scale * ce_mean;                  // TMA.py:148
      c_ce = lp.ce_scale / R;
      break;
    case LOSS_NEG_CE:
      loss = -ce_mean;                               // UPA.py:150
      c_ce = -1.f / R;
      break;
  }
  if (threadIdx.x == 0) {
    scalars[LS_LOSS] = loss;
    scalars[LS_CE] = ce_mean;
    scalars[LS_AUX0] = (lp.kind == LOSS_UPA) ? upa_angle : mse_mean;
    scalars[LS_AUX1] = upa_dist;
    scalars[LS_UAD] = nact > 0.f ? uad / nact : 0.f;
    scalars[LS_NTOK] = static_cast<float>(R);
    scalars[LS_NACT] = nact;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += LH_THREADS) {
    float* rs = row_stats + r * RS_N;
    const int target = meta[r * 3];
    float c_e = 0.f, sc = 1.f;
    if ((lp.kind == LOSS_UADA || lp.kind == LOSS_UADA_DDP) && target > 2 && nact > 0.f) {
      sc = 1.f / 256.f;
      const float es = rs[RS_E1] / 256.f;
      const float tt = (target <= ACT_ZERO) ? 1.f : 0.f;
      c_e = 2.f * lp.mse_weight * lp.mse_weight * (es - tt) / nact;
    } else if (lp.kind == LOSS_UPA && meta[r * 3 + 2] < 3) {
      const int b = meta[r * 3 + 1], j = meta[r * 3 + 2];
      float xh[3] = {0, 0, 0}, x[3] = {0, 0, 0};
      for (int q = 0; q < R; ++q)
        if (meta[q * 3 + 1] == b && meta[q * 3 + 2] < 3) {
          const int jj = meta[q * 3 + 2];
          xh[jj] = (row_stats[q * RS_N + RS_E1] - 1.f) / 255.f;
          x[jj] = (static_cast<float>(meta[q * 3] - 31743) - 1.f) / 255.f;
        }
      const float dot = xh[0] * x[0] + xh[1] * x[1] + xh[2] * x[2];
      const float nh = fmaxf(sqrtf(xh[0] * xh[0] + xh[1] * xh[1] + xh[2] * xh[2]), 1e-8f);
      const float nx = fmaxf(sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]), 1e-8f);
      const float cosv = dot / (nh * nx);
      const float dcos = x[j] / (nh * nx) - cosv * xh[j] / (nh * nh);
      const float d0 = xh[0] - x[0], d1 = xh[1] - x[1], d2 = xh[2] - x[2];
      const float nd = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
      const float ddist = nd > 0.f ? (xh[j] - x[j]) / nd : 0.f;
      const float mn = mean_norm + 1e-3f;
      const float dl_dxh = lp.alpha * dcos / B - lp.belta / (mn * mn) * ddist / B;
      c_e = dl_dxh / 255.f;
      sc = 1.f;
    }
    rs[RS_CCE] = c_ce;
    rs[RS_CE_COEF] = c_e;
    rs[RS_SCALE] = sc;
  }
}

constexpr int LG_SPLIT = 16;
__global__ void __launch_bounds__(LH_THREADS) loss_grad_kernel(const float* __restrict__ logits, const int* __restrict__ meta, int V,
                                                               const float* __restrict__ row_stats, bf16* __restrict__ dlogits) {
  // grid (R, LG_SPLIT): a row's vocabulary is split over LG_SPLIT CTAs -- with one CTA per row every thread walked 125
  // dependent global loads (66 us for 32 rows, all of it latency on the step's critical path between forward and backward)
  const int r = blockIdx.x;
  const float* z = logits + static_cast<int64_t>(r) * V;
  bf16* dz = dlogits + static_cast<int64_t>(r) * V;
  const float* rs = row_stats + r * RS_N;
  const int target = meta[r * 3];
  const float c_ce = rs[RS_CCE], c_e = rs[RS_CE_COEF], sc = rs[RS_SCALE];
  const float lse_full = rs[RS_LSEFULL], lse256 = rs[RS_LSE256], e_sc = rs[RS_E1] * sc;
  const int per = (V + gridDim.y - 1) / gridDim.y;
  const int n_end = min(V, static_cast<int>(blockIdx.y + 1) * per);
#pragma unroll 4
  for (int n = blockIdx.y * per + threadIdx.x; n < n_end; n += LH_THREADS) {
    float g = 0.f;
    if (c_ce != 0.f) g = c_ce * (expf(z[n] - lse_full) - (n == target ? 1.f : 0.f));
    if (c_e != 0.f && n >= ACT_LO && n < ACT_LO + ACT_N) {
      const float p = expf(z[n] - lse256);
      g += c_e * p * (static_cast<float>(n - ACT_LO + 1) * sc - e_sc);
    }
    dz[n] = f2b(g);
  }
}

}  // namespace

size_t loss_head_row_stats_floats(int R) { return static_cast<size_t>(R) * RS_N; }

int loss_head_fwd_bwd(const float* logits, const int* meta, int R, int V, int B, const LossParams& lp, float* row_stats,
                      bf16* dlogits, float* scalars, int* pred_ids, cudaStream_t s) {
  VLA_REQUIRE(R > 0, "loss head: no supervised rows (every label is -100)");
  VLA_REQUIRE(V >= ACT_LO + ACT_N, "loss head: vocabulary %d does not contain the 256 action ids", V);
  VLA_REQUIRE(lp.kind >= LOSS_UADA && lp.kind <= LOSS_NEG_CE, "loss head: bad loss kind %d", lp.kind);
  loss_rows_kernel<<<R, LH_THREADS, 0, s>>>(logits, meta, V, row_stats);
  VLA_LAUNCH_CHECK();
  loss_reduce_kernel<<<1, LH_THREADS, 0, s>>>(meta, R, B, lp, row_stats, scalars, pred_ids);
  VLA_LAUNCH_CHECK();
  loss_grad_kernel<<<dim3(R, LG_SPLIT), LH_THREADS, 0, s>>>(logits, meta, V, row_stats, dlogits);
  VLA_LAUNCH_CHECK();
  g_vla_launch_count += 3;
  return 0;
}
