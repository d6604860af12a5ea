// LayerNorm (timm ViT blocks, eps 1e-6) and RMSNorm (HF LlamaRMSNorm) forward / input-gradient kernels.
// One CTA per row, the row lives in registers (16-byte vector loads), fp32 statistics, bf16 in/out with the
// eager path's rounding points. The backward kernels fuse the residual-stream gradient add
// (dx_out = bf16(dres + bf16(dnorm))), which is what autograd's AccumulateGrad does in bf16.
#include "kernels.h"

namespace {

constexpr int NORM_THREADS = 128;
constexpr int MAX_VEC = 4;   // up to 128*4*8 = 4096 columns per row

template <typename F>
__device__ __forceinline__ void for_each_vec(int d8, F f) {
#pragma unroll
  for (int it = 0; it < MAX_VEC; ++it) {
    const int v = threadIdx.x + it * NORM_THREADS;
    if (v < d8) f(it, v);
  }
}

__device__ __forceinline__ void load8(const bf16* p, float (&x)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = unpack_bf16x2(w[t]);
    x[2 * t] = f.x;
    x[2 * t + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(bf16* p, const float (&x)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                                            pack_bf16x2(x[6], x[7]));
}

__device__ __forceinline__ void layernorm_fwd_row(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                  const bf16* __restrict__ b, bf16* __restrict__ y,
                                                  float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                  int d, float eps, const int64_t row, float* red) {
  const int d8 = d / 8;
  float xv[MAX_VEC][8];
  float s = 0.f;
  for_each_vec(d8, [&](int it, int v) {
    load8(x + row * d + v * 8, xv[it]);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += xv[it][k];
  });
  const float mean = block_sum(s, red) / d;
  float q = 0.f;
  for_each_vec(d8, [&](int it, int v) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float c = xv[it][k] - mean;
      q += c * c;
    }
  });
  const float var = block_sum(q, red) / d;
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  for_each_vec(d8, [&](int it, int v) {
    float wv[8], bv[8], o[8];
    load8(w + v * 8, wv);
    load8(b + v * 8, bv);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (xv[it][k] - mean) * rstd * wv[k] + bv[k];
    store8(y + row * d + v * 8, o);
  });
}

__global__ void __launch_bounds__(NORM_THREADS) layernorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                      const bf16* __restrict__ b, bf16* __restrict__ y,
                                                                      float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                      int d, float eps) {
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  layernorm_fwd_row(x, w, b, y, mean_out, rstd_out, d, eps, blockIdx.x, red);
}

// Two independent LayerNorm problems in one launch (the DINOv2 and the SigLIP norm of the same depth): rows [0, M0) belong to
// problem 0, the rest to problem 1.
struct LnFwdArgs {
  const bf16 *x, *w, *b;
  bf16* y;
  float *mean, *rstd;
  int d;
};
struct LnBwdArgs {
  const bf16 *dy, *x, *w;
  const float *mean, *rstd;
  const bf16* dres;
  bf16* dx;
  int d;
  const bf16* gamma2;
  bf16* scaled;
};
__global__ void __launch_bounds__(NORM_THREADS) layernorm_fwd2_kernel(const LnFwdArgs a0, const LnFwdArgs a1, int M0, float eps) {
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  const bool second = static_cast<int>(blockIdx.x) >= M0;
  const LnFwdArgs& a = second ? a1 : a0;
  layernorm_fwd_row(a.x, a.w, a.b, a.y, a.mean, a.rstd, a.d, eps, second ? blockIdx.x - M0 : blockIdx.x, red);
}

__device__ __forceinline__ void layernorm_bwd_row(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                  const bf16* __restrict__ w, const float* __restrict__ mean_in,
                                                  const float* __restrict__ rstd_in, const bf16* __restrict__ dres,
                                                  bf16* __restrict__ dx, int d, const bf16* __restrict__ gamma2,
                                                  bf16* __restrict__ scaled, const int64_t row, float* red) {
  const int d8 = d / 8;
  const float mean = mean_in[row], rstd = rstd_in[row];
  float g[MAX_VEC][8], xh[MAX_VEC][8];
  float s1 = 0.f, s2 = 0.f;
  for_each_vec(d8, [&](int it, int v) {
    float dyv[8], wv[8], xv[8];
    load8(dy + row * d + v * 8, dyv);
    load8(w + v * 8, wv);
    load8(x + row * d + v * 8, xv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      g[it][k] = dyv[k] * wv[k];
      xh[it][k] = (xv[k] - mean) * rstd;
      s1 += g[it][k];
      s2 += g[it][k] * xh[it][k];
    }
  });
  // the residual-stream gradient is fetched before the block reductions so that its latency hides behind them
  uint4 rv[MAX_VEC];
  if (dres) for_each_vec(d8, [&](int it, int v) { rv[it] = *reinterpret_cast<const uint4*>(dres + row * d + v * 8); });
  const float m1 = block_sum(s1, red) / d;
  const float m2 = block_sum(s2, red) / d;
  for_each_vec(d8, [&](int it, int v) {
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = rstd * (g[it][k] - m1 - xh[it][k] * m2);
    if (dres) {
      const uint32_t w4[4] = {rv[it].x, rv[it].y, rv[it].z, rv[it].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 r = unpack_bf16x2(w4[t]);
        o[2 * t] = r.x + rbf(o[2 * t]);
        o[2 * t + 1] = r.y + rbf(o[2 * t + 1]);
      }
    }
    store8(dx + row * d + v * 8, o);
    if (scaled) {   // the LayerScale backward of the branch this gradient enters next: bf16(bf16(dx) * gamma), fused here
      float gv[8], sc[8];
      load8(gamma2 + v * 8, gv);
#pragma unroll
      for (int k = 0; k < 8; ++k) sc[k] = rbf(o[k]) * gv[k];
      store8(scaled + row * d + v * 8, sc);
    }
  });
}

__global__ void __launch_bounds__(NORM_THREADS) layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                                      const bf16* __restrict__ w, const float* __restrict__ mean_in,
                                                                      const float* __restrict__ rstd_in, const bf16* __restrict__ dres,
                                                                      bf16* __restrict__ dx, int d, const bf16* __restrict__ gamma2,
                                                                      bf16* __restrict__ scaled) {
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  layernorm_bwd_row(dy, x, w, mean_in, rstd_in, dres, dx, d, gamma2, scaled, blockIdx.x, red);
}
__global__ void __launch_bounds__(NORM_THREADS) layernorm_bwd2_kernel(const LnBwdArgs a0, const LnBwdArgs a1, int M0) {
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  const bool second = static_cast<int>(blockIdx.x) >= M0;
  const LnBwdArgs& a = second ? a1 : a0;
  layernorm_bwd_row(a.dy, a.x, a.w, a.mean, a.rstd, a.dres, a.dx, a.d, a.gamma2, a.scaled, second ? blockIdx.x - M0 : blockIdx.x, red);
}

__global__ void __launch_bounds__(NORM_THREADS) rmsnorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                    bf16* __restrict__ y, float* __restrict__ rstd_out, int d,
                                                                    float eps) {
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  const int64_t row = blockIdx.x;
  const int d8 = d / 8;
  float xv[MAX_VEC][8];
  float q = 0.f;
  for_each_vec(d8, [&](int it, int v) {
    load8(x + row * d + v * 8, xv[it]);
#pragma unroll
    for (int k = 0; k < 8; ++k) q += xv[it][k] * xv[it][k];
  });
  const float rstd = rsqrtf(block_sum(q, red) / d + eps);
  if (threadIdx.x == 0) rstd_out[row] = rstd;
  for_each_vec(d8, [&](int it, int v) {
    float wv[8], o[8];
    load8(w + v * 8, wv);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = wv[k] * rbf(xv[it][k] * rstd);   // weight * hidden.to(bf16)
    store8(y + row * d + v * 8, o);
  });
}

__global__ void __launch_bounds__(NORM_THREADS) rmsnorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                                    const bf16* __restrict__ w, const float* __restrict__ rstd_in,
                                                                    const bf16* __restrict__ dres, bf16* __restrict__ dx, int d) {
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  const int64_t row = blockIdx.x;
  const int d8 = d / 8;
  const float rstd = rstd_in[row];
  float g[MAX_VEC][8], xv[MAX_VEC][8];
  float s = 0.f;
  for_each_vec(d8, [&](int it, int v) {
    float dyv[8], wv[8];
    load8(dy + row * d + v * 8, dyv);
    load8(w + v * 8, wv);
    load8(x + row * d + v * 8, xv[it]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      g[it][k] = rbf(dyv[k] * wv[k]);   // grad of (weight * h_bf16) wrt h_bf16, a bf16 tensor in autograd
      s += g[it][k] * xv[it][k];
    }
  });
  uint4 rv[MAX_VEC];   // residual-stream gradient, fetched before the block reduction (latency hidden behind it)
  if (dres) for_each_vec(d8, [&](int it, int v) { rv[it] = *reinterpret_cast<const uint4*>(dres + row * d + v * 8); });
  const float mgx = block_sum(s, red) / d;
  const float r3 = rstd * rstd * rstd;
  for_each_vec(d8, [&](int it, int v) {
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = rstd * g[it][k] - xv[it][k] * r3 * mgx;
    if (dres) {
      const uint32_t w4[4] = {rv[it].x, rv[it].y, rv[it].z, rv[it].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 r = unpack_bf16x2(w4[t]);
        o[2 * t] = r.x + rbf(o[2 * t]);
        o[2 * t + 1] = r.y + rbf(o[2 * t + 1]);
      }
    }
    store8(dx + row * d + v * 8, o);
  });
}

int check_dims(int64_t M, int d, const char* what) {
  if (vla_ablated("norm")) return -1;   // diagnostics: skip the launch (callers return 0)
  VLA_REQUIRE(M > 0 && d > 0 && d % 8 == 0 && d <= NORM_THREADS * MAX_VEC * 8, "%s: unsupported row width %d", what, d);
  return 0;
}

}  // namespace

int layernorm_fwd(const bf16* x, const bf16* w, const bf16* b, bf16* y, float* mean, float* rstd, int64_t M, int d, float eps,
                  cudaStream_t s) {
  if (int rc = check_dims(M, d, "layernorm_fwd")) return rc < 0 ? 0 : rc;
  VLA_CHECK_CUDA(vla_launch(layernorm_fwd_kernel, dim3(static_cast<unsigned>(M)), dim3(NORM_THREADS), 0, s, x, w, b, y, mean, rstd, d, eps));
  if (vla_doubled("norm")) {
    VLA_CHECK_CUDA(vla_launch(layernorm_fwd_kernel, dim3(static_cast<unsigned>(M)), dim3(NORM_THREADS), 0, s, x, w, b, y, mean, rstd, d, eps));
  }
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
int layernorm_bwd(const bf16* dy, const bf16* x, const bf16* w, const float* mean, const float* rstd, const bf16* dres,
                  bf16* dx, int64_t M, int d, cudaStream_t s, const bf16* gamma2, bf16* scaled) {
  if (int rc = check_dims(M, d, "layernorm_bwd")) return rc < 0 ? 0 : rc;
  VLA_REQUIRE((gamma2 == nullptr) == (scaled == nullptr), "layernorm_bwd: gamma2 and scaled go together");
  VLA_CHECK_CUDA(vla_launch(layernorm_bwd_kernel, dim3(static_cast<unsigned>(M)), dim3(NORM_THREADS), 0, s, dy, x, w, mean, rstd, dres, dx, d,
                            gamma2, scaled));
  if (vla_doubled("norm")) {
    VLA_CHECK_CUDA(vla_launch(layernorm_bwd_kernel, dim3(static_cast<unsigned>(M)), dim3(NORM_THREADS), 0, s, dy, x, w, mean, rstd, dres, dx, d,
                              gamma2, scaled));
  }
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
int layernorm_fwd2(const LnFwdProblem& p0, const LnFwdProblem& p1, float eps, cudaStream_t s) {
  if (int rc = check_dims(p0.M, p0.d, "layernorm_fwd2")) return rc < 0 ? 0 : rc;
  if (int rc = check_dims(p1.M, p1.d, "layernorm_fwd2")) return rc;
  const LnFwdArgs a0{p0.x, p0.w, p0.b, p0.y, p0.mean, p0.rstd, p0.d}, a1{p1.x, p1.w, p1.b, p1.y, p1.mean, p1.rstd, p1.d};
  VLA_CHECK_CUDA(vla_launch(layernorm_fwd2_kernel, dim3(static_cast<unsigned>(p0.M + p1.M)), dim3(NORM_THREADS), 0, s, a0, a1,
                            static_cast<int>(p0.M), eps));
  if (vla_doubled("norm")) {
    VLA_CHECK_CUDA(vla_launch(layernorm_fwd2_kernel, dim3(static_cast<unsigned>(p0.M + p1.M)), dim3(NORM_THREADS), 0, s, a0, a1,
                              static_cast<int>(p0.M), eps));
  }
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
int layernorm_bwd2(const LnBwdProblem& p0, const LnBwdProblem& p1, cudaStream_t s) {
  if (int rc = check_dims(p0.M, p0.d, "layernorm_bwd2")) return rc < 0 ? 0 : rc;
  if (int rc = check_dims(p1.M, p1.d, "layernorm_bwd2")) return rc;
  VLA_REQUIRE((p0.gamma2 == nullptr) == (p0.scaled == nullptr) && (p1.gamma2 == nullptr) == (p1.scaled == nullptr),
              "layernorm_bwd2: gamma2 and scaled go together");
  const LnBwdArgs a0{p0.dy, p0.x, p0.w, p0.mean, p0.rstd, p0.dres, p0.dx, p0.d, p0.gamma2, p0.scaled},
      a1{p1.dy, p1.x, p1.w, p1.mean, p1.rstd, p1.dres, p1.dx, p1.d, p1.gamma2, p1.scaled};
  VLA_CHECK_CUDA(vla_launch(layernorm_bwd2_kernel, dim3(static_cast<unsigned>(p0.M + p1.M)), dim3(NORM_THREADS), 0, s, a0, a1,
                            static_cast<int>(p0.M)));
  if (vla_doubled("norm")) {
    VLA_CHECK_CUDA(vla_launch(layernorm_bwd2_kernel, dim3(static_cast<unsigned>(p0.M + p1.M)), dim3(NORM_THREADS), 0, s, a0, a1,
                              static_cast<int>(p0.M)));
  }
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
int rmsnorm_fwd(const bf16* x, const bf16* w, bf16* y, float* rstd, int64_t M, int d, float eps, cudaStream_t s) {
  if (int rc = check_dims(M, d, "rmsnorm_fwd")) return rc < 0 ? 0 : rc;
  VLA_CHECK_CUDA(vla_launch(rmsnorm_fwd_kernel, dim3(static_cast<unsigned>(M)), dim3(NORM_THREADS), 0, s, x, w, y, rstd, d, eps));
  if (vla_doubled("norm")) {
    VLA_CHECK_CUDA(vla_launch(rmsnorm_fwd_kernel, dim3(static_cast<unsigned>(M)), dim3(NORM_THREADS), 0, s, x, w, y, rstd, d, eps));
  }
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
int rmsnorm_bwd(const bf16* dy, const bf16* x, const bf16* w, const float* rstd, const bf16* dres, bf16* dx, int64_t M, int d,
                cudaStream_t s) {
  if (int rc = check_dims(M, d, "rmsnorm_bwd")) return rc < 0 ? 0 : rc;
  VLA_CHECK_CUDA(vla_launch(rmsnorm_bwd_kernel, dim3(static_cast<unsigned>(M)), dim3(NORM_THREADS), 0, s, dy, x, w, rstd, dres, dx, d));
  if (vla_doubled("norm")) {
    VLA_CHECK_CUDA(vla_launch(rmsnorm_bwd_kernel, dim3(static_cast<unsigned>(M)), dim3(NORM_THREADS), 0, s, dy, x, w, rstd, dres, dx, d));
  }
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
