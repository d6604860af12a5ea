// Layout and elementwise kernels between the GEMMs (all HBM-bound, vectorised where the layout allows).
// Each mirrors the bf16 rounding points of the eager PyTorch op it replaces (timm Block / HF LlamaMLP /
// apply_rotary_pos_emb / the multimodal splice of modeling_prismatic.py:380-401).
#include "kernels.h"

namespace {

constexpr int EW_THREADS = 256;
inline unsigned blocks_for(int64_t n, int threads = EW_THREADS) { return static_cast<unsigned>(ceil_div64(n, threads)); }

__global__ void im2col_kernel(const bf16* __restrict__ px, bf16* __restrict__ a0, bf16* __restrict__ a1, int B, int H,
                              int W, int P, int kpad) {
  const int gw = W / P, gh = H / P, np = gw * gh, kk = 3 * P * P;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * np * kpad) return;
  const int k = static_cast<int>(idx % kpad);
  const int64_t row = idx / kpad;
  bf16 v0 = f2b(0.f), v1 = v0;
  if (k < kk) {
    const int p = static_cast<int>(row % np), b = static_cast<int>(row / np);
    const int c = k / (P * P), ky = (k / P) % P, kx = k % P;
    const int y = (p / gw) * P + ky, x = (p % gw) * P + kx;
    const int64_t plane = static_cast<int64_t>(H) * W;
    const bf16* base = px + static_cast<int64_t>(b) * 6 * plane + static_cast<int64_t>(y) * W + x;
    v0 = base[c * plane];
    v1 = base[(3 + c) * plane];
  }
  a0[idx] = v0;
  a1[idx] = v1;
}

__global__ void col2im_kernel(const bf16* __restrict__ da0, const bf16* __restrict__ da1, bf16* __restrict__ dpx, int B,
                              int H, int W, int P, int kpad) {
  const int gw = W / P, np = gw * (H / P);
  const int64_t plane = static_cast<int64_t>(H) * W;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * 6 * plane) return;
  const int x = static_cast<int>(idx % W), y = static_cast<int>((idx / W) % H);
  const int ch = static_cast<int>((idx / plane) % 6), b = static_cast<int>(idx / (6 * plane));
  const int c = ch % 3;
  const int64_t row = static_cast<int64_t>(b) * np + (y / P) * gw + (x / P);
  const int k = c * P * P + (y % P) * P + (x % P);
  dpx[idx] = (ch < 3 ? da0 : da1)[row * kpad + k];
}

__global__ void prefix_tokens_kernel(const bf16* __restrict__ cls, const bf16* __restrict__ reg, bf16* __restrict__ x,
                                     int B, int ntok, int npre, int d) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * npre * d) return;
  const int c = static_cast<int>(idx % d), t = static_cast<int>((idx / d) % npre), b = static_cast<int>(idx / (static_cast<int64_t>(d) * npre));
  x[(static_cast<int64_t>(b) * ntok + t) * d + c] = (t == 0) ? cls[c] : reg[static_cast<int64_t>(t - 1) * d + c];
}

// 16-byte vectorised strided row-block copy
__global__ void copy_rows_kernel(const bf16* __restrict__ src, int64_t lds, int rows_src, int src_off,
                                 bf16* __restrict__ dst, int64_t ldd, int rows_dst, int dst_off, int B, int rows,
                                 int cols8) {
  pdl_trigger();
  pdl_wait();
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * rows * cols8) return;
  const int c = static_cast<int>(idx % cols8), r = static_cast<int>((idx / cols8) % rows);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(cols8) * rows));
  const uint4 v = *reinterpret_cast<const uint4*>(src + (static_cast<int64_t>(b) * rows_src + src_off + r) * lds + c * 8);
  *reinterpret_cast<uint4*>(dst + (static_cast<int64_t>(b) * rows_dst + dst_off + r) * ldd + c * 8) = v;
}

__global__ void embed_splice_kernel(const int64_t* __restrict__ ids, int ld_ids, const bf16* __restrict__ table,
                                    bf16* __restrict__ x, int B, int T, int P, int d8) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * T * d8) return;
  const int c = static_cast<int>(idx % d8), t = static_cast<int>((idx / d8) % T);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(d8) * T));
  const int64_t id = ids[b * ld_ids + t];
  const int pos = (t == 0) ? 0 : P + t;   // [BOS | P patch rows | rest of the text]
  const int64_t d = static_cast<int64_t>(d8) * 8;
  *reinterpret_cast<uint4*>(x + (static_cast<int64_t>(b) * (T + P) + pos) * d + c * 8) =
      *reinterpret_cast<const uint4*>(table + id * d + c * 8);
}

__global__ void gelu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ pre, bf16* __restrict__ dx, int64_t n8) {
  pdl_trigger();
  pdl_wait();
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= n8) return;
  const uint4 g = reinterpret_cast<const uint4*>(dy)[idx], p = reinterpret_cast<const uint4*>(pre)[idx];
  const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, pw[4] = {p.x, p.y, p.z, p.w};
  uint32_t o[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 a = unpack_bf16x2(gw[t]), x = unpack_bf16x2(pw[t]);
    o[t] = pack_bf16x2(a.x * gelu_erf_grad(x.x), a.y * gelu_erf_grad(x.y));
  }
  reinterpret_cast<uint4*>(dx)[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void scale_cols_kernel(const bf16* __restrict__ x, const bf16* __restrict__ gamma, bf16* __restrict__ y,
                                  int64_t rows, int cols8) {
  pdl_trigger();
  pdl_wait();
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= rows * cols8) return;
  const int c = static_cast<int>(idx % cols8);
  const uint4 a = reinterpret_cast<const uint4*>(x)[idx], g = reinterpret_cast<const uint4*>(gamma)[c];
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
  uint32_t o[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 u = unpack_bf16x2(aw[t]), v = unpack_bf16x2(gw[t]);
    o[t] = pack_bf16x2(u.x * v.x, u.y * v.y);
  }
  reinterpret_cast<uint4*>(y)[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_fast(x); }

__global__ void swiglu_fwd_kernel(const bf16* __restrict__ gu, bf16* __restrict__ act, int64_t M, int F8) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= M * F8) return;
  const int c = static_cast<int>(idx % F8);
  const int64_t m = idx / F8;
  const int64_t F = static_cast<int64_t>(F8) * 8;
  const int64_t gcol = static_cast<int64_t>(c / 8) * 128 + (c % 8) * 8;   // interleaved [gate 64 | up 64] groups
  const uint4 g = *reinterpret_cast<const uint4*>(gu + m * 2 * F + gcol);
  const uint4 u = *reinterpret_cast<const uint4*>(gu + m * 2 * F + gcol + 64);
  const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, uw[4] = {u.x, u.y, u.z, u.w};
  uint32_t o[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 a = unpack_bf16x2(gw[t]), b = unpack_bf16x2(uw[t]);
    o[t] = pack_bf16x2(rbf(silu_f(a.x)) * b.x, rbf(silu_f(a.y)) * b.y);
  }
  *reinterpret_cast<uint4*>(act + m * F + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void swiglu_bwd_kernel(const bf16* __restrict__ dact, const bf16* __restrict__ gu, bf16* __restrict__ dgu,
                                  int64_t M, int F8) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= M * F8) return;
  const int c = static_cast<int>(idx % F8);
  const int64_t m = idx / F8;
  const int64_t F = static_cast<int64_t>(F8) * 8;
  const int64_t gcol = static_cast<int64_t>(c / 8) * 128 + (c % 8) * 8;   // interleaved [gate 64 | up 64] groups
  const uint4 d = *reinterpret_cast<const uint4*>(dact + m * F + c * 8);
  const uint4 g = *reinterpret_cast<const uint4*>(gu + m * 2 * F + gcol);
  const uint4 u = *reinterpret_cast<const uint4*>(gu + m * 2 * F + gcol + 64);
  const uint32_t dw[4] = {d.x, d.y, d.z, d.w}, gw[4] = {g.x, g.y, g.z, g.w}, uw[4] = {u.x, u.y, u.z, u.w};
  uint32_t og[4], ou[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 dd = unpack_bf16x2(dw[t]), gg = unpack_bf16x2(gw[t]), uu = unpack_bf16x2(uw[t]);
    float r[2][2];
    const float dv[2] = {dd.x, dd.y}, gv[2] = {gg.x, gg.y}, uv[2] = {uu.x, uu.y};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float sig = sigmoid_fast(gv[e]);
      const float s_b = rbf(gv[e] * sig);                 // saved bf16 output of silu
      const float ds = rbf(dv[e] * uv[e]);                // grad wrt silu output
      r[0][e] = ds * sig * (1.f + gv[e] * (1.f - sig));   // grad wrt gate
      r[1][e] = dv[e] * s_b;                              // grad wrt up
    }
    og[t] = pack_bf16x2(r[0][0], r[0][1]);
    ou[t] = pack_bf16x2(r[1][0], r[1][1]);
  }
  *reinterpret_cast<uint4*>(dgu + m * 2 * F + gcol) = make_uint4(og[0], og[1], og[2], og[3]);
  *reinterpret_cast<uint4*>(dgu + m * 2 * F + gcol + 64) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
}

// one thread per (row, q|k, head, 8 consecutive rotation pairs): two 16-byte loads, two 16-byte stores
__global__ void rope_kernel(bf16* __restrict__ qkv, const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                            int64_t M, int L, int H, int hd, float sgn) {
  const int half = hd / 2, oct = half / 8;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= M * 2 * H * oct) return;
  const int o = static_cast<int>(idx % oct);
  const int h = static_cast<int>((idx / oct) % H);
  const int which = static_cast<int>((idx / (static_cast<int64_t>(oct) * H)) % 2);
  const int64_t m = idx / (static_cast<int64_t>(oct) * H * 2);
  const int pos = static_cast<int>(m % L);
  bf16* p = qkv + m * 3 * H * hd + static_cast<int64_t>(which) * H * hd + h * hd + o * 8;
  const uint4 lo = *reinterpret_cast<const uint4*>(p), hi = *reinterpret_cast<const uint4*>(p + half);
  const float4 c0 = *reinterpret_cast<const float4*>(cos_tab + pos * half + o * 8);
  const float4 c1 = *reinterpret_cast<const float4*>(cos_tab + pos * half + o * 8 + 4);
  const float4 s0 = *reinterpret_cast<const float4*>(sin_tab + pos * half + o * 8);
  const float4 s1 = *reinterpret_cast<const float4*>(sin_tab + pos * half + o * 8 + 4);
  const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w}, hw[4] = {hi.x, hi.y, hi.z, hi.w};
  uint32_t ol[4], oh[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 a = unpack_bf16x2(lw[t]), b = unpack_bf16x2(hw[t]);
    const float x1[2] = {a.x, a.y}, x2[2] = {b.x, b.y};
    float r1[2], r2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float cc = c[2 * t + e], ss = sgn * sn[2 * t + e];
      r1[e] = rbf(x1[e] * cc) + rbf(-x2[e] * ss);   // q*cos + rotate_half(q)*sin, each product a bf16 tensor in HF
      r2[e] = rbf(x2[e] * cc) + rbf(x1[e] * ss);
    }
    ol[t] = pack_bf16x2(r1[0], r1[1]);
    oh[t] = pack_bf16x2(r2[0], r2[1]);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  *reinterpret_cast<uint4*>(p + half) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
}

__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, int64_t n8) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= n8) return;
  const uint4 x = reinterpret_cast<const uint4*>(a)[idx], y = reinterpret_cast<const uint4*>(b)[idx];
  const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
  uint32_t o[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 u = unpack_bf16x2(xw[t]), v = unpack_bf16x2(yw[t]);
    o[t] = pack_bf16x2(u.x + v.x, u.y + v.y);
  }
  reinterpret_cast<uint4*>(out)[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void transpose_kernel(const bf16* __restrict__ w, int64_t ldi, bf16* __restrict__ wt, int64_t ldo, int rows, int cols) {
  __shared__ bf16 tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int y = by + r, x = bx + threadIdx.x;
    if (y < rows && x < cols) tile[r][threadIdx.x] = w[static_cast<int64_t>(y) * ldi + x];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int x = bx + r, y = by + threadIdx.x;   // wt[x][y]
    if (x < cols && y < rows) wt[static_cast<int64_t>(x) * ldo + y] = tile[threadIdx.x][r];
  }
}

__global__ void gather_rows_kernel(const bf16* __restrict__ src, const int* __restrict__ rows, bf16* __restrict__ dst,
                                   int R, int d8, int scatter) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(R) * d8) return;
  const int c = static_cast<int>(idx % d8), r = static_cast<int>(idx / d8);
  const int64_t d = static_cast<int64_t>(d8) * 8;
  const int64_t g = rows[r];
  if (scatter)
    *reinterpret_cast<uint4*>(dst + g * d + c * 8) = *reinterpret_cast<const uint4*>(src + r * d + c * 8);
  else
    *reinterpret_cast<uint4*>(dst + r * d + c * 8) = *reinterpret_cast<const uint4*>(src + g * d + c * 8);
}

}  // namespace

#define EW_DONE()      \
  VLA_LAUNCH_CHECK();  \
  ++g_vla_launch_count; \
  return 0

int im2col_patches(const bf16* px, bf16* a_dino, bf16* a_sig, int B, int H, int W, int P, int kpad, cudaStream_t s) {
  VLA_REQUIRE(H % P == 0 && W % P == 0 && kpad >= 3 * P * P, "im2col: bad geometry H=%d W=%d P=%d kpad=%d", H, W, P, kpad);
  const int64_t n = static_cast<int64_t>(B) * (H / P) * (W / P) * kpad;
  im2col_kernel<<<blocks_for(n), EW_THREADS, 0, s>>>(px, a_dino, a_sig, B, H, W, P, kpad);
  EW_DONE();
}
int col2im_patches(const bf16* da_dino, const bf16* da_sig, bf16* dpx, int B, int H, int W, int P, int kpad, cudaStream_t s) {
  const int64_t n = static_cast<int64_t>(B) * 6 * H * W;
  col2im_kernel<<<blocks_for(n), EW_THREADS, 0, s>>>(da_dino, da_sig, dpx, B, H, W, P, kpad);
  EW_DONE();
}
int write_prefix_tokens(const bf16* cls, const bf16* reg, bf16* x, int B, int ntok, int npre, int d, cudaStream_t s) {
  if (npre == 0) return 0;
  prefix_tokens_kernel<<<blocks_for(static_cast<int64_t>(B) * npre * d), EW_THREADS, 0, s>>>(cls, reg, x, B, ntok, npre, d);
  EW_DONE();
}
int copy_rows(const bf16* src, int64_t lds, int rows_src, int src_off, bf16* dst, int64_t ldd, int rows_dst, int dst_off,
              int B, int rows, int cols, cudaStream_t s) {
  VLA_REQUIRE(cols % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "copy_rows: cols/ld must be multiples of 8");
  VLA_CHECK_CUDA(vla_launch(copy_rows_kernel, dim3(blocks_for(static_cast<int64_t>(B) * rows * (cols / 8))), dim3(EW_THREADS), 0, s,
                            src, lds, rows_src, src_off, dst, ldd, rows_dst, dst_off, B, rows, cols / 8));
  EW_DONE();
}
int embed_tokens_splice(const int64_t* ids, int ld_ids, const bf16* table, bf16* x, int B, int T, int P, int d,
                        cudaStream_t s) {
  VLA_REQUIRE(d % 8 == 0, "embed: hidden must be a multiple of 8");
  embed_splice_kernel<<<blocks_for(static_cast<int64_t>(B) * T * (d / 8)), EW_THREADS, 0, s>>>(ids, ld_ids, table, x, B, T, P, d / 8);
  EW_DONE();
}
int gelu_bwd(const bf16* dy, const bf16* pre, bf16* dx, int64_t n, cudaStream_t s) {
  VLA_REQUIRE(n % 8 == 0, "gelu_bwd: n %% 8 != 0");
  VLA_CHECK_CUDA(vla_launch(gelu_bwd_kernel, dim3(blocks_for(n / 8)), dim3(EW_THREADS), 0, s, dy, pre, dx, n / 8));
  EW_DONE();
}
int scale_cols(const bf16* x, const bf16* gamma, bf16* y, int64_t rows, int cols, cudaStream_t s) {
  VLA_REQUIRE(cols % 8 == 0, "scale_cols: cols %% 8 != 0");
  VLA_CHECK_CUDA(vla_launch(scale_cols_kernel, dim3(blocks_for(rows * (cols / 8))), dim3(EW_THREADS), 0, s, x, gamma, y, rows, cols / 8));
  EW_DONE();
}
int swiglu_fwd(const bf16* gu, bf16* act, int64_t M, int F, cudaStream_t s) {
  VLA_REQUIRE(F % 64 == 0, "swiglu: F must be a multiple of 64 (interleaved gate|up groups)");
  swiglu_fwd_kernel<<<blocks_for(M * (F / 8)), EW_THREADS, 0, s>>>(gu, act, M, F / 8);
  EW_DONE();
}
int swiglu_bwd(const bf16* dact, const bf16* gu, bf16* dgu, int64_t M, int F, cudaStream_t s) {
  VLA_REQUIRE(F % 64 == 0, "swiglu: F must be a multiple of 64 (interleaved gate|up groups)");
  swiglu_bwd_kernel<<<blocks_for(M * (F / 8)), EW_THREADS, 0, s>>>(dact, gu, dgu, M, F / 8);
  EW_DONE();
}
int rope_inplace(bf16* qkv, const float* cos_tab, const float* sin_tab, int64_t M, int L, int H, int hd, int dir,
                 cudaStream_t s) {
  VLA_REQUIRE(hd % 16 == 0, "rope: head dim must be a multiple of 16");
  rope_kernel<<<blocks_for(M * 2 * H * (hd / 16)), EW_THREADS, 0, s>>>(qkv, cos_tab, sin_tab, M, L, H, hd, dir >= 0 ? 1.f : -1.f);
  EW_DONE();
}
int add_bf16(const bf16* a, const bf16* b, bf16* out, int64_t n, cudaStream_t s) {
  VLA_REQUIRE(n % 8 == 0, "add: n %% 8 != 0");
  add_kernel<<<blocks_for(n / 8), EW_THREADS, 0, s>>>(a, b, out, n / 8);
  EW_DONE();
}
int transpose_bf16(const bf16* w, int64_t ldi, bf16* wt, int64_t ldo, int rows, int cols, cudaStream_t s) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, s>>>(w, ldi, wt, ldo, rows, cols);
  EW_DONE();
}
int gather_rows(const bf16* src, const int* rows, bf16* dst, int R, int d, cudaStream_t s) {
  VLA_REQUIRE(d % 8 == 0, "gather_rows: d %% 8 != 0");
  if (R == 0) return 0;
  gather_rows_kernel<<<blocks_for(static_cast<int64_t>(R) * (d / 8)), EW_THREADS, 0, s>>>(src, rows, dst, R, d / 8, 0);
  EW_DONE();
}
int scatter_rows(const bf16* src, const int* rows, bf16* dst, int R, int d, cudaStream_t s) {
  VLA_REQUIRE(d % 8 == 0, "scatter_rows: d %% 8 != 0");
  if (R == 0) return 0;
  gather_rows_kernel<<<blocks_for(static_cast<int64_t>(R) * (d / 8)), EW_THREADS, 0, s>>>(src, rows, dst, R, d / 8, 1);
  EW_DONE();
}
