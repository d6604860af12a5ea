// F1 / F2: fused patch paste + affine warp + composite + dual normalise (+ bf16 cast), and its backward.
//
// Replaces RandomPatchTransform.apply_random_patch_batch / paste_patch_fix / random_paste_patch / im_process
// (VLAAttacker/white_patch/appply_random_transform.py:104-197) and the `.to(torch.bfloat16)` at UADA.py:142,
// i.e. ~14 ATen launches PER IMAGE (ToTensor, ones*-100, slice-assign, affine_grid, grid_sample, where,
// 2x normalize, cat) become one launch for the whole batch; the backward (grid_sampler_2d_backward, where,
// CopySlices) becomes one launch that recomputes the sampling coordinates instead of saving them.
//
// HBM-bound: 3 B/pixel read (u8 HWC) + 12 B/pixel written (6 x bf16) forward; 12 B/pixel read backward (only
// pixels whose taps touch the patch read their gradient).
#include "kernels.h"

namespace {

struct Taps {
  int x0, y0;      // north-west tap
  float nw, ne, sw, se;
};

// F.affine_grid(align_corners=False) + F.grid_sample(bilinear, padding_mode='border', align_corners=False)
// source coordinate of output pixel (i, j)  (appply_random_transform.py:93-102)
__device__ __forceinline__ Taps warp_taps(const float* __restrict__ th, int i, int j, int H, int W) {
  const float xn = (2.f * j + 1.f) / W - 1.f;
  const float yn = (2.f * i + 1.f) / H - 1.f;
  const float xs = th[0] * xn + th[1] * yn + th[2];
  const float ys = th[3] * xn + th[4] * yn + th[5];
  float ix = ((xs + 1.f) * W - 1.f) * 0.5f;
  float iy = ((ys + 1.f) * H - 1.f) * 0.5f;
  ix = fminf(fmaxf(ix, 0.f), static_cast<float>(W - 1));
  iy = fminf(fmaxf(iy, 0.f), static_cast<float>(H - 1));
  Taps t;
  const float fx = floorf(ix), fy = floorf(iy);
  t.x0 = static_cast<int>(fx);
  t.y0 = static_cast<int>(fy);
  const float ex = fx + 1.f - ix, ey = fy + 1.f - iy;   // (ix_se - ix), (iy_se - iy)
  const float wx = ix - fx, wy = iy - fy;
  t.nw = ex * ey;
  t.ne = wx * ey;
  t.sw = ex * wy;
  t.se = wx * wy;
  return t;
}

// value of the -100 canvas with the patch pasted at (px, py); out-of-image taps contribute 0 (never weighted)
__device__ __forceinline__ float canvas_at(const float* __restrict__ patch_c, int yy, int xx, int px, int py, int ph,
                                           int pw, int H, int W) {
  if (yy >= H || xx >= W) return 0.f;
  const int u = yy - py, v = xx - px;
  if (u >= 0 && u < ph && v >= 0 && v < pw) return __ldg(patch_c + u * pw + v);
  return -100.f;
}

__global__ void patch_frontend_fwd_kernel(const uint8_t* __restrict__ obs, const float* __restrict__ patch,
                                          const int* __restrict__ xy, const float* __restrict__ theta,
                                          bf16* __restrict__ out, int B, int H, int W, int ph, int pw, int mode,
                                          FrontendNorm nrm) {
  const int wq = (W + 3) / 4;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * H * wq) return;
  const int jq = static_cast<int>(idx % wq);
  const int i = static_cast<int>((idx / wq) % H);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(wq) * H));
  const int j0 = jq * 4;
  const int nj = min(4, W - j0);

  float im[3][4];
  const uint8_t* op = obs + (static_cast<int64_t>(b) * H + i) * W * 3 + j0 * 3;
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < 3; ++c) im[c][q] = (q < nj) ? static_cast<float>(op[q * 3 + c]) / 255.f : 0.f;   // ToTensor

  if (mode != FE_MODE_NONE) {
    const int px = xy[b * 2 + 0], py = xy[b * 2 + 1];
    const float* th = theta + b * 6;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q >= nj) break;
      const int j = j0 + q;
      if (mode == FE_MODE_WARP) {
        const Taps t = warp_taps(th, i, j, H, W);
        // skip the 4-tap gather when no tap can touch the patch rectangle (canvas == -100 exactly)
        const bool near = (t.x0 + 1 >= px) && (t.x0 < px + pw) && (t.y0 + 1 >= py) && (t.y0 < py + ph);
        if (near) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* pc = patch + c * ph * pw;
            float v = canvas_at(pc, t.y0, t.x0, px, py, ph, pw, H, W) * t.nw;
            v += canvas_at(pc, t.y0, t.x0 + 1, px, py, ph, pw, H, W) * t.ne;
            v += canvas_at(pc, t.y0 + 1, t.x0, px, py, ph, pw, H, W) * t.sw;
            v += canvas_at(pc, t.y0 + 1, t.x0 + 1, px, py, ph, pw, H, W) * t.se;
            if (!(v < -20.f)) im[c][q] = v;   // torch.where(canvas < -20, im, canvas)
          }
        }
      } else {
        const int u = i - py, v = j - px;
        if (u >= 0 && u < ph && v >= 0 && v < pw) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float cv = __ldg(patch + (c * ph + u) * pw + v);
            const bool keep = (mode == FE_MODE_FIX) ? (cv != -100.f) : !(cv < -20.f);
            if (keep) im[c][q] = cv;
          }
        }
      }
    }
  }

  const int64_t plane = static_cast<int64_t>(H) * W;
  bf16* ob = out + static_cast<int64_t>(b) * 6 * plane + static_cast<int64_t>(i) * W + j0;
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float y[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) y[q] = (im[c][q] - nrm.mean[s][c]) / nrm.std[s][c];
      bf16* dst = ob + (s * 3 + c) * plane;
      if (nj == 4 && (W % 4 == 0)) {
        uint2 pk = make_uint2(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]));
        *reinterpret_cast<uint2*>(dst) = pk;
      } else {
        for (int q = 0; q < nj; ++q) dst[q] = f2b(y[q]);
      }
    }
}

// d patch[c,u,v] += sum over images / pixels kept by the composite of (bilinear weight) * (g0/std0 + g1/std1)
__global__ void patch_frontend_bwd_kernel(const bf16* __restrict__ dout, const float* __restrict__ patch,
                                          const int* __restrict__ xy, const float* __restrict__ theta,
                                          float* __restrict__ dpatch, int B, int H, int W, int ph, int pw, int mode,
                                          FrontendNorm nrm) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * H * W) return;
  const int j = static_cast<int>(idx % W);
  const int i = static_cast<int>((idx / W) % H);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(W) * H));
  const int px = xy[b * 2 + 0], py = xy[b * 2 + 1];
  const int64_t plane = static_cast<int64_t>(H) * W;
  const bf16* gb = dout + static_cast<int64_t>(b) * 6 * plane + static_cast<int64_t>(i) * W + j;

  if (mode == FE_MODE_WARP) {
    const Taps t = warp_taps(theta + b * 6, i, j, H, W);
    const bool near = (t.x0 + 1 >= px) && (t.x0 < px + pw) && (t.y0 + 1 >= py) && (t.y0 < py + ph);
    if (!near) return;
    const int tx[4] = {t.x0, t.x0 + 1, t.x0, t.x0 + 1};
    const int ty[4] = {t.y0, t.y0, t.y0 + 1, t.y0 + 1};
    const float tw[4] = {t.nw, t.ne, t.sw, t.se};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pc = patch + c * ph * pw;
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) v += canvas_at(pc, ty[k], tx[k], px, py, ph, pw, H, W) * tw[k];
      if (v < -20.f) continue;   // the composite kept the clean image: no gradient to the canvas
      const float g = b2f(gb[c * plane]) / nrm.std[0][c] + b2f(gb[(3 + c) * plane]) / nrm.std[1][c];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int u = ty[k] - py, vv = tx[k] - px;
        if (ty[k] < H && tx[k] < W && u >= 0 && u < ph && vv >= 0 && vv < pw && tw[k] != 0.f)
          atomicAdd(dpatch + (c * ph + u) * pw + vv, tw[k] * g);
      }
    }
  } else if (mode != FE_MODE_NONE) {
    const int u = i - py, v = j - px;
    if (u < 0 || u >= ph || v < 0 || v >= pw) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float cv = __ldg(patch + (c * ph + u) * pw + v);
      const bool keep = (mode == FE_MODE_FIX) ? (cv != -100.f) : !(cv < -20.f);
      if (!keep) continue;
      const float g = b2f(gb[c * plane]) / nrm.std[0][c] + b2f(gb[(3 + c) * plane]) / nrm.std[1][c];
      atomicAdd(dpatch + (c * ph + u) * pw + v, g);
    }
  }
}


// Eval-time paste (RandomPatchTransform.simulation_random_patch, appply_random_transform.py:43-78): the patch is
// quantised to uint8 (ToPILImage: floor(p * 255)), pasted at a FIXED position on a -100 canvas, optionally warped by a
// FIXED affine map, composited with `canvas < 0 ? image : canvas` and truncated back to uint8 HWC (numpy astype).
// One thread per pixel; 3 B read + 3 B written per pixel.
__global__ void patch_sim_paste_kernel(const uint8_t* __restrict__ img, const float* __restrict__ patch,
                                       const int* __restrict__ xy, const float* __restrict__ theta, uint8_t* __restrict__ out,
                                       int B, int H, int W, int ph, int pw, int geometry) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(B) * H * W) return;
  const int j = static_cast<int>(idx % W);
  const int i = static_cast<int>((idx / W) % H);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(W) * H));
  const int px = xy[b * 2 + 0], py = xy[b * 2 + 1];
  const uint8_t* ip = img + idx * 3;
  uint8_t* op = out + idx * 3;
  auto q_at = [&](const float* pc, int yy, int xx) -> float {   // quantised canvas
    if (yy >= H || xx >= W) return 0.f;
    const int u = yy - py, v = xx - px;
    if (u >= 0 && u < ph && v >= 0 && v < pw) return truncf(__ldg(pc + u * pw + v) * 255.f);
    return -100.f;
  };
  Taps t;
  bool near = true;
  if (geometry) {
    t = warp_taps(theta + b * 6, i, j, H, W);
    near = (t.x0 + 1 >= px) && (t.x0 < px + pw) && (t.y0 + 1 >= py) && (t.y0 < py + ph);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* pc = patch + c * ph * pw;
    float v = -100.f;
    if (geometry) {
      if (near) {
        v = q_at(pc, t.y0, t.x0) * t.nw;
        v += q_at(pc, t.y0, t.x0 + 1) * t.ne;
        v += q_at(pc, t.y0 + 1, t.x0) * t.sw;
        v += q_at(pc, t.y0 + 1, t.x0 + 1) * t.se;
      }
    } else {
      v = q_at(pc, i, j);
    }
    op[c] = (v < 0.f) ? ip[c] : static_cast<uint8_t>(v);   // torch.where(canvas < 0, image, canvas).astype(uint8)
  }
}
}  // namespace

int patch_frontend_fwd(const uint8_t* obs, const float* patch, const int* xy, const float* theta, bf16* out, int B,
                       int H, int W, int ph, int pw, int mode, const FrontendNorm& nrm, cudaStream_t stream) {
  VLA_REQUIRE(B > 0 && H > 0 && W > 0, "patch_frontend_fwd: empty batch");
  VLA_REQUIRE(mode >= FE_MODE_WARP && mode <= FE_MODE_NONE, "patch_frontend_fwd: bad mode %d", mode);
  VLA_REQUIRE(mode == FE_MODE_NONE || (ph <= H && pw <= W && ph > 0 && pw > 0), "patch larger than the image");
  const int64_t n = static_cast<int64_t>(B) * H * ((W + 3) / 4);
  const int threads = 256;
  patch_frontend_fwd_kernel<<<static_cast<unsigned>(ceil_div64(n, threads)), threads, 0, stream>>>(
      obs, patch, xy, theta, out, B, H, W, ph, pw, mode, nrm);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

int patch_frontend_bwd(const bf16* dout, const float* patch, const int* xy, const float* theta, float* dpatch, int B,
                       int H, int W, int ph, int pw, int mode, const FrontendNorm& nrm, cudaStream_t stream) {
  VLA_REQUIRE(B > 0 && H > 0 && W > 0, "patch_frontend_bwd: empty batch");
  VLA_CHECK_CUDA(cudaMemsetAsync(dpatch, 0, sizeof(float) * 3 * ph * pw, stream));
  if (mode == FE_MODE_NONE) return 0;
  const int64_t n = static_cast<int64_t>(B) * H * W;
  const int threads = 256;
  patch_frontend_bwd_kernel<<<static_cast<unsigned>(ceil_div64(n, threads)), threads, 0, stream>>>(
      dout, patch, xy, theta, dpatch, B, H, W, ph, pw, mode, nrm);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

int patch_sim_paste(const uint8_t* img, const float* patch, const int* xy, const float* theta, uint8_t* out, int B, int H,
                    int W, int ph, int pw, int geometry, cudaStream_t stream) {
  VLA_REQUIRE(B > 0 && H > 0 && W > 0 && ph > 0 && pw > 0 && ph <= H && pw <= W, "patch_sim_paste: bad shape B=%d %dx%d patch %dx%d", B,
              H, W, ph, pw);
  const int64_t n = static_cast<int64_t>(B) * H * W;
  patch_sim_paste_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(img, patch, xy, theta, out, B, H, W, ph, pw, geometry);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
