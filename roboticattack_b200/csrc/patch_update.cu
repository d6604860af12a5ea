// F4: fused patch update tail -- gradient scaling (1/world after the all-reduce), optional L1 grad-norm clip,
// transformers.AdamW rule or sign-PGD, clamp to [0,1].  One launch instead of ~7 tiny ATen launches.
// Reference: UADA.py:107-115,155-157; UADA_ddp.py:167-174,208-209; UPA.py:157-160 (clip_grad_norm_ L1 1e-3);
// TMA.py:164-175 (AdamW / sign-PGD).  AdamW here is HF's: eps added to sqrt(v) BEFORE the bias correction.
#include <math.h>

#include "kernels.h"

namespace {

constexpr int PU_THREADS = 1024;

__global__ void __launch_bounds__(PU_THREADS) patch_update_kernel(float* __restrict__ patch, const float* __restrict__ grad,
                                                                  float* __restrict__ m, float* __restrict__ v, int n, int step,
                                                                  float lr, float beta1, float beta2, float eps, int kind,
                                                                  float grad_scale, float clip_l1, float* __restrict__ scalars,
                                                                  const StepState* __restrict__ st, float* __restrict__ zero_after) {
  __shared__ float red[32];
  if (st != nullptr) {   // graph-replayable form: the step counter and the learning rate are read from device memory
    step = st->adam_t + 1;
    lr = st->lr;
  }
  float gs = 0.f, ga = 0.f;
  for (int i = threadIdx.x; i < n; i += PU_THREADS) {
    const float g = grad[i] * grad_scale;
    gs += g;
    ga += fabsf(g);
  }
  gs = block_sum(gs, red);
  ga = block_sum(ga, red);
  if (threadIdx.x == 0 && scalars) scalars[LS_GRAD_MEAN] = gs / n;   // patch.grad.mean() logged before the step
  float coef = 1.f;
  if (clip_l1 > 0.f) coef = fminf(clip_l1 / (ga + 1e-6f), 1.f);     // torch.nn.utils.clip_grad_norm_(norm_type=1)
  // step_size = lr * sqrt(1 - b2^t) / (1 - b1^t)
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  const float step_size = lr * sqrtf(bc2) / bc1;
  for (int i = threadIdx.x; i < n; i += PU_THREADS) {
    const float g = grad[i] * grad_scale * coef;
    float p = patch[i];
    if (kind == OPT_ADAMW) {
      const float mi = beta1 * m[i] + (1.f - beta1) * g;
      const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
      m[i] = mi;
      v[i] = vi;
      p -= step_size * (mi / (sqrtf(vi) + eps));
    } else {
      const float sg = (g > 0.f) ? 1.f : ((g < 0.f) ? -1.f : 0.f);
      p -= lr * sg;
    }
    patch[i] = fminf(fmaxf(p, 0.f), 1.f);
    if (zero_after) zero_after[i] = 0.f;   // the accumulation buffer of TMA / UPA (optimizer.zero_grad after a stepping iteration)
  }
}

// ---- device-side step state of vla_attack_step (see include/vla_b200.h) ----
__global__ void step_begin_kernel(const StepState* __restrict__ st, const int* __restrict__ xy_all, const float* __restrict__ th_all,
                                  int* __restrict__ xy_cur, float* __restrict__ th_cur, int B, int n_place,
                                  float* __restrict__ scal_cur) {
  if (threadIdx.x == 0) scal_cur[LS_GRAD_MEAN] = 0.f;   // stays 0 on an accumulate-only iteration (no update kernel)
  int p = st->place;
  p = p < 0 ? 0 : (p >= n_place ? n_place - 1 : p);
  for (int i = threadIdx.x; i < B * 2; i += blockDim.x) xy_cur[i] = xy_all[static_cast<size_t>(p) * B * 2 + i];
  for (int i = threadIdx.x; i < B * 6; i += blockDim.x) th_cur[i] = th_all[static_cast<size_t>(p) * B * 6 + i];
}
__global__ void step_end_kernel(StepState* __restrict__ st, const float* __restrict__ scal_cur, float* __restrict__ scal_hist,
                                int adam_inc) {
  const int p = st->place;
  if (scal_hist != nullptr && threadIdx.x < LOSS_NUM_SCALARS) scal_hist[static_cast<size_t>(p) * LOSS_NUM_SCALARS + threadIdx.x] = scal_cur[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    st->place = p + 1;
    st->adam_t += adam_inc;
  }
}
__global__ void accumulate_kernel(float* __restrict__ acc, const float* __restrict__ g, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acc[i] += g[i];
}

}  // namespace

int patch_update(float* patch, const float* grad, float* m, float* v, int n, int step, float lr, float beta1, float beta2,
                 float eps, int kind, float grad_scale, float clip_l1, float* scalars, cudaStream_t s) {
  VLA_REQUIRE(n > 0, "patch_update: empty patch");
  VLA_REQUIRE(kind == OPT_ADAMW || kind == OPT_PGD, "patch_update: bad optimiser kind %d", kind);
  VLA_REQUIRE(kind != OPT_ADAMW || (step >= 1 && m && v), "patch_update: AdamW needs step >= 1 and moment buffers");
  patch_update_kernel<<<1, PU_THREADS, 0, s>>>(patch, grad, m, v, n, step, lr, beta1, beta2, eps, kind, grad_scale, clip_l1, scalars,
                                               nullptr, nullptr);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

int patch_update_dev(float* patch, const float* grad, float* m, float* v, int n, const StepState* st, float beta1, float beta2,
                     float eps, int kind, float grad_scale, float clip_l1, float* scalars, float* zero_after, cudaStream_t s) {
  VLA_REQUIRE(n > 0 && st != nullptr, "patch_update_dev: empty patch / no step state");
  VLA_REQUIRE(kind == OPT_ADAMW || kind == OPT_PGD, "patch_update: bad optimiser kind %d", kind);
  VLA_REQUIRE(kind != OPT_ADAMW || (m && v), "patch_update: AdamW needs moment buffers");
  patch_update_kernel<<<1, PU_THREADS, 0, s>>>(patch, grad, m, v, n, 0, 0.f, beta1, beta2, eps, kind, grad_scale, clip_l1, scalars, st,
                                               zero_after);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}

int step_begin(const StepState* st, const int* xy_all, const float* th_all, int* xy_cur, float* th_cur, int B, int n_place,
               float* scal_cur, cudaStream_t s) {
  step_begin_kernel<<<1, 128, 0, s>>>(st, xy_all, th_all, xy_cur, th_cur, B, n_place, scal_cur);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
int step_end(StepState* st, const float* scal_cur, float* scal_hist, int adam_inc, cudaStream_t s) {
  step_end_kernel<<<1, 32, 0, s>>>(st, scal_cur, scal_hist, adam_inc);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
int accumulate_f32(float* acc, const float* g, int n, cudaStream_t s) {
  accumulate_kernel<<<ceil_div(n, 256), 256, 0, s>>>(acc, g, n);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
