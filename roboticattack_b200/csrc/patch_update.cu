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
                                                                  float grad_scale, float clip_l1, float* __restrict__ scalars) {
  __shared__ float red[32];
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
  }
}

}  // namespace

int patch_update(float* patch, const float* grad, float* m, float* v, int n, int step, float lr, float beta1, float beta2,
                 float eps, int kind, float grad_scale, float clip_l1, float* scalars, cudaStream_t s) {
  VLA_REQUIRE(n > 0, "patch_update: empty patch");
  VLA_REQUIRE(kind == OPT_ADAMW || kind == OPT_PGD, "patch_update: bad optimiser kind %d", kind);
  VLA_REQUIRE(kind != OPT_ADAMW || (step >= 1 && m && v), "patch_update: AdamW needs step >= 1 and moment buffers");
  patch_update_kernel<<<1, PU_THREADS, 0, s>>>(patch, grad, m, v, n, step, lr, beta1, beta2, eps, kind, grad_scale, clip_l1, scalars);
  VLA_LAUNCH_CHECK();
  ++g_vla_launch_count;
  return 0;
}
