// Error state + misc C-ABI entry points (see include/vla_b200.h).
#include <stdarg.h>
#include <string.h>

#include "../../include/vla_b200.h"
#include "common.cuh"
#include "gemm.h"

static thread_local char g_err[1024] = "";

void vla_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* vla_last_error(void) { return g_err; }
extern "C" int vla_abi_version(void) { return VLA_B200_ABI_VERSION; }
extern "C" long long vla_launch_count(void) { return g_vla_launch_count; }

extern "C" int vla_gemm_bf16_tn(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldc, int M,
                                int N, int K, const void* bias, const void* gamma, const void* resid, int64_t ldr,
                                int act, void* preact_out, int out_f32, void* stream) {
  GemmEpilogue e;
  e.bias = static_cast<const bf16*>(bias);
  e.gamma = static_cast<const bf16*>(gamma);
  e.resid = static_cast<const bf16*>(resid);
  e.ldr = ldr;
  e.act = act;
  e.preact_out = static_cast<bf16*>(preact_out);
  e.out_f32 = out_f32;
  return gemm_bf16_tn(static_cast<const bf16*>(A), lda, static_cast<const bf16*>(W), ldw, out, ldc, M, N, K, e,
                      static_cast<cudaStream_t>(stream));
}
