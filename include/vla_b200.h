/* placeholder, filled in below */
#pragma once
#include <stdint.h>
#define VLA_B200_ABI_VERSION 1
#ifdef __cplusplus
extern "C" {
#endif
const char* vla_last_error(void);
int vla_abi_version(void);
long long vla_launch_count(void);
int vla_gemm_bf16_tn(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldc, int M, int N, int K,
                     const void* bias, const void* gamma, const void* resid, int64_t ldr, int act, void* preact_out,
                     int out_f32, void* stream);
#ifdef __cplusplus
}
#endif
