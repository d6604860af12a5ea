// Microbenchmark: peak of the legacy mma.sync.m16n8k16 bf16 path on this GPU (register-resident operands).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void k(float* out, int iters) {
  uint32_t a[4] = {threadIdx.x, 1, 2, 3}, b0 = 5, b1 = 7;
  float d[8][4] = {};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int j = 0; j < 8; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  for (int warps : {4, 8, 16, 32}) {
    int iters = 20000;
    k<<<148 * 2, warps * 32 / 2>>>(out, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<148 * 2, warps * 32 / 2>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 148.0 * warps * iters * 8 * (16.0 * 8 * 16 * 2);
    printf("warps/SM=%d: %.1f TFLOP/s (%.2f ms)\n", warps, fl / ms / 1e9, ms);
  }
  return 0;
}
