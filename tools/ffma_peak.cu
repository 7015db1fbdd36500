// FFMA-chain peak on the box (SURVEY.md 8d asks for a measured FP32 roof).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void ffma(float* out, int iters) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  float b = 1.0001f, c = 0.5f + blockIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
  int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    for (int bps = 1; bps <= 8; bps *= 2) {
      cudaEventRecord(a);
      ffma<<<148 * bps, 256>>>(out, iters);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      double flops = 2.0 * 16 * iters * 256.0 * 148 * bps;
      printf("rep %d ctas/sm %d: %.3f ms  %.2f TFLOP/s\n", rep, bps, ms, flops / ms / 1e9);
    }
  }
  return 0;
}
