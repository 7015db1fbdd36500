// FP32 issue / pipe microbenchmarks on the box (SURVEY.md 8d asks for a measured FP32 roof).
//   ffma      : scalar FFMA chains (the FP32 roof used by bench.py / DESIGN.md)
//   ffma2     : packed fma.rn.f32x2 chains (SASS FFMA2) — does one instruction retire two FMAs at the same issue cost?
//   cmac_lds  : the inner loop shape of the Kronecker kernels: one 8-byte shared-memory load per complex MAC, as 4 FFMA or as
//               2 FFMA2 — how much of the roof survives when load instructions share the issue slots
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_peak tools/ffma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 unpack2(f32x2 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ void fma2(f32x2& acc, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

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

__global__ void ffma2(float* out, int iters) {
  f32x2 a[16];
  for (int i = 0; i < 16; ++i) a[i] = pack2(threadIdx.x * 1e-3f + i, 0.25f * i);
  const f32x2 b = pack2(1.0001f, 0.9999f), c = pack2(0.5f + blockIdx.x * 1e-6f, 0.25f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) fma2(a[i], b, c);   // a = b * c + a
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) { const float2 v = unpack2(a[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// acc[q] += u * x[q] (complex), x from shared memory (one LDS.64 per complex MAC), u in registers
template <bool PACKED>
__global__ void cmac_lds(float* out, int iters) {
  __shared__ float2 sx[32 * 25];
  for (int i = threadIdx.x; i < 32 * 25; i += blockDim.x) sx[i] = make_float2(1e-3f * i, 1.f - 1e-3f * i);
  __syncthreads();
  const float2 u = make_float2(1.0001f + 1e-6f * threadIdx.x, 0.0002f);
  const int row = (threadIdx.x >> 3) & 31;
  float s = 0;
  if (PACKED) {
    f32x2 P[25], Q[25];
    for (int q = 0; q < 25; ++q) { P[q] = pack2(0.f, 0.f); Q[q] = pack2(0.f, 0.f); }
    const f32x2 ur = pack2(u.x, u.x), ui = pack2(u.y, u.y);
    for (int it = 0; it < iters; ++it) {
      const float2* x = sx + ((row + it) & 31) * 25;
#pragma unroll
      for (int q = 0; q < 25; ++q) {
        const float2 xv = x[q];
        const f32x2 xp = pack2(xv.x, xv.y);
        fma2(P[q], xp, ur);
        fma2(Q[q], xp, ui);
      }
    }
    for (int q = 0; q < 25; ++q) { const float2 p = unpack2(P[q]), r = unpack2(Q[q]); s += p.x - r.y + p.y + r.x; }
  } else {
    float2 acc[25];
    for (int q = 0; q < 25; ++q) acc[q] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
      const float2* x = sx + ((row + it) & 31) * 25;
#pragma unroll
      for (int q = 0; q < 25; ++q) {
        const float2 xv = x[q];
        acc[q].x = fmaf(u.x, xv.x, acc[q].x); acc[q].x = fmaf(-u.y, xv.y, acc[q].x);
        acc[q].y = fmaf(u.x, xv.y, acc[q].y); acc[q].y = fmaf(u.y, xv.x, acc[q].y);
      }
    }
    for (int q = 0; q < 25; ++q) s += acc[q].x + acc[q].y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char* name, F launch, double flops_per_thread_iter, int iters, int threads) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 2; ++rep) {
    for (int bps = 1; bps <= 8; bps *= 2) {
      cudaEventRecord(a);
      launch(148 * bps, threads, iters);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      const double flops = flops_per_thread_iter * iters * (double)threads * 148 * bps;
      printf("%-22s rep %d ctas/sm %d (%d thr): %.3f ms  %.2f TFLOP/s\n", name, rep, bps, threads, ms, flops / ms / 1e9);
    }
  }
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
  run("ffma", [&](int g, int t, int it) { ffma<<<g, t>>>(out, it); }, 2.0 * 16, 20000, 256);
  run("ffma2 (f32x2)", [&](int g, int t, int it) { ffma2<<<g, t>>>(out, it); }, 4.0 * 16, 20000, 256);
  run("cmac+lds 4xFFMA", [&](int g, int t, int it) { cmac_lds<false><<<g, t>>>(out, it); }, 8.0 * 25, 4000, 256);
  run("cmac+lds 2xFFMA2", [&](int g, int t, int it) { cmac_lds<true><<<g, t>>>(out, it); }, 8.0 * 25, 4000, 256);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
