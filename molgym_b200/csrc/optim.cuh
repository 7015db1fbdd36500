// optim.cuh — the optimizer tail of one PPO epoch on the flat buffers: global gradient norm
// (tools/util.py:61-69 compute_gradient_norm), clipping (torch.nn.utils.clip_grad_norm_, ppo.py:144) and the Adam / AMSGrad
// update (tools/util.py:197-205, torch.optim.Adam semantics) as two launches instead of ~10^3 tiny ones.
#pragma once
#include "common.cuh"

namespace mgb {

constexpr int kNormThreads = 256;
constexpr int kNormMaxBlocks = 296;

// sum of squares of g[0..n) in float64: per-CTA partials, the LAST CTA to finish adds them in index order (deterministic) and
// writes norm[0] = sqrt(sum), norm[1] = sum; `scratch` holds kNormMaxBlocks doubles + one counter (zero before the first launch,
// reset by the kernel).
__global__ void __launch_bounds__(kNormThreads)
k_grad_norm(const float* __restrict__ g, long long n, double* __restrict__ scratch, double* __restrict__ norm) {
  __shared__ double red[kNormThreads / 32];
  __shared__ bool last;
  double acc = 0.0;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += (double)g[i] * g[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + kNormMaxBlocks);
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < kNormThreads / 32; ++w) tot += red[w];
    scratch[blockIdx.x] = tot;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double tot = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) tot += reinterpret_cast<volatile double*>(scratch)[b];
    norm[0] = sqrt(tot);
    norm[1] = tot;
    *counter = 0u;
  }
}

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay;
  float step_size, bias2_sqrt;   // lr / (1 - beta1^t), sqrt(1 - beta2^t): formed in float64 on the host like torch does
  float max_norm;            // <= 0: no clipping
  int amsgrad, maximize;
};

// torch.optim.Adam (single-tensor formulation, float32 state):
//   g = clip_coef * grad (+ weight_decay * p);  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2
//   denom = sqrt(max_v or v) / sqrt(1 - b2^t) + eps;  p -= lr / (1 - b1^t) * m / denom
// clip_coef = min(1, max_norm / (norm + 1e-6)) as in clip_grad_norm_ (read from the device: no host round trip).
__global__ void __launch_bounds__(256)
k_adam_step(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float* __restrict__ vmax,
            long long n, const double* __restrict__ norm, AdamArgs a) {
  float coef = 1.f;
  if (a.max_norm > 0.f && norm) {
    const float c = a.max_norm / ((float)norm[0] + 1e-6f);
    coef = c < 1.f ? c : 1.f;
  }
  const float step_size = a.step_size;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * coef;
    if (a.maximize) gi = -gi;
    const float pi = p[i];
    if (a.weight_decay != 0.f) gi = fmaf(a.weight_decay, pi, gi);
    const float mi = m[i] + (gi - m[i]) * (1.f - a.beta1);          // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(a.beta2, v[i], (1.f - a.beta2) * gi * gi);   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    m[i] = mi;
    v[i] = vi;
    float vv = vi;
    if (a.amsgrad) {
      vv = fmaxf(vmax[i], vi);
      vmax[i] = vv;
    }
    const float denom = sqrtf(vv) / a.bias2_sqrt + a.eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

}  // namespace mgb
