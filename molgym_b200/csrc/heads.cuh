// heads.cuh — policy/value heads of the covariant actor-critic: forward and backward.
//
// Reference: molgym/agents/covariant/agent.py:220-316 (focus / element / distance / orientation heads, value),
// molgym/modules.py:26-50 (masked softmax, MLP), molgym/agents/covariant/gmm.py:8-18, so3_tools.py:47-132,
// spherical_dists.py:79-115,182-215,273-286, covariant/modules.py:180-190 (CormorantMixer) and
// torch.distributions.Categorical / Normal / MixtureSameFamily arithmetic.
#pragma once
#include "cov_forward.cuh"

namespace mgb {

constexpr int kRowTile = 8;
constexpr int kHeadThreads = 128;     // row-MLP kernels
constexpr int kPolicyThreads = 512;   // per-canvas policy forward (one CTA per canvas; the GEMV phases are latency-bound)
constexpr int kPolicyBwdThreads = 256;
// resident CTAs per SM the policy kernels are compiled for.  Backward: 2 (128 registers; measured C3 b1024 341 -> 225 us, C2 unchanged).
// Forward: a template parameter — 1 for minibatches of at most one canvas per SM (no spills: C2 46 us against 55 us), 2 above that
// (64 registers, some spills, but two canvases' latency chains overlap: C3 b1024 236 -> 184 us).
#ifndef MGB_POLICY_BWD_MIN_CTAS
#define MGB_POLICY_BWD_MIN_CTAS 2
#endif
constexpr float kF32Eps = 1.1920928955078125e-07f;
constexpr float kLogSqrt2Pi = 0.9189385332046727f;
constexpr float kLog4Pi = 2.5310242469692907f;

enum RowMode { kRowsAll = 0, kRowsActive = 1, kRowsValid = 2 };
__device__ __forceinline__ bool row_on(int mode, const int* __restrict__ n_atoms, int N, long long r) {
  if (mode == kRowsAll) return true;
  const int b = (int)(r / N), i = (int)(r % N), n = n_atoms[b];
  return mode == kRowsActive ? (i < (n > 1 ? n : 1)) : (i < n);
}

// ------------------------------------------------------------------------------------------------------------
// Two-layer MLPs applied to atom rows (phi_focus, phi_trans):  H = relu(X W0^T + b0),  Y = H W1^T + b1.
// grid = (row tiles, 2): y = 0 focus head on active rows, y = 1 value-transform on valid atoms.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kHeadThreads)
k_rows_mlp_fwd(const CovDesc* __restrict__ dp, const float* __restrict__ P, const float* __restrict__ Wt,
               const int* __restrict__ n_atoms, long long rows, const float* __restrict__ X, float* __restrict__ H0,
               float* __restrict__ Y0, float* __restrict__ H1, float* __restrict__ Y1) {
  const CovDesc& d = *dp;
  const MlpDesc& M = blockIdx.y == 0 ? d.focus : d.trans;
  const int mode = blockIdx.y == 0 ? kRowsActive : kRowsValid;
  float* H = blockIdx.y == 0 ? H0 : H1;
  float* Y = blockIdx.y == 0 ? Y0 : Y1;
  const int K = M.in, Wd = M.hidden, No = M.out;
  MGB_DYN_SMEM(float, sm);
  float* sx = sm;               // [kRowTile][K]
  float* sh = sm + kRowTile * K;  // [kRowTile][Wd]
  __shared__ int s_on[kRowTile];
  const long long r0 = (long long)blockIdx.x * kRowTile;
  if (threadIdx.x < kRowTile) {
    const long long r = r0 + threadIdx.x;
    s_on[threadIdx.x] = (r < rows && row_on(mode, n_atoms, d.N, r)) ? 1 : 0;
  }
  __syncthreads();
  int any = 0;
  for (int q = 0; q < kRowTile; ++q) any |= s_on[q];
  if (!any) return;
  for (int idx = threadIdx.x; idx < kRowTile * K; idx += blockDim.x) {
    const int q = idx / K;
    sx[idx] = s_on[q] ? X[(r0 + q) * K + (idx % K)] : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < Wd; o += blockDim.x) {
    float acc[kRowTile];
    const float bias = P[M.b0 + o];
    MGB_UNROLL
    for (int q = 0; q < kRowTile; ++q) acc[q] = bias;
    const float* w = Wt + M.W0t + o;
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
      const float wk = w[(long long)k * Wd];
      MGB_UNROLL
      for (int q = 0; q < kRowTile; ++q) acc[q] = fmaf(wk, sx[q * K + k], acc[q]);
    }
    MGB_UNROLL
    for (int q = 0; q < kRowTile; ++q) {
      const float h = fmaxf(acc[q], 0.f);
      sh[q * Wd + o] = h;
      if (s_on[q]) H[(r0 + q) * Wd + o] = h;
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < No; o += blockDim.x) {
    float acc[kRowTile];
    const float bias = P[M.b1 + o];
    MGB_UNROLL
    for (int q = 0; q < kRowTile; ++q) acc[q] = bias;
    const float* w = Wt + M.W1t + o;
#pragma unroll 8
    for (int k = 0; k < Wd; ++k) {
      const float wk = w[(long long)k * No];
      MGB_UNROLL
      for (int q = 0; q < kRowTile; ++q) acc[q] = fmaf(wk, sh[q * Wd + k], acc[q]);
    }
    MGB_UNROLL
    for (int q = 0; q < kRowTile; ++q)
      if (s_on[q]) Y[(r0 + q) * No + o] = acc[q];
  }
}

// Backward of the same: dY -> dH (masked by relu) -> dX (+=).  dH is kept for the weight gradient.
//   focus: dY0 [rows,1];  trans: dY = dvf[b,:] broadcast over the valid atoms of canvas b, also written to dY1 [rows,Wd].
__global__ void __launch_bounds__(kHeadThreads)
k_rows_mlp_bwd(const CovDesc* __restrict__ dp, const float* __restrict__ P, const int* __restrict__ n_atoms, long long rows,
               const float* __restrict__ H0, const float* __restrict__ dY0, float* __restrict__ dH0,
               const float* __restrict__ H1, const float* __restrict__ dvf, float* __restrict__ dY1, float* __restrict__ dH1,
               float* __restrict__ dX) {
  const CovDesc& d = *dp;
  MGB_DYN_SMEM(float, sm);
  const int Wd = d.focus.hidden, K = d.focus.in;
  float* sdy = sm;                         // [kRowTile][Wd]  (trans) or [kRowTile] (focus)
  float* sdh = sm + kRowTile * Wd;         // [2][kRowTile][Wd]
  __shared__ int s_on[2][kRowTile];
  const long long r0 = (long long)blockIdx.x * kRowTile;
  if (threadIdx.x < 2 * kRowTile) {
    const int which = threadIdx.x / kRowTile, q = threadIdx.x % kRowTile;
    const long long r = r0 + q;
    s_on[which][q] = (r < rows && row_on(which == 0 ? kRowsActive : kRowsValid, n_atoms, d.N, r)) ? 1 : 0;
  }
  __syncthreads();
  int any = 0;
  for (int q = 0; q < kRowTile; ++q) any |= s_on[0][q];
  if (!any) return;
  // trans: dY rows
  for (int idx = threadIdx.x; idx < kRowTile * Wd; idx += blockDim.x) {
    const int q = idx / Wd, o = idx % Wd;
    float v = 0.f;
    if (s_on[1][q]) {
      v = dvf[((r0 + q) / d.N) * Wd + o];
      dY1[(r0 + q) * Wd + o] = v;
    }
    sdy[idx] = v;
  }
  __syncthreads();
  for (int h = threadIdx.x; h < Wd; h += blockDim.x) {
    // focus: dH0[r][h] = relu'(H0) * W1[0][h] * dY0[r]
    const float w1 = P[d.focus.W1 + h];
    float acc[kRowTile];
    MGB_UNROLL
    for (int q = 0; q < kRowTile; ++q) acc[q] = 0.f;
    // trans: dH1[r][h] = relu'(H1) * sum_o W1[o][h] dY[r][o]
    const float* w = P + d.trans.W1 + h;
#pragma unroll 8
    for (int o = 0; o < Wd; ++o) {
      const float wo = w[(long long)o * Wd];
      MGB_UNROLL
      for (int q = 0; q < kRowTile; ++q) acc[q] = fmaf(wo, sdy[q * Wd + o], acc[q]);
    }
    MGB_UNROLL
    for (int q = 0; q < kRowTile; ++q) {
      float g0 = 0.f, g1 = 0.f;
      if (s_on[0][q]) {
        g0 = H0[(r0 + q) * Wd + h] > 0.f ? w1 * dY0[r0 + q] : 0.f;
        dH0[(r0 + q) * Wd + h] = g0;
      }
      if (s_on[1][q]) {
        g1 = H1[(r0 + q) * Wd + h] > 0.f ? acc[q] : 0.f;
        dH1[(r0 + q) * Wd + h] = g1;
      }
      sdh[q * Wd + h] = g0;
      sdh[(kRowTile + q) * Wd + h] = g1;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float acc[kRowTile];
    MGB_UNROLL
    for (int q = 0; q < kRowTile; ++q) acc[q] = 0.f;
    const float* w0 = P + d.focus.W0 + k;
    const float* w1 = P + d.trans.W0 + k;
#pragma unroll 4
    for (int h = 0; h < Wd; ++h) {
      const float a = w0[(long long)h * K], bq = w1[(long long)h * K];
      MGB_UNROLL
      for (int q = 0; q < kRowTile; ++q) acc[q] = fmaf(a, sdh[q * Wd + h], fmaf(bq, sdh[(kRowTile + q) * Wd + h], acc[q]));
    }
    MGB_UNROLL
    for (int q = 0; q < kRowTile; ++q)
      if (s_on[0][q]) dX[(r0 + q) * K + k] += acc[q];
  }
}

// ------------------------------------------------------------------------------------------------------------
// The same two MLPs with the weight matrices resident in shared memory (TMA bulk copies, everything in flight at once)
// and only the rows that matter: tiles of RT rows taken from the compact lists of active rows (focus head) / valid atoms
// (value transform).  grid = (ceil(B*N / RT), 2); CTAs beyond the end of their list leave at once.
// Needs K % 4 == 0, Wd % 4 == 0 and the matrices to fit (rows_mlp_smem_bytes); the generic kernels above are the fallback.
// ------------------------------------------------------------------------------------------------------------
template <int RT>
__host__ __device__ inline size_t rows_mlp_fwd_smem_bytes(int K, int Wd, int No_max) {
  return sizeof(float) * ((size_t)K * Wd + (size_t)Wd * No_max + (size_t)RT * (K + Wd) + 16 + RT);
}
template <int RT>
__global__ void __launch_bounds__(kHeadThreads)
k_rows_mlp_fwd_smem(const CovDesc* __restrict__ dp, const float* __restrict__ P, const float* __restrict__ Wt, int B,
                    const int* __restrict__ act_off, const int* __restrict__ act_list, const int* __restrict__ atom_off,
                    const int* __restrict__ atom_list, const float* __restrict__ X, float* __restrict__ H0, float* __restrict__ Y0,
                    float* __restrict__ H1, float* __restrict__ Y1) {
  const CovDesc& d = *dp;
  const bool focus = blockIdx.y == 0;
  const MlpDesc& M = focus ? d.focus : d.trans;
  const int n_rows = focus ? act_off[B] : atom_off[B];
  const int r0 = blockIdx.x * RT;
  if (r0 >= n_rows) return;
  const int* list = focus ? act_list : atom_list;
  float* H = focus ? H0 : H1;
  float* Y = focus ? Y0 : Y1;
  const int K = M.in, Wd = M.hidden, No = M.out;
  MGB_DYN_SMEM(float, sm);
  SmemBarrier* bar = reinterpret_cast<SmemBarrier*>(sm);     // [2]
  int* s_row = reinterpret_cast<int*>(sm + 16);              // [RT]
  float* sW0 = sm + 16 + RT;                                 // [K][Wd]   (16-byte aligned: RT is a multiple of 4)
  float* sW1 = sW0 + (size_t)K * Wd;                         // [Wd][No]
  float* sx = sW1 + (size_t)Wd * No;                         // [RT][K]
  float* sh = sx + RT * K;                                   // [RT][Wd]
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  if ((int)threadIdx.x < RT) s_row[threadIdx.x] = r0 + (int)threadIdx.x < n_rows ? list[r0 + threadIdx.x] : -1;
  __syncthreads();
  const bool a0 = smem_fill_begin(sW0, Wt + M.W0t, K * Wd, &bar[0]);
  const bool a1 = smem_fill_begin(sW1, Wt + M.W1t, Wd * No, &bar[1]);
  for (int idx = threadIdx.x; idx < RT * (K / 4); idx += blockDim.x) {
    const int q = idx / (K / 4), k4 = idx - q * (K / 4);
    const int row = s_row[q];
    reinterpret_cast<float4*>(sx)[idx] = row >= 0 ? reinterpret_cast<const float4*>(X + (long long)row * K)[k4] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  smem_fill_end(a0, &bar[0], 0);
  for (int o = threadIdx.x; o < Wd; o += blockDim.x) {
    float acc[RT];
    const float bias = P[M.b0 + o];
    MGB_UNROLL
    for (int q = 0; q < RT; ++q) acc[q] = bias;
    for (int k = 0; k < K; k += 4) {
      const float w0 = sW0[(k + 0) * Wd + o], w1 = sW0[(k + 1) * Wd + o], w2 = sW0[(k + 2) * Wd + o], w3 = sW0[(k + 3) * Wd + o];
      MGB_UNROLL
      for (int q = 0; q < RT; ++q) {
        const float4 x = *reinterpret_cast<const float4*>(sx + q * K + k);
        acc[q] = fmaf(w3, x.w, fmaf(w2, x.z, fmaf(w1, x.y, fmaf(w0, x.x, acc[q]))));
      }
    }
    MGB_UNROLL
    for (int q = 0; q < RT; ++q) {
      const float h = fmaxf(acc[q], 0.f);
      sh[q * Wd + o] = h;
      if (s_row[q] >= 0) H[(long long)s_row[q] * Wd + o] = h;
    }
  }
  __syncthreads();
  smem_fill_end(a1, &bar[1], 0);
  if (No >= 32) {
    for (int o = threadIdx.x; o < No; o += blockDim.x) {
      float acc[RT];
      const float bias = P[M.b1 + o];
      MGB_UNROLL
      for (int q = 0; q < RT; ++q) acc[q] = bias;
      for (int k = 0; k < Wd; k += 4) {
        const float w0 = sW1[(k + 0) * No + o], w1 = sW1[(k + 1) * No + o], w2 = sW1[(k + 2) * No + o], w3 = sW1[(k + 3) * No + o];
        MGB_UNROLL
        for (int q = 0; q < RT; ++q) {
          const float4 x = *reinterpret_cast<const float4*>(sh + q * Wd + k);
          acc[q] = fmaf(w3, x.w, fmaf(w2, x.z, fmaf(w1, x.y, fmaf(w0, x.x, acc[q]))));
        }
      }
      MGB_UNROLL
      for (int q = 0; q < RT; ++q)
        if (s_row[q] >= 0) Y[(long long)s_row[q] * No + o] = acc[q];
    }
  } else {
    // few outputs (the focus logit): a warp per (row, output), lanes over k, shuffle reduction
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int pr = warp; pr < RT * No; pr += nwarps) {
      const int q = pr / No, o = pr - q * No;
      float acc = 0.f;
      for (int k = lane; k < Wd; k += 32) acc = fmaf(sW1[k * No + o], sh[q * Wd + k], acc);
      acc = warp_sum(acc);
      if (lane == 0 && s_row[q] >= 0) Y[(long long)s_row[q] * No + o] = acc + P[M.b1 + o];
    }
  }
}

// Backward of the same.  blockIdx.y = 0: focus head on the active rows (dY0 [rows]); 1: value transform on the valid atoms
// (dY = dvf[b,:] for every atom of canvas b, also written to dY1 for the weight gradient).  Both add into dX (atomics).
template <int RT>
__host__ __device__ inline size_t rows_mlp_bwd_smem_bytes(int K, int Wd) {
  return sizeof(float) * ((size_t)K * Wd + (size_t)Wd * Wd + (size_t)RT * 2 * Wd + 16 + RT);
}
template <int RT>
__global__ void __launch_bounds__(kHeadThreads)
k_rows_mlp_bwd_smem(const CovDesc* __restrict__ dp, const float* __restrict__ P, int B, const int* __restrict__ act_off,
                    const int* __restrict__ act_list, const int* __restrict__ atom_off, const int* __restrict__ atom_list,
                    const float* __restrict__ H0, const float* __restrict__ dY0, float* __restrict__ dH0,
                    const float* __restrict__ H1, const float* __restrict__ dvf, float* __restrict__ dY1, float* __restrict__ dH1,
                    float* __restrict__ dX) {
  const CovDesc& d = *dp;
  const bool focus = blockIdx.y == 0;
  const MlpDesc& M = focus ? d.focus : d.trans;
  const int n_rows = focus ? act_off[B] : atom_off[B];
  const int r0 = blockIdx.x * RT;
  if (r0 >= n_rows) return;
  const int* list = focus ? act_list : atom_list;
  const int K = M.in, Wd = M.hidden, N = d.N;
  MGB_DYN_SMEM(float, sm);
  SmemBarrier* bar = reinterpret_cast<SmemBarrier*>(sm);     // [2]
  int* s_row = reinterpret_cast<int*>(sm + 16);              // [RT]
  float* sW0 = sm + 16 + RT;                                 // [Wd][K]   reference layout of W0
  float* sW1 = sW0 + (size_t)K * Wd;                         // [Wd][Wd]  reference layout of W1 (trans only)
  float* sdy = sW1 + (focus ? 0 : (size_t)Wd * Wd);          // [RT][Wd]
  float* sdh = sdy + RT * Wd;                                // [RT][Wd]
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  if ((int)threadIdx.x < RT) s_row[threadIdx.x] = r0 + (int)threadIdx.x < n_rows ? list[r0 + threadIdx.x] : -1;
  __syncthreads();
  const bool a0 = smem_fill_begin(sW0, P + M.W0, K * Wd, &bar[0]);
  bool a1 = false;
  if (!focus) a1 = smem_fill_begin(sW1, P + M.W1, Wd * Wd, &bar[1]);
  if (focus) {
    for (int idx = threadIdx.x; idx < RT * Wd; idx += blockDim.x) {
      const int q = idx / Wd, h = idx - q * Wd;
      const int row = s_row[q];
      float g = 0.f;
      if (row >= 0) {
        g = H0[(long long)row * Wd + h] > 0.f ? P[M.W1 + h] * dY0[row] : 0.f;
        dH0[(long long)row * Wd + h] = g;
      }
      sdh[idx] = g;
    }
  } else {
    for (int idx = threadIdx.x; idx < RT * Wd; idx += blockDim.x) {
      const int q = idx / Wd, o = idx - q * Wd;
      const int row = s_row[q];
      float v = 0.f;
      if (row >= 0) {
        v = dvf[(long long)(row / N) * Wd + o];
        dY1[(long long)row * Wd + o] = v;
      }
      sdy[idx] = v;
    }
    __syncthreads();
    smem_fill_end(a1, &bar[1], 0);
    for (int h = threadIdx.x; h < Wd; h += blockDim.x) {
      float acc[RT];
      MGB_UNROLL
      for (int q = 0; q < RT; ++q) acc[q] = 0.f;
      for (int o = 0; o < Wd; o += 4) {
        const float w0 = sW1[(o + 0) * Wd + h], w1 = sW1[(o + 1) * Wd + h], w2 = sW1[(o + 2) * Wd + h], w3 = sW1[(o + 3) * Wd + h];
        MGB_UNROLL
        for (int q = 0; q < RT; ++q) {
          const float4 g = *reinterpret_cast<const float4*>(sdy + q * Wd + o);
          acc[q] = fmaf(w3, g.w, fmaf(w2, g.z, fmaf(w1, g.y, fmaf(w0, g.x, acc[q]))));
        }
      }
      MGB_UNROLL
      for (int q = 0; q < RT; ++q) {
        const int row = s_row[q];
        float g = 0.f;
        if (row >= 0) {
          g = H1[(long long)row * Wd + h] > 0.f ? acc[q] : 0.f;
          dH1[(long long)row * Wd + h] = g;
        }
        sdh[q * Wd + h] = g;
      }
    }
  }
  __syncthreads();
  smem_fill_end(a0, &bar[0], 0);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float acc[RT];
    MGB_UNROLL
    for (int q = 0; q < RT; ++q) acc[q] = 0.f;
    for (int h = 0; h < Wd; h += 4) {
      const float w0 = sW0[(h + 0) * K + k], w1 = sW0[(h + 1) * K + k], w2 = sW0[(h + 2) * K + k], w3 = sW0[(h + 3) * K + k];
      MGB_UNROLL
      for (int q = 0; q < RT; ++q) {
        const float4 g = *reinterpret_cast<const float4*>(sdh + q * Wd + h);
        acc[q] = fmaf(w3, g.w, fmaf(w2, g.z, fmaf(w1, g.y, fmaf(w0, g.x, acc[q]))));
      }
    }
    MGB_UNROLL
    for (int q = 0; q < RT; ++q)
      if (s_row[q] >= 0 && acc[q] != 0.f) atomicAdd(dX + (long long)s_row[q] * K + k, acc[q]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Grouped weight gradient: for each problem p, dW[o][k] += sum_rows dY[r][o] X[r][k], db[o] += sum_rows dY[r][o].
// grid = (row chunks, work items); a work item is (problem, tile of 32 outputs); threads over k (k == K is the bias
// column), 32 accumulators in registers, dY tile staged in shared memory, X rows streamed coalesced from HBM.
// ------------------------------------------------------------------------------------------------------------
struct DwProblem {
  const float* X;
  const float* dY;
  long long rows;
  int K, No, mode;
  long long dW, db;  // float offsets into the gradient buffer
};
struct DwWork { int prob, o0; };
// the whole problem list travels as a kernel parameter (2.6 KB of the 4 KB parameter space): no staging launch
struct DwProblemList {
  DwProblem p[32];
  DwWork w[96];
  int n, nw;
};
constexpr int kDwThreads = 256;
constexpr int kDwTileO = 32;
constexpr int kDwRowChunk = 32;

__global__ void __launch_bounds__(kDwThreads)
k_dw_grouped(const MGB_GRID_CONSTANT DwProblemList list, const int* __restrict__ n_atoms, int N, float* __restrict__ grad) {
  const DwWork wk = list.w[blockIdx.y];
  const DwProblem pr = list.p[wk.prob];
  const long long per = (pr.rows + gridDim.x - 1) / gridDim.x;
  const long long r_begin = per * blockIdx.x, r_end = (r_begin + per < pr.rows) ? r_begin + per : pr.rows;
  if (r_begin >= r_end) return;
  __shared__ float sdy[kDwRowChunk][kDwTileO];
  __shared__ int s_on[kDwRowChunk];
  const int K = pr.K, No = pr.No, o0 = wk.o0;
  for (int k0 = 0; k0 < K + 1; k0 += blockDim.x) {
    const int k = k0 + threadIdx.x;
    float acc[kDwTileO];
    MGB_UNROLL
    for (int q = 0; q < kDwTileO; ++q) acc[q] = 0.f;
    for (long long rc = r_begin; rc < r_end; rc += kDwRowChunk) {
      const int nr = (int)((r_end - rc) < kDwRowChunk ? (r_end - rc) : kDwRowChunk);
      __syncthreads();
      if ((int)threadIdx.x < nr) s_on[threadIdx.x] = row_on(pr.mode, n_atoms, N, rc + threadIdx.x) ? 1 : 0;
      __syncthreads();
      for (int idx = threadIdx.x; idx < nr * kDwTileO; idx += blockDim.x) {
        const int q = idx / kDwTileO, o = o0 + idx % kDwTileO;
        sdy[q][idx % kDwTileO] = (s_on[q] && o < No) ? pr.dY[(rc + q) * No + o] : 0.f;
      }
      __syncthreads();
      if (k <= K) {
#pragma unroll 4
        for (int q = 0; q < nr; ++q) {
          const float x = !s_on[q] ? 0.f : (k < K ? pr.X[(rc + q) * K + k] : 1.f);
          MGB_UNROLL
          for (int o = 0; o < kDwTileO; ++o) acc[o] = fmaf(sdy[q][o], x, acc[o]);
        }
      }
    }
    if (k <= K) {
      MGB_UNROLL
      for (int o = 0; o < kDwTileO; ++o) {
        if (o0 + o < No && acc[o] != 0.f) {
          if (k < K) atomicAdd(grad + pr.dW + (long long)(o0 + o) * K + k, acc[o]);
          else if (pr.db >= 0) atomicAdd(grad + pr.db + o0 + o, acc[o]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Per-canvas policy head.  Shared-memory scratch carved out by PolicySmem.
// ------------------------------------------------------------------------------------------------------------
struct PolicySmem {
  float* fl;      // [N] focus logits -> probabilities p' (normalised twice like Categorical(probs=...))
  float* flog;    // [N] clamped log-probabilities
  float* finv;    // [lat]
  float* he;      // [Wd]
  float* el;      // [Z] element logits -> probabilities
  float* elog;    // [Z]
  float2* ecov;   // [25][CPE]
  float* einv;    // [latE]
  float* hd;      // [Wd]
  float* yd;      // [2G]
  float2* cat;    // [totM]
  float2* cond;   // [25][CPE]
  float2* alm;    // [25] normalised sum over channels
  float* vf;      // [Wd]
  float* hv;      // [Wd]
  float* red;     // [64] reduction scratch
  float* misc;    // [32]
};
__host__ __device__ inline int policy_smem_floats(const CovDesc& d) {
  return 2 * d.N + d.lat + d.Wd + 2 * d.Z + 2 * kM * d.CPE + d.latE + d.Wd + 2 * d.G + 2 * d.totM + 2 * kM * d.CPE + 2 * kM +
         2 * d.Wd + 64 + 32 + 16;
}
// saved per canvas by k_policy_fwd for k_policy_bwd: the PolicySmem block followed by PolicyScalars (padded to 16 bytes)
#define MGB_POLICY_SCALARS_FLOATS 32
__host__ __device__ inline int policy_state_floats(const CovDesc& d) { return ((policy_smem_floats(d) + 3) & ~3) + MGB_POLICY_SCALARS_FLOATS + 4; }
__device__ __forceinline__ PolicySmem policy_smem_carve(const CovDesc& d, float* base) {
  PolicySmem s;
  float* p = base;
  s.cat = reinterpret_cast<float2*>(p); p += 2 * d.totM;
  s.ecov = reinterpret_cast<float2*>(p); p += 2 * kM * d.CPE;
  s.cond = reinterpret_cast<float2*>(p); p += 2 * kM * d.CPE;
  s.alm = reinterpret_cast<float2*>(p); p += 2 * kM;
  s.fl = p; p += d.N;
  s.flog = p; p += d.N;
  s.finv = p; p += d.lat;
  s.he = p; p += d.Wd;
  s.el = p; p += d.Z;
  s.elog = p; p += d.Z;
  s.einv = p; p += d.latE;
  s.hd = p; p += d.Wd;
  s.yd = p; p += 2 * d.G;
  s.vf = p; p += d.Wd;
  s.hv = p; p += d.Wd;
  s.red = p; p += 64;
  s.misc = p;
  return s;
}

// masked softmax + Categorical(probs=...) arithmetic on a short vector held in shared memory (single thread; the
// vectors have <= 40 entries).  logits -> probs p' in place, clamped log-probs in `logp`.  Returns entropy.
//   q = exp(z - max) / (sum + 1e-12) * mask   (torch_scatter scatter_softmax + modules.py:27)
//   p' = q / sum(q) ;  L = log(clamp(p', eps, 1 - eps)) ; entropy = -sum p' L
struct SoftmaxAux { float mx, s1, s2; };
__device__ inline float categorical_fwd(float* z, float* logp, const bool* mask, int n, SoftmaxAux* aux) {
  float mx = -3.4028234663852886e38f;
  for (int i = 0; i < n; ++i) if (mask[i]) mx = fmaxf(mx, z[i]);
  float s1 = 0.f;
  for (int i = 0; i < n; ++i) { z[i] = mask[i] ? expf(z[i] - mx) : 0.f; s1 += z[i]; }
  s1 += 1e-12f;
  float s2 = 0.f;
  for (int i = 0; i < n; ++i) { z[i] = z[i] / s1; s2 += z[i]; }
  float ent = 0.f;
  for (int i = 0; i < n; ++i) {
    z[i] = z[i] / s2;
    logp[i] = logf(fminf(fmaxf(z[i], kF32Eps), 1.f - kF32Eps));
    ent -= z[i] * logp[i];
  }
  if (aux) { aux->mx = mx; aux->s1 = s1; aux->s2 = s2; }
  return ent;
}
// Backward of categorical_fwd: given g_logp (cotangent of L[pick]) and g_ent, p' and L as produced above,
// writes dz (cotangent of the raw logits) into dz[].
__device__ inline void categorical_bwd(const float* p, const float* logp, const bool* mask, int n, int pick, float g_logp,
                                       float g_ent, const SoftmaxAux& aux, float* dz) {
  // dL/dp'_i
  float dotq = 0.f;
  for (int i = 0; i < n; ++i) {
    const bool pass = p[i] >= kF32Eps && p[i] <= 1.f - kF32Eps;  // clamp passes gradient inside [min, max]
    float dL = (i == pick ? g_logp : 0.f) - g_ent * p[i];         // through L_i
    float dp = (pass ? dL / p[i] : 0.f) - g_ent * logp[i];        // + direct -g_ent * L_i
    dz[i] = dp;
  }
  // p' = q / s2  ->  dq_i = (dp_i - sum_j dp_j p'_j) / s2
  for (int i = 0; i < n; ++i) dotq += dz[i] * p[i];
  for (int i = 0; i < n; ++i) dz[i] = (dz[i] - dotq) / aux.s2;
  // q = e / s1 * mask, s1 = sum e + 1e-12 -> de_i = mask_i dq_i / s1 - sum_j (mask_j dq_j q_j) / s1 ; e = exp(z - mx)
  float dots = 0.f;
  for (int i = 0; i < n; ++i) {
    if (!mask[i]) dz[i] = 0.f;
    dots += dz[i] * (p[i] * aux.s2);  // q_i = p'_i * s2
  }
  for (int i = 0; i < n; ++i) {
    const float q = p[i] * aux.s2, e = q * aux.s1;
    dz[i] = mask[i] ? (dz[i] / aux.s1 - dots / aux.s1) * e : 0.f;
  }
}

// y[o] = b[o] + sum_k W[o][k] x[k]  (reference layout [n_out][K]); a warp owns 4 outputs at a time, lanes stride k
// (coalesced), shuffle reduction.  x in shared memory.  All threads of the block must call.
__device__ __forceinline__ void gemv_rows(const float* __restrict__ W, const float* __restrict__ bias, const float* x, int K,
                                          int n_out, bool relu, float* y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int o0 = warp * 4; o0 < n_out; o0 += nwarps * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* w = W + (long long)o0 * K;
    const int no = min(4, n_out - o0);
    // two k-steps x four outputs = eight independent loads in flight per lane (rows beyond n_out are read clamped)
    for (int k = lane; k < K; k += 64) {
      const int k1 = k + 32 < K ? k + 32 : k;
      float wv[2][4];
      MGB_UNROLL
      for (int q = 0; q < 4; ++q) {
        const long long row = (long long)(q < no ? q : no - 1) * K;
        wv[0][q] = w[row + k];
        wv[1][q] = w[row + k1];
      }
      const float x0 = x[k], x1 = k + 32 < K ? x[k1] : 0.f;
      MGB_UNROLL
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(wv[1][q], x1, fmaf(wv[0][q], x0, acc[q]));
    }
    MGB_UNROLL
    for (int q = 0; q < 4; ++q) {
      const float v = warp_sum(acc[q]);
      if (lane == 0 && q < no) {
        const float r = v + bias[o0 + q];
        y[o0 + q] = relu ? fmaxf(r, 0.f) : r;
      }
    }
  }
}

struct PolicyScalars {
  int n, focus, element;
  int focus_valid;
  float dist;
  float ent_f, ent_e, logp_f, logp_e, logp_d, logp_o, v;
  float k_raw, inv_sqrt_k, log_z, lse_max, lse_sum;  // spherical
  float2 s_o;                                          // s(orientation)
  SoftmaxAux aux_f, aux_e;
};
static_assert(sizeof(PolicyScalars) <= sizeof(float) * MGB_POLICY_SCALARS_FLOATS, "PolicyScalars outgrew its slot in the saved state");

// s(x) = sum_lm a_lm Y_lm(x)
__device__ __forceinline__ float2 sph_sum(const float2* a, const float2* y) {
  float2 s = make_float2(0.f, 0.f);
  MGB_UNROLL
  for (int q = 0; q < kM; ++q) cfma(s, a[q], y[q]);
  return s;
}
// same with the Lebedev table layout [25][n_grid] (lanes over grid points read coalesced)
__device__ __forceinline__ float2 sph_sum_grid(const float2* a, const float2* __restrict__ y, int stride) {
  float2 yv[kM];
  MGB_UNROLL
  for (int q = 0; q < kM; ++q) yv[q] = y[(long long)q * stride];
  float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
  MGB_UNROLL
  for (int q = 0; q < kM; ++q) {
    if (q & 1) cfma(s1, a[q], yv[q]); else cfma(s0, a[q], yv[q]);
  }
  return make_float2(s0.x + s1.x, s0.y + s1.y);
}

// Mixer cat vector (covariant/modules.py:180-190): ag = d * ecov ; sq = CG(ag x ag) ; cat_l = [ag | sq | ecov]
__device__ __forceinline__ void mixer_build_cat(const CovDesc& d, const float2* ecov, float dist, float2* scratch_ag, float2* cat) {
  const int CPE = d.CPE;
  for (int idx = threadIdx.x; idx < kM * CPE; idx += blockDim.x) {
    const int lm = idx / CPE, c = idx % CPE, l = ell_of_lm(lm);
    const float2 e = ecov[idx];
    const float2 ag = make_float2(dist * e.x, dist * e.y);
    scratch_ag[idx] = ag;
    const int base = d.offM[l] + (lm - l * l) * d.catM[l];
    cat[base + c] = ag;
    cat[base + d.inM_block[l] * CPE + c] = e;
  }
  __syncthreads();
  cg_gather<true>(d.mix_sq, CPE, scratch_ag, cat);
  __syncthreads();
}

// Everything the reference's step() does after the Cormorant body, for canvas b.  Leaves intermediates in `s`.
__device__ inline void policy_forward(const CovDesc& d, const float* __restrict__ P, const float* __restrict__ Wt, int b,
                                      const int* __restrict__ n_atoms, const float* __restrict__ bags,
                                      const float* __restrict__ actions, const float* __restrict__ A_last,
                                      const float* __restrict__ inv, const float* __restrict__ flogit,
                                      const float* __restrict__ trans, PolicySmem& s, PolicyScalars& ps,
                                      const float2* __restrict__ lse_saved) {
  const int N = d.N, Z = d.Z, CPE = d.CPE, Wd = d.Wd, G = d.G, tau = d.Cout;
  const int n = n_atoms[b];
  const float* act = actions + (long long)b * 6;
  ps.n = n;
  ps.focus = (int)rintf(act[0]);
  ps.element = (int)rintf(act[1]);
  ps.dist = act[2];
  ps.focus_valid = ps.focus < n;
  const int nact = n > 1 ? n : 1;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s.fl[i] = i < nact ? flogit[(long long)b * N + i] : 0.f;
  for (int k = threadIdx.x; k < d.lat; k += blockDim.x) s.finv[k] = inv[((long long)b * N + ps.focus) * d.lat + k];
  for (int idx = threadIdx.x; idx < kM * CPE; idx += blockDim.x) {
    const int lm = idx / CPE, c = idx % CPE;
    float2 v = make_float2(0.f, 0.f);
    if (ps.focus_valid)
      v = reinterpret_cast<const float2*>(A_last)[(((long long)b * N + ps.focus) * kM + lm) * tau + ps.element * CPE + c];
    s.ecov[idx] = v;
  }
  for (int h = threadIdx.x; h < Wd; h += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += trans[((long long)b * N + i) * Wd + h];
    s.vf[h] = acc;
  }
  __syncthreads();
  // --- focus (agent.py:223-240)
  if (threadIdx.x == 0) {
    bool mask[64];
    for (int i = 0; i < N; ++i) mask[i] = i < nact;
    ps.ent_f = categorical_fwd(s.fl, s.flog, mask, N, &ps.aux_f);
    s.misc[0] = ps.ent_f; s.misc[1] = ps.aux_f.mx; s.misc[2] = ps.aux_f.s1; s.misc[3] = ps.aux_f.s2;
  }
  // --- element (agent.py:243-259)
  gemv_rows(P + d.element.W0, P + d.element.b0, s.finv, d.lat, Wd, true, s.he);
  atomic_scalars_row(s.ecov, CPE, CPE, s.einv, threadIdx.x, blockDim.x);
  __syncthreads();
  gemv_rows(P + d.element.W1, P + d.element.b1, s.he, Wd, Z, false, s.el);
  // --- distance (agent.py:263-276)
  gemv_rows(P + d.dist.W0, P + d.dist.b0, s.einv, d.latE, Wd, true, s.hd);
  // --- value (agent.py:313-316)
  gemv_rows(P + d.value.W0, P + d.value.b0, s.vf, Wd, Wd, true, s.hv);
  __syncthreads();
  gemv_rows(P + d.dist.W1, P + d.dist.b1, s.hd, Wd, 2 * G, false, s.yd);
  if (threadIdx.x == 0) {
    bool mask[MGB_MAX_SPECIES];
    for (int z = 0; z < Z; ++z) mask[z] = bags[(long long)b * Z + z] > 0.f;
    ps.ent_e = categorical_fwd(s.el, s.elog, mask, Z, &ps.aux_e);
    s.misc[4] = ps.ent_e; s.misc[5] = ps.aux_e.mx; s.misc[6] = ps.aux_e.s1; s.misc[7] = ps.aux_e.s2;
  }
  {  // v = b1 + W1 hv  (block reduction)
    float part = 0.f;
    for (int h = threadIdx.x; h < Wd; h += blockDim.x) part += P[d.value.W1 + h] * s.hv[h];
    ps.v = block_sum(part, s.red) + P[d.value.b1];
  }
  __syncthreads();
  ps.ent_f = s.misc[0]; ps.aux_f.mx = s.misc[1]; ps.aux_f.s1 = s.misc[2]; ps.aux_f.s2 = s.misc[3];
  ps.ent_e = s.misc[4]; ps.aux_e.mx = s.misc[5]; ps.aux_e.s1 = s.misc[6]; ps.aux_e.s2 = s.misc[7];
  ps.logp_f = s.flog[ps.focus];
  ps.logp_e = s.elog[ps.element];
  {  // GMM log-prob (gmm.py:8-18): every thread computes the same few numbers
    float lse_g = -3.0e38f;
    for (int k = 0; k < G; ++k) lse_g = fmaxf(lse_g, s.yd[k]);
    float sg = 0.f;
    for (int k = 0; k < G; ++k) sg += expf(s.yd[k] - lse_g);
    lse_g += logf(sg);
    const float hw = 0.5f * (d.dmax - d.dmin), ctr = 0.5f * (d.dmin + d.dmax);
    float t[8], tm = -3.0e38f;
    for (int k = 0; k < G; ++k) {
      const float mu = tanhf(s.yd[G + k]) * hw + ctr;
      const float sd = fmaxf(expf(P[d.p_logstd + k]), 1e-6f);
      const float df = ps.dist - mu;
      t[k] = -(df * df) / (2.f * sd * sd) - logf(sd) - kLogSqrt2Pi + (s.yd[k] - lse_g);
      tm = fmaxf(tm, t[k]);
    }
    float st = 0.f;
    for (int k = 0; k < G; ++k) st += expf(t[k] - tm);
    ps.logp_d = tm + logf(st);
  }
  // --- condition on distance + spherical distribution (agent.py:279-292)
  mixer_build_cat(d, s.ecov, ps.dist, s.cond /* scratch for ag */, s.cat);
  mix_rows<4, 2>(d.units_hidden, d.n_units_hidden, d.catM, d.offM, d.offWM, CPE,
                 reinterpret_cast<const float2*>(P + d.p_mixW), s.cat, s.cond);
  __syncthreads();
  if (threadIdx.x < kM) {
    float2 a = make_float2(0.f, 0.f);
    for (int c = 0; c < CPE; ++c) { a.x += s.cond[threadIdx.x * CPE + c].x; a.y += s.cond[threadIdx.x * CPE + c].y; }
    s.alm[threadIdx.x] = a;
  }
  __syncthreads();
  float k_raw = 0.f;
  for (int q = 0; q < kM; ++q) k_raw += s.alm[q].x * s.alm[q].x + s.alm[q].y * s.alm[q].y;
  ps.k_raw = k_raw;
  ps.inv_sqrt_k = 1.f / sqrtf(fmaxf(k_raw, 1e-10f));
  __syncthreads();
  if (threadIdx.x < kM) {
    s.alm[threadIdx.x].x *= ps.inv_sqrt_k;
    s.alm[threadIdx.x].y *= ps.inv_sqrt_k;
  }
  __syncthreads();
  float2 a_loc[kM];
  MGB_UNROLL
  for (int q = 0; q < kM; ++q) a_loc[q] = s.alm[q];
  {
    float ox = act[3], oy = act[4], oz = act[5];
    const float nr = sqrtf(ox * ox + oy * oy + oz * oz);
    if (nr > 0.f) { ox /= nr; oy /= nr; oz /= nr; } else { ox = oy = oz = 0.f; }
    float2 y[kM];
    sph_harm_l4(ox, oy, oz, false, false, y);
    ps.s_o = sph_sum(a_loc, y);
  }
  const float so2 = ps.s_o.x * ps.s_o.x + ps.s_o.y * ps.s_o.y;
  if (d.has_beta) {
    // log Z = log 4 pi + logsumexp_g(-beta |s(x_g)|^2 + log w_g)   (spherical_dists.py:208-215)
    float mx, sum;
    if (lse_saved) {
      mx = lse_saved[b].x;
      sum = lse_saved[b].y;
    } else {
      float m_run = -3.0e38f, s_run = 0.f;   // online logsumexp over this thread's grid points
      for (int g = threadIdx.x; g < d.n_grid; g += blockDim.x) {
        const float2 sg = sph_sum_grid(a_loc, reinterpret_cast<const float2*>(d.leb_y) + g, d.n_grid);
        const float t = -d.beta * (sg.x * sg.x + sg.y * sg.y) + d.leb_logw[g];
        if (t > m_run) { s_run = s_run * expf(m_run - t) + 1.f; m_run = t; } else { s_run += expf(t - m_run); }
      }
      mx = block_max(m_run, s.red);
      sum = block_sum(s_run * expf(m_run - mx), s.red);
    }
    ps.lse_max = mx;
    ps.lse_sum = sum;
    ps.log_z = kLog4Pi + mx + logf(sum);
    ps.logp_o = -d.beta * so2 - ps.log_z;
  } else {
    ps.log_z = 0.f;
    const float p = (n == 0) ? (1.f / kFourPi) : so2;   // SO3Distribution.prob with the `empty` override
    ps.logp_o = logf(fmaxf(p, 1e-10f));
  }
}

// optional outputs of a canvas (probabilities, mixture parameters, normalised coefficients) from the shared-memory state
__device__ __forceinline__ void policy_write_extras(const CovDesc& d, const float* __restrict__ P, int b, const PolicySmem& s,
                                                    const PolicyScalars& ps, const mgb_cov_outputs& out) {
    if (out.focus_probs) for (int i = threadIdx.x; i < d.N; i += blockDim.x) out.focus_probs[(long long)b * d.N + i] = s.fl[i];
  if (out.element_probs) for (int z = threadIdx.x; z < d.Z; z += blockDim.x) out.element_probs[(long long)b * d.Z + z] = s.el[z];
  if (out.coefficients)
    for (int idx = threadIdx.x; idx < kM * d.CPE; idx += blockDim.x) {
      reinterpret_cast<float2*>(out.coefficients)[(long long)b * kM * d.CPE + idx] =
          make_float2(s.cond[idx].x * ps.inv_sqrt_k, s.cond[idx].y * ps.inv_sqrt_k);
    }
  if (out.gmm && threadIdx.x == 0) {
    const int G = d.G;
    float lse = -3.0e38f;
    for (int k = 0; k < G; ++k) lse = fmaxf(lse, s.yd[k]);
    float sg = 0.f;
    for (int k = 0; k < G; ++k) sg += expf(s.yd[k] - lse);
    lse += logf(sg);
    for (int k = 0; k < G; ++k) {
      out.gmm[((long long)b * 3 + 0) * G + k] = s.yd[k] - lse;
      out.gmm[((long long)b * 3 + 1) * G + k] = tanhf(s.yd[G + k]) * 0.5f * (d.dmax - d.dmin) + 0.5f * (d.dmin + d.dmax);
      out.gmm[((long long)b * 3 + 2) * G + k] = fmaxf(expf(P[d.p_logstd + k]), 1e-6f);
    }
  }
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(kPolicyThreads, MIN_CTAS)
k_policy_fwd(const CovDesc* __restrict__ dp, const float* __restrict__ P, const float* __restrict__ Wt, int B,
             const int* __restrict__ n_atoms, const float* __restrict__ bags, const float* __restrict__ actions,
             const float* __restrict__ A_last, const float* __restrict__ inv, const float* __restrict__ flogit,
             const float* __restrict__ trans, float2* __restrict__ lse_out, float* __restrict__ state, mgb_cov_outputs out) {
  const CovDesc& d = *dp;
  MGB_DYN_SMEM(float, sm);
  PolicySmem s = policy_smem_carve(d, sm);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    PolicyScalars ps;
    __syncthreads();
    policy_forward(d, P, Wt, b, n_atoms, bags, actions, A_last, inv, flogit, trans, s, ps, nullptr);
    if (state) {   // everything the backward needs: the shared-memory intermediates and the per-canvas scalars
      __syncthreads();
      float* dst = state + (long long)b * policy_state_floats(d);
      const int nf = policy_smem_floats(d);
      for (int idx = threadIdx.x; idx < nf; idx += blockDim.x) dst[idx] = sm[idx];
      if (threadIdx.x == 0) *reinterpret_cast<PolicyScalars*>(dst + ((nf + 3) & ~3)) = ps;
    }
    if (threadIdx.x == 0) {
      lse_out[b] = make_float2(ps.lse_max, ps.lse_sum);
      out.logp[b] = ((ps.logp_f + ps.logp_e) + ps.logp_d) + ps.logp_o;
      out.ent[b] = ps.ent_f + ps.ent_e;
      out.v[b] = ps.v;
      if (out.logp_parts) {
        out.logp_parts[b * 4 + 0] = ps.logp_f; out.logp_parts[b * 4 + 1] = ps.logp_e;
        out.logp_parts[b * 4 + 2] = ps.logp_d; out.logp_parts[b * 4 + 3] = ps.logp_o;
      }
      if (out.log_z) out.log_z[b] = ps.log_z;
    }
    policy_write_extras(d, P, b, s, ps, out);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Rollout mode (actions=None, agent.py:229-292): the four sub-actions are drawn on the device, one CTA per canvas, then the
// chosen action is evaluated by policy_forward — the same code an evaluate-mode step() runs, so its logp is reproduced exactly
// when the stored action is re-evaluated by ppo.compute_loss.
//   mode 1 (self.training): Categorical samples for focus / element, a mixture sample for the distance (clamped at 0.001),
//                           rejection sampling against the uniform proposal for the orientation (spherical_dists.py:116-150,227-262)
//   mode 2 (greedy):        argmax of the categoricals, best of 128 mixture samples (gmm.py:20-27), best of 128 (beta) / 256
//                           accepted orientation samples (spherical_dists.py:152-158,264-271)
// Random numbers: Philox4x32-10 keyed by the seed the host draws from torch's generator, counter = (canvas, thread, draw, stream).
// ------------------------------------------------------------------------------------------------------------
struct Philox {
  unsigned k0, k1, c0, c1, c2, c3;
  __device__ Philox(unsigned long long seed, unsigned canvas, unsigned thread, unsigned stream)
      : k0((unsigned)seed), k1((unsigned)(seed >> 32)), c0(canvas), c1(thread), c2(0u), c3(stream) {}
  __device__ static void mulhilo(unsigned a, unsigned b, unsigned& hi, unsigned& lo) {
    const unsigned long long p = (unsigned long long)a * b;
    hi = (unsigned)(p >> 32); lo = (unsigned)p;
  }
  // four uniforms in (0, 1); every call advances the draw counter
  __device__ void next4(float* u) {
    unsigned x0 = c0, x1 = c1, x2 = c2, x3 = c3, a = k0, b = k1;
    for (int r = 0; r < 10; ++r) {
      unsigned hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, x0, hi0, lo0);
      mulhilo(0xCD9E8D57u, x2, hi1, lo1);
      const unsigned y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
      x0 = y0; x1 = y1; x2 = y2; x3 = y3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    ++c2;
    const unsigned x[4] = {x0, x1, x2, x3};
    for (int q = 0; q < 4; ++q) u[q] = ((float)(x[q] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  }
};

// index drawn from the probabilities p[0..n) by inversion (single thread)
__device__ inline int categorical_draw(const float* p, int n, float u) {
  float cum = 0.f;
  int last = 0;
  for (int i = 0; i < n; ++i) {
    if (p[i] > 0.f) {
      cum += p[i];
      last = i;
      if (u < cum) return i;
    }
  }
  return last;
}
__device__ inline int argmax_first(const float* p, int n) {
  int best = 0;
  for (int i = 1; i < n; ++i) if (p[i] > p[best]) best = i;
  return best;
}

// block-wide (max value, smallest index attaining it); scratch >= 64 floats
__device__ __forceinline__ void block_argmax(float& val, int& idx, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, val, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
  }
  __syncthreads();
  if (lane == 0) { scratch[w] = val; reinterpret_cast<int*>(scratch)[32 + w] = idx; }
  __syncthreads();
  float v = lane < nw ? scratch[lane] : -3.0e38f;
  int i = lane < nw ? reinterpret_cast<int*>(scratch)[32 + lane] : 0x7fffffff;
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
  val = v; idx = i;
  __syncthreads();
}

__device__ __forceinline__ void sphere_point(float u1, float u2, float* x) {   // spherical_dists.py:49-61
  const float ct = 1.f - 2.f * u1, st = sqrtf(fmaxf(0.f, 1.f - ct * ct)), phi = kTwoPi * u2;
  x[0] = st * cosf(phi); x[1] = st * sinf(phi); x[2] = ct;
}
__device__ __forceinline__ void fibonacci_point(int i, int n, float* x) {      // so3_tools.py:8-20
  const float ct = 1.f - 2.f * ((float)i + 0.5f) / (float)n, st = sqrtf(fmaxf(0.f, 1.f - ct * ct));
  const float phi = kTwoPi * (float)i / 1.618033988749895f;
  x[0] = st * cosf(phi); x[1] = st * sinf(phi); x[2] = ct;
}
// |s(x)|^2 with s = sum_lm a_lm Y_lm(x) ('qm' harmonics of the unit vector x)
__device__ __forceinline__ float s2_of(const float2* a_loc, const float* x) {
  float2 y[kM];
  sph_harm_l4(x[0], x[1], x[2], false, false, y);
  const float2 sv = sph_sum(a_loc, y);
  return sv.x * sv.x + sv.y * sv.y;
}

__global__ void __launch_bounds__(kPolicyThreads)
k_policy_sample(const CovDesc* __restrict__ dp, const float* __restrict__ P, const float* __restrict__ Wt, int B,
                const int* __restrict__ n_atoms, const float* __restrict__ bags, const float* __restrict__ A_last,
                const float* __restrict__ inv, const float* __restrict__ flogit, const float* __restrict__ trans, int mode,
                unsigned long long seed, float* __restrict__ actions, float2* __restrict__ lse_out, mgb_cov_outputs out) {
  const CovDesc& d = *dp;
  MGB_DYN_SMEM(float, sm);
  PolicySmem s = policy_smem_carve(d, sm);
  const int N = d.N, Z = d.Z, CPE = d.CPE, Wd = d.Wd, G = d.G, tau = d.Cout;
  const bool greedy = mode == 2;
  __shared__ int s_pick[2];
  __shared__ float s_val[4];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const int n = n_atoms[b], nact = n > 1 ? n : 1;
    float* act = actions + (long long)b * 6;
    Philox rng(seed, (unsigned)b, threadIdx.x, 0u);
    float u4[4];
    // ---- focus (agent.py:223-234)
    for (int i = threadIdx.x; i < N; i += blockDim.x) s.fl[i] = i < nact ? flogit[(long long)b * N + i] : 0.f;
    __syncthreads();
    if (threadIdx.x == 0) {
      bool mask[64];
      for (int i = 0; i < N; ++i) mask[i] = i < nact;
      categorical_fwd(s.fl, s.flog, mask, N, nullptr);
      rng.next4(u4);
      s_pick[0] = greedy ? argmax_first(s.fl, N) : categorical_draw(s.fl, N, u4[0]);
    }
    __syncthreads();
    const int focus = s_pick[0];
    const bool focus_valid = focus < n;
    // ---- element (agent.py:243-253)
    for (int k = threadIdx.x; k < d.lat; k += blockDim.x) s.finv[k] = inv[((long long)b * N + focus) * d.lat + k];
    __syncthreads();
    gemv_rows(P + d.element.W0, P + d.element.b0, s.finv, d.lat, Wd, true, s.he);
    __syncthreads();
    gemv_rows(P + d.element.W1, P + d.element.b1, s.he, Wd, Z, false, s.el);
    __syncthreads();
    if (threadIdx.x == 0) {
      bool mask[MGB_MAX_SPECIES];
      for (int z = 0; z < Z; ++z) mask[z] = bags[(long long)b * Z + z] > 0.f;
      categorical_fwd(s.el, s.elog, mask, Z, nullptr);
      rng.next4(u4);
      s_pick[1] = greedy ? argmax_first(s.el, Z) : categorical_draw(s.el, Z, u4[0]);
    }
    __syncthreads();
    const int element = s_pick[1];
    // ---- distance (agent.py:256-276)
    for (int idx = threadIdx.x; idx < kM * CPE; idx += blockDim.x) {
      const int lm = idx / CPE, c = idx % CPE;
      float2 v = make_float2(0.f, 0.f);
      if (focus_valid) v = reinterpret_cast<const float2*>(A_last)[(((long long)b * N + focus) * kM + lm) * tau + element * CPE + c];
      s.ecov[idx] = v;
    }
    __syncthreads();
    atomic_scalars_row(s.ecov, CPE, CPE, s.einv, threadIdx.x, blockDim.x);
    __syncthreads();
    gemv_rows(P + d.dist.W0, P + d.dist.b0, s.einv, d.latE, Wd, true, s.hd);
    __syncthreads();
    gemv_rows(P + d.dist.W1, P + d.dist.b1, s.hd, Wd, 2 * G, false, s.yd);
    __syncthreads();
    {
      // mixture parameters (gmm.py:8-18): every thread holds them
      float lw[8], mu[8], sd[8], lse = -3.0e38f;
      for (int k = 0; k < G; ++k) lse = fmaxf(lse, s.yd[k]);
      float sg = 0.f;
      for (int k = 0; k < G; ++k) sg += expf(s.yd[k] - lse);
      lse += logf(sg);
      const float hw = 0.5f * (d.dmax - d.dmin), ctr = 0.5f * (d.dmin + d.dmax);
      for (int k = 0; k < G; ++k) {
        lw[k] = s.yd[k] - lse;
        mu[k] = tanhf(s.yd[G + k]) * hw + ctr;
        sd[k] = fmaxf(expf(P[d.p_logstd + k]), 1e-6f);
      }
      // one mixture sample per thread: component by inversion, Box-Muller normal
      rng.next4(u4);
      int comp = G - 1;
      float cum = 0.f;
      for (int k = 0; k < G; ++k) { cum += expf(lw[k]); if (u4[0] < cum) { comp = k; break; } }
      const float z = sqrtf(-2.f * logf(u4[1])) * cosf(kTwoPi * u4[2]);
      const float sample = mu[comp] + sd[comp] * z;
      if (!greedy) {
        if (threadIdx.x == 0) s_val[0] = fmaxf(sample, 0.001f);   // agent.py:273: ensure the sampled distance is > 0
      } else {
        float tm = -3.0e38f, t[8];
        for (int k = 0; k < G; ++k) {
          const float df = sample - mu[k];
          t[k] = -(df * df) / (2.f * sd[k] * sd[k]) - logf(sd[k]) - kLogSqrt2Pi + lw[k];
          tm = fmaxf(tm, t[k]);
        }
        float st = 0.f;
        for (int k = 0; k < G; ++k) st += expf(t[k] - tm);
        float score = (int)threadIdx.x < 128 ? tm + logf(st) : -3.0e38f;   // best of 128 samples (gmm.py:20-27)
        int who = threadIdx.x;
        block_argmax(score, who, s.red);
        if ((int)threadIdx.x == who) s_val[0] = sample;
      }
    }
    __syncthreads();
    const float dist = s_val[0];
    // ---- condition on the distance and normalise (agent.py:279-284, so3_tools.py:73-79)
    mixer_build_cat(d, s.ecov, dist, s.cond /* scratch for ag */, s.cat);
    mix_rows<4, 2>(d.units_hidden, d.n_units_hidden, d.catM, d.offM, d.offWM, CPE, reinterpret_cast<const float2*>(P + d.p_mixW), s.cat, s.cond);
    __syncthreads();
    if (threadIdx.x < kM) {
      float2 a = make_float2(0.f, 0.f);
      for (int c = 0; c < CPE; ++c) { a.x += s.cond[threadIdx.x * CPE + c].x; a.y += s.cond[threadIdx.x * CPE + c].y; }
      s.alm[threadIdx.x] = a;
    }
    __syncthreads();
    float k_raw = 0.f;
    for (int q = 0; q < kM; ++q) k_raw += s.alm[q].x * s.alm[q].x + s.alm[q].y * s.alm[q].y;
    const float inv_sqrt_k = 1.f / sqrtf(fmaxf(k_raw, 1e-10f));
    float2 a_loc[kM];
    MGB_UNROLL
    for (int q = 0; q < kM; ++q) a_loc[q] = make_float2(s.alm[q].x * inv_sqrt_k, s.alm[q].y * inv_sqrt_k);
    __syncthreads();
    // ---- orientation (agent.py:284-292): log p(x) up to the constant log Z, which cancels in the acceptance ratio
    //   beta: log p = -beta |s|^2;  beta None: p = |s|^2 (1 / 4 pi on the empty canvas)
    const bool empty = n == 0;
    auto score_of = [&](const float* x) {
      const float s2 = s2_of(a_loc, x);
      if (d.has_beta) return -d.beta * s2;
      return empty ? 1.f / kFourPi : s2;
    };
    float gmax = -3.0e38f;   // maximum of the score over the Fibonacci grid (4096 points with beta, else 1024)
    {
      const int ngrid = d.has_beta ? 4096 : 1024;
      for (int g = threadIdx.x; g < ngrid; g += blockDim.x) {
        float x[3];
        fibonacci_point(g, ngrid, x);
        gmax = fmaxf(gmax, score_of(x));
      }
      gmax = block_max(gmax, s.red);
    }
    const int want = greedy ? (d.has_beta ? 128 : 256) : 1;
    int have = 0;
    float best = -3.0e38f, best_x[3] = {0.f, 0.f, 1.f};
    int best_rank = 0x7fffffff;
    for (int round = 0; round < 4096 && have < want; ++round) {
      rng.next4(u4);
      float x[3];
      sphere_point(u4[0], u4[1], x);
      const float sc = score_of(x);
      // acceptance: u < p(x) / (M * uniform) with M = max_grid p / uniform  ->  u < p(x) / max_grid p
      const float thr = d.has_beta ? expf(sc - gmax) : (gmax > 0.f ? sc / gmax : 1.f);
      const bool acc = u4[2] < thr;
      // rank of this thread's candidate among the accepted ones of the round (candidate order = thread order)
      const unsigned bal = __ballot_sync(0xffffffffu, acc);
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
      __syncthreads();
      if (lane == 0) reinterpret_cast<int*>(s.red)[w] = __popc(bal);
      __syncthreads();
      int before = 0, total = 0;
      for (int q = 0; q < nw; ++q) { const int cq = reinterpret_cast<int*>(s.red)[q]; if (q < w) before += cq; total += cq; }
      const int rank = have + before + __popc(bal & ((1u << lane) - 1u));
      if (acc && rank < want) {
        const float key = greedy ? sc : 0.f;   // sampling keeps the first accepted candidate; greedy the best of the first `want`
        if (key > best || (key == best && rank < best_rank)) { best = key; best_rank = rank; best_x[0] = x[0]; best_x[1] = x[1]; best_x[2] = x[2]; }
      }
      have += total;
    }
    {
      float val = best_rank == 0x7fffffff ? -3.0e38f : best;
      int who = best_rank == 0x7fffffff ? 0x7fffffff : best_rank;
      // the winner: greedy -> highest score (ties: lowest rank); sampling -> rank 0
      float v2 = val;
      int key = who;
      block_argmax(v2, key, s.red);
      if (who == key && val == v2 && who != 0x7fffffff) { s_val[1] = best_x[0]; s_val[2] = best_x[1]; s_val[3] = best_x[2]; }
      if (key == 0x7fffffff && threadIdx.x == 0) { s_val[1] = 0.f; s_val[2] = 0.f; s_val[3] = 1.f; }   // nothing accepted (cannot happen with a finite score)
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      act[0] = (float)focus; act[1] = (float)element; act[2] = dist;
      act[3] = s_val[1]; act[4] = s_val[2]; act[5] = s_val[3];
    }
    __syncthreads();
    // ---- evaluate the chosen action with the code of the evaluate-mode step
    PolicyScalars ps;
    policy_forward(d, P, Wt, b, n_atoms, bags, actions, A_last, inv, flogit, trans, s, ps, nullptr);
    if (threadIdx.x == 0) {
      lse_out[b] = make_float2(ps.lse_max, ps.lse_sum);
      out.logp[b] = ((ps.logp_f + ps.logp_e) + ps.logp_d) + ps.logp_o;
      out.ent[b] = ps.ent_f + ps.ent_e;
      out.v[b] = ps.v;
      if (out.logp_parts) {
        out.logp_parts[b * 4 + 0] = ps.logp_f; out.logp_parts[b * 4 + 1] = ps.logp_e;
        out.logp_parts[b * 4 + 2] = ps.logp_d; out.logp_parts[b * 4 + 3] = ps.logp_o;
      }
      if (out.log_z) out.log_z[b] = ps.log_z;
    }
    policy_write_extras(d, P, b, s, ps, out);
  }
}

}  // namespace mgb
