// mlp_tc.cuh — the per-atom two-layer MLPs (focus head phi_focus, value transform phi_trans: molgym/modules.py:37-50 as used at
// molgym/agents/covariant/agent.py:223-226,313-316) on the tensor cores.
//
// These are the path's real GEMMs (rows = active atoms of the minibatch, K = 48 Z invariants, N = network width).  fp32 parity at
// 1e-5 rules out plain TF32 (10-bit mantissa), so every product is formed from the split  x = hi + lo  (hi = tf32(x),
// lo = tf32(x - hi)):  A B ~= A_lo B_hi + A_hi B_lo + A_hi B_hi  — three mma.sync.m16n8k8 TF32 instructions with fp32
// accumulation ("3xTF32"), relative error ~2^-21 per product.  A CTA keeps the weight matrices in shared memory (row strides
// chosen so that the fragment loads are bank-conflict free) and walks row tiles of 16; warp w owns output columns
// [w N/4, (w+1) N/4).  The kernel emulator build evaluates the same fragments with shuffles (tests/cusim).
#pragma once
#include "heads.cuh"

namespace mgb {

constexpr int kTcThreads = 256;   // eight warps: warp w owns the output columns [w N/8, (w+1) N/8) of the 16-row tile
constexpr int kTcRows = 16;

__device__ __forceinline__ unsigned tf32_of(float x) {
#ifndef MGB_CUSIM
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
#else
  unsigned u = __float_as_uint(x);
  u += 0x1000u;            // round to nearest (ties away), 13 dropped bits
  return u & 0xffffe000u;
#endif
}
struct Tf32Pair { unsigned hi, lo; };
__device__ __forceinline__ Tf32Pair tf32_split(float x) {
  Tf32Pair p;
  p.hi = tf32_of(x);
  p.lo = tf32_of(x - __uint_as_float(p.hi));
  return p;
}

// D(16x8) += A(16x8) B(8x8), fragments as in the PTX ISA: g = lane / 4, t = lane % 4;
//   a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  b0 (k = t, n = g) b1 (k = t+4, n = g);  d0 (g, 2t) d1 (g, 2t+1) d2 (g+8, 2t) d3 (g+8, 2t+1)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
#ifndef MGB_CUSIM
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
#else
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int k = 0; k < 8; ++k) {
    const int srcA = g * 4 + (k & 3);
    const float ag = __uint_as_float(__shfl_sync(0xffffffffu, k < 4 ? a[0] : a[2], srcA));
    const float ag8 = __uint_as_float(__shfl_sync(0xffffffffu, k < 4 ? a[1] : a[3], srcA));
    const float b0 = __uint_as_float(__shfl_sync(0xffffffffu, k < 4 ? b[0] : b[1], (2 * t) * 4 + (k & 3)));
    const float b1 = __uint_as_float(__shfl_sync(0xffffffffu, k < 4 ? b[0] : b[1], (2 * t + 1) * 4 + (k & 3)));
    d[0] = fmaf(ag, b0, d[0]); d[1] = fmaf(ag, b1, d[1]); d[2] = fmaf(ag8, b0, d[2]); d[3] = fmaf(ag8, b1, d[3]);
  }
#endif
}
// 3xTF32: small terms first
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const Tf32Pair (&a)[4], const Tf32Pair (&b)[2]) {
  const unsigned ah[4] = {a[0].hi, a[1].hi, a[2].hi, a[3].hi}, al[4] = {a[0].lo, a[1].lo, a[2].lo, a[3].lo};
  const unsigned bh[2] = {b[0].hi, b[1].hi}, bl[2] = {b[0].lo, b[1].lo};
  mma_tf32(d, al, bh);
  mma_tf32(d, ah, bl);
  mma_tf32(d, ah, bh);
}

// row strides (floats): A tiles [16][K] read as (g, k0 + t): stride = 4 mod 32 would do; 20 mod 32 keeps 16-byte row alignment for
// K % 4 == 0 ... any stride with {20 g + t} distinct works; B tiles read as (n0 + g, k0 + t) from an [N][K] matrix likewise.
__host__ __device__ inline int tc_stride_nk(int K) { return K + ((20 - (K & 31)) & 31); }     // = 20 mod 32: (20 g + t) distinct for g < 8, t < 4
// B tiles read as (k0 + t, n0 + g) from a [K][N] matrix: stride = 8 mod 32
__host__ __device__ inline int tc_stride_kn(int N) { return N + ((8 - (N & 31)) & 31); }

// one 16-row tile: acc[nt][4] += A[16][K] (shared, stride sa) * B, for the warp's n-tiles n0 + 8 nt.
//   BT = false: B[k][n] = Bm[n * sb + k]  ([N][K] matrix: y = x W^T with W in the reference layout [out][in])
//   BT = true : B[k][n] = Bm[k * sb + n]  ([K][N] matrix: dx = dy W with W [out][in])
template <int NT, bool BT>
__device__ __forceinline__ void tc_tile_gemm(const float* __restrict__ A, int sa, const float* __restrict__ Bm, int sb, int K, int n0, int nt_count,
                                             float (&acc)[NT][4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int k0 = 0; k0 < K; k0 += 8) {
    Tf32Pair a[4];
    a[0] = tf32_split(A[g * sa + k0 + t]);
    a[1] = tf32_split(A[(g + 8) * sa + k0 + t]);
    a[2] = tf32_split(A[g * sa + k0 + t + 4]);
    a[3] = tf32_split(A[(g + 8) * sa + k0 + t + 4]);
    MGB_UNROLL
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nt_count) {
        const int n = n0 + 8 * nt + g;
        Tf32Pair b[2];
        b[0] = tf32_split(BT ? Bm[(k0 + t) * sb + n] : Bm[n * sb + k0 + t]);
        b[1] = tf32_split(BT ? Bm[(k0 + t + 4) * sb + n] : Bm[n * sb + k0 + t + 4]);
        mma_3xtf32(acc[nt], a, b);
      }
    }
  }
}

// Stage a row-major [rows][cols] matrix from global memory into shared memory with row stride `stride`: one bulk copy (TMA) per
// row, issued by the threads in parallel and all completing on the same mbarrier (`bar` initialised by the caller; the caller
// posts the expected byte count once with tc_stage_expect).  Rows are 16-byte aligned on both sides (checked by the host).
__device__ __forceinline__ void tc_stage_expect(SmemBarrier* bar, unsigned bytes) {
  if (threadIdx.x == 0) mbar_expect(bar, bytes);
}
__device__ __forceinline__ void tc_stage_matrix(float* dst, int stride, const float* __restrict__ src, int rows, int cols, SmemBarrier* bar) {
#ifndef MGB_CUSIM
  for (int r = threadIdx.x; r < rows; r += blockDim.x) bulk_g2s(dst + r * stride, src + (long long)r * cols, (unsigned)(cols * 4), bar);
#else
  for (int idx = threadIdx.x; idx < rows * cols; idx += blockDim.x) dst[(idx / cols) * stride + idx % cols] = src[idx];
#endif
}

__host__ __device__ inline size_t rows_mlp_tc_fwd_smem_bytes(int K, int Wd, bool second) {
  return sizeof(float) * ((size_t)Wd * tc_stride_nk(K) + (second ? (size_t)Wd * tc_stride_nk(Wd) : 0) + (size_t)kTcRows * tc_stride_nk(K) +
                          (size_t)kTcRows * tc_stride_nk(Wd) + kTcRows + 4);
}

// Forward.  grid = (CTAs, 2): y = 0 focus head on the active rows, y = 1 value transform on the valid atoms; a CTA keeps its weights
// and walks row tiles blockIdx.x, blockIdx.x + gridDim.x, ...
template <int NT>   // n-tiles per warp = Wd / 64
__global__ void __launch_bounds__(kTcThreads)
k_rows_mlp_fwd_tc(const CovDesc* __restrict__ dp, const float* __restrict__ P, int B, const int* __restrict__ act_off,
                  const int* __restrict__ act_list, const int* __restrict__ atom_off, const int* __restrict__ atom_list,
                  const float* __restrict__ X, float* __restrict__ H0, float* __restrict__ Y0, float* __restrict__ H1, float* __restrict__ Y1) {
  const CovDesc& d = *dp;
  const bool focus = blockIdx.y == 0;
  const MlpDesc& M = focus ? d.focus : d.trans;
  const int n_rows = focus ? act_off[B] : atom_off[B];
  if ((int)(blockIdx.x * kTcRows) >= n_rows) return;
  const int* list = focus ? act_list : atom_list;
  float* H = focus ? H0 : H1;
  float* Y = focus ? Y0 : Y1;
  const int K = M.in, Wd = M.hidden, No = M.out;
  const int sk = tc_stride_nk(K), sw = tc_stride_nk(Wd);
  MGB_DYN_SMEM(float, sm);
  float* sW0 = sm;                                   // [Wd][sk]   reference layout of W0 ([out][in])
  float* sW1 = sW0 + (size_t)Wd * sk;                // [Wd][sw]   (value transform only)
  float* sx = sW1 + (focus ? 0 : (size_t)Wd * sw);   // [16][sk]
  float* sh = sx + kTcRows * sk;                     // [16][sw]
  int* s_row = reinterpret_cast<int*>(sh + kTcRows * sw);   // [16]
  __shared__ SmemBarrier s_bar;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  tc_stage_expect(&s_bar, (unsigned)(4 * (Wd * K + (focus ? 0 : No * Wd))));
  __syncthreads();   // the expectation is posted before any copy can complete
  tc_stage_matrix(sW0, sk, P + M.W0, Wd, K, &s_bar);
  if (!focus) tc_stage_matrix(sW1, sw, P + M.W1, No, Wd, &s_bar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int n0 = warp * 8 * NT;
  bool staged = false;
  for (int r0 = blockIdx.x * kTcRows; r0 < n_rows; r0 += gridDim.x * kTcRows) {
    __syncthreads();   // previous tile consumed
    if ((int)threadIdx.x < kTcRows) s_row[threadIdx.x] = r0 + (int)threadIdx.x < n_rows ? list[r0 + threadIdx.x] : -1;
    __syncthreads();
    for (int idx = threadIdx.x; idx < kTcRows * (K / 4); idx += blockDim.x) {
      const int q = idx / (K / 4), k4 = idx - q * (K / 4);
      const int row = s_row[q];
      *reinterpret_cast<float4*>(sx + q * sk + 4 * k4) =
          row >= 0 ? reinterpret_cast<const float4*>(X + (long long)row * K)[k4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (!staged) { mbar_wait(&s_bar, 0); staged = true; }   // the weights have landed (the first row tile was gathered meanwhile)
    float acc[NT][4];
    MGB_UNROLL
    for (int nt = 0; nt < NT; ++nt) {
      const float b0 = P[M.b0 + n0 + 8 * nt + 2 * t], b1 = P[M.b0 + n0 + 8 * nt + 2 * t + 1];
      acc[nt][0] = b0; acc[nt][1] = b1; acc[nt][2] = b0; acc[nt][3] = b1;
    }
    tc_tile_gemm<NT, false>(sx, sk, sW0, sk, K, n0, NT, acc);
    const int ra = s_row[g], rb = s_row[g + 8];
    MGB_UNROLL
    for (int nt = 0; nt < NT; ++nt) {
      const int col = n0 + 8 * nt + 2 * t;
      const float h0 = fmaxf(acc[nt][0], 0.f), h1 = fmaxf(acc[nt][1], 0.f), h2 = fmaxf(acc[nt][2], 0.f), h3 = fmaxf(acc[nt][3], 0.f);
      *reinterpret_cast<float2*>(sh + g * sw + col) = make_float2(h0, h1);
      *reinterpret_cast<float2*>(sh + (g + 8) * sw + col) = make_float2(h2, h3);
      if (ra >= 0) *reinterpret_cast<float2*>(H + (long long)ra * Wd + col) = make_float2(h0, h1);
      if (rb >= 0) *reinterpret_cast<float2*>(H + (long long)rb * Wd + col) = make_float2(h2, h3);
    }
    __syncthreads();
    if (!focus) {
      MGB_UNROLL
      for (int nt = 0; nt < NT; ++nt) {
        const float b0 = P[M.b1 + n0 + 8 * nt + 2 * t], b1 = P[M.b1 + n0 + 8 * nt + 2 * t + 1];
        acc[nt][0] = b0; acc[nt][1] = b1; acc[nt][2] = b0; acc[nt][3] = b1;
      }
      tc_tile_gemm<NT, false>(sh, sw, sW1, sw, Wd, n0, NT, acc);
      MGB_UNROLL
      for (int nt = 0; nt < NT; ++nt) {
        const int col = n0 + 8 * nt + 2 * t;
        if (ra >= 0) *reinterpret_cast<float2*>(Y + (long long)ra * No + col) = make_float2(acc[nt][0], acc[nt][1]);
        if (rb >= 0) *reinterpret_cast<float2*>(Y + (long long)rb * No + col) = make_float2(acc[nt][2], acc[nt][3]);
      }
    } else {
      // the focus logit: one output per row — a warp per two rows, lanes over the hidden units
      for (int q = warp * 2; q < warp * 2 + 2; ++q) {
        float part = 0.f;
        for (int k = lane; k < Wd; k += 32) part = fmaf(P[M.W1 + k], sh[q * sw + k], part);
        part = warp_sum(part);
        if (lane == 0 && s_row[q] >= 0) Y[s_row[q]] = part + P[M.b1];
      }
    }
  }
}

// Backward.  y = 0: focus head (dY0 [rows]), y = 1: value transform (dY = dvf[b, :] for every atom of canvas b, also written to dY1 for
// the weight gradient).  dH = relu'(H) * (dY W1), dX += dH W0 (atomics: both heads add into the same rows).
__host__ __device__ inline size_t rows_mlp_tc_bwd_smem_bytes(int K, int Wd, bool second) {
  return sizeof(float) * ((size_t)Wd * tc_stride_kn(K) + (second ? (size_t)Wd * tc_stride_kn(Wd) : 0) + 2 * (size_t)kTcRows * tc_stride_nk(Wd) +
                          kTcRows + 4);
}
template <int NT, int NTX>   // n-tiles per warp for the hidden width (Wd / 64) / for the input width (ceil(K / 64))
__global__ void __launch_bounds__(kTcThreads)
k_rows_mlp_bwd_tc(const CovDesc* __restrict__ dp, const float* __restrict__ P, int B, const int* __restrict__ act_off,
                  const int* __restrict__ act_list, const int* __restrict__ atom_off, const int* __restrict__ atom_list,
                  const float* __restrict__ H0, const float* __restrict__ dY0, float* __restrict__ dH0, const float* __restrict__ H1,
                  const float* __restrict__ dvf, float* __restrict__ dY1, float* __restrict__ dH1, float* __restrict__ dX) {
  const CovDesc& d = *dp;
  const bool focus = blockIdx.y == 0;
  const MlpDesc& M = focus ? d.focus : d.trans;
  const int n_rows = focus ? act_off[B] : atom_off[B];
  if ((int)(blockIdx.x * kTcRows) >= n_rows) return;
  const int* list = focus ? act_list : atom_list;
  const int K = M.in, Wd = M.hidden, N = d.N;
  const int sk = tc_stride_kn(K), sw = tc_stride_kn(Wd), sa = tc_stride_nk(Wd);
  MGB_DYN_SMEM(float, sm);
  float* sW0 = sm;                                   // [Wd][sk]  W0 [hidden][in]: B[k = h][n = input]
  float* sW1 = sW0 + (size_t)Wd * sk;                // [Wd][sw]  W1 [out][hidden]: B[k = out][n = hidden]  (value transform only)
  float* sdy = sW1 + (focus ? 0 : (size_t)Wd * sw);  // [16][sa]
  float* sdh = sdy + kTcRows * sa;                   // [16][sa]
  int* s_row = reinterpret_cast<int*>(sdh + kTcRows * sa);
  __shared__ SmemBarrier s_bar;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  tc_stage_expect(&s_bar, (unsigned)(4 * (Wd * K + (focus ? 0 : Wd * Wd))));
  __syncthreads();
  tc_stage_matrix(sW0, sk, P + M.W0, Wd, K, &s_bar);
  if (!focus) tc_stage_matrix(sW1, sw, P + M.W1, Wd, Wd, &s_bar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int n0 = warp * 8 * NT;
  const int ktiles = K / 8, per = (ktiles + 7) / 8, x0 = warp * per, xcount = max(0, min(per, ktiles - x0));   // input columns of this warp
  bool staged = false;
  for (int r0 = blockIdx.x * kTcRows; r0 < n_rows; r0 += gridDim.x * kTcRows) {
    __syncthreads();
    if ((int)threadIdx.x < kTcRows) s_row[threadIdx.x] = r0 + (int)threadIdx.x < n_rows ? list[r0 + threadIdx.x] : -1;
    __syncthreads();
    if (focus) {
      for (int idx = threadIdx.x; idx < kTcRows * Wd; idx += blockDim.x) {
        const int q = idx / Wd, h = idx - q * Wd;
        const int row = s_row[q];
        float gq = 0.f;
        if (row >= 0) {
          gq = H0[(long long)row * Wd + h] > 0.f ? P[M.W1 + h] * dY0[row] : 0.f;
          dH0[(long long)row * Wd + h] = gq;
        }
        sdh[q * sa + h] = gq;
      }
    } else {
      for (int idx = threadIdx.x; idx < kTcRows * Wd; idx += blockDim.x) {
        const int q = idx / Wd, o = idx - q * Wd;
        const int row = s_row[q];
        float v = 0.f;
        if (row >= 0) {
          v = dvf[(long long)(row / N) * Wd + o];
          dY1[(long long)row * Wd + o] = v;
        }
        sdy[q * sa + o] = v;
      }
      __syncthreads();
      if (!staged) { mbar_wait(&s_bar, 0); staged = true; }
      float acc[NT][4];
      MGB_UNROLL
      for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = 0.f; acc[nt][1] = 0.f; acc[nt][2] = 0.f; acc[nt][3] = 0.f; }
      tc_tile_gemm<NT, true>(sdy, sa, sW1, sw, Wd, n0, NT, acc);   // dH1 = dY W1
      const int ra = s_row[g], rb = s_row[g + 8];
      MGB_UNROLL
      for (int nt = 0; nt < NT; ++nt) {
        const int col = n0 + 8 * nt + 2 * t;
        float2 ga = make_float2(0.f, 0.f), gb = make_float2(0.f, 0.f);
        if (ra >= 0) {
          const float2 h = *reinterpret_cast<const float2*>(H1 + (long long)ra * Wd + col);
          ga = make_float2(h.x > 0.f ? acc[nt][0] : 0.f, h.y > 0.f ? acc[nt][1] : 0.f);
          *reinterpret_cast<float2*>(dH1 + (long long)ra * Wd + col) = ga;
        }
        if (rb >= 0) {
          const float2 h = *reinterpret_cast<const float2*>(H1 + (long long)rb * Wd + col);
          gb = make_float2(h.x > 0.f ? acc[nt][2] : 0.f, h.y > 0.f ? acc[nt][3] : 0.f);
          *reinterpret_cast<float2*>(dH1 + (long long)rb * Wd + col) = gb;
        }
        *reinterpret_cast<float2*>(sdh + g * sa + col) = ga;
        *reinterpret_cast<float2*>(sdh + (g + 8) * sa + col) = gb;
      }
    }
    __syncthreads();
    if (!staged) { mbar_wait(&s_bar, 0); staged = true; }
    float accx[NTX][4];
    MGB_UNROLL
    for (int nt = 0; nt < NTX; ++nt) { accx[nt][0] = 0.f; accx[nt][1] = 0.f; accx[nt][2] = 0.f; accx[nt][3] = 0.f; }
    tc_tile_gemm<NTX, true>(sdh, sa, sW0, sk, Wd, 8 * x0, xcount, accx);   // dX = dH W0
    const int ra = s_row[g], rb = s_row[g + 8];
    MGB_UNROLL
    for (int nt = 0; nt < NTX; ++nt) {
      if (nt < xcount) {
        const int col = 8 * (x0 + nt) + 2 * t;
        if (ra >= 0) {
          if (accx[nt][0] != 0.f) atomicAdd(dX + (long long)ra * K + col, accx[nt][0]);
          if (accx[nt][1] != 0.f) atomicAdd(dX + (long long)ra * K + col + 1, accx[nt][1]);
        }
        if (rb >= 0) {
          if (accx[nt][2] != 0.f) atomicAdd(dX + (long long)rb * K + col, accx[nt][2]);
          if (accx[nt][3] != 0.f) atomicAdd(dX + (long long)rb * K + col + 1, accx[nt][3]);
        }
      }
    }
  }
}

}  // namespace mgb
