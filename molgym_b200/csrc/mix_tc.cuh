// mix_tc.cuh — the channel mix of an atom level (cormorant MixReps / CatMixReps as used at molgym/agents/covariant/
// gnn.py via cg_lib: A_{s+1}[l][m][c'] = sum_k W_l[c'][k] cat_l[m][k], complex) and its input gradient on the tensor cores.
//
// Per ell this is a real GEMM with a long inner dimension:  [rows = (atom, m)] x [2 K_l]  times  [2 K_l] x [2 Cout]
//     X = (.. x_re[k], x_im[k] ..),   Bm[2k][2c'] = W_re, Bm[2k][2c'+1] = W_im, Bm[2k+1][2c'] = -W_im, Bm[2k+1][2c'+1] = W_re
// (K_l up to 350 complex = 700 reals at the default width).  The FFMA version needs one shared-memory weight load per two
// FFMA2 and ran at ~18 % of the HBM rate the cat rows could stream at.  Here the products are 3xTF32 mma.sync m16n8k8
// (mlp_tc.cuh: fp32-level accuracy); what decides the speed is how often an operand is split into its (hi, lo) TF32 pair:
//   forward : one persistent CTA of 16 warps per SM.  The warps split the inner dimension; each keeps the weight fragments of its
//             k-slice, split once, in registers for the CTA's lifetime.  The 16-row cat tiles arrive by per-row bulk copies
//             (TMA) into a three-deep ring, every element is split by exactly one warp, partial 16 x 2 Cout tiles are summed
//             through shared memory.
//   backward: dcat = dA W^H is the same matrix read transposed ([2 Cout] inner, [2 K_l] outer).  The weights sit in shared memory
//             as pre-split (hi, lo) pairs; the warps split the output columns and store them straight to HBM.
// CTAs are dealt to the five ells in proportion to their share of the work.
#pragma once
#include "mlp_tc.cuh"

namespace mgb {

constexpr int kMixTcThreads = 512;
constexpr int kMixTcRows = 16;
constexpr int kMixTcWarps = kMixTcThreads / 32;
constexpr int kMixTcBufs = 3;      // row tiles in flight per CTA (forward)
constexpr int kMixTcKsw = 6;       // k-steps (of 8 reals) per warp whose weight fragments stay in registers: K_l <= 384

// smallest stride >= cols with stride = mod (mod `of`)
__host__ __device__ inline int mix_tc_stride(int cols, int mod, int of = 32) { return cols + ((mod - (cols % of) + of) % of); }
__host__ __device__ inline int mix_tc_ksteps(int K) { return (2 * K + 7) / 8; }

__host__ __device__ inline size_t mix_tc_fwd_smem_bytes(int Kmax, int NT) {
  const int sa = mix_tc_stride(8 * mix_tc_ksteps(Kmax), 4);
  return sizeof(float) * ((size_t)kMixTcBufs * kMixTcRows * sa + 2 * (size_t)kMixTcWarps * kMixTcRows * (8 * NT + 8)) +
         sizeof(long long) * (kMixTcBufs + 1) * kMixTcRows;
}
__host__ __device__ inline size_t mix_tc_bwd_smem_bytes(int Kmax, int NT) {
  const int sb = mix_tc_stride(8 * mix_tc_ksteps(Kmax), 8, 16);
  return sizeof(float) * ((size_t)2 * 8 * NT * sb + (size_t)kMixTcRows * mix_tc_stride(8 * NT, 4)) + sizeof(long long) * 2 * kMixTcRows;
}

// CTA -> (ell, index among the ell's CTAs, number of CTAs of the ell); gridDim.x >= kNL
__device__ __forceinline__ void mix_tc_assign(const LevelDesc& L, int G, int b, int c0, int& l_out, int& ci, int& cn) {
  long long work[kNL], tot = 0;
  for (int l = 0; l < kNL; ++l) { work[l] = (long long)(2 * l + 1) * (L.catA[l] + c0); tot += work[l]; }   // c0: the per-tile constant cost in units of k
  int n[kNL], used = 0;
  for (int l = 0; l < kNL; ++l) { n[l] = 1 + (int)((long long)(G - kNL) * work[l] / tot); used += n[l]; }
  for (int r = G - used, l = kNL - 1; r > 0; --r) { n[l] += 1; l = l ? l - 1 : kNL - 1; }
  int start = 0;
  for (int l = 0; l < kNL; ++l) {
    if (b < start + n[l] || l == kNL - 1) { l_out = l; ci = b - start; cn = n[l]; return; }
    start += n[l];
  }
}

// entry (kk, n) of the real-expanded weights of one ell: kk = input real (2k + part), n = output real (2c + comp); W [Cout][K] complex
__device__ __forceinline__ float mix_tc_weight(const float2* __restrict__ W2, int K, int Cout, int kk, int n) {
  const int k = kk >> 1, c = n >> 1;
  if (k >= K || c >= Cout) return 0.f;
  const float2 w = W2[(long long)c * K + k];
  return (kk & 1) ? ((n & 1) ? w.x : -w.y) : ((n & 1) ? w.y : w.x);
}

template <int NT>
__global__ void __launch_bounds__(kMixTcThreads, 1)
k_mix_rows_tc_fwd(const CovDesc* __restrict__ dp, int level, const float* __restrict__ P, const int* __restrict__ atom_off,
                  const int* __restrict__ atom_list, int B, const float* __restrict__ cat, float* __restrict__ out, int c0) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  int l, ci, cn;
  mix_tc_assign(L, gridDim.x, blockIdx.x, c0, l, ci, cn);
  const int K = L.catA[l], nm = 2 * l + 1, Cout = L.Cout;
  const int rows = atom_off[B] * nm, tiles = (rows + kMixTcRows - 1) / kMixTcRows;
  if (ci >= tiles) return;
  constexpr int NP = 8 * NT, RS = NP + 8;
  const int KT = mix_tc_ksteps(K), sa = mix_tc_stride(8 * KT, 4);
  MGB_DYN_SMEM(float, sm);
  float* sA = sm;                                               // [bufs][16][sa]
  float* red = sA + (size_t)kMixTcBufs * kMixTcRows * sa;        // [2][warps][16][RS]: the partial tiles of two consecutive row tiles
  long long* s_orow = reinterpret_cast<long long*>(red + 2 * (size_t)kMixTcWarps * kMixTcRows * RS);   // [bufs + 1][16] output row (complex index), -1 = none
  __shared__ SmemBarrier s_bar[kMixTcBufs];
  if (threadIdx.x == 0)
    for (int b = 0; b < kMixTcBufs; ++b) mbar_init(&s_bar[b], kMixTcRows);
  for (int idx = threadIdx.x; idx < kMixTcBufs * kMixTcRows * (sa - 2 * K); idx += blockDim.x) {   // the k padding of the row tiles (never written by the copies)
    const int q = idx / (sa - 2 * K), kk = 2 * K + idx % (sa - 2 * K);
    sA[(size_t)q * sa + kk] = 0.f;
  }
  __syncthreads();
  const unsigned row_bytes = 8u * (unsigned)K;
  // tile number `seq` of this CTA goes to buffer seq % bufs; its row table to slot seq % (bufs + 1) — the slot of the tile that is
  // being summed while the copy is issued stays intact
  auto issue = [&](int tile, int seq) {   // threads 0..15: one bulk copy per row of the tile
    const int buf = seq % kMixTcBufs;
    if ((int)threadIdx.x < kMixTcRows) {
      const int row = tile * kMixTcRows + (int)threadIdx.x;
      long long orow = -1;
      if (tile < tiles && row < rows) {
        const int a = row / nm, m = row - a * nm;
        const long long slot = atom_list[a];
        orow = (slot * kM + l * l + m) * Cout;
        mbar_expect(&s_bar[buf], row_bytes);
        bulk_g2s(sA + ((size_t)buf * kMixTcRows + threadIdx.x) * sa, cat + 2 * (slot * L.totA + L.offA[l] + (long long)m * K), row_bytes, &s_bar[buf]);
      } else {
        mbar_expect(&s_bar[buf], 0u);
      }
      s_orow[(seq % (kMixTcBufs + 1)) * kMixTcRows + threadIdx.x] = orow;
    }
  };
  for (int b = 0; b < kMixTcBufs; ++b) issue(ci + b * cn, b);
  // the warp's slice of the inner dimension and its weight fragments (split once, kept in registers for every tile)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int per = (KT + kMixTcWarps - 1) / kMixTcWarps, ks0 = min(KT, warp * per), nks = min(KT, ks0 + per) - ks0;
  Tf32Pair bf[kMixTcKsw][NT][2];
  {
    const float2* W2 = reinterpret_cast<const float2*>(P + L.p_atomW + 2ll * L.offWA[l]);
    MGB_UNROLL
    for (int s = 0; s < kMixTcKsw; ++s) {
      MGB_UNROLL
      for (int nt = 0; nt < NT; ++nt) {
        const int kk = 8 * (ks0 + s) + t, n = 8 * nt + g;
        bf[s][nt][0] = tf32_split(s < nks ? mix_tc_weight(W2, K, Cout, kk, n) : 0.f);
        bf[s][nt][1] = tf32_split(s < nks ? mix_tc_weight(W2, K, Cout, kk + 4, n) : 0.f);
      }
    }
  }
  int it = 0;
  for (int tile = ci; tile < tiles; tile += cn, ++it) {
    const int buf = it % kMixTcBufs;
    mbar_wait(&s_bar[buf], (unsigned)((it / kMixTcBufs) & 1));
    float acc[NT][4];
    MGB_UNROLL
    for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = 0.f; acc[nt][1] = 0.f; acc[nt][2] = 0.f; acc[nt][3] = 0.f; }
    const float* A = sA + (size_t)buf * kMixTcRows * sa + 8 * ks0 + t;
    MGB_UNROLL
    for (int s = 0; s < kMixTcKsw; ++s) {
      if (s < nks) {
        Tf32Pair a[4];
        a[0] = tf32_split(A[g * sa + 8 * s]);
        a[1] = tf32_split(A[(g + 8) * sa + 8 * s]);
        a[2] = tf32_split(A[g * sa + 8 * s + 4]);
        a[3] = tf32_split(A[(g + 8) * sa + 8 * s + 4]);
        MGB_UNROLL
        for (int nt = 0; nt < NT; ++nt) mma_3xtf32(acc[nt], a, bf[s][nt]);
      }
    }
    float* rbuf = red + (size_t)(it & 1) * kMixTcWarps * kMixTcRows * RS;
    float* rw = rbuf + (size_t)warp * kMixTcRows * RS;
    MGB_UNROLL
    for (int nt = 0; nt < NT; ++nt) {
      *reinterpret_cast<float2*>(rw + g * RS + 8 * nt + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<float2*>(rw + (g + 8) * RS + 8 * nt + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
    }
    __syncthreads();   // the one barrier per tile: the tile is consumed, the partial sums are complete
    if (tile + kMixTcBufs * cn < tiles) issue(tile + kMixTcBufs * cn, it + kMixTcBufs);
    const long long* orows = s_orow + (it % (kMixTcBufs + 1)) * kMixTcRows;
    for (int idx = threadIdx.x; idx < kMixTcRows * NP; idx += blockDim.x) {
      const int q = idx / NP, n = idx - q * NP;
      const long long orow = orows[q];
      if (orow >= 0 && n < 2 * Cout) {
        float s = 0.f;
        MGB_UNROLL
        for (int w8 = 0; w8 < kMixTcWarps; ++w8) s += rbuf[((size_t)w8 * kMixTcRows + q) * RS + n];
        out[2 * orow + n] = s;
      }
    }
  }
}

// Backward: dcat[r][kk] = sum_n dA[r][n] Bm[kk][n] — the inner dimension is the 2 Cout output reals, the warps split the 2 K_l
// output columns.  The weights stay in shared memory ALREADY split into (hi, lo) pairs (one 8-byte load per fragment element).
template <int NT>
__global__ void __launch_bounds__(kMixTcThreads, 1)
k_mix_rows_tc_bwd(const CovDesc* __restrict__ dp, int level, const float* __restrict__ P, const int* __restrict__ atom_off,
                  const int* __restrict__ atom_list, int B, const float* __restrict__ dA_out, float* __restrict__ dcat, int c0) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  int l, ci, cn;
  mix_tc_assign(L, gridDim.x, blockIdx.x, c0, l, ci, cn);
  const int K = L.catA[l], nm = 2 * l + 1, Cout = L.Cout;
  const int rows = atom_off[B] * nm, tiles = (rows + kMixTcRows - 1) / kMixTcRows;
  if (ci >= tiles) return;
  constexpr int NP = 8 * NT;
  const int KT = mix_tc_ksteps(K), sb = mix_tc_stride(8 * KT, 8, 16), sg = mix_tc_stride(NP, 4);
  MGB_DYN_SMEM(float, sm);
  uint2* sB = reinterpret_cast<uint2*>(sm);                         // [NP][sb] (hi, lo): B[k = output real of the mix][n = cat real]
  float* sG = sm + 2 * (size_t)NP * sb;                             // [16][sg]  the incoming gradient rows
  long long* s_dst = reinterpret_cast<long long*>(sG + (size_t)kMixTcRows * sg);   // [16] float offset of the dcat row, -1 = none
  long long* s_src = s_dst + kMixTcRows;                                           // [16] float offset of the dA row
  {
    const float2* W2 = reinterpret_cast<const float2*>(P + L.p_atomW + 2ll * L.offWA[l]);
    for (int idx = threadIdx.x; idx < NP * sb; idx += blockDim.x) {
      const int n = idx / sb, kk = idx - n * sb;
      const Tf32Pair p = tf32_split(mix_tc_weight(W2, K, Cout, kk, n));
      sB[idx] = make_uint2(p.hi, p.lo);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int per = (KT + kMixTcWarps - 1) / kMixTcWarps, nt0 = min(KT, warp * per), nt1 = min(KT, nt0 + per);
  for (int tile = ci; tile < tiles; tile += cn) {
    __syncthreads();   // previous tile consumed (first pass: the weights are staged)
    if ((int)threadIdx.x < kMixTcRows) {
      const int row = tile * kMixTcRows + (int)threadIdx.x;
      long long dst = -1, src = 0;
      if (row < rows) {
        const int a = row / nm, m = row - a * nm;
        const long long slot = atom_list[a];
        dst = 2 * (slot * L.totA + L.offA[l] + (long long)m * K);
        src = 2 * (slot * kM + l * l + m) * Cout;
      }
      s_dst[threadIdx.x] = dst;
      s_src[threadIdx.x] = src;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < kMixTcRows * sg; idx += blockDim.x) {
      const int q = idx / sg, n = idx - q * sg;
      sG[idx] = (s_dst[q] >= 0 && n < 2 * Cout) ? dA_out[s_src[q] + n] : 0.f;
    }
    __syncthreads();
    Tf32Pair a[NT][4];
    MGB_UNROLL
    for (int ks = 0; ks < NT; ++ks) {
      a[ks][0] = tf32_split(sG[g * sg + 8 * ks + t]);
      a[ks][1] = tf32_split(sG[(g + 8) * sg + 8 * ks + t]);
      a[ks][2] = tf32_split(sG[g * sg + 8 * ks + t + 4]);
      a[ks][3] = tf32_split(sG[(g + 8) * sg + 8 * ks + t + 4]);
    }
    const long long da = s_dst[g], db = s_dst[g + 8];
    for (int nt = nt0; nt < nt1; ++nt) {
      const int n0 = 8 * nt;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      MGB_UNROLL
      for (int ks = 0; ks < NT; ++ks) {
        const uint2 u0 = sB[(8 * ks + t) * sb + n0 + g], u1 = sB[(8 * ks + t + 4) * sb + n0 + g];
        Tf32Pair b[2];
        b[0].hi = u0.x; b[0].lo = u0.y; b[1].hi = u1.x; b[1].lo = u1.y;
        mma_3xtf32(acc, a[ks], b);
      }
      const int col = n0 + 2 * t;
      if (col < 2 * K) {
        if (da >= 0) *reinterpret_cast<float2*>(dcat + da + col) = make_float2(acc[0], acc[1]);
        if (db >= 0) *reinterpret_cast<float2*>(dcat + db + col) = make_float2(acc[2], acc[3]);
      }
    }
  }
}

}  // namespace mgb
