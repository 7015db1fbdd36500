// edge.cuh — the Cormorant edge level (cormorant CormorantEdgeLevel: DotMatrix + CatMixRepsScalar + MaskLevel; called from
// molgym/agents/covariant/modules.py:110) as thread-per-pair kernels.
//
//   E[b,i,j,l,c'] = s_ij * sum_k WE_l[c',k] catE_ijl[k],   catE_l = [E_prev_l (C) | dot(A_i, A_j) (nLin*C) | radial_l (C)]
//
// Layout of the work: the flat list of valid (b, i, j) pairs (pair_off) is walked with ONE THREAD PER PAIR; all lanes of a
// warp work on the same ell, so the mixing weights are warp-uniform and are read from shared memory as broadcast loads
// while the accumulators stay in registers.  The dot matrix is its own per-canvas kernel (it is symmetric in (i, j) and
// only needs the atom representations).  Weight cotangents are reductions over pairs and run as a separate kernel.
#pragma once
#include "cov_forward.cuh"

namespace mgb {

constexpr int kEdgeC = 10;          // register tile over the (<= 10) hidden channels
constexpr int kPairThreads = 128;

// radial basis of one pair: f[t], t = trig*4 + p (RadPolyTrig), and optionally d f[t] / d(arg_t)
__device__ __forceinline__ void rad_features_all(const PairGeom& g, const float* __restrict__ scales, const float* __restrict__ phases,
                                                 float* f, float* dfdarg) {
  const float inv = g.mrad ? 1.f / g.r : 0.f;
  const float pw[4] = {1.f, inv, inv * inv, inv * inv * inv};
  MGB_UNROLL
  for (int tt = 0; tt < kTrig; ++tt) {
    const float arg = __fadd_rn(__fmul_rn(__fmul_rn(kTwoPi, scales[tt]), g.r), phases[tt]);
    const float sv = g.mrad ? sinf(arg) : 0.f;
    const float cv = (dfdarg && g.mrad) ? cosf(arg) : 0.f;
    MGB_UNROLL
    for (int p = 0; p < 4; ++p) {
      f[tt * 4 + p] = sv * pw[p];
      if (dfdarg) dfdarg[tt * 4 + p] = cv * pw[p];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Dot matrix D[b,i,j,l',c] = sum_m (-1)^m A_i[l',m,c] A_j[l',-m,c]  (cormorant DotMatrix).  One CTA per canvas.
// ------------------------------------------------------------------------------------------------------------
template <int NLIN>
__global__ void __launch_bounds__(256)
k_dot_fwd(const CovDesc* __restrict__ dp, int level, const int* __restrict__ n_atoms, const float* __restrict__ A_in,
          float* __restrict__ D) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int N = d.N, C = L.C, b = blockIdx.x;
  constexpr int NLM = NLIN * NLIN;
  const int n = n_atoms[b];
  if (n == 0) return;
  MGB_DYN_SMEM(float2, sA);   // [n][NLM][C]
  const float2* Ab = reinterpret_cast<const float2*>(A_in) + (long long)b * N * NLM * C;
  for (int idx = threadIdx.x; idx < n * NLM * C; idx += blockDim.x) sA[idx] = Ab[idx];
  __syncthreads();
  float2* Db = reinterpret_cast<float2*>(D) + (long long)b * N * N * kNL * C;
  // D is symmetric in (i, j): a task is (pair i <= j, channel c) — all NLIN ells of it in one straight-line pass (the index decode,
  // one division and one square root, is paid once per 25 complex products), both orientations written
  const int npairs = n * (n + 1) / 2, total = npairs * C;
  const float tn = (float)(2 * n + 1);
  for (int task = threadIdx.x; task < total; task += blockDim.x) {
    const int pr = task / C, c = task - pr * C;
    int i = (int)((tn - sqrtf(tn * tn - 8.f * (float)pr)) * 0.5f);
    i = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
    while (i > 0 && pr < i * n - i * (i - 1) / 2) --i;                 // row i holds the pairs [i n - i (i - 1) / 2, ...) of j = i .. n - 1
    while (pr >= (i + 1) * n - (i + 1) * i / 2) ++i;
    const int j = i + (pr - (i * n - i * (i - 1) / 2));
    const float2* ai = sA + (i * NLM) * C + c;
    const float2* aj = sA + (j * NLM) * C + c;
    float2* dij = Db + ((long long)i * N + j) * kNL * C + c;
    float2* dji = Db + ((long long)j * N + i) * kNL * C + c;
    MGB_UNROLL
    for (int lp = 0; lp < NLIN; ++lp) {
      float2 acc = make_float2(0.f, 0.f);
      MGB_UNROLL
      for (int m = -lp; m <= lp; ++m) {
        const float2 v = cmul(ai[(lp * lp + lp + m) * C], aj[(lp * lp + lp - m) * C]);
        if (m & 1) { acc.x -= v.x; acc.y -= v.y; } else { acc.x += v.x; acc.y += v.y; }
      }
      dij[lp * C] = acc;
      if (i != j) dji[lp * C] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Forward: grid = (pair blocks, 5 ells), one thread per pair.
// ------------------------------------------------------------------------------------------------------------
template <int NLIN>
__global__ void __launch_bounds__(kPairThreads)
k_edge_pairs_fwd(const CovDesc* __restrict__ dp, int level, int B, const float* __restrict__ P, const float* __restrict__ Wt,
                 const float* __restrict__ pos, const int* __restrict__ n_atoms, const int* __restrict__ pair_off,
                 const float* __restrict__ D, const float* __restrict__ E_prev, float* __restrict__ E_out,
                 int* __restrict__ pair_slot /* level 0 only: [pairs] dense slot (b*N+i)*N+j of every flat pair, else NULL */) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int N = d.N, C = L.C, C2 = 2 * C, l = blockIdx.y, K = L.catE[l];
  const int total = pair_off[B];
  if ((int)(blockIdx.x * blockDim.x) >= total) return;
  MGB_DYN_SMEM(float2, smem);
  float2* sW = smem;                                          // [K][kEdgeC]   WE_l transposed ([k][c'])
  float* sRad = reinterpret_cast<float*>(sW + K * kEdgeC);    // [2C][32] radial linear of this ell + [2C] bias
  {
    const float2* src = reinterpret_cast<const float2*>(Wt + d.wt_edge[level]) + L.offE[l];
    for (int idx = threadIdx.x; idx < K * kEdgeC; idx += blockDim.x) {
      const int k = idx / kEdgeC, c = idx - k * kEdgeC;
      sW[idx] = c < C ? src[k * C + c] : make_float2(0.f, 0.f);
    }
    for (int idx = threadIdx.x; idx < C2 * kRadFeat; idx += blockDim.x) sRad[idx] = P[L.p_radW + (long long)l * C2 * kRadFeat + idx];
    for (int idx = threadIdx.x; idx < C2; idx += blockDim.x) sRad[C2 * kRadFeat + idx] = P[L.p_radb + l * C2 + idx];
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const PairId id = decode_pair(p, B, pair_off, n_atoms);
  const PairGeom g = pair_geom(pos + (long long)id.b * N * 3, id.i, id.j, d.cut_rad, d.cut_width);
  const long long pair = ((long long)id.b * N + id.i) * N + id.j;
  if (pair_slot && l == 0) pair_slot[p] = (int)pair;
  float2 acc[kEdgeC];
  MGB_UNROLL
  for (int c = 0; c < kEdgeC; ++c) acc[c] = make_float2(0.f, 0.f);
  int kk = 0;
  const bool pairable = (C & 1) == 0;   // even channel count: the pair's rows are read / written two complex channels (16 bytes) at a time
  if (L.has_prev) {
    const float2* x = reinterpret_cast<const float2*>(E_prev) + pair * kNL * C + l * C;
    if (pairable) {
      for (int k = 0; k < C; k += 2) {
        const float4 xv = *reinterpret_cast<const float4*>(x + k);
        MGB_UNROLL
        for (int c = 0; c < kEdgeC; ++c) {
          cfma(acc[c], sW[(kk + k) * kEdgeC + c], make_float2(xv.x, xv.y));
          cfma(acc[c], sW[(kk + k + 1) * kEdgeC + c], make_float2(xv.z, xv.w));
        }
      }
    } else {
      for (int k = 0; k < C; ++k) {
        const float2 xv = x[k];
        MGB_UNROLL
        for (int c = 0; c < kEdgeC; ++c) cfma(acc[c], sW[(kk + k) * kEdgeC + c], xv);
      }
    }
    kk += C;
  }
  if (l < NLIN) {
    const float2* x = reinterpret_cast<const float2*>(D) + pair * kNL * C;
    if (pairable) {
#pragma unroll 2
      for (int k = 0; k < NLIN * C; k += 2) {
        const float4 xv = *reinterpret_cast<const float4*>(x + k);
        MGB_UNROLL
        for (int c = 0; c < kEdgeC; ++c) {
          cfma(acc[c], sW[(kk + k) * kEdgeC + c], make_float2(xv.x, xv.y));
          cfma(acc[c], sW[(kk + k + 1) * kEdgeC + c], make_float2(xv.z, xv.w));
        }
      }
    } else {
#pragma unroll 2
      for (int k = 0; k < NLIN * C; ++k) {
        const float2 xv = x[k];
        MGB_UNROLL
        for (int c = 0; c < kEdgeC; ++c) cfma(acc[c], sW[(kk + k) * kEdgeC + c], xv);
      }
    }
    kk += NLIN * C;
  }
  {
    float f[kRadFeat];
    rad_features_all(g, P + L.p_scales, P + L.p_phases, f, nullptr);
    for (int k = 0; k < C; ++k) {
      float re = sRad[C2 * kRadFeat + 2 * k], im = sRad[C2 * kRadFeat + 2 * k + 1];
      const float* wr = sRad + (2 * k) * kRadFeat;
      MGB_UNROLL
      for (int t = 0; t < kRadFeat; ++t) { re = fmaf(wr[t], f[t], re); im = fmaf(wr[kRadFeat + t], f[t], im); }
      const float2 xv = make_float2(re, im);
      MGB_UNROLL
      for (int c = 0; c < kEdgeC; ++c) cfma(acc[c], sW[(kk + k) * kEdgeC + c], xv);
    }
  }
  float2* Eo = reinterpret_cast<float2*>(E_out) + pair * kNL * C + l * C;
  if (pairable) {
    MGB_UNROLL
    for (int c = 0; c < kEdgeC; c += 2)
      if (c < C) *reinterpret_cast<float4*>(Eo + c) = make_float4(acc[c].x * g.s, acc[c].y * g.s, acc[c + 1].x * g.s, acc[c + 1].y * g.s);
  } else {
    MGB_UNROLL
    for (int c = 0; c < kEdgeC; ++c)
      if (c < C) Eo[c] = make_float2(acc[c].x * g.s, acc[c].y * g.s);
  }
}

// scale / phase cotangents of a CTA: warp reduction, shared-memory reduction over the warps, then ONE global atomic per CTA
// and parameter (all CTAs of all ells hit the same 16 floats: per-warp atomics serialise in L2)
__device__ __forceinline__ void edge_flush_scale_phase(const float* dsc, const float* dph, float* __restrict__ g_scales,
                                                       float* __restrict__ g_phases) {
  __shared__ float s_red[2 * kTrig];
  if (threadIdx.x < 2 * kTrig) s_red[threadIdx.x] = 0.f;
  __syncthreads();
  MGB_UNROLL
  for (int t = 0; t < kTrig; ++t) {
    const float a = warp_sum(dsc[t]), b2 = warp_sum(dph[t]);
    if ((threadIdx.x & 31) == 0) {
      if (a != 0.f) atomicAdd(&s_red[t], a);
      if (b2 != 0.f) atomicAdd(&s_red[kTrig + t], b2);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * kTrig) {
    const float v = s_red[threadIdx.x];
    if (v != 0.f) atomicAdd(threadIdx.x < kTrig ? g_scales + threadIdx.x : g_phases + (threadIdx.x - kTrig), v);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Backward, per-pair part: one thread per pair, loop over the five ells.
//   reads  dE[pair][l][c']                      (cotangent of this level's edge scalars)
//   writes dE_prev[pair][l][k] (assigned), dD[pair][l'*C + c] (assigned),
//          scratch for the weight-gradient kernel: dpre[pair][l][c'] = s dE, R[pair][l][2C] (radial filter values),
//          dR[pair][l][2C], f[pair][32];  accumulates the radial scale / phase cotangents (per-thread, then atomics).
// ------------------------------------------------------------------------------------------------------------
struct EdgeScratch {
  float* dpre;   // [pairs][5][C][2]
  float* R;      // [pairs][5][2C]
  float* dR;     // [pairs][5][2C]
  float* f;      // [pairs][32]
};

template <int NLIN, bool SPLIT>   // SPLIT: grid.y = 5, one thread per (pair, ell); dD goes to slice ell (summed by k_dot_bwd)
__global__ void __launch_bounds__(kPairThreads)
k_edge_pairs_bwd(const CovDesc* __restrict__ dp, int level, int B, const float* __restrict__ P, const float* __restrict__ pos,
                 const int* __restrict__ n_atoms, const int* __restrict__ pair_off, const float* __restrict__ dE,
                 float* __restrict__ dE_prev, float* __restrict__ dD, long long slice_stride /* complex */, EdgeScratch sc,
                 float* __restrict__ grad) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int N = d.N, C = L.C, C2 = 2 * C;
  const int total = pair_off[B];
  if ((int)(blockIdx.x * blockDim.x) >= total) return;
  MGB_DYN_SMEM(float2, smem);
  float2* sW = smem;                                              // [sumCatE][kEdgeC]: WE in the REFERENCE layout per ell, [k][c'] padded
  float* sRad = reinterpret_cast<float*>(sW + L.sumCatE * kEdgeC);   // [5][2C][32] + [5][2C]
  {
    int off = 0;
    for (int l = 0; l < kNL; ++l) {
      const float2* src = reinterpret_cast<const float2*>(P + L.p_edgeW) + L.offE[l];   // [c'][k]
      const int K = L.catE[l];
      for (int idx = threadIdx.x; idx < K * kEdgeC; idx += blockDim.x) {
        const int k = idx / kEdgeC, c = idx - k * kEdgeC;
        sW[off * kEdgeC + idx] = c < C ? src[c * K + k] : make_float2(0.f, 0.f);
      }
      off += K;
    }
    for (int idx = threadIdx.x; idx < kNL * C2 * kRadFeat; idx += blockDim.x) sRad[idx] = P[L.p_radW + idx];
    for (int idx = threadIdx.x; idx < kNL * C2; idx += blockDim.x) sRad[kNL * C2 * kRadFeat + idx] = P[L.p_radb + idx];
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  float dsc[kTrig], dph[kTrig];
  MGB_UNROLL
  for (int t = 0; t < kTrig; ++t) { dsc[t] = 0.f; dph[t] = 0.f; }
  if (p < total) {
    const PairId id = decode_pair(p, B, pair_off, n_atoms);
    const PairGeom g = pair_geom(pos + (long long)id.b * N * 3, id.i, id.j, d.cut_rad, d.cut_width);
    const long long pair = ((long long)id.b * N + id.i) * N + id.j;
    const bool pairable = (C & 1) == 0;   // even channel count: two complex channels per 16-byte store (rows are 8 C bytes)
    float f[kRadFeat], df[kRadFeat];
    rad_features_all(g, P + L.p_scales, P + L.p_phases, f, nullptr);
    const int l_begin = SPLIT ? (int)blockIdx.y : 0, l_end = SPLIT ? (int)blockIdx.y + 1 : kNL;
    MGB_UNROLL
    for (int t = 0; t < kRadFeat; ++t) df[t] = 0.f;
    if (l_begin == 0) {   // 16-byte stores: the thread's rows of the scratch are contiguous and 16-byte aligned
      MGB_UNROLL
      for (int t = 0; t < kRadFeat; t += 4)
        *reinterpret_cast<float4*>(sc.f + (long long)p * kRadFeat + t) = make_float4(f[t], f[t + 1], f[t + 2], f[t + 3]);
    }
    float2 dDacc[NLIN * kEdgeC];
    MGB_UNROLL
    for (int k = 0; k < NLIN * kEdgeC; ++k) dDacc[k] = make_float2(0.f, 0.f);
    int off = 0;
    for (int l = 0; l < l_begin; ++l) off += L.catE[l];
    for (int l = l_begin; l < l_end; ++l) {
      float2 dpre[kEdgeC];
      const float2* g_in = reinterpret_cast<const float2*>(dE) + pair * kNL * C + l * C;
      if (pairable) {
        MGB_UNROLL
        for (int c = 0; c < kEdgeC; c += 2) {
          const float4 v = c < C ? *reinterpret_cast<const float4*>(g_in + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          dpre[c] = make_float2(v.x * g.s, v.y * g.s);
          dpre[c + 1] = make_float2(v.z * g.s, v.w * g.s);
        }
      } else {
        MGB_UNROLL
        for (int c = 0; c < kEdgeC; ++c) dpre[c] = c < C ? make_float2(g_in[c].x * g.s, g_in[c].y * g.s) : make_float2(0.f, 0.f);
      }
      {
        float2* o = reinterpret_cast<float2*>(sc.dpre) + ((long long)p * kNL + l) * C;
        if (pairable) {
          MGB_UNROLL
          for (int c = 0; c < kEdgeC; c += 2)
            if (c < C) *reinterpret_cast<float4*>(o + c) = make_float4(dpre[c].x, dpre[c].y, dpre[c + 1].x, dpre[c + 1].y);
        } else {
          MGB_UNROLL
          for (int c = 0; c < kEdgeC; ++c)
            if (c < C) o[c] = dpre[c];
        }
      }
      const float2* Wl = sW + off * kEdgeC;
      int kk = 0;
      if (L.has_prev) {
        float2* o = reinterpret_cast<float2*>(dE_prev) + pair * kNL * C + l * C;
        if (pairable) {
          for (int k = 0; k < C; k += 2) {
            float2 a = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
            MGB_UNROLL
            for (int c = 0; c < kEdgeC; ++c) {
              cfmacl(a, Wl[(kk + k) * kEdgeC + c], dpre[c]);
              cfmacl(a2, Wl[(kk + k + 1) * kEdgeC + c], dpre[c]);
            }
            *reinterpret_cast<float4*>(o + k) = make_float4(a.x, a.y, a2.x, a2.y);
          }
        } else {
          for (int k = 0; k < C; ++k) {
            float2 a = make_float2(0.f, 0.f);
            MGB_UNROLL
            for (int c = 0; c < kEdgeC; ++c) cfmacl(a, Wl[(kk + k) * kEdgeC + c], dpre[c]);
            o[k] = a;
          }
        }
        kk += C;
      }
      if (l < NLIN) {
        MGB_UNROLL
        for (int lp = 0; lp < NLIN; ++lp)
          MGB_UNROLL
          for (int cc = 0; cc < kEdgeC; ++cc) {
            if (cc < C) {
              MGB_UNROLL
              for (int c = 0; c < kEdgeC; ++c) cfmacl(dDacc[lp * kEdgeC + cc], Wl[(kk + lp * C + cc) * kEdgeC + c], dpre[c]);
            }
          }
        kk += NLIN * C;
      }
      // radial part: R_l (forward value, recomputed) and its cotangent
      for (int k = 0; k < C; ++k) {
        float2 a = make_float2(0.f, 0.f);
        MGB_UNROLL
        for (int c = 0; c < kEdgeC; ++c) cfmacl(a, Wl[(kk + k) * kEdgeC + c], dpre[c]);
        const float* wr = sRad + ((l * C2) + 2 * k) * kRadFeat;
        float re = sRad[kNL * C2 * kRadFeat + l * C2 + 2 * k], im = sRad[kNL * C2 * kRadFeat + l * C2 + 2 * k + 1];
        MGB_UNROLL
        for (int t = 0; t < kRadFeat; ++t) {
          re = fmaf(wr[t], f[t], re);
          im = fmaf(wr[kRadFeat + t], f[t], im);
          df[t] = fmaf(wr[t], a.x, fmaf(wr[kRadFeat + t], a.y, df[t]));
        }
        *reinterpret_cast<float2*>(sc.R + ((long long)p * kNL + l) * C2 + 2 * k) = make_float2(re, im);
        *reinterpret_cast<float2*>(sc.dR + ((long long)p * kNL + l) * C2 + 2 * k) = a;
      }
      off += L.catE[l];
    }
    float2* od = reinterpret_cast<float2*>(dD) + (SPLIT ? (long long)blockIdx.y * slice_stride : 0ll) + pair * kNL * C;
    if (!SPLIT || l_begin < NLIN) {
      if (pairable) {
        MGB_UNROLL
        for (int lp = 0; lp < NLIN; ++lp)
          MGB_UNROLL
          for (int cc = 0; cc < kEdgeC; cc += 2)
            if (cc < C)
              *reinterpret_cast<float4*>(od + lp * C + cc) =
                  make_float4(dDacc[lp * kEdgeC + cc].x, dDacc[lp * kEdgeC + cc].y, dDacc[lp * kEdgeC + cc + 1].x, dDacc[lp * kEdgeC + cc + 1].y);
      } else {
        MGB_UNROLL
        for (int lp = 0; lp < NLIN; ++lp)
          MGB_UNROLL
          for (int cc = 0; cc < kEdgeC; ++cc)
            if (cc < C) od[lp * C + cc] = dDacc[lp * kEdgeC + cc];
      }
    }
    {   // f is no longer needed: reuse its registers for d f[t] / d arg_t = cos(arg) r^-p
      float tmp[kRadFeat];
      rad_features_all(g, P + L.p_scales, P + L.p_phases, tmp, f);
    }
    MGB_UNROLL
    for (int t = 0; t < kRadFeat; ++t) {
      const float gv = df[t] * f[t];
      dph[t >> 2] += gv;
      dsc[t >> 2] = fmaf(gv, kTwoPi * g.r, dsc[t >> 2]);
    }
  }
  // scale / phase cotangents: warp reduction, one atomic per warp and parameter
  edge_flush_scale_phase(dsc, dph, grad + L.p_scales, grad + L.p_phases);
}

// ------------------------------------------------------------------------------------------------------------
// Small minibatches (a few thousand pairs): the same two per-pair kernels with kEdgeCs threads per (pair, ell), so that
// the work fills the machine and every thread's serial chain is 5x shorter.
//   forward : thread (pair, ell, g) owns the output channels c' = g*kEdgeCg .. and the radial filters of the same indices
//             (exchanged through shared memory);
//   backward: thread (pair, ell, g) owns a fifth of every segment of dcat_l (previous edge / dot / radial) — the radial
//             feature cotangent is linear in the owned radial outputs, so the scale / phase cotangents need no exchange.
// ------------------------------------------------------------------------------------------------------------
constexpr int kEdgeCs = 5;                        // threads per (pair, ell)
constexpr int kEdgeCg = kEdgeC / kEdgeCs;         // channels per thread
constexpr int kPairCsPairs = 32;                  // pairs per CTA
constexpr int kPairCsThreads = kPairCsPairs * kEdgeCs;
constexpr int kEdgeKMaxFwd = 7 * kEdgeC;           // [E_prev (C) | dot (5 C) | radial (C)]

template <int NLIN>
__global__ void __launch_bounds__(kPairCsThreads)
k_edge_pairs_fwd_cs(const CovDesc* __restrict__ dp, int level, int B, const float* __restrict__ P, const float* __restrict__ Wt,
                    const float* __restrict__ pos, const int* __restrict__ n_atoms, const int* __restrict__ pair_off,
                    const float* __restrict__ D, const float* __restrict__ E_prev, float* __restrict__ E_out,
                    int* __restrict__ pair_slot) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int N = d.N, C = L.C, C2 = 2 * C, l = blockIdx.y, K = L.catE[l];
  const int total = pair_off[B];
  if ((int)(blockIdx.x * kPairCsPairs) >= total) return;
  MGB_DYN_SMEM(float2, smem);
  float2* sW = smem;                                          // [K][kEdgeC]   WE_l transposed ([k][c'])
  float2* sX = sW + kEdgeKMaxFwd * kEdgeC;                    // [pairs][K]    the pair's cat vector [E_prev | dot | radial]
  float* sRad = reinterpret_cast<float*>(sX + kPairCsPairs * kEdgeKMaxFwd);   // [2C][32] radial linear of this ell + [2C] bias
  {
    const float2* src = reinterpret_cast<const float2*>(Wt + d.wt_edge[level]) + L.offE[l];
    for (int idx = threadIdx.x; idx < K * kEdgeC; idx += blockDim.x) {
      const int k = idx / kEdgeC, c = idx - k * kEdgeC;
      sW[idx] = c < C ? src[k * C + c] : make_float2(0.f, 0.f);
    }
    for (int idx = threadIdx.x; idx < C2 * kRadFeat; idx += blockDim.x) sRad[idx] = P[L.p_radW + (long long)l * C2 * kRadFeat + idx];
    for (int idx = threadIdx.x; idx < C2; idx += blockDim.x) sRad[C2 * kRadFeat + idx] = P[L.p_radb + l * C2 + idx];
  }
  const int pl = threadIdx.x / kEdgeCs, g = threadIdx.x - pl * kEdgeCs;
  const int p = blockIdx.x * kPairCsPairs + pl;
  const bool valid = p < total;
  const int kprev = L.has_prev ? C : 0, kdot = (l < NLIN) ? NLIN * C : 0;
  PairGeom geo;
  long long pair = 0;
  float2* xrow = sX + pl * kEdgeKMaxFwd;
  constexpr int kXPer = (kEdgeKMaxFwd - kEdgeC + kEdgeCs - 1) / kEdgeCs;   // previous-edge + dot entries fetched per thread
  float2 xin[kXPer];
  if (valid) {
    const PairId id = decode_pair(p, B, pair_off, n_atoms);
    pair = ((long long)id.b * N + id.i) * N + id.j;
    // the pair's previous edge scalars and dot matrix: the five threads of the pair fetch interleaved entries, all loads in
    // flight before the radial arithmetic below
    const float2* xe = reinterpret_cast<const float2*>(E_prev) + pair * kNL * C + l * C;
    const float2* xd = reinterpret_cast<const float2*>(D) + pair * kNL * C;
    MGB_UNROLL
    for (int q = 0; q < kXPer; ++q) {
      const int k = g + q * kEdgeCs;
      if (k < kprev) xin[q] = xe[k];
      else if (k < kprev + kdot) xin[q] = xd[k - kprev];
    }
    geo = pair_geom(pos + (long long)id.b * N * 3, id.i, id.j, d.cut_rad, d.cut_width);
    if (pair_slot && l == 0 && g == 0) pair_slot[p] = (int)pair;
  }
  __syncthreads();   // weights staged
  if (valid) {
    float f[kRadFeat];
    rad_features_all(geo, P + L.p_scales, P + L.p_phases, f, nullptr);
    MGB_UNROLL
    for (int q = 0; q < kEdgeCg; ++q) {
      const int k = g * kEdgeCg + q;
      float re = 0.f, im = 0.f;
      if (k < C) {
        re = sRad[C2 * kRadFeat + 2 * k]; im = sRad[C2 * kRadFeat + 2 * k + 1];
        const float* wr = sRad + (2 * k) * kRadFeat;
        MGB_UNROLL
        for (int t = 0; t < kRadFeat; ++t) { re = fmaf(wr[t], f[t], re); im = fmaf(wr[kRadFeat + t], f[t], im); }
        xrow[kprev + kdot + k] = make_float2(re, im);
      }
    }
    MGB_UNROLL
    for (int q = 0; q < kXPer; ++q) {
      const int k = g + q * kEdgeCs;
      if (k < kprev + kdot) xrow[k] = xin[q];
    }
  }
  __syncthreads();
  if (!valid) return;
  float2 acc[kEdgeCg];
  MGB_UNROLL
  for (int q = 0; q < kEdgeCg; ++q) acc[q] = make_float2(0.f, 0.f);
  const float2* w = sW + g * kEdgeCg;
#pragma unroll 5
  for (int k = 0; k < K; ++k) {
    const float2 xv = xrow[k];
    MGB_UNROLL
    for (int q = 0; q < kEdgeCg; ++q) cfma(acc[q], w[k * kEdgeC + q], xv);
  }
  float2* Eo = reinterpret_cast<float2*>(E_out) + pair * kNL * C + l * C;
  MGB_UNROLL
  for (int q = 0; q < kEdgeCg; ++q)
    if (g * kEdgeCg + q < C) Eo[g * kEdgeCg + q] = make_float2(acc[q].x * geo.s, acc[q].y * geo.s);
}

template <int NLIN>
__global__ void __launch_bounds__(kPairCsThreads)
k_edge_pairs_bwd_cs(const CovDesc* __restrict__ dp, int level, int B, const float* __restrict__ P, const float* __restrict__ pos,
                    const int* __restrict__ n_atoms, const int* __restrict__ pair_off, const float* __restrict__ dE,
                    float* __restrict__ dE_prev, float* __restrict__ dD, long long slice_stride /* complex */, EdgeScratch sc,
                    float* __restrict__ grad) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int N = d.N, C = L.C, C2 = 2 * C, l = blockIdx.y, K = L.catE[l];
  const int total = pair_off[B];
  if ((int)(blockIdx.x * kPairCsPairs) >= total) return;
  MGB_DYN_SMEM(float2, smem);
  float2* sW = smem;                                              // [K][kEdgeC]: WE_l of the REFERENCE layout read as [k][c'], padded
  float* sRad = reinterpret_cast<float*>(sW + K * kEdgeC);        // [2C][32] + [2C]
  {
    const float2* src = reinterpret_cast<const float2*>(P + L.p_edgeW) + L.offE[l];   // [c'][k]
    for (int idx = threadIdx.x; idx < K * kEdgeC; idx += blockDim.x) {
      const int k = idx / kEdgeC, c = idx - k * kEdgeC;
      sW[idx] = c < C ? src[c * K + k] : make_float2(0.f, 0.f);
    }
    for (int idx = threadIdx.x; idx < C2 * kRadFeat; idx += blockDim.x) sRad[idx] = P[L.p_radW + (long long)l * C2 * kRadFeat + idx];
    for (int idx = threadIdx.x; idx < C2; idx += blockDim.x) sRad[C2 * kRadFeat + idx] = P[L.p_radb + l * C2 + idx];
  }
  __syncthreads();
  const int pl = threadIdx.x / kEdgeCs, g = threadIdx.x - pl * kEdgeCs;
  const int p = blockIdx.x * kPairCsPairs + pl;
  float dsc[kTrig], dph[kTrig];
  MGB_UNROLL
  for (int t = 0; t < kTrig; ++t) { dsc[t] = 0.f; dph[t] = 0.f; }
  if (p < total) {
    const PairId id = decode_pair(p, B, pair_off, n_atoms);
    const PairGeom geo = pair_geom(pos + (long long)id.b * N * 3, id.i, id.j, d.cut_rad, d.cut_width);
    const long long pair = ((long long)id.b * N + id.i) * N + id.j;
    float f[kRadFeat], df[kRadFeat];
    rad_features_all(geo, P + L.p_scales, P + L.p_phases, f, nullptr);
    MGB_UNROLL
    for (int t = 0; t < kRadFeat; ++t) {
      df[t] = 0.f;
      if (l == 0 && g == 0) sc.f[(long long)p * kRadFeat + t] = f[t];
    }
    float2 dpre[kEdgeC];
    const float2* g_in = reinterpret_cast<const float2*>(dE) + pair * kNL * C + l * C;
    MGB_UNROLL
    for (int c = 0; c < kEdgeC; ++c) {
      dpre[c] = c < C ? make_float2(g_in[c].x * geo.s, g_in[c].y * geo.s) : make_float2(0.f, 0.f);
      if (c < C && g == 0) reinterpret_cast<float2*>(sc.dpre)[((long long)p * kNL + l) * C + c] = dpre[c];
    }
    int kk = 0;
    if (L.has_prev) {
      float2* o = reinterpret_cast<float2*>(dE_prev) + pair * kNL * C + l * C;
      MGB_UNROLL
      for (int q = 0; q < kEdgeCg; ++q) {
        const int k = g * kEdgeCg + q;
        if (k < C) {
          float2 a = make_float2(0.f, 0.f);
          MGB_UNROLL
          for (int c = 0; c < kEdgeC; ++c) cfmacl(a, sW[(kk + k) * kEdgeC + c], dpre[c]);
          o[k] = a;
        }
      }
      kk += C;
    }
    if (l < NLIN) {
      float2* od = reinterpret_cast<float2*>(dD) + (long long)l * slice_stride + pair * kNL * C;
      const int per = (NLIN * C + kEdgeCs - 1) / kEdgeCs;
      const int k0 = g * per, k1 = min(NLIN * C, k0 + per);
      for (int k = k0; k < k1; ++k) {
        float2 a = make_float2(0.f, 0.f);
        MGB_UNROLL
        for (int c = 0; c < kEdgeC; ++c) cfmacl(a, sW[(kk + k) * kEdgeC + c], dpre[c]);
        od[k] = a;
      }
      kk += NLIN * C;
    }
    MGB_UNROLL
    for (int q = 0; q < kEdgeCg; ++q) {
      const int k = g * kEdgeCg + q;
      if (k < C) {
        float2 a = make_float2(0.f, 0.f);
        MGB_UNROLL
        for (int c = 0; c < kEdgeC; ++c) cfmacl(a, sW[(kk + k) * kEdgeC + c], dpre[c]);
        const float* wr = sRad + (2 * k) * kRadFeat;
        float re = sRad[C2 * kRadFeat + 2 * k], im = sRad[C2 * kRadFeat + 2 * k + 1];
        MGB_UNROLL
        for (int t = 0; t < kRadFeat; ++t) {
          re = fmaf(wr[t], f[t], re);
          im = fmaf(wr[kRadFeat + t], f[t], im);
          df[t] = fmaf(wr[t], a.x, fmaf(wr[kRadFeat + t], a.y, df[t]));
        }
        *reinterpret_cast<float2*>(sc.R + ((long long)p * kNL + l) * C2 + 2 * k) = make_float2(re, im);
        *reinterpret_cast<float2*>(sc.dR + ((long long)p * kNL + l) * C2 + 2 * k) = a;
      }
    }
    {   // f is no longer needed: reuse its registers for d f[t] / d arg_t = cos(arg) r^-p
      float tmp[kRadFeat];
      rad_features_all(geo, P + L.p_scales, P + L.p_phases, tmp, f);
    }
    MGB_UNROLL
    for (int t = 0; t < kRadFeat; ++t) {
      const float gv = df[t] * f[t];
      dph[t >> 2] += gv;
      dsc[t >> 2] = fmaf(gv, kTwoPi * geo.r, dsc[t >> 2]);
    }
  }
  edge_flush_scale_phase(dsc, dph, grad + L.p_scales, grad + L.p_phases);
}

// ------------------------------------------------------------------------------------------------------------
// Backward, weight cotangents: reductions over pairs, run as small GEMMs out of shared memory.
//   dWE_l[c'][k]   += sum_p conj(cat_l[p][k]) dpre_l[p][c']       cat_l[p] = [E_prev[p][l] | D[p] | R[p][l]]
//   dWrad_l[o][t]  += sum_p dR_l[p][o] f[p][t],   db_l[o] += sum_p dR_l[p][o]
// grid = (pair chunks, 5 ells), 128 threads.  A CTA walks its chunk in tiles of kEdgeDwTile pairs: the tile's cat rows
// (gathered through pair_slot, coalesced over k), dpre, dR and f are staged in shared memory, then
//   threads 0 .. K_l-1 (warps 0-2) own one k: 10 complex accumulators, x from shared memory (lanes over k), dpre as broadcasts;
//   warp 3 owns the radial features, lane = t: 2C real accumulators (+ the bias of output o = lane).
// One flush of atomics per CTA at the end.
// ------------------------------------------------------------------------------------------------------------
constexpr int kEdgeDwThreads = 128;
constexpr int kEdgeDwTile = 32;
constexpr int kEdgeKMax = 7 * kEdgeC;   // [E_prev (C) | dot (5 C) | radial (C)]
static_assert(kRadFeat == 32, "warp 3 of k_edge_dw maps one lane to one radial feature");

// shared memory of one staging buffer (floats): cat tile + dpre + dR + f
constexpr int kEdgeDwBufFloats = kEdgeDwTile * (2 * kEdgeKMax + 2 * kEdgeC + 2 * kEdgeC + kRadFeat);

template <int NLIN>
__global__ void __launch_bounds__(kEdgeDwThreads)
k_edge_dw(const CovDesc* __restrict__ dp, int level, int B, const int* __restrict__ pair_off, const int* __restrict__ pair_slot,
          const float* __restrict__ E_prev, const float* __restrict__ D, EdgeScratch sc, float* __restrict__ grad) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int C = L.C, C2 = 2 * C, l = blockIdx.y, K = L.catE[l];
  const int total = pair_off[B];
  const int per = max((int)((total + gridDim.x - 1) / gridDim.x), kEdgeDwTile);   // at least one full tile per CTA: fewer atomics
  const int p_begin = per * blockIdx.x, p_end = min(total, p_begin + per);
  if (p_begin >= p_end) return;
  // two staging buffers filled with cp.async: the next tile of pairs is in flight while the current one is reduced
  MGB_DYN_SMEM(float, smem);
  const int tid = threadIdx.x;
  const bool edge_thread = tid < K, rad_thread = tid >= 96;
  const int kprev = L.has_prev ? C : 0, kdot = (l < NLIN) ? NLIN * C : 0;
  auto cat_of = [&](int buf) { return reinterpret_cast<float2*>(smem + buf * kEdgeDwBufFloats); };                       // [tile][kEdgeKMax]
  auto dpre_of = [&](int buf) { return cat_of(buf) + kEdgeDwTile * kEdgeKMax; };                                          // [tile][kEdgeC]
  auto dR_of = [&](int buf) { return reinterpret_cast<float*>(dpre_of(buf) + kEdgeDwTile * kEdgeC); };                    // [tile][2 kEdgeC]
  auto f_of = [&](int buf) { return dR_of(buf) + kEdgeDwTile * 2 * kEdgeC; };                                             // [tile][32]
  // channels beyond C stay zero for the whole kernel
  for (int idx = tid; idx < 2 * kEdgeDwTile * kEdgeC; idx += blockDim.x) {
    const int buf = idx / (kEdgeDwTile * kEdgeC), r = idx - buf * kEdgeDwTile * kEdgeC;
    dpre_of(buf)[r] = make_float2(0.f, 0.f);
    dR_of(buf)[2 * r] = 0.f;
    dR_of(buf)[2 * r + 1] = 0.f;
  }
  __syncthreads();
  const float2* Ep = reinterpret_cast<const float2*>(E_prev);
  const float2* Dp = reinterpret_cast<const float2*>(D);
  const float2* Rp = reinterpret_cast<const float2*>(sc.R);
  // stage one tile: thread k < K copies column k of the tile's cat rows (no index arithmetic beyond the pair slot), warp 3
  // copies dpre / dR / f
  auto stage = [&](int p0, int np, int buf) {
    if (edge_thread) {
      float2* dst = cat_of(buf) + tid;
      if (tid < kprev) {
        const float2* src = Ep + l * C + tid;
        for (int q = 0; q < np; ++q) cp_async8(dst + q * kEdgeKMax, src + (long long)pair_slot[p0 + q] * kNL * C);
      } else if (tid < kprev + kdot) {
        const float2* src = Dp + (tid - kprev);
        for (int q = 0; q < np; ++q) cp_async8(dst + q * kEdgeKMax, src + (long long)pair_slot[p0 + q] * kNL * C);
      } else {
        const float2* src = Rp + ((long long)p0 * kNL + l) * C + (tid - kprev - kdot);
        for (int q = 0; q < np; ++q) cp_async8(dst + q * kEdgeKMax, src + (long long)q * kNL * C);
      }
    } else if (rad_thread) {
      const int lane = tid - 96;
      const float2* sp = reinterpret_cast<const float2*>(sc.dpre) + ((long long)p0 * kNL + l) * C;
      const float2* sr = reinterpret_cast<const float2*>(sc.dR) + ((long long)p0 * kNL + l) * C;
      for (int idx = lane; idx < np * C; idx += 32) {
        const int q = idx / C, c = idx - q * C;
        cp_async8(dpre_of(buf) + q * kEdgeC + c, sp + (long long)q * kNL * C + c);
        cp_async8(reinterpret_cast<float2*>(dR_of(buf)) + q * kEdgeC + c, sr + (long long)q * kNL * C + c);
      }
      const float2* sf = reinterpret_cast<const float2*>(sc.f + (long long)p0 * kRadFeat);
      for (int idx = lane; idx < np * (kRadFeat / 2); idx += 32) cp_async8(reinterpret_cast<float2*>(f_of(buf)) + idx, sf + idx);
    }
    cp_async_commit();
  };
  float2 acc[kEdgeC];
  MGB_UNROLL
  for (int c = 0; c < kEdgeC; ++c) acc[c] = make_float2(0.f, 0.f);
  float racc[2 * kEdgeC], bacc = 0.f;
  MGB_UNROLL
  for (int o = 0; o < 2 * kEdgeC; ++o) racc[o] = 0.f;
  stage(p_begin, min(kEdgeDwTile, p_end - p_begin), 0);
  int buf = 0;
  for (int p0 = p_begin; p0 < p_end; p0 += kEdgeDwTile, buf ^= 1) {
    const int np = min(kEdgeDwTile, p_end - p0);
    cp_async_wait_all();
    __syncthreads();   // this tile is visible to everyone, and everyone is done with the other buffer
    if (p0 + kEdgeDwTile < p_end) stage(p0 + kEdgeDwTile, min(kEdgeDwTile, p_end - p0 - kEdgeDwTile), buf ^ 1);
    if (edge_thread) {
      const float2* s_cat = cat_of(buf) + tid;
      const float2* s_dpre = dpre_of(buf);
#pragma unroll 2
      for (int q = 0; q < np; ++q) {
        const float2 x = s_cat[q * kEdgeKMax];
        MGB_UNROLL
        for (int c = 0; c < kEdgeC; ++c) cfmacl(acc[c], x, s_dpre[q * kEdgeC + c]);
      }
    } else if (rad_thread) {
      const int t = tid - 96;
      const float* s_dR = dR_of(buf);
      const float* s_f = f_of(buf);
#pragma unroll 2
      for (int q = 0; q < np; ++q) {
        const float f = s_f[q * kRadFeat + t];
        MGB_UNROLL
        for (int o = 0; o < 2 * kEdgeC; ++o) racc[o] = fmaf(s_dR[q * 2 * kEdgeC + o], f, racc[o]);
        if (t < 2 * kEdgeC) bacc += s_dR[q * 2 * kEdgeC + t];
      }
    }
  }
  if (edge_thread) {
    MGB_UNROLL
    for (int c = 0; c < kEdgeC; ++c) {
      if (c < C && (acc[c].x != 0.f || acc[c].y != 0.f))
        atomic_add2(reinterpret_cast<float2*>(grad + L.p_edgeW) + L.offE[l] + c * K + tid, acc[c]);
    }
  } else if (rad_thread) {
    const int t = tid - 96;
    MGB_UNROLL
    for (int o = 0; o < 2 * kEdgeC; ++o)
      if (o < C2 && racc[o] != 0.f) atomicAdd(grad + L.p_radW + ((long long)l * C2 + o) * kRadFeat + t, racc[o]);
    if (t < C2 && bacc != 0.f) atomicAdd(grad + L.p_radb + l * C2 + t, bacc);
  }
}

}  // namespace mgb
