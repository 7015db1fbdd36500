// cov_forward.cuh — forward kernels of the covariant (Cormorant) actor-critic body.
//
// What is computed follows the reference's CovariantAC.step (molgym/agents/covariant/agent.py:209-220) →
// Cormorant.forward (molgym/agents/covariant/modules.py:97-135) → cormorant's CormorantCG levels
// (restated in oracle/thirdparty/cormorant/models/cormorant_cg.py).  How it is computed is new: one CTA per
// atom, spherical harmonics / radial features recomputed on the fly, the Kronecker sums kept in registers,
// Clebsch-Gordan contraction and channel mixing done out of shared memory.
#pragma once
#include "model.cuh"

namespace mgb {

// ------------------------------------------------------------------------------------------------------------
// Per-step parameter preparation: transposed copies of the weights that forward kernels read "lanes over outputs".
// segment s: src [rows][cols][elem] -> dst [cols][rows][elem]
// ------------------------------------------------------------------------------------------------------------
struct TransposeSeg {
  long long src, dst;  // float offsets (params / scratch)
  int rows, cols, elem;
};

__global__ void k_prep_params(const TransposeSeg* __restrict__ segs, const float* __restrict__ P, float* __restrict__ Wt) {
  const TransposeSeg s = segs[blockIdx.x];
  const int n = s.rows * s.cols;
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
    const int r = idx / s.cols, c = idx % s.cols;
    for (int e = 0; e < s.elem; ++e) Wt[s.dst + ((long long)c * s.rows + r) * s.elem + e] = P[s.src + (long long)idx * s.elem + e];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Input featurisation + InputLinear (covariant/modules.py:116-135 + cormorant InputLinear): one CTA per canvas.
// Writes n_atoms[b], X[b,i,S_in] (kept for the weight gradient) and A0[b,i,c] (complex, zero on padded atoms).
// ------------------------------------------------------------------------------------------------------------
__global__ void k_input_fwd(const CovDesc* __restrict__ dp, const float* __restrict__ P, const int* __restrict__ charges,
                            const float* __restrict__ bags, int* __restrict__ n_atoms, float* __restrict__ X,
                            float* __restrict__ A0) {
  const CovDesc& d = *dp;
  const int b = blockIdx.x, N = d.N, Z = d.Z, S = d.S_in, C2 = 2 * d.C;
  MGB_DYN_SMEM(float, sx);  // [N][S]
  __shared__ int s_n;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 0; i < N; ++i) n += charges[b * N + i] > 0 ? 1 : 0;
    s_n = n;
    n_atoms[b] = n;
  }
  for (int idx = threadIdx.x; idx < N * S; idx += blockDim.x) {
    const int i = idx / S, s = idx % S;
    const int q = charges[b * N + i];
    float v;
    if (s < 3 * Z) {
      const int z = s / 3, p = s % 3;
      const float qs = (float)q / d.charge_scale;
      const float pw = p == 0 ? 1.f : (p == 1 ? qs : qs * qs);
      v = (q == d.zs[z]) ? pw : 0.f;
    } else {
      v = bags[b * Z + (s - 3 * Z)] / d.bag_scale;
    }
    sx[idx] = v;
    X[(long long)b * N * S + idx] = v;
  }
  __syncthreads();
  const int n = s_n;
  for (int idx = threadIdx.x; idx < N * C2; idx += blockDim.x) {
    const int i = idx / C2, o = idx % C2;
    float acc = 0.f;
    if (i < n) {
      acc = P[d.p_inb + o];
      const float* w = P + d.p_inW + (long long)o * S;
      for (int s = 0; s < S; ++s) acc = fmaf(w[s], sx[i * S + s], acc);
    }
    A0[(long long)b * N * C2 + idx] = acc;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Pair geometry shared by the edge/atom kernels.
// ------------------------------------------------------------------------------------------------------------
struct PairGeom {
  float dx, dy, dz, r;
  float s;     // soft cutoff sigmoid((rc - r)/w)   (cormorant MaskLevel 'soft', agent.py:66-69)
  bool mrad;   // r > 0 (RadPolyTrig mask)
};
__device__ __forceinline__ PairGeom pair_geom(const float* __restrict__ pos_b, int i, int j, float cut_rad, float cut_width) {
  PairGeom g;
  g.dx = pos_b[i * 3 + 0] - pos_b[j * 3 + 0];
  g.dy = pos_b[i * 3 + 1] - pos_b[j * 3 + 1];
  g.dz = pos_b[i * 3 + 2] - pos_b[j * 3 + 2];
  g.r = sqrtf(g.dx * g.dx + g.dy * g.dy + g.dz * g.dz);
  g.mrad = g.r > 0.f;
  g.s = 1.f / (1.f + expf(-(cut_rad - g.r) / cut_width));
  return g;
}

// Radial basis feature t = trig*4 + p : sin(2 pi scale r + phase) * r^-p  (cormorant RadPolyTrig, basis_set=(3,3))
__device__ __forceinline__ float rad_feature(int t, const PairGeom& g, const float* __restrict__ scales,
                                             const float* __restrict__ phases, float* dval_darg) {
  const int tt = t >> 2, p = t & 3;
  if (!g.mrad) {
    if (dval_darg) *dval_darg = 0.f;
    return 0.f;
  }
  const float arg = __fadd_rn(__fmul_rn(__fmul_rn(kTwoPi, scales[tt]), g.r), phases[tt]);
  const float inv = 1.f / g.r;
  const float pw = p == 0 ? 1.f : (p == 1 ? inv : (p == 2 ? inv * inv : inv * inv * inv));
  if (dval_darg) *dval_darg = cosf(arg) * pw;
  return sinf(arg) * pw;
}

// ------------------------------------------------------------------------------------------------------------
// Edge level: E[b,i,j,l,c'] = s_ij * sum_k WE_l[c',k] catE_ijl[k],  catE = [E_prev | dot(A_i, A_j) | radial_l]
// (cormorant CormorantEdgeLevel: DotMatrix + CatMixRepsScalar + MaskLevel).  One CTA per (b, i), one warp per j.
// ------------------------------------------------------------------------------------------------------------
constexpr int kEdgeThreads = 128;

// Fills the per-warp cat buffer for one pair; returns with the warp synchronised.  catbuf: [sumCatE] complex,
// l-major.  f: [32] radial features (written here).  Also used by the backward kernel.
template <int NLIN>
__device__ __forceinline__ void edge_build_cat(const LevelDesc& L, const float* __restrict__ P, const float* __restrict__ Wt_rad,
                                               const PairGeom& g, const float2* __restrict__ sAi,
                                               const float2* __restrict__ Aj, const float2* __restrict__ Eprev_ij,
                                               float2* catbuf, float* f, int lane) {
  const int C = L.C;
  f[lane] = rad_feature(lane, g, P + L.p_scales, P + L.p_phases, nullptr);
  __syncwarp();
  // radial filters: R_l[o] = b_l[o] + sum_t W_l[o][t] f[t]   (Wt_rad: [l][t][2C])
  const int C2 = 2 * C;
  int off_l[kNL];
  {
    int o = 0;
    for (int l = 0; l < kNL; ++l) { off_l[l] = o; o += L.catE[l]; }
  }
  for (int idx = lane; idx < kNL * C2; idx += 32) {
    const int l = idx / C2, o = idx % C2;
    float acc = P[L.p_radb + l * C2 + o];
    const float* w = Wt_rad + (long long)l * kRadFeat * C2 + o;
    for (int t = 0; t < kRadFeat; ++t) acc = fmaf(w[t * C2], f[t], acc);
    const int krad = L.catE[l] - C;  // radial block is last
    reinterpret_cast<float*>(catbuf + off_l[l] + krad)[o] = acc;
  }
  // dot matrix D[l',c] = sum_m (-1)^m A_i[l',m,c] A_j[l',-m,c]
  for (int idx = lane; idx < NLIN * C; idx += 32) {
    const int lp = idx / C, c = idx % C;
    float2 acc = make_float2(0.f, 0.f);
    for (int m = -lp; m <= lp; ++m) {
      const float2 a = sAi[lm_index(lp, m) * C + c];
      const float2 bj = Aj[lm_index(lp, -m) * C + c];
      float2 pr = cmul(a, bj);
      if (m & 1) { acc.x -= pr.x; acc.y -= pr.y; } else { acc.x += pr.x; acc.y += pr.y; }
    }
    const int kdot = L.has_prev ? C : 0;
    for (int l = 0; l < NLIN; ++l) catbuf[off_l[l] + kdot + idx] = acc;
  }
  if (L.has_prev) {
    for (int idx = lane; idx < kNL * C; idx += 32) {
      const int l = idx / C, c = idx % C;
      catbuf[off_l[l] + c] = Eprev_ij[idx];
    }
  }
  __syncwarp();
}

template <int NLIN>
__global__ void __launch_bounds__(kEdgeThreads)
k_edge_fwd(const CovDesc* __restrict__ dp, int level, const float* __restrict__ P, const float* __restrict__ Wt,
           const float* __restrict__ pos, const int* __restrict__ n_atoms, const float* __restrict__ A_in,
           const float* __restrict__ E_prev, float* __restrict__ E_out) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int N = d.N, C = L.C;
  const int b = blockIdx.x / N, i = blockIdx.x % N;
  const int n = n_atoms[b];
  if (i >= n) return;
  constexpr int NLM = NLIN * NLIN;
  MGB_DYN_SMEM(float2, smem);
  float2* sAi = smem;                                  // [NLM][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float2* catbuf = sAi + NLM * C + warp * (L.sumCatE + 16);
  float* f = reinterpret_cast<float*>(catbuf + L.sumCatE);
  const float2* Ab = reinterpret_cast<const float2*>(A_in) + (long long)b * N * NLM * C;
  for (int idx = threadIdx.x; idx < NLM * C; idx += blockDim.x) sAi[idx] = Ab[(long long)i * NLM * C + idx];
  __syncthreads();
  const float* Wt_rad = Wt + d.wt_edge[level] + 2ll * L.totE;  // radial transposes follow the edge weights
  const float2* WEt = reinterpret_cast<const float2*>(Wt + d.wt_edge[level]);
  const float* pos_b = pos + (long long)b * N * 3;
  for (int j = warp; j < n; j += nwarps) {
    const PairGeom g = pair_geom(pos_b, i, j, d.cut_rad, d.cut_width);
    const long long pair = ((long long)b * N + i) * N + j;
    const float2* Eprev_ij = L.has_prev ? reinterpret_cast<const float2*>(E_prev) + pair * kNL * C : nullptr;
    edge_build_cat<NLIN>(L, P, Wt_rad, g, sAi, Ab + (long long)j * NLM * C, Eprev_ij, catbuf, f, lane);
    float2* Eo = reinterpret_cast<float2*>(E_out) + pair * kNL * C;
    int off = 0, l = 0, nextl = C;
    for (int idx = lane; idx < kNL * C; idx += 32) {
      l = idx / C;
      off = 0;
      for (int q = 0; q < l; ++q) off += L.catE[q];
      (void)nextl;
      const int cp = idx % C;
      const float2* w = WEt + L.offE[l] + cp;  // [k][c']
      const float2* x = catbuf + off;
      float2 acc = make_float2(0.f, 0.f);
      const int K = L.catE[l];
      for (int k = 0; k < K; ++k) cfma(acc, w[k * C], x[k]);
      Eo[idx] = make_float2(acc.x * g.s, acc.y * g.s);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------------
// Channel mixing out of shared memory: out[l, m, c'] = sum_k W_l[c', k] cat_l[m][k]  (complex), one warp per unit of
// <= NM rows, lanes over k, CO output channels per pass.  W in the reference layout [c'][k][2] (coalesced over k).
// ------------------------------------------------------------------------------------------------------------
template <int CO, int NM>
__device__ __forceinline__ void mix_rows(const MixUnit* __restrict__ units, int n_units, const int* catA, const int* offA,
                                         const int* offW, int Cout, const float2* __restrict__ W,
                                         const float2* __restrict__ sCat, float2* __restrict__ out /* [25][Cout] */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int u = warp; u < n_units; u += nwarps) {
    const MixUnit un = units[u];
    const int K = catA[un.l];
    const float2* cat0 = sCat + offA[un.l] + un.m0 * K;
    for (int c0 = 0; c0 < Cout; c0 += CO) {
      float2 acc[NM][CO];
      MGB_UNROLL
      for (int q = 0; q < NM; ++q)
        MGB_UNROLL
        for (int c = 0; c < CO; ++c) acc[q][c] = make_float2(0.f, 0.f);
      const float2* Wl = W + offW[un.l] + (long long)c0 * K;
      for (int k = lane; k < K; k += 32) {
        float2 x[NM];
        MGB_UNROLL
        for (int q = 0; q < NM; ++q) x[q] = (q < un.nm) ? cat0[q * K + k] : make_float2(0.f, 0.f);
        MGB_UNROLL
        for (int c = 0; c < CO; ++c) {
          if (c0 + c < Cout) {
            const float2 w = Wl[c * K + k];
            MGB_UNROLL
            for (int q = 0; q < NM; ++q) cfma(acc[q][c], w, x[q]);
          }
        }
      }
      MGB_UNROLL
      for (int q = 0; q < NM; ++q)
        MGB_UNROLL
        for (int c = 0; c < CO; ++c) {
          float vx = warp_sum(acc[q][c].x), vy = warp_sum(acc[q][c].y);
          if (lane == 0 && q < un.nm && c0 + c < Cout)
            out[(lm_index(un.l, -un.l) + un.m0 + q) * Cout + c0 + c] = make_float2(vx, vy);
        }
    }
  }
}

// CG gather out of shared memory: cat[dest(o), c] = sum_terms coef * (T[lm1][lm2][c])   or  A[lm1][c]*A[lm2][c]
// SQUARE=false: sT is [n_pair][C];  SQUARE=true: sT is the rep A [nlm][C] and the pair product is formed on the fly.
template <bool SQUARE>
__device__ __forceinline__ void cg_gather(const CgTable& t, int C, const float2* __restrict__ sT, const int* catA,
                                          const int* offA, int block0_of_l[kNL], float2* __restrict__ sCat, float scale) {
  const int total = t.n_out * C;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int o = idx / C, c = idx % C;
    float2 acc = make_float2(0.f, 0.f);
    const int t0 = t.term_start[o], t1 = t.term_start[o + 1];
    for (int q = t0; q < t1; ++q) {
      const float cf = t.term_coef[q];
      float2 v;
      if (SQUARE) v = cmul(sT[t.term_lm1[q] * C + c], sT[t.term_lm2[q] * C + c]);
      else v = sT[(t.term_lm1[q] * t.nlm2 + t.term_lm2[q]) * C + c];
      acc.x = fmaf(cf, v.x, acc.x);
      acc.y = fmaf(cf, v.y, acc.y);
    }
    const int l = t.out_l[o];
    sCat[offA[l] + t.out_m[o] * catA[l] + (block0_of_l[l] + t.out_block[o]) * C + c] = make_float2(acc.x * scale, acc.y * scale);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Atom level (cormorant CormorantAtomLevel): for atom i
//   T[c, lm1, lm2] = sum_j E_ij[l1, c] Y_lm1(r_ij) A_j[lm2, c]           (registers, thread = (lm1, c))
//   ag = CG T ; sq = CG (A_i x A_i) ; cat_l = [ag | A_i | sq] ; A_out[l, m, c'] = sum_k W_l[c', k] cat_l[m][k]
// One CTA per (b, i).  The cat vector is also written to HBM: the weight-gradient kernel reads it back.
// ------------------------------------------------------------------------------------------------------------
constexpr int kAtomThreads = 256;
constexpr int kJChunk = 8;

__host__ __device__ inline int atom_smem_floats(const LevelDesc& L) {
  const int nlm2 = L.nlm_in;
  const int stage = kJChunk * (kM + kNL * L.C + nlm2 * L.C) * 2;
  const int tsz = kM * nlm2 * L.C * 2;
  return (stage > tsz ? stage : tsz) + L.totA * 2 + nlm2 * L.C * 2;
}

// Stage one chunk of neighbours j0..j0+nj-1 of atom i into shared memory: Y_ij (conj, 'unit' norm, un-normalised
// argument: SphericalHarmonicsRel(conj=True), covariant/modules.py:52-56), E_ij, A_j.
template <int NLM2>
__device__ __forceinline__ void stage_neighbours(const CovDesc& d, const LevelDesc& L, const float* __restrict__ pos_b,
                                                 const float2* __restrict__ Ab, const float2* __restrict__ E_i, int i, int j0,
                                                 int nj, float2* sY, float2* sE, float2* sAj) {
  const int C = L.C;
  if ((int)threadIdx.x < nj) {
    const int j = j0 + threadIdx.x;
    float2 y[kM];
    sph_harm_l4(pos_b[i * 3 + 0] - pos_b[j * 3 + 0], pos_b[i * 3 + 1] - pos_b[j * 3 + 1], pos_b[i * 3 + 2] - pos_b[j * 3 + 2], true,
                true, y);
    for (int q = 0; q < kM; ++q) sY[threadIdx.x * kM + q] = y[q];
  }
  for (int idx = threadIdx.x; idx < nj * kNL * C; idx += blockDim.x) sE[idx] = E_i[(long long)j0 * kNL * C + idx];
  for (int idx = threadIdx.x; idx < nj * NLM2 * C; idx += blockDim.x) sAj[idx] = Ab[(long long)j0 * NLM2 * C + idx];
}

template <int NLM2, int CO, int NM>
__global__ void __launch_bounds__(kAtomThreads)
k_atom_fwd(const CovDesc* __restrict__ dp, int level, const float* __restrict__ P, const float* __restrict__ pos,
           const int* __restrict__ n_atoms, const float* __restrict__ A_in, const float* __restrict__ E,
           float* __restrict__ cat_out, float* __restrict__ A_out) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int N = d.N, C = L.C;
  const int b = blockIdx.x / N, i = blockIdx.x % N;
  const int n = n_atoms[b];
  if (i >= n) return;
  MGB_DYN_SMEM(float2, smem);
  const int stage = kJChunk * (kM + kNL * C + NLM2 * C), tsz = kM * NLM2 * C;
  float2* sT = smem;
  float2* sY = smem;
  float2* sE = sY + kJChunk * kM;
  float2* sAj = sE + kJChunk * kNL * C;
  float2* sCat = smem + (stage > tsz ? stage : tsz);
  float2* sAi = sCat + L.totA;
  const float2* Ab = reinterpret_cast<const float2*>(A_in) + (long long)b * N * NLM2 * C;
  const float2* E_i = reinterpret_cast<const float2*>(E) + ((long long)b * N + i) * N * kNL * C;
  const float* pos_b = pos + (long long)b * N * 3;
  for (int idx = threadIdx.x; idx < NLM2 * C; idx += blockDim.x) sAi[idx] = Ab[(long long)i * NLM2 * C + idx];

  const bool owner = (int)threadIdx.x < kM * C;
  const int lm1 = owner ? threadIdx.x / C : 0, c = owner ? threadIdx.x % C : 0;
  const int l1 = ell_of_lm(lm1);
  float2 acc[NLM2];
  MGB_UNROLL
  for (int q = 0; q < NLM2; ++q) acc[q] = make_float2(0.f, 0.f);
  for (int j0 = 0; j0 < n; j0 += kJChunk) {
    const int nj = min(kJChunk, n - j0);
    __syncthreads();
    stage_neighbours<NLM2>(d, L, pos_b, Ab, E_i, i, j0, nj, sY, sE, sAj);
    __syncthreads();
    if (owner) {
      for (int jj = 0; jj < nj; ++jj) {
        const float2 u = cmul(sE[(jj * kNL + l1) * C + c], sY[jj * kM + lm1]);
        const float2* a = sAj + jj * NLM2 * C + c;
        MGB_UNROLL
        for (int q = 0; q < NLM2; ++q) cfma(acc[q], u, a[q * C]);
      }
    }
  }
  __syncthreads();
  if (owner) {
    MGB_UNROLL
    for (int q = 0; q < NLM2; ++q) sT[(lm1 * NLM2 + q) * C + c] = acc[q];
  }
  __syncthreads();
  int zero_blocks[kNL] = {0, 0, 0, 0, 0};
  int sq_blocks[kNL];
  for (int l = 0; l < kNL; ++l) sq_blocks[l] = L.sq_block[l];
  cg_gather<false>(L.ag, C, sT, L.catA, L.offA, zero_blocks, sCat, 1.f);
  cg_gather<true>(L.sq, C, sAi, L.catA, L.offA, sq_blocks, sCat, 1.f);
  for (int idx = threadIdx.x; idx < NLM2 * C; idx += blockDim.x) {
    const int lm = idx / C, cc = idx % C, l = ell_of_lm(lm);
    sCat[L.offA[l] + (lm - l * l) * L.catA[l] + L.in_block[l] * C + cc] = sAi[idx];
  }
  __syncthreads();
  if (cat_out) {
    float2* co = reinterpret_cast<float2*>(cat_out) + ((long long)b * N + i) * L.totA;
    for (int idx = threadIdx.x; idx < L.totA; idx += blockDim.x) co[idx] = sCat[idx];
  }
  const bool last = (level == d.K - 1);
  mix_rows<CO, NM>(last ? d.units_out : d.units_hidden, last ? d.n_units_out : d.n_units_hidden, L.catA, L.offA, L.offWA,
                   L.Cout, reinterpret_cast<const float2*>(P + L.p_atomW), sCat,
                   reinterpret_cast<float2*>(A_out) + ((long long)b * N + i) * kM * L.Cout);
}

// ------------------------------------------------------------------------------------------------------------
// Invariants (so3_tools.py:147-190): inv[b,i,:] = [l=0 (re,im) per channel | per l: (Re sum_m (-1)^m a_m a_-m, sum |a|^2)]
// Rows i >= n_atoms are written as zeros (the reference sees zero representations on padded atoms).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_scalars_row(const float2* __restrict__ a /* [25][tau] */, int tau, int stride, float* out,
                                                   int tid, int nthreads) {
  for (int idx = tid; idx < (kL + 2) * tau; idx += nthreads) {
    const int blk = idx / tau, t = idx % tau;
    float v0, v1;
    if (blk == 0) {
      const float2 z = a[t];
      v0 = z.x; v1 = z.y;
    } else {
      const int l = blk - 1;
      float pr = 0.f, nr = 0.f;
      for (int m = -l; m <= l; ++m) {
        const float2 p = a[lm_index(l, m) * stride + t], q = a[lm_index(l, -m) * stride + t];
        const float sg = (m & 1) ? -1.f : 1.f;
        pr += sg * (p.x * q.x - p.y * q.y);
        nr += p.x * p.x + p.y * p.y;
      }
      v0 = pr; v1 = nr;
    }
    out[idx * 2 + 0] = v0;
    out[idx * 2 + 1] = v1;
  }
}

__global__ void k_scalars_fwd(const CovDesc* __restrict__ dp, const int* __restrict__ n_atoms, const float* __restrict__ A,
                              float* __restrict__ inv) {
  const CovDesc& d = *dp;
  const int N = d.N, tau = d.Cout;
  const int b = blockIdx.x / N, i = blockIdx.x % N;
  float* out = inv + (long long)blockIdx.x * d.lat;
  if (i >= n_atoms[b]) {
    for (int idx = threadIdx.x; idx < d.lat; idx += blockDim.x) out[idx] = 0.f;
    return;
  }
  atomic_scalars_row(reinterpret_cast<const float2*>(A) + (long long)blockIdx.x * kM * tau, tau, tau, out, threadIdx.x, blockDim.x);
}

}  // namespace mgb
