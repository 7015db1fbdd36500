// cov_backward.cuh — hand-written backward of the covariant actor-critic (the reference relies on torch autograd
// through molgym/agents/covariant/agent.py:209-334; nothing here has a counterpart file in the reference).
//
// Complex tensors are stored as (re, im) pairs and cotangents likewise; for a holomorphic product z = a * b the
// cotangents are  da += conj(b) * dz,  db += conj(a) * dz.
#pragma once
#include "heads.cuh"

namespace mgb {

__device__ __forceinline__ void smem_add2(float2* p, float2 v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); }

// g = sum over the (zero-padded, fixed trip count) table entries of pair p: coef * dcat[dst + c]
template <int NPAD>
__device__ __forceinline__ float2 pair_scatter(const int2* __restrict__ tab, int p, const float2* __restrict__ sDcat, int c) {
  const int2* e = tab + (long long)p * NPAD;
  int2 ent[NPAD];
  MGB_UNROLL
  for (int q = 0; q < NPAD; ++q) ent[q] = e[q];
  float2 g = make_float2(0.f, 0.f);
  MGB_UNROLL
  for (int q = 0; q < NPAD; ++q) {
    const float cf = __int_as_float(ent[q].y);
    const float2 dc = sDcat[ent[q].x + c];
    g.x = fmaf(cf, dc.x, g.x);
    g.y = fmaf(cf, dc.y, g.y);
  }
  return g;
}

// cotangent of atomic_scalars_row: d a[lm][t] += ...   (a: [25][stride] complex, dinv: [(L+2)*tau*2])
__device__ __forceinline__ float2 scalars_bwd_elem(const float2* __restrict__ a, int tau, int stride, const float* __restrict__ dinv,
                                                   int lm, int t) {
  const int l = ell_of_lm(lm), m = lm - l * l - l;
  const float2 p = a[lm * stride + t], q = a[lm_index(l, -m) * stride + t];
  const float sg = (m & 1) ? -1.f : 1.f;
  const float dprod = dinv[((1 + l) * tau + t) * 2 + 0], dnorm = dinv[((1 + l) * tau + t) * 2 + 1];
  float2 g = make_float2(2.f * sg * q.x * dprod + 2.f * p.x * dnorm, -2.f * sg * q.y * dprod + 2.f * p.y * dnorm);
  if (lm == 0) { g.x += dinv[t * 2 + 0]; g.y += dinv[t * 2 + 1]; }
  return g;
}

__global__ void k_scalars_bwd(const CovDesc* __restrict__ dp, const int* __restrict__ n_atoms, const float* __restrict__ A,
                              const float* __restrict__ dinv, float* __restrict__ dA) {
  const CovDesc& d = *dp;
  const int N = d.N, tau = d.Cout;
  const int b = blockIdx.x / N, i = blockIdx.x % N;
  if (i >= n_atoms[b]) return;
  const float2* a = reinterpret_cast<const float2*>(A) + (long long)blockIdx.x * kM * tau;
  float2* da = reinterpret_cast<float2*>(dA) + (long long)blockIdx.x * kM * tau;
  const float* di = dinv + (long long)blockIdx.x * d.lat;
  for (int idx = threadIdx.x; idx < kM * tau; idx += blockDim.x) {
    const float2 g = scalars_bwd_elem(a, tau, tau, di, idx / tau, idx % tau);
    da[idx].x += g.x;
    da[idx].y += g.y;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Per-canvas policy head backward.  Persistent CTAs: mixer-weight and log-std cotangents are accumulated on chip
// over the canvases a CTA visits and flushed once.
// ------------------------------------------------------------------------------------------------------------
struct PolicyBwdOut {
  float* finv; float* he; float* einv; float* hd; float* vf; float* hv;   // activations for the weight gradients
  float* dhe; float* dye; float* dhd; float* dyd; float* dhv; float* dyv;
  float* dvf; float* dflogit; float* dinv; float* dA_last;
};

// y[k] = sum_h W[h][k] x[h]  (reference layout [rows=h][cols=k], lanes over k)
__device__ __forceinline__ void gemv_n(const float* __restrict__ W, const float* x, int rows, int cols, float* y) {
  for (int k = threadIdx.x; k < cols; k += blockDim.x) {
    float acc = 0.f;
#pragma unroll 16
    for (int h = 0; h < rows; ++h) acc = fmaf(W[(long long)h * cols + k], x[h], acc);
    y[k] = acc;
  }
}

__global__ void __launch_bounds__(kPolicyBwdThreads, MGB_POLICY_BWD_MIN_CTAS)
k_policy_bwd(const CovDesc* __restrict__ dp, const float* __restrict__ P, const float* __restrict__ Wt, int B,
             const int* __restrict__ n_atoms, const float* __restrict__ bags, const float* __restrict__ actions,
             const float* __restrict__ A_last, const float* __restrict__ inv, const float* __restrict__ flogit,
             const float* __restrict__ trans, const float* __restrict__ state, const float* __restrict__ g_logp,
             const float* __restrict__ g_ent, const float* __restrict__ g_v, PolicyBwdOut o, float* __restrict__ mix_stage,
             float* __restrict__ grad) {
  const CovDesc& d = *dp;
  MGB_DYN_SMEM(float, sm);
  PolicySmem s = policy_smem_carve(d, sm);
  const int Wd = d.Wd, Z = d.Z, G = d.G, CPE = d.CPE, N = d.N;
  float* extra = sm + policy_smem_floats(d);
  float* s_dh = extra;                                                // [Wd] scratch
  float* s_dx = s_dh + Wd;                                            // [max(lat, Wd)]
  float2* s_dcat = reinterpret_cast<float2*>(s_dx + (d.lat > Wd ? d.lat : Wd));   // [totM]
  float2* s_dWM = s_dcat + d.totM;                                    // [sum_l catM] one per (l, k): identical for every c'
  float2* s_decov = s_dWM + d.totWM;                                  // [25][CPE]
  float2* s_ag = s_decov + kM * CPE;                                  // [25][CPE]
  float2* s_da = s_ag + kM * CPE;                                     // [25]
  float* s_small = reinterpret_cast<float*>(s_da + kM);               // [64]
  for (int idx = threadIdx.x; idx < d.totWM; idx += blockDim.x) s_dWM[idx] = make_float2(0.f, 0.f);
  float acc_logstd[8];
  for (int k = 0; k < 8; ++k) acc_logstd[k] = 0.f;

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    PolicyScalars ps;
    __syncthreads();
    {   // the forward's intermediates (PolicySmem block + scalars), saved by k_policy_fwd
      const float* src = state + (long long)b * policy_state_floats(d);
      const int nf = policy_smem_floats(d);
      for (int idx = threadIdx.x; idx < nf; idx += blockDim.x) sm[idx] = src[idx];
      ps = *reinterpret_cast<const PolicyScalars*>(src + ((nf + 3) & ~3));
    }
    __syncthreads();
    const float gl = g_logp[b], ge = g_ent[b], gv = g_v[b];
    const int nact = ps.n > 1 ? ps.n : 1;
    // ---- value head
    for (int h = threadIdx.x; h < Wd; h += blockDim.x) {
      const float dh = s.hv[h] > 0.f ? P[d.value.W1 + h] * gv : 0.f;
      s_dh[h] = dh;
      o.dhv[(long long)b * Wd + h] = dh;
      o.hv[(long long)b * Wd + h] = s.hv[h];
      o.vf[(long long)b * Wd + h] = s.vf[h];
    }
    if (threadIdx.x == 0) o.dyv[b] = gv;
    __syncthreads();
    gemv_n(P + d.value.W0, s_dh, Wd, Wd, s_dx);
    __syncthreads();
    for (int k = threadIdx.x; k < Wd; k += blockDim.x) o.dvf[(long long)b * Wd + k] = s_dx[k];
    // ---- focus + element categoricals
    if (threadIdx.x == 0) {
      bool mask[64];
      float dz[64];
      for (int i = 0; i < N; ++i) mask[i] = i < nact;
      categorical_bwd(s.fl, s.flog, mask, N, ps.focus, gl, ge, ps.aux_f, dz);
      for (int i = 0; i < N; ++i) o.dflogit[(long long)b * N + i] = dz[i];
      bool emask[MGB_MAX_SPECIES];
      for (int z = 0; z < Z; ++z) emask[z] = bags[(long long)b * Z + z] > 0.f;
      categorical_bwd(s.el, s.elog, emask, Z, ps.element, gl, ge, ps.aux_e, s_small);
      for (int z = 0; z < Z; ++z) o.dye[(long long)b * Z + z] = s_small[z];
    }
    __syncthreads();
    for (int h = threadIdx.x; h < Wd; h += blockDim.x) {
      float acc = 0.f;
      for (int z = 0; z < Z; ++z) acc = fmaf(P[d.element.W1 + (long long)z * Wd + h], s_small[z], acc);
      const float dh = s.he[h] > 0.f ? acc : 0.f;
      s_dh[h] = dh;
      o.dhe[(long long)b * Wd + h] = dh;
      o.he[(long long)b * Wd + h] = s.he[h];
    }
    for (int k = threadIdx.x; k < d.lat; k += blockDim.x) o.finv[(long long)b * d.lat + k] = s.finv[k];
    __syncthreads();
    gemv_n(P + d.element.W0, s_dh, Wd, d.lat, s_dx);
    __syncthreads();
    for (int k = threadIdx.x; k < d.lat; k += blockDim.x) o.dinv[((long long)b * N + ps.focus) * d.lat + k] += s_dx[k];
    __syncthreads();
    // ---- distance (GMM) head
    if (threadIdx.x == 0) {
      float lse_g = -3.0e38f;
      for (int k = 0; k < G; ++k) lse_g = fmaxf(lse_g, s.yd[k]);
      float sg = 0.f;
      for (int k = 0; k < G; ++k) sg += expf(s.yd[k] - lse_g);
      lse_g += logf(sg);
      const float hw = 0.5f * (d.dmax - d.dmin), ctr = 0.5f * (d.dmin + d.dmax);
      float t[8], tm = -3.0e38f, mu[8], sd[8], th[8], raw_sd[8];
      for (int k = 0; k < G; ++k) {
        th[k] = tanhf(s.yd[G + k]);
        mu[k] = th[k] * hw + ctr;
        raw_sd[k] = expf(P[d.p_logstd + k]);
        sd[k] = fmaxf(raw_sd[k], 1e-6f);
        const float df = ps.dist - mu[k];
        t[k] = -(df * df) / (2.f * sd[k] * sd[k]) - logf(sd[k]) - kLogSqrt2Pi + (s.yd[k] - lse_g);
        tm = fmaxf(tm, t[k]);
      }
      float st = 0.f;
      for (int k = 0; k < G; ++k) st += expf(t[k] - tm);
      float wsum = 0.f, w[8];
      for (int k = 0; k < G; ++k) { w[k] = expf(t[k] - tm) / st * gl; wsum += w[k]; }
      for (int k = 0; k < G; ++k) {
        const float df = ps.dist - mu[k];
        const float dlogit = w[k] - expf(s.yd[k] - lse_g) * wsum;
        const float dmu = w[k] * df / (sd[k] * sd[k]);
        const float dsd = w[k] * (df * df / (sd[k] * sd[k] * sd[k]) - 1.f / sd[k]);
        s_small[16 + k] = dlogit;
        s_small[16 + G + k] = dmu * hw * (1.f - th[k] * th[k]);
        s_small[40 + k] = raw_sd[k] >= 1e-6f ? dsd * raw_sd[k] : 0.f;
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < G) acc_logstd[threadIdx.x] += s_small[40 + threadIdx.x];  // thread k owns log-std k
    for (int q = threadIdx.x; q < 2 * G; q += blockDim.x) o.dyd[(long long)b * 2 * G + q] = s_small[16 + q];
    for (int h = threadIdx.x; h < Wd; h += blockDim.x) {
      float acc = 0.f;
      for (int q = 0; q < 2 * G; ++q) acc = fmaf(P[d.dist.W1 + (long long)q * Wd + h], s_small[16 + q], acc);
      const float dh = s.hd[h] > 0.f ? acc : 0.f;
      s_dh[h] = dh;
      o.dhd[(long long)b * Wd + h] = dh;
      o.hd[(long long)b * Wd + h] = s.hd[h];
    }
    for (int k = threadIdx.x; k < d.latE; k += blockDim.x) o.einv[(long long)b * d.latE + k] = s.einv[k];
    __syncthreads();
    gemv_n(P + d.dist.W0, s_dh, Wd, d.latE, s_dx);   // s_dx[0..latE) = d einv
    // ---- orientation: cotangent of the normalised coefficients a~_lm
    {
      float2 da[kM];
      MGB_UNROLL
      for (int q = 0; q < kM; ++q) da[q] = make_float2(0.f, 0.f);
      float2 a_loc[kM];
      MGB_UNROLL
      for (int q = 0; q < kM; ++q) a_loc[q] = s.alm[q];
      if (d.has_beta) {
        // -g * dlogZ/da~ : softmax-weighted over the Lebedev grid
        for (int g = threadIdx.x; g < d.n_grid; g += blockDim.x) {
          const float2* y = reinterpret_cast<const float2*>(d.leb_y) + g;
          float2 yv[kM];
          MGB_UNROLL
          for (int q = 0; q < kM; ++q) yv[q] = y[(long long)q * d.n_grid];
          const float2 sg = sph_sum(a_loc, yv);
          const float wgt = expf(-d.beta * (sg.x * sg.x + sg.y * sg.y) + d.leb_logw[g] - ps.lse_max) / ps.lse_sum;
          const float cf = -gl * wgt * (-2.f * d.beta);
          const float2 z = make_float2(cf * sg.x, cf * sg.y);
          MGB_UNROLL
          for (int q = 0; q < kM; ++q) cfmacl(da[q], yv[q], z);   // conj(Y) * z
        }
      }
      // block-reduce the 50 numbers: shuffles inside each warp, then one pass over the per-warp partials
      {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        float* part = reinterpret_cast<float*>(s_dcat);   // [nwarps][50] scratch (the mixer cotangent is written later)
        MGB_UNROLL
        for (int q = 0; q < kM; ++q) {
          const float rx = warp_sum(da[q].x), ry = warp_sum(da[q].y);
          if (lane == 0) { part[warp * 2 * kM + 2 * q] = rx; part[warp * 2 * kM + 2 * q + 1] = ry; }
        }
        __syncthreads();
        if ((int)threadIdx.x < 2 * kM) {
          float acc = 0.f;
          for (int w = 0; w < nwarps; ++w) acc += part[w * 2 * kM + threadIdx.x];
          reinterpret_cast<float*>(s_da)[threadIdx.x] = acc;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        float ox = actions[(long long)b * 6 + 3], oy = actions[(long long)b * 6 + 4], oz = actions[(long long)b * 6 + 5];
        const float nr = sqrtf(ox * ox + oy * oy + oz * oz);
        if (nr > 0.f) { ox /= nr; oy /= nr; oz /= nr; } else { ox = oy = oz = 0.f; }
        float2 y[kM];
        sph_harm_l4(ox, oy, oz, false, false, y);
        float cf;
        if (d.has_beta) {
          cf = gl * (-2.f * d.beta);
        } else {
          const float so2 = ps.s_o.x * ps.s_o.x + ps.s_o.y * ps.s_o.y;
          cf = (ps.n == 0 || so2 < 1e-10f) ? 0.f : gl * 2.f / so2;
        }
        const float2 z = make_float2(cf * ps.s_o.x, cf * ps.s_o.y);
        // through a~ = a / sqrt(max(k, 1e-10))
        float dot = 0.f;
        for (int q = 0; q < kM; ++q) {
          cfmacl(s_da[q], y[q], z);
          dot += s_da[q].x * a_loc[q].x + s_da[q].y * a_loc[q].y;   // sum d a~ . a~
        }
        // a_raw = a~ / inv_sqrt_k ; dk = -0.5 k^-3/2 sum(da~ . a_raw) = -0.5 inv^2 * dot(da~, a~)
        const float dk = ps.k_raw >= 1e-10f ? -0.5f * ps.inv_sqrt_k * ps.inv_sqrt_k * dot : 0.f;
        for (int q = 0; q < kM; ++q) {
          const float2 araw = make_float2(a_loc[q].x / ps.inv_sqrt_k, a_loc[q].y / ps.inv_sqrt_k);
          s_da[q] = make_float2(s_da[q].x * ps.inv_sqrt_k + 2.f * araw.x * dk, s_da[q].y * ps.inv_sqrt_k + 2.f * araw.y * dk);
        }
      }
      __syncthreads();
    }
    // ---- mixer backward: cond[lm][c'] = sum_k W_l[c'][k] cat_l[m][k], d cond[lm][c'] = s_da[lm] for every c'
    for (int l = 0; l < kNL; ++l) {
      const int Kc = d.catM[l];
      const float2* Wl = reinterpret_cast<const float2*>(P + d.p_mixW) + d.offWM[l];
      for (int k = threadIdx.x; k < Kc; k += blockDim.x) {
        float2 wsum = make_float2(0.f, 0.f);
        for (int c = 0; c < CPE; ++c) { wsum.x += Wl[c * Kc + k].x; wsum.y += Wl[c * Kc + k].y; }
        float2 dw = make_float2(0.f, 0.f);
        for (int m = 0; m < 2 * l + 1; ++m) {
          const float2 g = s_da[l * l + m];
          const float2 x = s.cat[d.offM[l] + m * Kc + k];
          float2 dc = make_float2(0.f, 0.f);
          cfmacl(dc, wsum, g);            // conj(sum_c' W) * g
          s_dcat[d.offM[l] + m * Kc + k] = dc;
          cfmacl(dw, x, g);               // conj(cat) * g
        }
        s_dWM[d.offWM[l] / CPE + k].x += dw.x;
        s_dWM[d.offWM[l] / CPE + k].y += dw.y;
      }
    }
    // ag = dist * ecov (recomputed)
    for (int idx = threadIdx.x; idx < kM * CPE; idx += blockDim.x)
      s_ag[idx] = make_float2(ps.dist * s.ecov[idx].x, ps.dist * s.ecov[idx].y);
    __syncthreads();
    // d ecov = d in + dist * (d ag_block + square-backward) + invariants backward
    for (int idx = threadIdx.x; idx < kM * CPE; idx += blockDim.x) {
      const int x = idx / CPE, c = idx % CPE, l = ell_of_lm(x);
      const int base = d.offM[l] + (x - l * l) * d.catM[l];
      float2 dag = s_dcat[base + c];
      for (int y = 0; y < kM; ++y) {
        const float2 g = pair_scatter<2 * kCgPad>(d.mix_sq.pad_sym, x * kM + y, s_dcat, c);
        cfmacl(dag, s_ag[y * CPE + c], g);   // conj(ag_y) * g
      }
      float2 de = s_dcat[base + d.inM_block[l] * CPE + c];
      de.x += ps.dist * dag.x; de.y += ps.dist * dag.y;
      const float2 gs = scalars_bwd_elem(s.ecov, CPE, CPE, s_dx, x, c);
      de.x += gs.x; de.y += gs.y;
      if (ps.focus_valid) {
        float2* dst = reinterpret_cast<float2*>(o.dA_last) + (((long long)b * N + ps.focus) * kM + x) * d.Cout + ps.element * CPE + c;
        dst->x += de.x; dst->y += de.y;
      }
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < d.totWM / CPE; idx += blockDim.x) {   // compact staging, expanded by k_mixer_dw_finish
    const float2 v = s_dWM[idx];
    if (v.x != 0.f) atomicAdd(mix_stage + 2 * idx, v.x);
    if (v.y != 0.f) atomicAdd(mix_stage + 2 * idx + 1, v.y);
  }
  if ((int)threadIdx.x < G && acc_logstd[threadIdx.x] != 0.f) atomicAdd(grad + d.p_logstd + threadIdx.x, acc_logstd[threadIdx.x]);
}
// grad[mixer W_l[c'][k]] += stage[l][k] for every c'
__global__ void k_mixer_dw_finish(const CovDesc* __restrict__ dp, const float* __restrict__ stage, float* __restrict__ grad) {
  const CovDesc& d = *dp;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < d.totWM; idx += gridDim.x * blockDim.x) {
    int l = 0;
    while (l + 1 < kNL && idx >= d.offWM[l + 1]) ++l;
    const int Kc = d.catM[l], k = (idx - d.offWM[l]) % Kc;
    const int src = d.offWM[l] / d.CPE + k;
    grad[d.p_mixW + 2ll * idx] += stage[2 * src];
    grad[d.p_mixW + 2ll * idx + 1] += stage[2 * src + 1];
  }
}
__host__ __device__ inline int policy_bwd_extra_floats(const CovDesc& d) {
  return d.Wd + (d.lat > d.Wd ? d.lat : d.Wd) + 2 * d.totM + 2 * d.totWM + 2 * kM * d.CPE * 2 + 2 * kM + 64 + 16;
}

// ------------------------------------------------------------------------------------------------------------
// Atom level backward for atom i (see k_atom_fwd).  Inputs: dOut = d A_{k+1}[b,i].  Produces
//   dE_k[b,i,j,:,:] for all j (assigned, or added when the next level's edge network already wrote its share),
//   dA_k[b,j] += ... for all j (global atomics; dA_k zero-initialised), including the own-atom terms.
// Phases: (A) dcat = W^H dOut into shared memory; (B) row pass: dT[x][.] in registers -> dE_ij; (C) column pass:
// dT[.][x] in registers -> dA_j.  The neighbours are staged twice so that only one 25-vector of dT lives in registers.
// ------------------------------------------------------------------------------------------------------------
constexpr int kAtomBwdThreads = 256;
constexpr int kEdgeCMax = 10;   // hidden channels <= 10 (checked at plan creation)
constexpr int kJChunkBwd = 8;
// Level 0 (scalar inputs, NLM2 = 1) has almost no arithmetic per neighbour: its chunks are 32 neighbours, so that a canvas row is
// one or two passes with eight loads in flight per thread instead of 3-5 latency-exposed passes (C5 b256: 315 us per launch with chunks of 8)
constexpr int kJChunkBwd0 = 32;
__host__ __device__ constexpr int atom_bwd_chunk(int nlm2) { return nlm2 == 1 ? kJChunkBwd0 : kJChunkBwd; }
// The neighbour chunks are software-pipelined through registers: while a chunk is consumed out of shared memory the next one
// is already in flight from L2 (kAtomPre float2 per thread cover a chunk of E_ij and A_j rows for C <= 10).
constexpr int kAtomPre = (kJChunkBwd * (kNL + kM) * kEdgeCMax + kAtomBwdThreads - 1) / kAtomBwdThreads;
static_assert(kJChunkBwd0 * (kNL + 1) * kEdgeCMax <= kAtomPre * kAtomBwdThreads, "level-0 chunk must fit the register staging");

__host__ __device__ inline int atom_bwd_smem_floats(const LevelDesc& L, int N) {
  const int nlm2 = L.nlm_in;
  const int stage = atom_bwd_chunk(nlm2) * (kNL * L.C + nlm2 * L.C + kM * L.C) * 2;
  return L.totA * 2 + stage + nlm2 * L.C * 2 + N * kM * 2;
}

// register-staged copy of one neighbour chunk: E_ij rows [nj][5][C] and (NLM2 > 0) A_j rows [nj][NLM2][C] are contiguous in
// HBM (neighbours j0 .. j0+nj-1 of atom i), so element e of the chunk is either E_i[j0*5C + e] or Ab[j0*NLM2*C + (e - nE)].
template <int NLM2>
__device__ __forceinline__ void chunk_fetch(const float2* __restrict__ Ab, const float2* __restrict__ E_i, int C, int j0, int nj,
                                            float2* pre) {
  const int nE = nj * kNL * C, nA = nj * NLM2 * C;
  MGB_UNROLL
  for (int q = 0; q < kAtomPre; ++q) {
    const int e = threadIdx.x + q * kAtomBwdThreads;
    if (e < nE) pre[q] = E_i[(long long)j0 * kNL * C + e];
    else if (e < nE + nA) pre[q] = Ab[(long long)j0 * NLM2 * C + (e - nE)];
  }
}
template <int NLM2>
__device__ __forceinline__ void chunk_commit(int C, int nj, const float2* pre, float2* sE, float2* sAj) {
  const int nE = nj * kNL * C, nA = nj * NLM2 * C;
  MGB_UNROLL
  for (int q = 0; q < kAtomPre; ++q) {
    const int e = threadIdx.x + q * kAtomBwdThreads;
    if (e < nE) sE[e] = pre[q];
    else if (e < nE + nA) sAj[e - nE] = pre[q];
  }
}

template <int NLM2, int CT>   // CT: compile-time channel count (0: read it from the level descriptor)
__global__ void __launch_bounds__(kAtomBwdThreads, 2)
k_atom_bwd(const CovDesc* __restrict__ dp, int level, const float* __restrict__ pos, const int* __restrict__ n_atoms,
           const int* __restrict__ atom_off, const int* __restrict__ atom_list, int B, const float* __restrict__ A_in,
           const float* __restrict__ E, const float* __restrict__ dcat, float* __restrict__ dA_in, float* __restrict__ dE,
           int accumulate_dE, int phases) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int N = d.N, C = CT ? CT : L.C;
  constexpr int JC = atom_bwd_chunk(NLM2);   // neighbours per staged chunk
  if ((int)blockIdx.x >= atom_off[B]) return;
  const int slot = atom_list[blockIdx.x];
  const int b = slot / N, i = slot - b * N;
  const int n = n_atoms[b];
  MGB_DYN_SMEM(float2, smem);
  float2* sDcat = smem;                         // [totA]  cotangent of the cat vector (written by k_mix_rows<.., true>)
  float2* sE = sDcat + L.totA;                  // [JC][5][C]
  float2* sAj = sE + JC * kNL * C;      // [JC][NLM2][C]
  float2* sU = sAj + JC * NLM2 * C;     // [JC][25][C]   column pass: E * Y ; row pass: per-thread dE contributions
  float2* sAi = sU + JC * kM * C;       // [NLM2][C]
  float2* sYall = sAi + NLM2 * C;               // [n][25]
  const float2* Ab = reinterpret_cast<const float2*>(A_in) + (long long)b * N * NLM2 * C;
  const float2* E_i = reinterpret_cast<const float2*>(E) + ((long long)b * N + i) * N * kNL * C;
  float2* dE_i = reinterpret_cast<float2*>(dE) + ((long long)b * N + i) * N * kNL * C;
  float2* dAb = reinterpret_cast<float2*>(dA_in) + (long long)b * N * NLM2 * C;
  const float* pos_b = pos + (long long)b * N * 3;
  // the atom's dcat slice (62 KB at the default width) arrives as ONE bulk copy (TMA, cp.async.bulk + mbarrier) issued by one
  // thread while the others stage A_i and evaluate the neighbour harmonics
  __shared__ SmemBarrier s_bar;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  const bool bulk = smem_fill_begin(reinterpret_cast<float*>(sDcat), dcat + 2ll * slot * L.totA, 2 * L.totA, &s_bar);
  for (int idx = threadIdx.x; idx < NLM2 * C; idx += blockDim.x) sAi[idx] = Ab[(long long)i * NLM2 * C + idx];
  neighbour_harmonics(pos_b, i, n, sYall);
  smem_fill_end(bulk, &s_bar, 0);
  __syncthreads();

  const bool owner = (int)threadIdx.x < kM * C;
  const int x = owner ? threadIdx.x / C : 0, c = owner ? threadIdx.x % C : 0;
  const int l1 = ell_of_lm(x);
  const bool col_owner = owner && x < NLM2;
  // small minibatches launch the two passes as separate kernels (kAtomPhaseA / kAtomPhaseB): the edge backward only waits
  // for the row pass, the column pass runs beside it on a side stream
  const bool do_rows = (phases & kAtomPhaseA) != 0, do_cols = (phases & kAtomPhaseB) != 0;
  // ---- (B) row pass
  if (do_rows) {
    float2 dTrow[NLM2];   // dT[x][y], y < NLM2
    if (owner) {
      MGB_UNROLL
      for (int y = 0; y < NLM2; ++y) dTrow[y] = pair_scatter<kCgPad>(L.ag.pad_pair, x * NLM2 + y, sDcat, c);
    }
    float2 pre[kAtomPre];
    chunk_fetch<NLM2>(Ab, E_i, C, 0, min(JC, n), pre);
    for (int j0 = 0; j0 < n; j0 += JC) {
      const int nj = min(JC, n - j0);
      __syncthreads();
      chunk_commit<NLM2>(C, nj, pre, sE, sAj);
      __syncthreads();
      if (j0 + JC < n) chunk_fetch<NLM2>(Ab, E_i, C, j0 + JC, min(JC, n - j0 - JC), pre);
      if (owner) {
        for (int jj = 0; jj < nj; ++jj) {
          // contribution of m1 = x to dE_ij[l1, c]: conj(Y[x]) * sum_y conj(A_j[y, c]) dT[x][y]
          float2 w0 = make_float2(0.f, 0.f), w1 = make_float2(0.f, 0.f);
          const float2* a = sAj + jj * NLM2 * C + c;
          MGB_UNROLL
          for (int y = 0; y < NLM2; ++y) {
            if (y & 1) cfmacl(w1, a[y * C], dTrow[y]); else cfmacl(w0, a[y * C], dTrow[y]);
          }
          w0.x += w1.x; w0.y += w1.y;
          float2 de = make_float2(0.f, 0.f);
          cfmacl(de, sYall[(j0 + jj) * kM + x], w0);
          sU[(jj * kM + x) * C + c] = de;
        }
      }
      __syncthreads();
      // sum the 2l+1 contributions of every (j, l, c) and write dE
      for (int idx = threadIdx.x; idx < nj * kNL * C; idx += blockDim.x) {
        const int jj = idx / (kNL * C), r = idx - jj * (kNL * C), l = r / C, cc = r - l * C;
        float2 v = make_float2(0.f, 0.f);
        for (int m = 0; m < 2 * l + 1; ++m) { const float2 t = sU[(jj * kM + l * l + m) * C + cc]; v.x += t.x; v.y += t.y; }
        float2* dst = dE_i + (long long)j0 * kNL * C + idx;
        if (accumulate_dE) { v.x += dst->x; v.y += dst->y; }
        *dst = v;
      }
    }
  }
  // ---- (C) column pass + own-atom terms
  if (do_cols) {
    float2 dTcol[kM];     // dT[y][x], y < 25   (threads x < NLM2)
    float2* sT = sAj;     // NLM2 == 1: dT[y][0] per channel, [25][C] (the column pass does not stage A_j)
    if (NLM2 == 1) {
      if (owner) sT[x * C + c] = pair_scatter<kCgPad>(L.ag.pad_pair, x, sDcat, c);
    }
    if (col_owner) {
      if (NLM2 != 1) {
        MGB_UNROLL
        for (int y = 0; y < kM; ++y) dTcol[y] = pair_scatter<kCgPad>(L.ag.pad_pair, y * NLM2 + x, sDcat, c);
      }
      // own atom: pass-through block and CG square
      const int base = L.offA[l1] + (x - l1 * l1) * L.catA[l1];
      float2 dai = sDcat[base + L.in_block[l1] * C + c];
      for (int y = 0; y < NLM2; ++y) {
        const float2 g = pair_scatter<2 * kCgPad>(L.sq.pad_sym, x * NLM2 + y, sDcat, c);   // entries of (x,y) and (y,x)
        cfmacl(dai, sAi[y * C + c], g);
      }
      atomic_add2(dAb + (long long)i * NLM2 * C + x * C + c, dai);
    }
    float2 pre[kAtomPre];
    chunk_fetch<0>(Ab, E_i, C, 0, min(JC, n), pre);
    for (int j0 = 0; j0 < n; j0 += JC) {
      const int nj = min(JC, n - j0);
      __syncthreads();
      chunk_commit<0>(C, nj, pre, sE, sAj);
      __syncthreads();
      if (j0 + JC < n) chunk_fetch<0>(Ab, E_i, C, j0 + JC, min(JC, n - j0 - JC), pre);
      if (owner)
        for (int jj = 0; jj < nj; ++jj) sU[(jj * kM + x) * C + c] = cmul(sE[(jj * kNL + l1) * C + c], sYall[(j0 + jj) * kM + x]);
      __syncthreads();
      if (NLM2 == 1) {
        // scalar inputs: dT[.][0] depends on the channel only (staged in sT); all threads share the (neighbour, channel) sums
        // instead of the C column owners walking every neighbour
        for (int idx = threadIdx.x; idx < nj * C; idx += blockDim.x) {
          const int jj = idx / C, cc = idx - jj * C;
          float2 v0 = make_float2(0.f, 0.f), v1 = make_float2(0.f, 0.f);
          const float2* u = sU + jj * kM * C + cc;
          MGB_UNROLL
          for (int y = 0; y < kM; ++y) {
            if (y & 1) cfmacl(v1, u[y * C], sT[y * C + cc]); else cfmacl(v0, u[y * C], sT[y * C + cc]);
          }
          atomic_add2(dAb + (long long)(j0 + jj) * C + cc, make_float2(v0.x + v1.x, v0.y + v1.y));
        }
      } else if (col_owner) {
        for (int jj = 0; jj < nj; ++jj) {
          // dA_j[x, c] += sum_y conj(E_ij[l(y), c] Y[y]) dT[y][x]
          float2 v0 = make_float2(0.f, 0.f), v1 = make_float2(0.f, 0.f);
          const float2* u = sU + jj * kM * C + c;
          MGB_UNROLL
          for (int y = 0; y < kM; ++y) {
            if (y & 1) cfmacl(v1, u[y * C], dTcol[y]); else cfmacl(v0, u[y * C], dTcol[y]);
          }
          atomic_add2(dAb + (long long)(j0 + jj) * NLM2 * C + x * C + c, make_float2(v0.x + v1.x, v0.y + v1.y));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Atom-mix weight gradient: dW_l[c'][k] += sum_{atoms, m} conj(cat[a][l][m][k]) dOut[a][lm][c'].
// grid = (chunks of the compact valid-atom list, 5 ells); lanes over k straight out of HBM (the forward saved cat,
// every row is one coalesced run), all output channels of a k in registers (one pass over cat), dOut rows of a group of
// atoms staged in shared memory and read as broadcasts.
// ------------------------------------------------------------------------------------------------------------
constexpr int kMixDwThreads = 128;
constexpr int kMixDwAtoms = 8;   // atoms per CTA pass (their dOut rows are staged together)

// grid = (chunks of the compact valid-atom list, 5 ells, 128-wide slices of k); thread = one k, all output channels of that k in
// registers, the 2l+1 rows of an atom loaded together (up to nine independent loads in flight per thread).
template <int CO>
__global__ void __launch_bounds__(kMixDwThreads)
k_mix_dw(const CovDesc* __restrict__ dp, int level, int B, const int* __restrict__ atom_off, const int* __restrict__ atom_list,
         const float* __restrict__ cat, const float* __restrict__ dA_out, int c_base, float* __restrict__ grad) {
  // this launch covers the output channels [c_base, min(c_base + CO, Cout))
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int Cout = L.Cout, l = blockIdx.y, K = L.catA[l], nm = 2 * l + 1;
  const int k = blockIdx.z * kMixDwThreads + threadIdx.x;
  if ((int)(blockIdx.z * kMixDwThreads) >= K) return;
  const int n_at = atom_off[B];
  const int per = max((int)((n_at + gridDim.x - 1) / gridDim.x), 2 * kMixDwAtoms);   // every CTA flushes 128 * Cout atomics: not too few atoms
  const int a0 = per * blockIdx.x, a1 = min(n_at, a0 + per);
  if (a0 >= a1) return;
  MGB_DYN_SMEM(float2, sd);   // [kMixDwAtoms][nm][CO], zero beyond Cout
  // pair accumulators (FFMA2): P[c] += g[c] * (x.re, x.re), Q[c] += g[c] * (x.im, x.im);  conj(x) * g = (P.lo + Q.hi, P.hi - Q.lo)
  f32x2 P[CO], Q[CO];
  MGB_UNROLL
  for (int c = 0; c < CO; ++c) { P[c] = pack2(0.f, 0.f); Q[c] = pack2(0.f, 0.f); }
  const float2* dO = reinterpret_cast<const float2*>(dA_out);
  const bool on = k < K;
  // the rows of the NEXT atom are fetched while the current atom's rows are consumed (software pipeline over the atom list)
  float2 xn[2 * kL + 1];
  {
    const float2* cr = reinterpret_cast<const float2*>(cat) + (long long)atom_list[a0] * L.totA + L.offA[l] + k;
    MGB_UNROLL
    for (int m = 0; m < 2 * kL + 1; ++m) xn[m] = (on && m < nm) ? cr[m * K] : make_float2(0.f, 0.f);
  }
  for (int ab = a0; ab < a1; ab += kMixDwAtoms) {
    const int cnt = min(kMixDwAtoms, a1 - ab);
    __syncthreads();
    for (int idx = threadIdx.x; idx < cnt * nm * CO; idx += blockDim.x) {
      const int a = idx / (nm * CO), rem = idx - a * nm * CO, m = rem / CO, c = rem - m * CO;
      sd[idx] = c_base + c < Cout ? dO[((long long)atom_list[ab + a] * kM + l * l + m) * Cout + c_base + c] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    for (int a = 0; a < cnt; ++a) {
      const float2* ga = sd + a * nm * CO;
      float2 xv[2 * kL + 1];
      MGB_UNROLL
      for (int m = 0; m < 2 * kL + 1; ++m) xv[m] = xn[m];
      if (ab + a + 1 < a1) {
        const float2* cr = reinterpret_cast<const float2*>(cat) + (long long)atom_list[ab + a + 1] * L.totA + L.offA[l] + k;
        MGB_UNROLL
        for (int m = 0; m < 2 * kL + 1; ++m) xn[m] = (on && m < nm) ? cr[m * K] : make_float2(0.f, 0.f);
      }
      MGB_UNROLL
      for (int m = 0; m < 2 * kL + 1; ++m) {
        if (m < nm) {
          const f32x2 xr = pack2(xv[m].x, xv[m].x), xi = pack2(xv[m].y, xv[m].y);
          MGB_UNROLL
          for (int c = 0; c < CO; ++c) {
            const f32x2 gp = as_pair(ga[m * CO + c]);
            fma2(P[c], gp, xr);
            fma2(Q[c], gp, xi);
          }
        }
      }
    }
  }
  if (on) {
    MGB_UNROLL
    for (int c = 0; c < CO; ++c) {
      const float2 p = unpack2(P[c]), q = unpack2(Q[c]);
      const float2 acc = make_float2(p.x + q.y, p.y - q.x);
      if (c_base + c < Cout && (acc.x != 0.f || acc.y != 0.f))
        atomic_add2(reinterpret_cast<float2*>(grad + L.p_atomW) + L.offWA[l] + (long long)(c_base + c) * K + k, acc);
    }
  }
}

// dA_k[b,i,l',m,c] += sum_j (dD_ij + dD_ji)[l',c] (-1)^m conj(A_j[l',-m,c]);  dD may come in n_slices partial slices
template <int NLIN>
__global__ void k_dot_bwd(const CovDesc* __restrict__ dp, int level, const int* __restrict__ n_atoms, const float* __restrict__ A_in,
                          const float* __restrict__ dD, int n_slices, long long slice_stride /* complex */, float* __restrict__ dA_in) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int N = d.N, C = L.C;
  constexpr int NLM = NLIN * NLIN;
  const int b = blockIdx.x / N, i = blockIdx.x % N;
  const int n = n_atoms[b];
  if (i >= n) return;
  const float2* Ab = reinterpret_cast<const float2*>(A_in) + (long long)b * N * NLM * C;
  const float2* dDb = reinterpret_cast<const float2*>(dD) + (long long)b * N * N * kNL * C;
  float2* dst = reinterpret_cast<float2*>(dA_in) + ((long long)b * N + i) * NLM * C;
  for (int idx = threadIdx.x; idx < NLM * C; idx += blockDim.x) {
    const int lm = idx / C, c = idx % C, l = ell_of_lm(lm), m = lm - l * l - l;
    float2 acc = make_float2(0.f, 0.f);
    const float2* arow = Ab + lm_index(l, -m) * C + c;
    const float2* drow = dDb + (long long)i * N * kNL * C + l * C + c;   // (i, j) entries: j strides kNL C
    const float2* dcol = dDb + (long long)i * kNL * C + l * C + c;       // (j, i) entries: j strides N kNL C
    int j = 0;
    if (n_slices == 1) {
      // four neighbours per pass: twelve independent loads in flight instead of a dependent chain per neighbour
      for (; j + 3 < n; j += 4) {
        float2 g1[4], g2[4], a[4];
        MGB_UNROLL
        for (int q = 0; q < 4; ++q) {
          g1[q] = drow[(long long)(j + q) * kNL * C];
          g2[q] = dcol[(long long)(j + q) * N * kNL * C];
          a[q] = arow[(long long)(j + q) * NLM * C];
        }
        MGB_UNROLL
        for (int q = 0; q < 4; ++q) cfmacl(acc, a[q], make_float2(g1[q].x + g2[q].x, g1[q].y + g2[q].y));
      }
    }
    for (; j < n; ++j) {
      float2 g = make_float2(0.f, 0.f);
      for (int s = 0; s < n_slices; ++s) {
        const float2 g1 = drow[s * slice_stride + (long long)j * kNL * C];
        const float2 g2 = dcol[s * slice_stride + (long long)j * N * kNL * C];
        g.x += g1.x + g2.x; g.y += g1.y + g2.y;
      }
      cfmacl(acc, arow[(long long)j * NLM * C], g);
    }
    const float sg = (m & 1) ? -1.f : 1.f;
    atomic_add2(dst + idx, make_float2(sg * acc.x, sg * acc.y));   // the atom level's column pass may still be adding to dA_k
  }
}

// ------------------------------------------------------------------------------------------------------------
// InputLinear weight gradient (cormorant InputLinear, covariant/modules.py:106): dW[o][s] += sum_rows dA0[r][o] X[r][s],
// db[o] += sum_rows dA0[r][o] over the valid atoms.  The last kernel of the backward: thread = one (o, s) entry
// (s == S_in is the bias), a CTA walks the atoms of a few canvases and flushes one atomic per entry.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_input_dw(const CovDesc* __restrict__ dp, int B, int per_cta, const int* __restrict__ n_atoms, const float* __restrict__ X,
           const float* __restrict__ dA0, float* __restrict__ grad) {
  const CovDesc& d = *dp;
  const int N = d.N, S = d.S_in, C2 = 2 * d.C;
  const int b0 = blockIdx.x * per_cta, b1 = min(B, b0 + per_cta);
  for (int e = threadIdx.x; e < C2 * (S + 1); e += blockDim.x) {
    const int o = e / (S + 1), sidx = e - o * (S + 1);
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) {
      const int n = n_atoms[b];
      for (int i = 0; i < n; ++i) {
        const long long r = (long long)b * N + i;
        const float x = sidx < S ? X[r * S + sidx] : 1.f;
        acc = fmaf(dA0[r * C2 + o], x, acc);
      }
    }
    if (acc != 0.f) atomicAdd(sidx < S ? grad + d.p_inW + (long long)o * S + sidx : grad + d.p_inb + o, acc);
  }
}

// ------------------------------------------------------------------------------------------------------------
// dst = (accumulate ? dst : 0) + scale * src over the flat gradient; `scale` is a device scalar (float64 or float32): the
// cotangent autograd hands to the fused PPO loss (agents/covariant/agent.py::_fused_backward) — one launch instead of a cast,
// a multiply and an add.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_scale_accumulate(float* __restrict__ dst, const float* __restrict__ src, const void* __restrict__ scale,
                                   int scale_is_double, long long n, int accumulate) {
  const float sc = scale_is_double ? (float)*reinterpret_cast<const double*>(scale) : *reinterpret_cast<const float*>(scale);
  const long long n4 = n >> 2;
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4* s4 = reinterpret_cast<const float4*>(src);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = s4[i];
    float4 y = accumulate ? d4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    y.x = fmaf(sc, x.x, y.x); y.y = fmaf(sc, x.y, y.y); y.z = fmaf(sc, x.z, y.z); y.w = fmaf(sc, x.w, y.w);
    d4[i] = y;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = fmaf(sc, src[i], accumulate ? dst[i] : 0.f);
}

// ------------------------------------------------------------------------------------------------------------
// PPO-clip loss (molgym/ppo.py:28-52) and its cotangents, float64 like the reference (adv / ret are float64).
// ------------------------------------------------------------------------------------------------------------
__global__ void k_ppo_loss(int B, const float* __restrict__ logp, const float* __restrict__ ent, const float* __restrict__ v,
                           const float* __restrict__ old_logp, const double* __restrict__ adv, const double* __restrict__ ret,
                           double clip, double vf_coef, double ent_coef, double invB, double* __restrict__ info,
                           float* __restrict__ g_logp, float* __restrict__ g_ent, float* __restrict__ g_v) {
  __shared__ double red[6][32];
  double acc[6] = {0, 0, 0, 0, 0, 0};   // policy, entropy, vf, kl, clipfrac, unused
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float ratio_f = expf(logp[b] - old_logp[b]);
    const double ratio = ratio_f, a = adv[b];
    const double lo = 1.0 - clip, hi = 1.0 + clip;
    const float ratio_c = fminf(fmaxf(ratio_f, (float)lo), (float)hi);   // clamp happens in float32 (ppo.py:35)
    const double o1 = ratio * a, o2 = (double)ratio_c * a;
    const bool in_range = ratio_f >= (float)lo && ratio_f <= (float)hi;
    acc[0] -= (o1 < o2 ? o1 : o2);
    acc[1] -= ent_coef * (double)ent[b];
    const double dv = (double)v[b] - ret[b];
    acc[2] += vf_coef * dv * dv;
    acc[3] += (double)(old_logp[b] - logp[b]);
    acc[4] += (ratio_f < (float)lo || ratio_f > (float)hi) ? 1.0 : 0.0;
    if (g_logp) {
      double dr;
      if (o1 < o2) dr = a;
      else if (o1 == o2) dr = 0.5 * a + (in_range ? 0.5 * a : 0.0);
      else dr = in_range ? a : 0.0;
      g_logp[b] = (float)(-invB * dr * ratio);
      g_ent[b] = (float)(-ent_coef * invB);
      g_v[b] = (float)(vf_coef * 2.0 * dv * invB);
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  for (int q = 0; q < 5; ++q) {
    double x = acc[q];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[q][w] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[5] = {0, 0, 0, 0, 0};
    for (int q = 0; q < 5; ++q)
      for (int k = 0; k < nw; ++k) tot[q] += red[q][k];
    info[1] = tot[0] * invB; info[2] = tot[1] * invB; info[3] = tot[2] * invB;
    info[0] = info[1] + info[2] + info[3];
    info[4] = tot[3] * invB; info[5] = tot[4] * invB; info[6] = 0; info[7] = 0;
  }
}

}  // namespace mgb
