// internal.cuh — the internal-coordinate (SchNet) actor-critic: forward and hand-written backward.
//
// Reference: molgym/agents/internal/agent.py:17-353 (SchNetAC) on schnetpack 0.3's representation.SchNet
// (n_atom_basis = network_width/2, n_filters 128, 3 interactions, 25 Gaussians on [0, 5 A], cosine cutoff; restated in
// oracle/thirdparty/schnetpack/representation.py).  The reference evaluates SchNet three times per observation at batch
// size one (canvas; canvas + hypothetical atom for both dihedral signs, agent.py:124-128,163-177); here every
// (canvas, variant) is a "molecule" with its own CTA.  z-matrix placement (zmat.py:99-133, float64) stays on the host.
#pragma once
#include "cov_backward.cuh"

namespace mgb {

constexpr int kSchFilters = 128;
constexpr int kSchGauss = 25;
constexpr int kSchIters = 3;
constexpr float kSchCutoff = 5.0f;
constexpr int kSchThreads = 128;
constexpr float kLn2 = 0.6931471805599453f;

struct SchLayer {
  long long W1, b1, W2, b2, in2f, Wo, bo, Wd, bd;   // float offsets into the flat parameter buffer
};
struct IntDesc {
  int N, M, Z, Wd, F, LB, lat;      // canvas size, atoms per molecule slot (N+1), species, width, atom features, bag latent
  int zs[MGB_MAX_SPECIES];
  float dmin, dmax;
  long long p_emb;                   // [100][F]
  SchLayer it[kSchIters];
  MlpDesc beta, focus, element, cont, kappa;
  long long crW0, crb0, crW1, crb1, crW2, crb2;
  long long p_logstd;
  long long n_params;
};

__device__ __forceinline__ float ssp(float x) {      // shifted softplus, softplus threshold 20 like torch
  return (x > 20.f ? x : log1pf(expf(x))) - kLn2;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct SchWs {
  int* n_mol;          // [3B]
  float* xs;           // [4][3B][M][F]
  float* agg;          // [3][3B][M][128]
  float* u;            // [3][3B][M][F]
  // backward row buffers (dW problems)
  float* z;            // [3][R][F]
  float* dv;           // [3][R][F]
  float* du;           // [3][R][F]
  float* dy;           // [3][R][128]
  float* gauss;        // [Pn][25]
  float* h1;           // [3][Pn][128]
  float* dw2o;         // [3][Pn][128]
  float* dpre1;        // [3][Pn][128]
  float* dxf;          // [3B][M][F] cotangent of the final features
};

// filter network of one pair: h1 = ssp(W1 g + b1) (shared memory), wf = W2 h1 + b2 (returned per thread f), cut = cosine cutoff
__device__ __forceinline__ void sch_gauss(float r, float* sg) {
  if (threadIdx.x < kSchGauss) {
    const float width = kSchCutoff / (kSchGauss - 1);
    const float df = r - width * threadIdx.x;
    sg[threadIdx.x] = expf(-0.5f / (width * width) * df * df);
  }
}
__device__ __forceinline__ float sch_cut(float r) { return r < kSchCutoff ? 0.5f * (cosf(r * 3.14159265358979323846f / kSchCutoff) + 1.f) : 0.f; }

// One CTA per molecule.  numbers[3B][M] (0 = empty slot, real atoms first), pos[3B][M][3].
__global__ void __launch_bounds__(kSchThreads)
k_sch_fwd(const IntDesc* __restrict__ dp, const float* __restrict__ P, const int* __restrict__ numbers, const float* __restrict__ pos,
          SchWs w, int n_mols) {
  const IntDesc& d = *dp;
  const int mol = blockIdx.x, M = d.M, F = d.F, f = threadIdx.x;
  MGB_DYN_SMEM(float, sm);
  float* sx = sm;                       // [M][F]
  float* sy = sx + M * F;               // [M][128]
  float* sagg = sy + M * kSchFilters;   // [M][128]
  float* sz = sagg + M * kSchFilters;   // [M][F]
  float* sh1 = sz + M * F;              // [128]
  float* sg = sh1 + kSchFilters;        // [32]
  float* spos = sg + 32;                // [M][3]
  __shared__ int s_m;
  if (threadIdx.x == 0) {
    int m = 0;
    for (int i = 0; i < M; ++i) m += numbers[mol * M + i] > 0 ? 1 : 0;
    s_m = m;
    w.n_mol[mol] = m;
  }
  for (int idx = threadIdx.x; idx < M * 3; idx += blockDim.x) spos[idx] = pos[(long long)mol * M * 3 + idx];
  __syncthreads();
  const int m = s_m;
  for (int idx = threadIdx.x; idx < M * F; idx += blockDim.x) {
    const int i = idx / F, k = idx - i * F;
    sx[idx] = i < m ? P[d.p_emb + (long long)numbers[mol * M + i] * F + k] : 0.f;
  }
  __syncthreads();
  const long long R = (long long)n_mols * M;
  for (int t = 0; t < kSchIters; ++t) {
    const SchLayer& Ly = d.it[t];
    for (int idx = threadIdx.x; idx < M * F; idx += blockDim.x) w.xs[((long long)t * R + (long long)mol * M) * F + idx] = sx[idx];
    // y = in2f x
    for (int j = 0; j < m; ++j) {
      float acc = 0.f;
      const float* wr = P + Ly.in2f + (long long)f * F;
      for (int k = 0; k < F; ++k) acc = fmaf(wr[k], sx[j * F + k], acc);
      sy[j * kSchFilters + f] = acc;
    }
    for (int i = 0; i < M; ++i) sagg[i * kSchFilters + f] = 0.f;
    __syncthreads();
    for (int i = 0; i < m; ++i)
      for (int j = i + 1; j < m; ++j) {
        const float dx = spos[i * 3] - spos[j * 3], dy = spos[i * 3 + 1] - spos[j * 3 + 1], dz = spos[i * 3 + 2] - spos[j * 3 + 2];
        const float r = sqrtf(dx * dx + dy * dy + dz * dz);
        sch_gauss(r, sg);
        __syncthreads();
        {
          float acc = P[Ly.b1 + f];
          const float* wr = P + Ly.W1 + (long long)f * kSchGauss;
          for (int g = 0; g < kSchGauss; ++g) acc = fmaf(wr[g], sg[g], acc);
          sh1[f] = ssp(acc);
        }
        __syncthreads();
        float wf = P[Ly.b2 + f];
        const float* wr = P + Ly.W2 + (long long)f * kSchFilters;
#pragma unroll 8
        for (int h = 0; h < kSchFilters; ++h) wf = fmaf(wr[h], sh1[h], wf);
        wf *= sch_cut(r);
        sagg[i * kSchFilters + f] = fmaf(sy[j * kSchFilters + f], wf, sagg[i * kSchFilters + f]);
        sagg[j * kSchFilters + f] = fmaf(sy[i * kSchFilters + f], wf, sagg[j * kSchFilters + f]);
        __syncthreads();
      }
    for (int i = 0; i < M; ++i) w.agg[((long long)t * R + (long long)mol * M + i) * kSchFilters + f] = sagg[i * kSchFilters + f];
    __syncthreads();
    // u = f2out(agg), z = ssp(u)
    for (int idx = threadIdx.x; idx < m * F; idx += blockDim.x) {
      const int i = idx / F, o = idx - i * F;
      float acc = P[Ly.bo + o];
      const float* wr = P + Ly.Wo + (long long)o * kSchFilters;
#pragma unroll 8
      for (int q = 0; q < kSchFilters; ++q) acc = fmaf(wr[q], sagg[i * kSchFilters + q], acc);
      w.u[((long long)t * R + (long long)mol * M + i) * F + o] = acc;
      sz[idx] = ssp(acc);
    }
    __syncthreads();
    // x += dense(z)
    for (int idx = threadIdx.x; idx < m * F; idx += blockDim.x) {
      const int i = idx / F, o = idx - i * F;
      float acc = P[Ly.bd + o];
      const float* wr = P + Ly.Wd + (long long)o * F;
      for (int k = 0; k < F; ++k) acc = fmaf(wr[k], sz[i * F + k], acc);
      sx[idx] += acc;
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < M * F; idx += blockDim.x) w.xs[((long long)kSchIters * R + (long long)mol * M) * F + idx] = sx[idx];
}

// Backward of k_sch_fwd.  Reads w.dxf (cotangent of the final features), writes the row buffers of the weight-gradient
// problems (all buffers are zero-initialised by the caller; only valid rows are written) and the embedding cotangent.
__global__ void __launch_bounds__(kSchThreads)
k_sch_bwd(const IntDesc* __restrict__ dp, const float* __restrict__ P, const int* __restrict__ numbers, const float* __restrict__ pos,
          SchWs w, int n_mols, float* __restrict__ grad) {
  const IntDesc& d = *dp;
  const int mol = blockIdx.x, M = d.M, F = d.F, f = threadIdx.x;
  MGB_DYN_SMEM(float, sm);
  float* sx = sm;                        // [M][F]
  float* sy = sx + M * F;                // [M][128]
  float* sdagg = sy + M * kSchFilters;   // [M][128]
  float* sdy = sdagg + M * kSchFilters;  // [M][128]
  float* sdx = sdy + M * kSchFilters;    // [M][F]
  float* sdu = sdx + M * F;              // [M][F]
  float* sh1 = sdu + M * F;              // [128]
  float* sdwf = sh1 + kSchFilters;       // [128]
  float* spre = sdwf + kSchFilters;      // [128]
  float* sg = spre + kSchFilters;        // [32]
  float* spos = sg + 32;                 // [M][3]
  const int m = w.n_mol[mol];
  const long long R = (long long)n_mols * M, Pn = R * M;
  for (int idx = threadIdx.x; idx < M * 3; idx += blockDim.x) spos[idx] = pos[(long long)mol * M * 3 + idx];
  for (int idx = threadIdx.x; idx < M * F; idx += blockDim.x) sdx[idx] = w.dxf[(long long)mol * M * F + idx];
  __syncthreads();
  for (int t = kSchIters - 1; t >= 0; --t) {
    const SchLayer& Ly = d.it[t];
    const long long row0 = (long long)t * R + (long long)mol * M;
    for (int idx = threadIdx.x; idx < M * F; idx += blockDim.x) sx[idx] = w.xs[row0 * F + idx];
    __syncthreads();
    // dense: dv = dx ; dz = Wd^T dv ; du = dz * ssp'(u)
    for (int idx = threadIdx.x; idx < m * F; idx += blockDim.x) {
      const int i = idx / F, k = idx - i * F;
      w.dv[row0 * F + idx] = sdx[idx];
      const float uu = w.u[row0 * F + idx];
      w.z[row0 * F + idx] = ssp(uu);
      float acc = 0.f;
      for (int o = 0; o < F; ++o) acc = fmaf(P[Ly.Wd + (long long)o * F + k], sdx[i * F + o], acc);
      const float du = acc * sigmoidf_(uu);
      sdu[idx] = du;
      w.du[row0 * F + idx] = du;
    }
    __syncthreads();
    // dagg = Wo^T du ; y = in2f x
    for (int i = 0; i < m; ++i) {
      float acc = 0.f;
      for (int o = 0; o < F; ++o) acc = fmaf(P[Ly.Wo + (long long)o * kSchFilters + f], sdu[i * F + o], acc);
      sdagg[i * kSchFilters + f] = acc;
      float yy = 0.f;
      const float* wr = P + Ly.in2f + (long long)f * F;
      for (int k = 0; k < F; ++k) yy = fmaf(wr[k], sx[i * F + k], yy);
      sy[i * kSchFilters + f] = yy;
      sdy[i * kSchFilters + f] = 0.f;
    }
    __syncthreads();
    for (int i = 0; i < m; ++i)
      for (int j = i + 1; j < m; ++j) {
        const float dx = spos[i * 3] - spos[j * 3], dy = spos[i * 3 + 1] - spos[j * 3 + 1], dz = spos[i * 3 + 2] - spos[j * 3 + 2];
        const float r = sqrtf(dx * dx + dy * dy + dz * dz);
        const long long prow = ((long long)mol * M + i) * M + j;
        sch_gauss(r, sg);
        __syncthreads();
        if (t == 0 && threadIdx.x < kSchGauss) w.gauss[prow * kSchGauss + threadIdx.x] = sg[threadIdx.x];
        {
          float acc = P[Ly.b1 + f];
          const float* wr = P + Ly.W1 + (long long)f * kSchGauss;
          for (int g = 0; g < kSchGauss; ++g) acc = fmaf(wr[g], sg[g], acc);
          spre[f] = acc;
          const float hh = ssp(acc);
          sh1[f] = hh;
          w.h1[((long long)t * Pn + prow) * kSchFilters + f] = hh;
        }
        __syncthreads();
        float wf = P[Ly.b2 + f];
        const float* wr = P + Ly.W2 + (long long)f * kSchFilters;
#pragma unroll 8
        for (int h = 0; h < kSchFilters; ++h) wf = fmaf(wr[h], sh1[h], wf);
        const float cut = sch_cut(r);
        const float gi = sdagg[i * kSchFilters + f], gj = sdagg[j * kSchFilters + f];
        sdy[j * kSchFilters + f] = fmaf(gi, wf * cut, sdy[j * kSchFilters + f]);
        sdy[i * kSchFilters + f] = fmaf(gj, wf * cut, sdy[i * kSchFilters + f]);
        const float dwf = (gi * sy[j * kSchFilters + f] + gj * sy[i * kSchFilters + f]) * cut;
        sdwf[f] = dwf;
        w.dw2o[((long long)t * Pn + prow) * kSchFilters + f] = dwf;
        __syncthreads();
        {
          float acc = 0.f;   // dh1[h = f] = sum_q W2[q][h] dwf[q]
#pragma unroll 8
          for (int q = 0; q < kSchFilters; ++q) acc = fmaf(P[Ly.W2 + (long long)q * kSchFilters + f], sdwf[q], acc);
          w.dpre1[((long long)t * Pn + prow) * kSchFilters + f] = acc * sigmoidf_(spre[f]);
        }
        __syncthreads();
      }
    for (int i = 0; i < m; ++i) w.dy[(row0 + i) * kSchFilters + f] = sdy[i * kSchFilters + f];
    __syncthreads();
    // dx += in2f^T dy
    for (int idx = threadIdx.x; idx < m * F; idx += blockDim.x) {
      const int j = idx / F, k = idx - j * F;
      float acc = sdx[idx];
      for (int q = 0; q < kSchFilters; ++q) acc = fmaf(P[Ly.in2f + (long long)q * F + k], sdy[j * kSchFilters + q], acc);
      sdu[idx] = acc;   // staging
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < m * F; idx += blockDim.x) sdx[idx] = sdu[idx];
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < m * F; idx += blockDim.x) {
    const int i = idx / F, k = idx - i * F;
    if (sdx[idx] != 0.f) atomicAdd(grad + d.p_emb + (long long)numbers[mol * M + i] * F + k, sdx[idx]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Heads (agent.py:186-353), one CTA per canvas.
// ------------------------------------------------------------------------------------------------------------
struct IntHeadBufs {   // rows for the weight-gradient problems (X, hidden H, dH, dOut)
  float* Xb; float* Hb; float* dHb; float* dOb;     // beta: [2B] rows
  float* Xf; float* Hf; float* dHf; float* dOf;     // focus: [B*N] rows
  float* Xe; float* He; float* dHe; float* dOe;     // element: [B]
  float* Xc; float* Hc; float* dHc; float* dOc;     // continuous: [B]
  float* Xk; float* Hk; float* dHk; float* dOk;     // kappa: [2B]
  float* Xv; float* H1; float* H2; float* dH1; float* dH2; float* dOv;   // critic: [B]
};
struct IntOutputs {
  float* logp; float* ent; float* v; float* logp_terms /* [B,6] */; float* focus_probs /* [B,N] */; float* element_probs /* [B,Z] */;
  float* means /* [B,3] */; float* kappa_logits /* [B,2] */;
};

constexpr int kIntHeadThreads = 128;

// hidden = relu(W0 x + b0) ; out = W1 hidden + b1   (x, hidden, out in shared memory)
__device__ __forceinline__ void mlp2_fwd(const MlpDesc& M, const float* __restrict__ P, const float* x, float* hid, float* out) {
  gemv_rows(P + M.W0, P + M.b0, x, M.in, M.hidden, true, hid);
  __syncthreads();
  gemv_rows(P + M.W1, P + M.b1, hid, M.hidden, M.out, false, out);
  __syncthreads();
}
// dhid = relu'(hid) * W1^T dout ; dx = W0^T dhid     (all shared memory)
__device__ __forceinline__ void mlp2_bwd(const MlpDesc& M, const float* __restrict__ P, const float* hid, const float* dout, float* dhid,
                                         float* dx) {
  for (int h = threadIdx.x; h < M.hidden; h += blockDim.x) {
    float acc = 0.f;
    for (int o = 0; o < M.out; ++o) acc = fmaf(P[M.W1 + (long long)o * M.hidden + h], dout[o], acc);
    dhid[h] = hid[h] > 0.f ? acc : 0.f;
  }
  __syncthreads();
  gemv_n(P + M.W0, dhid, M.hidden, M.in, dx);
  __syncthreads();
}
__device__ __forceinline__ void store_row(float* dst, long long row, int n, const float* src) {
  for (int k = threadIdx.x; k < n; k += blockDim.x) dst[row * n + k] = src[k];
}

struct IntSmem {
  float *count, *countn, *lb, *lbn, *hb, *hbn, *lat, *hf, *fl, *flog, *foc, *he, *el, *elog, *xc, *hc, *yc, *xk0, *xk1, *hk0, *hk1, *yk,
      *xv, *h1, *h2, *yv, *dout, *dhid, *dx, *dlb, *dlbn, *red, *misc;
};
__host__ __device__ inline int int_smem_floats(const IntDesc& d) {
  return 2 * d.Z + 2 * d.LB + 2 * d.Wd + d.N * d.lat + d.N * d.Wd + 2 * d.N + d.lat + d.Wd + 2 * d.Z + (d.lat + d.Z) + d.Wd + 4 + 2 * d.lat +
         2 * d.Wd + 4 + d.lat + 2 * d.Wd + 4 + 16 + d.Wd + (d.lat + d.Z) + 2 * d.LB + 64 + 64;
}
__device__ __forceinline__ IntSmem int_smem_carve(const IntDesc& d, float* p) {
  IntSmem s;
  auto take = [&](int n) { float* r = p; p += n; return r; };
  s.count = take(d.Z); s.countn = take(d.Z); s.lb = take(d.LB); s.lbn = take(d.LB); s.hb = take(d.Wd); s.hbn = take(d.Wd);
  s.lat = take(d.N * d.lat); s.hf = take(d.N * d.Wd); s.fl = take(d.N); s.flog = take(d.N); s.foc = take(d.lat); s.he = take(d.Wd);
  s.el = take(d.Z); s.elog = take(d.Z); s.xc = take(d.lat + d.Z); s.hc = take(d.Wd); s.yc = take(4);
  s.xk0 = take(d.lat); s.xk1 = take(d.lat); s.hk0 = take(d.Wd); s.hk1 = take(d.Wd); s.yk = take(4);
  s.xv = take(d.lat); s.h1 = take(d.Wd); s.h2 = take(d.Wd); s.yv = take(4);
  s.dout = take(16); s.dhid = take(d.Wd); s.dx = take(d.lat + d.Z); s.dlb = take(d.LB); s.dlbn = take(d.LB); s.red = take(64); s.misc = take(64);
  return s;
}

struct IntScalars {
  int n, nact, focus, element, kappa;
  float ent_f, ent_e, amask[6], logp_t[6], v;
  SoftmaxAux aux_f, aux_e;
};

__device__ inline void int_heads_forward(const IntDesc& d, const float* __restrict__ P, int b, int B, const SchWs& w,
                                         const float* __restrict__ bags, const float* __restrict__ actions, IntSmem& s, IntScalars& q) {
  const int N = d.N, M = d.M, F = d.F, LB = d.LB, Z = d.Z, Wd = d.Wd, lat = d.lat;
  const long long R = (long long)3 * B * M;
  const float* act = actions + (long long)b * 7;
  const int n = w.n_mol[b * 3];
  q.n = n; q.nact = n > 1 ? n : 1;
  q.focus = (int)rintf(act[1]); q.element = (int)rintf(act[2]); q.kappa = (int)rintf(act[6]);
  q.amask[0] = n >= 1; q.amask[1] = 1.f; q.amask[2] = n >= 1; q.amask[3] = n >= 2; q.amask[4] = n >= 3; q.amask[5] = n >= 3;
  const float* xfin = w.xs + (long long)kSchIters * R * F;
  for (int z = threadIdx.x; z < Z; z += blockDim.x) {
    s.count[z] = bags[(long long)b * Z + z];
    s.countn[z] = bags[(long long)b * Z + z] - (z == q.element ? 1.f : 0.f);
  }
  __syncthreads();
  mlp2_fwd(d.beta, P, s.count, s.hb, s.lb);
  mlp2_fwd(d.beta, P, s.countn, s.hbn, s.lbn);
  for (int idx = threadIdx.x; idx < N * lat; idx += blockDim.x) {
    const int a = idx / lat, k = idx - a * lat;
    s.lat[idx] = k < F ? (a < n ? xfin[((long long)(b * 3) * M + a) * F + k] : 0.f) : s.lb[k - F];
  }
  for (int k = threadIdx.x; k < lat; k += blockDim.x) {
    s.xk0[k] = k < F ? xfin[((long long)(b * 3 + 1) * M + n) * F + k] : s.lbn[k - F];
    s.xk1[k] = k < F ? xfin[((long long)(b * 3 + 2) * M + n) * F + k] : s.lbn[k - F];
  }
  __syncthreads();
  for (int a = 0; a < q.nact; ++a) mlp2_fwd(d.focus, P, s.lat + a * lat, s.hf + a * Wd, s.fl + a);
  if (threadIdx.x == 0) {
    bool mask[64];
    for (int a = 0; a < N; ++a) { mask[a] = a < q.nact; if (!mask[a]) s.fl[a] = 0.f; }
    q.ent_f = categorical_fwd(s.fl, s.flog, mask, N, &q.aux_f);
    s.misc[0] = q.ent_f; s.misc[1] = q.aux_f.mx; s.misc[2] = q.aux_f.s1; s.misc[3] = q.aux_f.s2;
  }
  for (int k = threadIdx.x; k < lat; k += blockDim.x) {
    const float v = s.lat[q.focus * lat + k];
    s.foc[k] = v;
    s.xc[k] = v;
  }
  for (int z = threadIdx.x; z < Z; z += blockDim.x) s.xc[lat + z] = z == q.element ? 1.f : 0.f;
  for (int k = threadIdx.x; k < lat; k += blockDim.x) {
    float acc = 0.f;
    if (k < F) { for (int a = 0; a < q.nact; ++a) acc += s.lat[a * lat + k]; } else acc = s.lb[k - F];
    s.xv[k] = acc;
  }
  __syncthreads();
  mlp2_fwd(d.element, P, s.foc, s.he, s.el);
  if (threadIdx.x == 0) {
    bool mask[MGB_MAX_SPECIES];
    for (int z = 0; z < Z; ++z) mask[z] = s.count[z] > 0.f;
    q.ent_e = categorical_fwd(s.el, s.elog, mask, Z, &q.aux_e);
    s.misc[4] = q.ent_e; s.misc[5] = q.aux_e.mx; s.misc[6] = q.aux_e.s1; s.misc[7] = q.aux_e.s2;
  }
  mlp2_fwd(d.cont, P, s.xc, s.hc, s.yc);
  mlp2_fwd(d.kappa, P, s.xk0, s.hk0, s.yk);
  mlp2_fwd(d.kappa, P, s.xk1, s.hk1, s.yk + 1);
  // critic: three layers
  gemv_rows(P + d.crW0, P + d.crb0, s.xv, lat, Wd, true, s.h1);
  __syncthreads();
  gemv_rows(P + d.crW1, P + d.crb1, s.h1, Wd, Wd, true, s.h2);
  __syncthreads();
  gemv_rows(P + d.crW2, P + d.crb2, s.h2, Wd, 1, false, s.yv);
  __syncthreads();
  q.ent_f = s.misc[0]; q.aux_f.mx = s.misc[1]; q.aux_f.s1 = s.misc[2]; q.aux_f.s2 = s.misc[3];
  q.ent_e = s.misc[4]; q.aux_e.mx = s.misc[5]; q.aux_e.s1 = s.misc[6]; q.aux_e.s2 = s.misc[7];
  q.v = s.yv[0];
  q.logp_t[0] = s.flog[q.focus];
  q.logp_t[1] = s.elog[q.element];
  const float width[3] = {d.dmax - d.dmin, 3.14159265358979323846f, 3.14159265358979323846f};
  const float center[3] = {0.5f * (d.dmax + d.dmin), 0.5f * 3.14159265358979323846f, 0.5f * 3.14159265358979323846f};
  for (int c = 0; c < 3; ++c) {
    const float mean = tanhf(s.yc[c]) * width[c] / 2.f + center[c];
    const float sd = expf(1e-6f + P[d.p_logstd + c]);
    const float df = act[3 + c] - mean;
    q.logp_t[2 + c] = -(df * df) / (2.f * sd * sd) - logf(sd) - kLogSqrt2Pi;
  }
  {
    const float mx = fmaxf(s.yk[0], s.yk[1]);
    const float lse = mx + logf(expf(s.yk[0] - mx) + expf(s.yk[1] - mx));
    q.logp_t[5] = s.yk[q.kappa] - lse;
  }
}

__global__ void __launch_bounds__(kIntHeadThreads)
k_int_heads_fwd(const IntDesc* __restrict__ dp, const float* __restrict__ P, int B, SchWs w, const float* __restrict__ bags,
                const float* __restrict__ actions, IntOutputs out) {
  const IntDesc& d = *dp;
  MGB_DYN_SMEM(float, sm);
  IntSmem s = int_smem_carve(d, sm);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    IntScalars q;
    __syncthreads();
    int_heads_forward(d, P, b, B, w, bags, actions, s, q);
    if (threadIdx.x == 0) {
      float lp = 0.f;
      for (int c = 0; c < 6; ++c) {
        lp += q.logp_t[c] * q.amask[c];
        if (out.logp_terms) out.logp_terms[b * 6 + c] = q.logp_t[c] * q.amask[c];
      }
      out.logp[b] = lp;
      out.ent[b] = q.ent_f * q.amask[0] + q.ent_e * q.amask[1];
      out.v[b] = q.v;
      if (out.means) {
        const float width[3] = {d.dmax - d.dmin, 3.14159265358979323846f, 3.14159265358979323846f};
        const float center[3] = {0.5f * (d.dmax + d.dmin), 0.5f * 3.14159265358979323846f, 0.5f * 3.14159265358979323846f};
        for (int c = 0; c < 3; ++c) out.means[b * 3 + c] = tanhf(s.yc[c]) * width[c] / 2.f + center[c];
      }
      if (out.kappa_logits) { out.kappa_logits[b * 2] = s.yk[0]; out.kappa_logits[b * 2 + 1] = s.yk[1]; }
    }
    if (out.focus_probs) for (int a = threadIdx.x; a < d.N; a += blockDim.x) out.focus_probs[(long long)b * d.N + a] = s.fl[a];
    if (out.element_probs) for (int z = threadIdx.x; z < d.Z; z += blockDim.x) out.element_probs[(long long)b * d.Z + z] = s.el[z];
  }
}

__global__ void __launch_bounds__(kIntHeadThreads)
k_int_heads_bwd(const IntDesc* __restrict__ dp, const float* __restrict__ P, int B, SchWs w, const float* __restrict__ bags,
                const float* __restrict__ actions, const float* __restrict__ g_logp, const float* __restrict__ g_ent,
                const float* __restrict__ g_v, IntHeadBufs hb, float* __restrict__ grad) {
  const IntDesc& d = *dp;
  MGB_DYN_SMEM(float, sm);
  IntSmem s = int_smem_carve(d, sm);
  float* scr = sm + int_smem_floats(d);                 // [Wd] scratch
  float* dfoc = scr + d.Wd;                             // [lat] cotangent of the focused latent state
  float* dfl = dfoc + d.lat;                            // [N] cotangent of the focus logits
  const int N = d.N, M = d.M, F = d.F, LB = d.LB, Z = d.Z, Wd = d.Wd, lat = d.lat;
  float acc_logstd = 0.f;   // thread c < 3 owns log_std c
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    IntScalars q;
    __syncthreads();
    int_heads_forward(d, P, b, B, w, bags, actions, s, q);
    const float gl = g_logp[b], ge = g_ent[b], gv = g_v[b];
    const float* act = actions + (long long)b * 7;
    float* dxf0 = w.dxf + (long long)(b * 3) * M * F;
    for (int k = threadIdx.x; k < LB; k += blockDim.x) { s.dlb[k] = 0.f; s.dlbn[k] = 0.f; }
    for (int k = threadIdx.x; k < lat; k += blockDim.x) dfoc[k] = 0.f;
    __syncthreads();
    // ---- critic (three layers)
    store_row(hb.Xv, b, lat, s.xv); store_row(hb.H1, b, Wd, s.h1); store_row(hb.H2, b, Wd, s.h2);
    if (threadIdx.x == 0) hb.dOv[b] = gv;
    for (int h = threadIdx.x; h < Wd; h += blockDim.x) {
      const float g2 = s.h2[h] > 0.f ? P[d.crW2 + h] * gv : 0.f;
      s.dhid[h] = g2;
      hb.dH2[(long long)b * Wd + h] = g2;
    }
    __syncthreads();
    gemv_n(P + d.crW1, s.dhid, Wd, Wd, scr);
    __syncthreads();
    for (int h = threadIdx.x; h < Wd; h += blockDim.x) {
      const float g1 = s.h1[h] > 0.f ? scr[h] : 0.f;
      s.dhid[h] = g1;
      hb.dH1[(long long)b * Wd + h] = g1;
    }
    __syncthreads();
    gemv_n(P + d.crW0, s.dhid, Wd, lat, s.dx);
    __syncthreads();
    for (int k = threadIdx.x; k < lat; k += blockDim.x) {
      if (k < F) { for (int a = 0; a < q.n; ++a) dxf0[a * F + k] += s.dx[k]; } else s.dlb[k - F] += s.dx[k];
    }
    __syncthreads();
    // ---- kappa (two applications of phi_kappa)
    {
      const float mx = fmaxf(s.yk[0], s.yk[1]);
      const float e0 = expf(s.yk[0] - mx), e1 = expf(s.yk[1] - mx);
      const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
      const float gk = gl * q.amask[5];
      const float dk[2] = {gk * ((q.kappa == 0 ? 1.f : 0.f) - p0), gk * ((q.kappa == 1 ? 1.f : 0.f) - p1)};
      for (int v = 0; v < 2; ++v) {
        const float* xk = v == 0 ? s.xk0 : s.xk1;
        const float* hk = v == 0 ? s.hk0 : s.hk1;
        if (threadIdx.x == 0) { s.dout[0] = dk[v]; hb.dOk[(long long)b * 2 + v] = dk[v]; }
        __syncthreads();
        mlp2_bwd(d.kappa, P, hk, s.dout, s.dhid, s.dx);
        store_row(hb.Xk, (long long)b * 2 + v, lat, xk); store_row(hb.Hk, (long long)b * 2 + v, Wd, hk);
        store_row(hb.dHk, (long long)b * 2 + v, Wd, s.dhid);
        for (int k = threadIdx.x; k < lat; k += blockDim.x) {
          if (k < F) w.dxf[((long long)(b * 3 + 1 + v) * M + q.n) * F + k] += s.dx[k]; else s.dlbn[k - F] += s.dx[k];
        }
        __syncthreads();
      }
    }
    // ---- continuous head (distance, angle, dihedral Normals, agent.py:243-281)
    if (threadIdx.x < 3) {
      const int c = threadIdx.x;
      const float width[3] = {d.dmax - d.dmin, 3.14159265358979323846f, 3.14159265358979323846f};
      const float center[3] = {0.5f * (d.dmax + d.dmin), 0.5f * 3.14159265358979323846f, 0.5f * 3.14159265358979323846f};
      const float th = tanhf(s.yc[c]);
      const float mean = th * width[c] / 2.f + center[c];
      const float sd = expf(1e-6f + P[d.p_logstd + c]);
      const float df = act[3 + c] - mean;
      const float g = gl * q.amask[2 + c];
      const float dmean = g * df / (sd * sd);
      const float dsd = g * (df * df / (sd * sd * sd) - 1.f / sd);
      acc_logstd += dsd * sd;
      s.dout[c] = dmean * width[c] / 2.f * (1.f - th * th);
      hb.dOc[(long long)b * 3 + c] = s.dout[c];
    }
    __syncthreads();
    mlp2_bwd(d.cont, P, s.hc, s.dout, s.dhid, s.dx);
    store_row(hb.Xc, b, lat + Z, s.xc); store_row(hb.Hc, b, Wd, s.hc); store_row(hb.dHc, b, Wd, s.dhid);
    for (int k = threadIdx.x; k < lat; k += blockDim.x) dfoc[k] += s.dx[k];
    __syncthreads();
    // ---- element head
    if (threadIdx.x == 0) {
      bool mask[MGB_MAX_SPECIES];
      for (int z = 0; z < Z; ++z) mask[z] = s.count[z] > 0.f;
      categorical_bwd(s.el, s.elog, mask, Z, q.element, gl * q.amask[1], ge * q.amask[1], q.aux_e, s.dout);
      for (int z = 0; z < Z; ++z) hb.dOe[(long long)b * Z + z] = s.dout[z];
    }
    __syncthreads();
    mlp2_bwd(d.element, P, s.he, s.dout, s.dhid, s.dx);
    store_row(hb.Xe, b, lat, s.foc); store_row(hb.He, b, Wd, s.he); store_row(hb.dHe, b, Wd, s.dhid);
    for (int k = threadIdx.x; k < lat; k += blockDim.x) dfoc[k] += s.dx[k];
    // ---- focus head over the active atoms
    if (threadIdx.x == 0) {
      bool mask[64];
      for (int a = 0; a < N; ++a) mask[a] = a < q.nact;
      categorical_bwd(s.fl, s.flog, mask, N, q.focus, gl * q.amask[0], ge * q.amask[0], q.aux_f, dfl);
    }
    __syncthreads();
    for (int a = 0; a < q.nact; ++a) {
      if (threadIdx.x == 0) { s.dout[0] = dfl[a]; hb.dOf[(long long)b * N + a] = dfl[a]; }
      __syncthreads();
      mlp2_bwd(d.focus, P, s.hf + a * Wd, s.dout, s.dhid, s.dx);
      store_row(hb.Xf, (long long)b * N + a, lat, s.lat + a * lat); store_row(hb.Hf, (long long)b * N + a, Wd, s.hf + a * Wd);
      store_row(hb.dHf, (long long)b * N + a, Wd, s.dhid);
      for (int k = threadIdx.x; k < lat; k += blockDim.x) {
        const float g = s.dx[k] + (a == q.focus ? dfoc[k] : 0.f);
        if (k < F) { if (a < q.n) dxf0[a * F + k] += g; } else s.dlb[k - F] += g;
      }
      __syncthreads();
    }
    // ---- bag latent (phi_beta applied to the bag and to the bag after the element is taken out)
    mlp2_bwd(d.beta, P, s.hb, s.dlb, s.dhid, s.dx);
    store_row(hb.Xb, (long long)b * 2, Z, s.count); store_row(hb.Hb, (long long)b * 2, Wd, s.hb);
    store_row(hb.dHb, (long long)b * 2, Wd, s.dhid); store_row(hb.dOb, (long long)b * 2, LB, s.dlb);
    __syncthreads();
    mlp2_bwd(d.beta, P, s.hbn, s.dlbn, s.dhid, s.dx);
    store_row(hb.Xb, (long long)b * 2 + 1, Z, s.countn); store_row(hb.Hb, (long long)b * 2 + 1, Wd, s.hbn);
    store_row(hb.dHb, (long long)b * 2 + 1, Wd, s.dhid); store_row(hb.dOb, (long long)b * 2 + 1, LB, s.dlbn);
    __syncthreads();
  }
  if (threadIdx.x < 3 && acc_logstd != 0.f) atomicAdd(grad + d.p_logstd + threadIdx.x, acc_logstd);
}
__host__ __device__ inline int int_bwd_extra_floats(const IntDesc& d) { return d.Wd + d.lat + d.N + 16; }

}  // namespace mgb
