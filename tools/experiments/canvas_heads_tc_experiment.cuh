// Experiment (measured and dropped, see tools/experiments/README.md): the per-canvas element / distance / value MLP chains as batched
// tensor-core kernels over the canvases.  Fragment of mlp_tc.cuh; not compiled.
// ------------------------------------------------------------------------------------------------------------
// The per-canvas heads as batched GEMMs (rows = canvases): element head phi_element on the focused atom's invariants
// (agent.py:243-247), distance head phi_d on the invariants of the chosen element block (agent.py:256-264), value head phi_v on
// the summed value transforms (agent.py:313-316).  In evaluate mode the focus / element sub-actions are inputs, so none of these
// depends on a softmax: they run as tensor-core kernels over all canvases and the one-CTA-per-canvas policy kernels keep only the
// softmax / mixture / mixer / quadrature arithmetic.  (Rollout mode draws the sub-actions one after the other and keeps the
// per-canvas chains: heads.cuh::k_policy_sample.)
// ------------------------------------------------------------------------------------------------------------
enum CanvasRoleKind { kRoleElement = 0, kRoleDist = 1, kRoleValue = 2 };
struct CanvasRole {
  int kind, K, No;
  long long W0, b0, W1, b1;   // float offsets into the flat parameter / gradient buffer
  float* X;                   // [B][K]  gathered inputs (kept for the weight gradient)
  float* H;                   // [B][Wd] hidden activations
  float* Y;                   // [B][No] outputs
  const float* dY;            // [B][No] output cotangents (backward)
  float* dH;                  // [B][Wd] hidden cotangents (backward; kept for the weight gradient)
  float* dX;                  // backward destination: element -> dinv [B,N,lat] (+=, row of the focused atom), dist -> [B][latE], value -> [B][Wd]
};
struct CanvasRoleList {
  CanvasRole r[3];
  int n;
};

__host__ __device__ inline size_t canvas_mlp_tc_smem_bytes(int Kmax, int Wd, bool backward) {
  const size_t w0 = (size_t)Wd * (backward ? tc_stride_kn(Kmax) : tc_stride_nk(Kmax));
  return sizeof(float) * (w0 + (size_t)kTcRows * tc_stride_nk(Kmax) + (size_t)kTcRows * tc_stride_nk(Wd) + 16 + 4);
}

template <int NT>
__global__ void __launch_bounds__(kTcThreads)
k_canvas_mlp_fwd_tc(const CovDesc* __restrict__ dp, const float* __restrict__ P, const MGB_GRID_CONSTANT CanvasRoleList roles, int B,
                    const int* __restrict__ n_atoms, const float* __restrict__ actions, const float* __restrict__ A_last,
                    const float* __restrict__ inv, const float* __restrict__ trans) {
  const CovDesc& d = *dp;
  const CanvasRole& R = roles.r[blockIdx.y];
  if ((int)(blockIdx.x * kTcRows) >= B) return;
  const int K = R.K, Wd = d.Wd, No = R.No, N = d.N, CPE = d.CPE, tau = d.Cout;
  const int sk = tc_stride_nk(K), sw = tc_stride_nk(Wd);
  MGB_DYN_SMEM(float, sm);
  float* sW0 = sm;                        // [Wd][sk]
  float* sx = sW0 + (size_t)Wd * sk;      // [16][sk]
  float* sh = sx + kTcRows * sk;          // [16][sw]
  __shared__ SmemBarrier s_bar;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  tc_stage_expect(&s_bar, (unsigned)(4 * Wd * K));
  __syncthreads();
  tc_stage_matrix(sW0, sk, P + R.W0, Wd, K, &s_bar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int n0 = warp * 8 * NT;
  bool staged = false;
  for (int r0 = blockIdx.x * kTcRows; r0 < B; r0 += gridDim.x * kTcRows) {
    __syncthreads();
    // ---- gather the input rows of the tile
    if (R.kind == kRoleElement) {
      for (int idx = threadIdx.x; idx < kTcRows * K; idx += blockDim.x) {
        const int q = idx / K, k = idx - q * K, b = r0 + q;
        float v = 0.f;
        if (b < B) {
          const int focus = (int)rintf(actions[(long long)b * 6]);
          v = inv[((long long)b * N + focus) * K + k];
          R.X[(long long)b * K + k] = v;
        }
        sx[q * sk + k] = v;
      }
    } else if (R.kind == kRoleDist) {
      // invariants of the chosen element block of the focused atom (so3_tools.py:147-190 on select_taus, agent.py:256-259)
      const int pairs = (kL + 2) * CPE;
      for (int idx = threadIdx.x; idx < kTcRows * pairs; idx += blockDim.x) {
        const int q = idx / pairs, r = idx - q * pairs, blk = r / CPE, c = r - blk * CPE, b = r0 + q;
        float v0 = 0.f, v1 = 0.f;
        if (b < B) {
          const int focus = (int)rintf(actions[(long long)b * 6]), element = (int)rintf(actions[(long long)b * 6 + 1]);
          if (focus < n_atoms[b]) {
            const float2* a = reinterpret_cast<const float2*>(A_last) + ((long long)b * N + focus) * kM * tau + element * CPE + c;
            if (blk == 0) {
              v0 = a[0].x; v1 = a[0].y;
            } else {
              const int l = blk - 1;
              for (int m = -l; m <= l; ++m) {
                const float2 p = a[lm_index(l, m) * tau], q2 = a[lm_index(l, -m) * tau];
                const float sg = (m & 1) ? -1.f : 1.f;
                v0 += sg * (p.x * q2.x - p.y * q2.y);
                v1 += p.x * p.x + p.y * p.y;
              }
            }
          }
          R.X[(long long)b * K + 2 * r] = v0;
          R.X[(long long)b * K + 2 * r + 1] = v1;
        }
        sx[q * sk + 2 * r] = v0;
        sx[q * sk + 2 * r + 1] = v1;
      }
    } else {
      for (int idx = threadIdx.x; idx < kTcRows * K; idx += blockDim.x) {
        const int q = idx / K, k = idx - q * K, b = r0 + q;
        float v = 0.f;
        if (b < B) {
          const int n = n_atoms[b];
          for (int i = 0; i < n; ++i) v += trans[((long long)b * N + i) * K + k];
          R.X[(long long)b * K + k] = v;
        }
        sx[q * sk + k] = v;
      }
    }
    __syncthreads();
    if (!staged) { mbar_wait(&s_bar, 0); staged = true; }
    // ---- hidden layer on the tensor cores
    float acc[NT][4];
    MGB_UNROLL
    for (int nt = 0; nt < NT; ++nt) {
      const float b0 = P[R.b0 + n0 + 8 * nt + 2 * t], b1 = P[R.b0 + n0 + 8 * nt + 2 * t + 1];
      acc[nt][0] = b0; acc[nt][1] = b1; acc[nt][2] = b0; acc[nt][3] = b1;
    }
    tc_tile_gemm<NT, false>(sx, sk, sW0, sk, K, n0, NT, acc);
    const int ba = r0 + g, bb = r0 + g + 8;
    MGB_UNROLL
    for (int nt = 0; nt < NT; ++nt) {
      const int col = n0 + 8 * nt + 2 * t;
      const float h0 = fmaxf(acc[nt][0], 0.f), h1 = fmaxf(acc[nt][1], 0.f), h2 = fmaxf(acc[nt][2], 0.f), h3 = fmaxf(acc[nt][3], 0.f);
      *reinterpret_cast<float2*>(sh + g * sw + col) = make_float2(h0, h1);
      *reinterpret_cast<float2*>(sh + (g + 8) * sw + col) = make_float2(h2, h3);
      if (ba < B) *reinterpret_cast<float2*>(R.H + (long long)ba * Wd + col) = make_float2(h0, h1);
      if (bb < B) *reinterpret_cast<float2*>(R.H + (long long)bb * Wd + col) = make_float2(h2, h3);
    }
    __syncthreads();
    // ---- the few outputs of the second layer: a warp per (row, output), lanes over the hidden units
    for (int pr = warp; pr < kTcRows * No; pr += kTcThreads / 32) {
      const int q = pr / No, o = pr - q * No;
      float part = 0.f;
      for (int k = lane; k < Wd; k += 32) part = fmaf(P[R.W1 + (long long)o * Wd + k], sh[q * sw + k], part);
      part = warp_sum(part);
      if (lane == 0 && r0 + q < B) R.Y[(long long)(r0 + q) * No + o] = part + P[R.b1 + o];
    }
  }
}

template <int NT, int NTX>
__global__ void __launch_bounds__(kTcThreads)
k_canvas_mlp_bwd_tc(const CovDesc* __restrict__ dp, const float* __restrict__ P, const MGB_GRID_CONSTANT CanvasRoleList roles, int B,
                    const float* __restrict__ actions) {
  const CovDesc& d = *dp;
  const CanvasRole& R = roles.r[blockIdx.y];
  if ((int)(blockIdx.x * kTcRows) >= B) return;
  const int K = R.K, Wd = d.Wd, No = R.No, N = d.N;
  const int sk = tc_stride_kn(K), sa = tc_stride_nk(Wd);
  MGB_DYN_SMEM(float, sm);
  float* sW0 = sm;                        // [Wd][sk]   W0 [hidden][in]: B[k = hidden][n = input]
  float* sdh = sW0 + (size_t)Wd * sk;     // [16][sa]
  __shared__ SmemBarrier s_bar;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  tc_stage_expect(&s_bar, (unsigned)(4 * Wd * K));
  __syncthreads();
  tc_stage_matrix(sW0, sk, P + R.W0, Wd, K, &s_bar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ktiles = K / 8, per = (ktiles + 7) / 8, x0 = warp * per, xcount = max(0, min(per, ktiles - x0));
  bool staged = false;
  for (int r0 = blockIdx.x * kTcRows; r0 < B; r0 += gridDim.x * kTcRows) {
    __syncthreads();
    // dH = relu'(H) * (W1^T dY): few outputs, plain FMAs
    for (int idx = threadIdx.x; idx < kTcRows * Wd; idx += blockDim.x) {
      const int q = idx / Wd, h = idx - q * Wd, b = r0 + q;
      float gq = 0.f;
      if (b < B) {
        if (R.H[(long long)b * Wd + h] > 0.f)
          for (int o = 0; o < No; ++o) gq = fmaf(P[R.W1 + (long long)o * Wd + h], R.dY[(long long)b * No + o], gq);
        R.dH[(long long)b * Wd + h] = gq;
      }
      sdh[q * sa + h] = gq;
    }
    __syncthreads();
    if (!staged) { mbar_wait(&s_bar, 0); staged = true; }
    float accx[NTX][4];
    MGB_UNROLL
    for (int nt = 0; nt < NTX; ++nt) { accx[nt][0] = 0.f; accx[nt][1] = 0.f; accx[nt][2] = 0.f; accx[nt][3] = 0.f; }
    tc_tile_gemm<NTX, true>(sdh, sa, sW0, sk, Wd, 8 * x0, xcount, accx);   // dX = dH W0
    const int ba = r0 + g, bb = r0 + g + 8;
    MGB_UNROLL
    for (int nt = 0; nt < NTX; ++nt) {
      if (nt < xcount) {
        const int col = 8 * (x0 + nt) + 2 * t;
        MGB_UNROLL
        for (int half = 0; half < 2; ++half) {
          const int b = half ? bb : ba;
          if (b >= B) continue;
          const float v0 = accx[nt][2 * half], v1 = accx[nt][2 * half + 1];
          if (R.kind == kRoleElement) {   // into the invariants' cotangent of the focused atom (other heads add to the same rows)
            const int focus = (int)rintf(actions[(long long)b * 6]);
            float* dst = R.dX + ((long long)b * N + focus) * K + col;
            if (v0 != 0.f) atomicAdd(dst, v0);
            if (v1 != 0.f) atomicAdd(dst + 1, v1);
          } else {
            *reinterpret_cast<float2*>(R.dX + (long long)b * K + col) = make_float2(v0, v1);
          }
        }
      }
    }
  }
}

}  // namespace mgb
