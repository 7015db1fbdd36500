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
  int pad_rows;        // dst row count (>= rows; the extra rows are written as zeros)
};

__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef MGB_CUSIM
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

__global__ void k_prep_params(const TransposeSeg* __restrict__ segs, int n_seg, long long total, const float* __restrict__ P,
                              float* __restrict__ Wt, long long n_params, const char* __restrict__ tables, long long table_bytes) {
  // The step starts with cold caches in training (the optimizer just rewrote the parameters; the benchmark flushes L2):
  // pull every parameter line and the constant tables (Clebsch-Gordan, Lebedev) into L2 while the transposes run.
  {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    const char* pb = reinterpret_cast<const char*>(P);
    for (long long off = tid * 128; off < n_params * 4; off += nthr * 128) prefetch_l2(pb + off);
    for (long long off = tid * 128; off < table_bytes; off += nthr * 128) prefetch_l2(tables + off);
  }
  // one thread per destination float (coalesced writes); segments tile [0, total) in order of their dst offsets
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_seg;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (segs[mid].dst <= idx) lo = mid; else hi = mid;
    }
    const TransposeSeg s = segs[lo];
    const int local = (int)(idx - s.dst), e = local % s.elem, t = local / s.elem;
    const int c = t / s.pad_rows, r = t - c * s.pad_rows;
    if (c >= s.cols) continue;   // alignment gap between two segments
    Wt[idx] = r < s.rows ? P[s.src + ((long long)r * s.cols + c) * s.elem + e] : 0.f;
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
// Flat list of valid (b, i, j) pairs: pair_off[b] = sum_{b' < b} n_b'^2, pair_off[B] = total.  One CTA.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_pair_offsets(int B, int N, const int* __restrict__ n_atoms, int* __restrict__ pair_off, int* __restrict__ atom_off,
                               int* __restrict__ atom_list, int* __restrict__ act_off, int* __restrict__ act_list) {
  // three exclusive prefix sums over the canvases: n^2 (pairs), n (valid atoms), max(n, 1) (rows the focus head looks at)
  __shared__ int wsum[3][32];
  const int per = (B + blockDim.x - 1) / blockDim.x;
  const int lo = threadIdx.x * per, hi = min(B, lo + per);
  int own[3] = {0, 0, 0};
  for (int b = lo; b < hi; ++b) {
    const int n = n_atoms[b];
    own[0] += n * n; own[1] += n; own[2] += n > 1 ? n : 1;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc[3];
  for (int q = 0; q < 3; ++q) {
    inc[q] = own[q];
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc[q], o);
      if (lane >= o) inc[q] += t;
    }
    if (lane == 31) wsum[q][wid] = inc[q];
  }
  __syncthreads();
  if (wid == 0) {
    for (int q = 0; q < 3; ++q) {
      const int v = lane < nw ? wsum[q][lane] : 0;
      int s = v;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
      }
      wsum[q][lane] = s - v;   // exclusive prefix of the warp totals
      if (lane == 31) (q == 0 ? pair_off : (q == 1 ? atom_off : act_off))[B] = s;
    }
  }
  __syncthreads();
  int run = wsum[0][wid] + inc[0] - own[0], runa = wsum[1][wid] + inc[1] - own[1], runc = wsum[2][wid] + inc[2] - own[2];
  for (int b = lo; b < hi; ++b) {
    const int n = n_atoms[b], nact = n > 1 ? n : 1;
    pair_off[b] = run;
    atom_off[b] = runa;
    act_off[b] = runc;
    for (int i = 0; i < n; ++i) atom_list[runa + i] = b * N + i;       // flat list of valid atoms (slot index b*N + i)
    for (int i = 0; i < nact; ++i) act_list[runc + i] = b * N + i;     // flat list of active rows
    run += n * n;
    runa += n;
    runc += nact;
  }
}

struct PairId { int b, i, j; };
__device__ __forceinline__ PairId decode_pair(int p, int B, const int* __restrict__ pair_off, const int* __restrict__ n_atoms) {
  int lo = 0, hi = B;   // largest b with pair_off[b] <= p
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pair_off[mid] <= p) lo = mid; else hi = mid;
  }
  const int r = p - pair_off[lo], n = n_atoms[lo];
  PairId id;
  id.b = lo; id.i = r / n; id.j = r - id.i * n;
  return id;
}

// (the edge level lives in edge.cuh)

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
      const int cmax = Cout - 1 - c0;   // rows beyond Cout are clamped (loaded twice, results discarded)
      // all CO weights of one k are loaded into distinct registers before any use, and the next k is prefetched while
      // the current one is consumed: the loop is otherwise bound by L2 latency (the weights stream from L2 / L1)
      float2 wn[CO];
      int k = lane;
      if (k < K) {
        MGB_UNROLL
        for (int c = 0; c < CO; ++c) wn[c] = Wl[(long long)(c < cmax ? c : cmax) * K + k];
      }
      for (; k < K; k += 32) {
        float2 w[CO];
        MGB_UNROLL
        for (int c = 0; c < CO; ++c) w[c] = wn[c];
        const int kn = (k + 32 < K) ? k + 32 : k;
        MGB_UNROLL
        for (int c = 0; c < CO; ++c) wn[c] = Wl[(long long)(c < cmax ? c : cmax) * K + kn];
        float2 x[NM];
        MGB_UNROLL
        for (int q = 0; q < NM; ++q) x[q] = cat0[(q < un.nm ? q : 0) * K + k];
        MGB_UNROLL
        for (int c = 0; c < CO; ++c)
          MGB_UNROLL
          for (int q = 0; q < NM; ++q) cfma(acc[q][c], w[c], x[q]);
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

// CG gather out of shared memory: cat[out_dst(o) + c] = sum_terms coef * T[a + c]   or  coef * A[a + c] * A[b + c]
// (tables resolved on the host against this use site, see resolve_cg_table).
template <bool SQUARE>
__device__ __forceinline__ void cg_gather(const CgTable& t, int C, const float2* __restrict__ sT, float2* __restrict__ sCat) {
  // thread = (slot, channel): a slot walks a contiguous run of whole outputs of the output-major table
  const int lanes = blockDim.x / C;
  const int s0 = threadIdx.x / C, c = threadIdx.x - s0 * C;
  if (s0 >= lanes) return;
  for (int s = s0; s < t.n_slots; s += lanes) {
    const int q0 = t.slot_start[s], q1 = t.slot_start[s + 1];
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int q = q0; q < q1; ++q) {
      const int4 e = t.flat[q];
      const float cf = __int_as_float(e.w);
      float2 v;
      if (SQUARE) v = cmul(sT[e.x + c], sT[e.y + c]);
      else v = sT[e.x + c];
      acc.x = fmaf(cf, v.x, acc.x);
      acc.y = fmaf(cf, v.y, acc.y);
      if (e.z & 1) {
        sCat[(e.z >> 1) + c] = acc;
        acc = make_float2(0.f, 0.f);
      }
    }
  }
}

// CG gather out of shared memory into the atom's cat vector (HBM), thread = (slot, channel); the slot's terms are fetched
// kGatherBatch at a time (independent, L1-resident loads) before they are consumed.
constexpr int kGatherBatch = 8;   // table entries fetched together
template <bool SQUARE>
__device__ __forceinline__ void gather25(const int2* __restrict__ tab, const int* __restrict__ slots, int C, const float2* __restrict__ src,
                                         float2* __restrict__ cat) {
  const int s0 = threadIdx.x / C, c = threadIdx.x - s0 * C;
  if (s0 >= kGatherSlots) return;
  const int q0 = slots[s0], q1 = slots[s0 + 1];
  float2 acc = make_float2(0.f, 0.f);
  for (int qb = q0; qb < q1; qb += kGatherBatch) {
    int2 e[kGatherBatch];
    MGB_UNROLL
    for (int k = 0; k < kGatherBatch; ++k) e[k] = __ldg(tab + (qb + k < q1 ? qb + k : q1 - 1));
    MGB_UNROLL
    for (int k = 0; k < kGatherBatch; ++k) {
      if (qb + k < q1) {
        const float cf = __int_as_float(e[k].y);
        float2 v;
        int dst, last;
        if (SQUARE) {
          v = cmul(src[(e[k].x & 0xff) + c], src[((e[k].x >> 8) & 0xff) + c]);
          last = (e[k].x >> 16) & 1; dst = (e[k].x >> 17) & 0x1fff;
        } else {
          v = src[(e[k].x & 0x1fff) + c];
          last = (e[k].x >> 13) & 1; dst = (e[k].x >> 14) & 0x1fff;
        }
        acc.x = fmaf(cf, v.x, acc.x);
        acc.y = fmaf(cf, v.y, acc.y);
        if (last) {
          cat[dst + c] = acc;
          acc = make_float2(0.f, 0.f);
        }
      }
    }
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
constexpr int kAtomPhaseA = 1;   // forward: CG aggregate over the neighbours (needs E_k)   backward: row pass -> dE
constexpr int kAtomPhaseB = 2;   // forward: CG square + pass-through (needs A_k only)       backward: column pass + own atom -> dA

__host__ __device__ inline int atom_smem_floats(const LevelDesc& L, int N) {
  const int nlm2 = L.nlm_in;
  const int stage = 2 * kJChunk * (kNL * L.C + nlm2 * L.C) * 2;   // two staging buffers
  const int tsz = kM * nlm2 * L.C * 2;
  return (stage > tsz ? stage : tsz) + nlm2 * L.C * 2 + N * kM * 2;
}

// Y_lm(r_i - r_j) for every neighbour j of atom i, once per CTA (conj, 'unit' norm, un-normalised argument:
// SphericalHarmonicsRel(conj=True), covariant/modules.py:52-56).  sYall: [n][25]
__device__ __forceinline__ void neighbour_harmonics(const float* __restrict__ pos_b, int i, int n, float2* sYall) {
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float2 y[kM];
    sph_harm_l4(pos_b[i * 3 + 0] - pos_b[j * 3 + 0], pos_b[i * 3 + 1] - pos_b[j * 3 + 1], pos_b[i * 3 + 2] - pos_b[j * 3 + 2], true,
                true, y);
    MGB_UNROLL
    for (int q = 0; q < kM; ++q) sYall[j * kM + q] = y[q];
  }
}

// Stage one chunk of neighbours j0..j0+nj-1 of atom i into shared memory: E_ij and (NLM2 > 0) A_j.
template <int NLM2>
__device__ __forceinline__ void stage_neighbours(const LevelDesc& L, const float2* __restrict__ Ab, const float2* __restrict__ E_i,
                                                 int j0, int nj, float2* sE, float2* sAj) {
  const int C = L.C;
  for (int idx = threadIdx.x; idx < nj * kNL * C; idx += blockDim.x) sE[idx] = E_i[(long long)j0 * kNL * C + idx];
  for (int idx = threadIdx.x; idx < nj * NLM2 * C; idx += blockDim.x) sAj[idx] = Ab[(long long)j0 * NLM2 * C + idx];
}

// asynchronous copy of one neighbour chunk into a staging buffer (E_ij rows [nj][5][C], A_j rows [nj][NLM2][C]; contiguous in HBM)
template <int NLM2>
__device__ __forceinline__ void chunk_copy_async(const float2* __restrict__ Ab, const float2* __restrict__ E_i, int C, int j0, int nj,
                                                 float2* sE, float2* sAj) {
  for (int idx = threadIdx.x; idx < nj * kNL * C; idx += blockDim.x) cp_async8(sE + idx, E_i + (long long)j0 * kNL * C + idx);
  for (int idx = threadIdx.x; idx < nj * NLM2 * C; idx += blockDim.x) cp_async8(sAj + idx, Ab + (long long)j0 * NLM2 * C + idx);
  cp_async_commit();
}

// cg_gather writes straight to the atom's cat vector in HBM (coalesced over the channel index).
template <int NLM2, int CT>   // CT: compile-time channel count (0: read it from the level descriptor)
__global__ void __launch_bounds__(kAtomThreads, 3)
k_atom_cat(const CovDesc* __restrict__ dp, int level, const float* __restrict__ pos, const int* __restrict__ n_atoms,
           const int* __restrict__ atom_off, const int* __restrict__ atom_list, int B, const float* __restrict__ A_in,
           const float* __restrict__ E, float* __restrict__ cat_out, int phases) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int N = d.N, C = CT ? CT : L.C;
  if ((int)blockIdx.x >= atom_off[B]) return;
  const int slot = atom_list[blockIdx.x];
  const int b = slot / N, i = slot - b * N;
  const int n = n_atoms[b];
  MGB_DYN_SMEM(float2, smem);
  const int stage = kJChunk * (kNL * C + NLM2 * C), tsz = kM * NLM2 * C;
  float2* sT = smem;
  float2* sE = smem;
  float2* sAj = sE + kJChunk * kNL * C;
  float2* sAi = smem + (2 * stage > tsz ? 2 * stage : tsz);
  float2* sYall = sAi + NLM2 * C;
  const float2* Ab = reinterpret_cast<const float2*>(A_in) + (long long)b * N * NLM2 * C;
  const float2* E_i = reinterpret_cast<const float2*>(E) + ((long long)b * N + i) * N * kNL * C;
  const float* pos_b = pos + (long long)b * N * 3;
  float2* co = reinterpret_cast<float2*>(cat_out) + (long long)slot * L.totA;
  for (int idx = threadIdx.x; idx < NLM2 * C; idx += blockDim.x) sAi[idx] = Ab[(long long)i * NLM2 * C + idx];
  if (!(phases & kAtomPhaseA)) {   // only the blocks that depend on A_i alone
    __syncthreads();
    if (NLM2 == kM) gather25<true>(L.gt.sq_flat8, L.gt.sq_slot, C, sAi, co); else cg_gather<true>(L.sq, C, sAi, co);
    for (int idx = threadIdx.x; idx < NLM2 * C; idx += blockDim.x) {
      const int lm = idx / C, cc = idx % C, l = ell_of_lm(lm);
      co[L.offA[l] + (lm - l * l) * L.catA[l] + L.in_block[l] * C + cc] = sAi[idx];
    }
    return;
  }
  neighbour_harmonics(pos_b, i, n, sYall);

  const bool owner = (int)threadIdx.x < kM * C;
  const int lm1 = owner ? threadIdx.x / C : 0, c = owner ? threadIdx.x % C : 0;
  const int l1 = ell_of_lm(lm1);
  float2 acc[NLM2];
  MGB_UNROLL
  for (int q = 0; q < NLM2; ++q) acc[q] = make_float2(0.f, 0.f);
  // neighbour chunks double-buffered with asynchronous copies: the next chunk is in flight from L2 while this one is consumed
  // (both buffers live in the space that T takes over after the loop)
  chunk_copy_async<NLM2>(Ab, E_i, C, 0, min(kJChunk, n), sE, sAj);
  int buf = 0;
  for (int j0 = 0; j0 < n; j0 += kJChunk, buf ^= 1) {
    const int nj = min(kJChunk, n - j0);
    cp_async_wait_all();
    __syncthreads();   // this chunk is visible to everyone, and everyone is done with the other buffer
    if (j0 + kJChunk < n)
      chunk_copy_async<NLM2>(Ab, E_i, C, j0 + kJChunk, min(kJChunk, n - j0 - kJChunk), sE + (buf ^ 1) * stage, sAj + (buf ^ 1) * stage);
    const float2* cE = sE + buf * stage;
    const float2* cA = sAj + buf * stage;
    if (owner) {
      for (int jj = 0; jj < nj; ++jj) {
        const float2 u = cmul(cE[(jj * kNL + l1) * C + c], sYall[(j0 + jj) * kM + lm1]);
        const float2* a = cA + jj * NLM2 * C + c;
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
  if (NLM2 == kM) gather25<false>(L.gt.ag_flat8, L.gt.ag_slot, C, sT, co); else cg_gather<false>(L.ag, C, sT, co);
  if (!(phases & kAtomPhaseB)) return;
  if (NLM2 == kM) gather25<true>(L.gt.sq_flat8, L.gt.sq_slot, C, sAi, co); else cg_gather<true>(L.sq, C, sAi, co);
  for (int idx = threadIdx.x; idx < NLM2 * C; idx += blockDim.x) {
    const int lm = idx / C, cc = idx % C, l = ell_of_lm(lm);
    co[L.offA[l] + (lm - l * l) * L.catA[l] + L.in_block[l] * C + cc] = sAi[idx];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Channel mix as a row-parallel kernel: one thread per (valid atom, l, m) row of the cat vector,
//   out[row][c'] = sum_k W_l[c'][k] cat[row][k]      (forward)
//   dcat[row][k] = sum_c' conj(W_l[c'][k]) dOut[row][c']   (backward)
// W_l transposed to [k][c'] sits in shared memory and is read as warp-uniform (broadcast) vector loads; the accumulators
// are registers; no cross-lane reduction.  grid = (row blocks, 5 ells).
// ------------------------------------------------------------------------------------------------------------
constexpr int kMixThreads = 128;
constexpr int kMixThreadsLarge = 512;   // backward (dcat) mix of large minibatches: 64 rows per CTA share one staging of W_l (C5 b256: 215 / 253 -> 180 / 178 us)
template <int CO> constexpr int kMixStride = (CO % 4 == 2) ? CO : ((CO % 2 == 0) ? CO + 2 : CO);
inline int mix_stride_of(int co) { return (co % 4 == 2) ? co : ((co % 2 == 0) ? co + 2 : co); }

template <int CO, bool BACKWARD, int KS, int THREADS = kMixThreads>
__global__ void __launch_bounds__(THREADS)
k_mix_rows(const CovDesc* __restrict__ dp, int level, const float* __restrict__ Wt, const int* __restrict__ atom_off,
           const int* __restrict__ atom_list, int B, const float* __restrict__ cat, const float* __restrict__ A_out,
           float* __restrict__ out) {
  // KS adjacent lanes share a row and split its k range (k = ks, ks + KS, ...): more parallelism for small minibatches and
  // coalesced reads of the row; the forward reduces the KS partial sums with shuffles, the backward needs no reduction.
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const int l = blockIdx.y, K = L.catA[l], nm = 2 * l + 1, Cout = L.Cout;
  const int rows = atom_off[B] * nm;
  constexpr int kRowsPerCta = THREADS / KS;
  if ((int)(blockIdx.x * kRowsPerCta) >= rows) return;   // uniform per CTA: the whole CTA leaves before any barrier
  // row stride of the staged weights: KS lanes of a quarter-warp read KS different k at once, so the stride (in float2) is
  // kept = 2 mod 4 — 16-byte aligned rows whose 16-byte slots fall into different bank groups (CO = 16 would otherwise put
  // all eight k of a request on the same banks)
  constexpr int CS = kMixStride<CO>;
  MGB_DYN_SMEM(float2, sW);   // [K][CS]
  {
    // k_prep_params left W_l transposed and padded to the stride CS in the scratch: one bulk copy (TMA) brings the whole
    // matrix in while the threads set up their rows
    __shared__ SmemBarrier s_bar;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    const float* src = Wt + d.wt_atom[level] + 2ll * L.offWAt[l];
    bool bulk = false;
    if (L.mixCS == CS) {
      bulk = smem_fill_begin(reinterpret_cast<float*>(sW), src, 2 * K * CS, &s_bar);
    } else {   // stride mismatch (should not happen: the plan pads for the instantiation it launches)
      const float2* src2 = reinterpret_cast<const float2*>(src);
      for (int idx = threadIdx.x; idx < K * CO; idx += blockDim.x) {
        const int k = idx / CO, c = idx - k * CO;
        sW[k * CS + c] = c < L.mixCS ? src2[k * L.mixCS + c] : make_float2(0.f, 0.f);
      }
    }
    smem_fill_end(bulk, &s_bar, 0);
  }
  __syncthreads();
  const int ks = threadIdx.x % KS;
  for (int rg = blockIdx.x; rg * kRowsPerCta < rows; rg += gridDim.x) {   // a CTA keeps its weights for several row groups
  const int row_raw = rg * kRowsPerCta + threadIdx.x / KS;
  const bool valid = row_raw < rows;
  const int row = valid ? row_raw : rows - 1;   // clamp: every lane takes part in the shuffles
  const int a = row / nm, m = row - a * nm;
  const long long slot = atom_list[a];
  const float2* crow = reinterpret_cast<const float2*>(cat) + slot * L.totA + L.offA[l] + m * K;
  const long long orow = (slot * kM + l * l + m) * Cout;
  if (!BACKWARD) {
    // pair accumulators (FFMA2, see common.cuh): P[c] += W[k][c] * (x.re, x.re), Q[c] += W[k][c] * (x.im, x.im)
    f32x2 P[CO], Q[CO];
    MGB_UNROLL
    for (int c = 0; c < CO; ++c) { P[c] = pack2(0.f, 0.f); Q[c] = pack2(0.f, 0.f); }
    int k = ks;
    // eight independent row loads in flight per thread (the rows come from L2 / HBM; the weights are broadcasts from shared memory)
    for (; k + 7 * KS < K; k += 8 * KS) {
      float2 x[8];
      MGB_UNROLL
      for (int q = 0; q < 8; ++q) x[q] = crow[k + q * KS];
      MGB_UNROLL
      for (int q = 0; q < 8; ++q) {
        const f32x2 xr = pack2(x[q].x, x[q].x), xi = pack2(x[q].y, x[q].y);
        MGB_UNROLL
        for (int c = 0; c < CO; ++c) {
          const f32x2 w = as_pair(sW[(k + q * KS) * CS + c]);
          fma2(P[c], w, xr);
          fma2(Q[c], w, xi);
        }
      }
    }
    for (; k < K; k += KS) {
      const float2 x = crow[k];
      const f32x2 xr = pack2(x.x, x.x), xi = pack2(x.y, x.y);
      MGB_UNROLL
      for (int c = 0; c < CO; ++c) {
        const f32x2 w = as_pair(sW[k * CS + c]);
        fma2(P[c], w, xr);
        fma2(Q[c], w, xi);
      }
    }
    float2 acc[CO];
    MGB_UNROLL
    for (int c = 0; c < CO; ++c) {
      acc[c] = cpair_mul(P[c], Q[c]);
      MGB_UNROLL
      for (int o = KS / 2; o > 0; o >>= 1) {
        acc[c].x += __shfl_xor_sync(0xffffffffu, acc[c].x, o);
        acc[c].y += __shfl_xor_sync(0xffffffffu, acc[c].y, o);
      }
    }
    if (valid && ks == 0) {
      float2* o = reinterpret_cast<float2*>(out) + orow;
      MGB_UNROLL
      for (int c = 0; c < CO; ++c)
        if (c < Cout) o[c] = acc[c];
    }
  } else {
    float2 g[CO];
    const float2* gi = reinterpret_cast<const float2*>(A_out) + orow;
    MGB_UNROLL
    for (int c = 0; c < CO; ++c) g[c] = c < Cout ? gi[c] : make_float2(0.f, 0.f);
    float2* o = reinterpret_cast<float2*>(out) + slot * L.totA + L.offA[l] + m * K;
    if (valid) {
      // conj(W[k][c]) * g[c] with pair accumulators: the broadcast pairs of g are formed once per row
      f32x2 gr[CO], gq[CO];
      MGB_UNROLL
      for (int c = 0; c < CO; ++c) { gr[c] = pack2(g[c].x, g[c].x); gq[c] = pack2(g[c].y, g[c].y); }
#pragma unroll 2
      for (int k = ks; k < K; k += KS) {
        f32x2 P = pack2(0.f, 0.f), Q = pack2(0.f, 0.f);
        MGB_UNROLL
        for (int c = 0; c < CO; ++c) {
          const f32x2 w = as_pair(sW[k * CS + c]);
          fma2(P, w, gr[c]);
          fma2(Q, w, gq[c]);
        }
        o[k] = cpair_mulc(P, Q);
      }
    }
  }
  }
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
