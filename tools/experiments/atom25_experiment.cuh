// atom25.cuh — the Cormorant atom level (cormorant CormorantAtomLevel; reached from molgym/agents/covariant/modules.py:110)
// for levels whose input representation carries all ells (levels >= 1: 25 components per channel), forward and backward.
//
// These kernels hold most of the path's FLOPs: per ordered pair (i, j) and channel a 25 x 25 complex outer product
//     T_i[x][y][c] += U_ij[x][c] * A_j[y][c],     U_ij[x][c] = E_ij[l(x)][c] * Y_x(r_i - r_j)
// forward, and the two transposed contractions backward (row pass -> dE_ij, column pass -> dA_j).  One CTA per atom.
//
// Register tiling.  Measured on B200 (tools/ffma_peak.cu): a loop with one 8-byte shared-memory load per complex MAC (4 FFMA)
// saturates the shared-memory pipe at 69 % of the FFMA roof, with FFMA or packed FFMA2 alike; the fix is fewer loads per MAC,
// not fewer instructions.  The 25 components are cut into four groups {0-6, 7-12, 13-18, 19-24}; thread (xg, yg, c) owns the
// 7 x 7 tile (x in group xg, y in group yg) of channel c: 49 complex accumulators (forward) or 49 dT operands (backward) in
// registers.  Per neighbour it loads 7 U values + 7 A_j values for 49 MACs forward — and backward the SAME dT tile serves the row
// pass (w[x] += conj(a[y]) dT[x][y]) and the column pass (v[y] += conj(u[x]) dT[x][y]): 98 MACs per 14 loads, one neighbour loop.
// Lane layout: yg = lane bits 0-1, xg = bits 2-3, channel parity = bit 4 (a warp = all 16 tiles of two channels), so the sums
// over the four yg (row pass) and the four xg (column pass) are two xor-shuffles each and every load is bank-conflict free.
//
// Backward: the atom's slice of the cat cotangent (written by k_mix_rows<.., true>) arrives in shared memory as one bulk copy
// (TMA); every thread then builds its dT tile straight into registers from a tile-major Clebsch-Gordan table (49 entries x 5
// zero-padded terms, all loads independent, the 16 tiles of one table row in one 128-byte line that stays in L1).  The same
// table, shifted to the CG-square blocks, gives the own-atom terms.  No table walk depends on a previous load: the first
// version of these kernels (slot-wise walks of un-padded term lists) spent more time waiting for table entries than computing.
#pragma once
#include "cov_backward.cuh"

namespace mgb {

constexpr int kA25JC = 8;            // neighbours per staged chunk (double-buffered)
constexpr int kA25T = 7;             // tile edge (groups of 7, 6, 6, 6 components)
__host__ __device__ inline int atom25_threads(int C) { return 32 * ((C + 1) / 2); }   // one warp per pair of channels

// tile coordinates of a thread
struct Tile25 {
  int xg, yg, c;
  bool owner;      // false: lane of an odd last channel pair (computes a clamped copy, stores nothing)
  int x0, nx, y0, ny;
};
__device__ __forceinline__ int grp_start(int g) { return g == 0 ? 0 : 1 + 6 * g; }   // 0, 7, 13, 19
__device__ __forceinline__ Tile25 tile25(int tid, int C) {
  Tile25 t;
  const int lane = tid & 31, warp = tid >> 5;
  t.yg = lane & 3;
  t.xg = (lane >> 2) & 3;
  const int c = 2 * warp + (lane >> 4);
  t.owner = c < C;
  t.c = t.owner ? c : C - 1;
  t.x0 = grp_start(t.xg); t.nx = t.xg == 0 ? 7 : 6;
  t.y0 = grp_start(t.yg); t.ny = t.yg == 0 ? 7 : 6;
  return t;
}

// asynchronous copy of `count` float2 from global to shared memory (16-byte pieces when both sides are aligned)
__device__ __forceinline__ void copy_async_f2(float2* dst, const float2* __restrict__ src, int count) {
#ifndef MGB_CUSIM
  if ((((unsigned long long)src | (unsigned long long)smem_addr(dst)) & 15ull) == 0ull) {
    const int n2 = count >> 1;
    for (int idx = threadIdx.x; idx < n2; idx += blockDim.x) {
      const unsigned s = smem_addr(dst + 2 * idx);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(src + 2 * idx) : "memory");
    }
    if ((count & 1) && threadIdx.x == 0) cp_async8(dst + count - 1, src + count - 1);
    return;
  }
#endif
  for (int idx = threadIdx.x; idx < count; idx += blockDim.x) cp_async8(dst + idx, src + idx);
}

// Sum of eight values over a quartet of lanes (lane ids differing in the two bits `bit`, 2 * bit): reduce-scatter, lane k of the
// quartet (k = its two bits, high bit first) ends with the complete sums of components 2k and 2k + 1.
__device__ __forceinline__ void reduce_scatter4(const float2 (&x)[8], int bit, int k, float2 (&out)[2]) {
  const bool hi = (k & 2) != 0, odd = (k & 1) != 0;
  float2 keep[4];
  MGB_UNROLL
  for (int q = 0; q < 4; ++q) {
    const float2 send = hi ? x[q] : x[q + 4], mine = hi ? x[q + 4] : x[q];
    keep[q] = make_float2(mine.x + __shfl_xor_sync(0xffffffffu, send.x, 2 * bit), mine.y + __shfl_xor_sync(0xffffffffu, send.y, 2 * bit));
  }
  MGB_UNROLL
  for (int q = 0; q < 2; ++q) {
    const float2 send = odd ? keep[q] : keep[q + 2], mine = odd ? keep[q + 2] : keep[q];
    out[q] = make_float2(mine.x + __shfl_xor_sync(0xffffffffu, send.x, bit), mine.y + __shfl_xor_sync(0xffffffffu, send.y, bit));
  }
}

// one staged chunk of neighbours: E_ij rows [JC][5][C], A_j rows [JC][25][C] (both contiguous in HBM), then U = E * Y [JC][25][C]
__host__ __device__ inline int atom25_chunk_f2(int C) { return kA25JC * (kNL * C + kM * C); }

__host__ __device__ inline int atom25_fwd_smem_floats(const LevelDesc& L, int N) {
  const int C = L.C;
  const int stage = (2 * atom25_chunk_f2(C) + kA25JC * kM * C) * 2;   // two staging buffers + U
  const int tsz = kM * kM * C * 2;
  return (stage > tsz ? stage : tsz) + kM * C * 2 + N * kM * 2;
}

// U[jj][x][c] = E[jj][l(x)][c] * Y_x(r_i - r_jj) for the staged chunk (whole CTA)
__device__ __forceinline__ void build_u(const float2* cE, const float2* sY, int C, int nj, float2* sU) {
  for (int idx = threadIdx.x; idx < nj * kM * C; idx += blockDim.x) {
    const int jj = idx / (kM * C), r = idx - jj * (kM * C), x = r / C, cc = r - x * C;
    sU[idx] = cmul(cE[(jj * kNL + ell_of_lm(x)) * C + cc], sY[jj * kM + x]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Forward: cat_i = [CG(T_i) | A_i | CG(A_i x A_i)] written to HBM (k_mix_rows applies the weights; k_mix_dw reads it back).
// phases: kAtomPhaseA = the neighbour aggregate, kAtomPhaseB = the blocks that depend on A_i alone (cov_forward.cuh).
// ------------------------------------------------------------------------------------------------------------
template <int CT>
__global__ void __launch_bounds__(160, 3)
k_atom25_fwd(const CovDesc* __restrict__ dp, int level, const float* __restrict__ pos, const int* __restrict__ n_atoms,
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
  const int chunk = atom25_chunk_f2(C), stage = 2 * chunk + kA25JC * kM * C, tsz = kM * kM * C;
  float2* sT = smem;
  float2* sU = smem + 2 * chunk;
  float2* sAi = smem + (stage > tsz ? stage : tsz);
  float2* sYall = sAi + kM * C;
  const float2* Ab = reinterpret_cast<const float2*>(A_in) + (long long)b * N * kM * C;
  const float2* E_i = reinterpret_cast<const float2*>(E) + ((long long)b * N + i) * N * kNL * C;
  const float* pos_b = pos + (long long)b * N * 3;
  float2* co = reinterpret_cast<float2*>(cat_out) + (long long)slot * L.totA;
  for (int idx = threadIdx.x; idx < kM * C; idx += blockDim.x) sAi[idx] = Ab[(long long)i * kM * C + idx];
  if (phases & kAtomPhaseA) {
    neighbour_harmonics(pos_b, i, n, sYall);
    const Tile25 t = tile25(threadIdx.x, C);
    float2 acc[kA25T][kA25T];
    MGB_UNROLL
    for (int xx = 0; xx < kA25T; ++xx)
      MGB_UNROLL
      for (int yy = 0; yy < kA25T; ++yy) acc[xx][yy] = make_float2(0.f, 0.f);
    auto issue = [&](int j0, int buf) {
      const int nj = min(kA25JC, n - j0);
      float2* sE = smem + buf * chunk;
      copy_async_f2(sE, E_i + (long long)j0 * kNL * C, nj * kNL * C);
      copy_async_f2(sE + kA25JC * kNL * C, Ab + (long long)j0 * kM * C, nj * kM * C);
      cp_async_commit();
    };
    if (n > 0) issue(0, 0);
    int buf = 0;
    for (int j0 = 0; j0 < n; j0 += kA25JC, buf ^= 1) {
      const int nj = min(kA25JC, n - j0);
      cp_async_wait_all();
      __syncthreads();   // this chunk is visible to everyone; the other buffer and U are free
      if (j0 + kA25JC < n) issue(j0 + kA25JC, buf ^ 1);
      const float2* cE = smem + buf * chunk;
      const float2* cA = cE + kA25JC * kNL * C;
      build_u(cE, sYall + j0 * kM, C, nj, sU);
      __syncthreads();
      for (int jj = 0; jj < nj; ++jj) {
        const float2* a = cA + (jj * kM + t.y0) * C + t.c;
        const float2* u = sU + (jj * kM + t.x0) * C + t.c;
        float2 av[kA25T];
        MGB_UNROLL
        for (int yy = 0; yy < kA25T; ++yy) av[yy] = yy < t.ny ? a[yy * C] : make_float2(0.f, 0.f);
        MGB_UNROLL
        for (int xx = 0; xx < kA25T; ++xx) {
          const float2 uv = xx < t.nx ? u[xx * C] : make_float2(0.f, 0.f);
          MGB_UNROLL
          for (int yy = 0; yy < kA25T; ++yy) cfma(acc[xx][yy], uv, av[yy]);
        }
      }
    }
    __syncthreads();   // the staging buffers become T
    if (t.owner) {
      MGB_UNROLL
      for (int xx = 0; xx < kA25T; ++xx)
        MGB_UNROLL
        for (int yy = 0; yy < kA25T; ++yy)
          if (xx < t.nx && yy < t.ny) sT[((t.x0 + xx) * kM + t.y0 + yy) * C + t.c] = acc[xx][yy];
    }
    __syncthreads();
    gather25<false>(L.t25.ag_flat8, L.t25.ag_slot, C, sT, co);
  } else {
    __syncthreads();
  }
  if (phases & kAtomPhaseB) {
    gather25<true>(L.t25.sq_flat8, L.t25.sq_slot, C, sAi, co);
    for (int idx = threadIdx.x; idx < kM * C; idx += blockDim.x) {
      const int lm = idx / C, cc = idx % C, l = ell_of_lm(lm);
      co[L.offA[l] + (lm - l * l) * L.catA[l] + L.in_block[l] * C + cc] = sAi[idx];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Backward.  Shared memory: the atom's cat cotangent (re-used as the staging area of the neighbour loop) | A_i | own-atom
// accumulator | harmonics of the neighbours.
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int atom25_bwd_stage_f2(int C) { return 2 * atom25_chunk_f2(C) + 2 * kA25JC * kM * C; }   // chunks, U, dE parts
__host__ __device__ inline int atom25_bwd_smem_floats(const LevelDesc& L, int N) {
  const int C = L.C;
  const int stage = atom25_bwd_stage_f2(C);
  return ((L.totA > stage ? L.totA : stage) + 2 * kM * C + N * kM) * 2;
}

// dT tile of thread t out of the cat cotangent in shared memory: tile[xx][yy] = sum_k coef * dcat[dst + shift(l) + c]
template <bool SHIFT>
__device__ __forceinline__ void scatter_tile(const Atom25Tables& T, const float2* __restrict__ sDcat, const Tile25& t, int lane16,
                                             float2 (&tile)[kA25T][kA25T]) {
  const int2* tab = T.tile_tab + lane16;
  MGB_UNROLL
  for (int xx = 0; xx < kA25T; ++xx)
    MGB_UNROLL
    for (int yy = 0; yy < kA25T; ++yy) {
      float2 acc = make_float2(0.f, 0.f);
      MGB_UNROLL
      for (int k = 0; k < kCgPad; ++k) {
        const int2 e = __ldg(tab + ((xx * kA25T + yy) * kCgPad + k) * 16);
        const int off = (e.x & 0x1fff) + (SHIFT ? T.sq_shift[(e.x >> 13) & 7] : 0);
        const float cf = __int_as_float(e.y);
        const float2 v = sDcat[off + t.c];
        acc.x = fmaf(cf, v.x, acc.x);
        acc.y = fmaf(cf, v.y, acc.y);
      }
      tile[xx][yy] = acc;
    }
}

template <int CT>
__global__ void __launch_bounds__(160, 2)
k_atom25_bwd(const CovDesc* __restrict__ dp, int level, const float* __restrict__ pos, const int* __restrict__ n_atoms,
             const int* __restrict__ atom_off, const int* __restrict__ atom_list, int B, const float* __restrict__ A_in,
             const float* __restrict__ E, const float* __restrict__ dcat, float* __restrict__ dA_in, float* __restrict__ dE,
             int accumulate_dE) {
  const CovDesc& d = *dp;
  const LevelDesc& L = d.lv[level];
  const Atom25Tables& T = L.t25;
  const int N = d.N, C = CT ? CT : L.C;
  if ((int)blockIdx.x >= atom_off[B]) return;
  const int slot = atom_list[blockIdx.x];
  const int b = slot / N, i = slot - b * N;
  const int n = n_atoms[b];
  MGB_DYN_SMEM(float2, smem);
  const int stage_f2 = atom25_bwd_stage_f2(C), chunk = atom25_chunk_f2(C);
  float2* sDcat = smem;                                 // [totA]  | staging of the neighbour loop
  float2* sAi = smem + (L.totA > stage_f2 ? L.totA : stage_f2);   // [25][C]
  float2* sAcc = sAi + kM * C;                          // [25][C]  own-atom cotangent
  float2* sYall = sAcc + kM * C;                        // [n][25]
  const float2* Ab = reinterpret_cast<const float2*>(A_in) + (long long)b * N * kM * C;
  const float2* E_i = reinterpret_cast<const float2*>(E) + ((long long)b * N + i) * N * kNL * C;
  float2* dE_i = reinterpret_cast<float2*>(dE) + ((long long)b * N + i) * N * kNL * C;
  float2* dAb = reinterpret_cast<float2*>(dA_in) + (long long)b * N * kM * C;
  const float* pos_b = pos + (long long)b * N * 3;

  __shared__ SmemBarrier s_bar;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  const bool bulk = smem_fill_begin(reinterpret_cast<float*>(sDcat), dcat + 2ll * slot * L.totA, 2 * L.totA, &s_bar);
  for (int idx = threadIdx.x; idx < kM * C; idx += blockDim.x) sAi[idx] = Ab[(long long)i * kM * C + idx];
  neighbour_harmonics(pos_b, i, n, sYall);
  smem_fill_end(bulk, &s_bar, 0);
  __syncthreads();

  const Tile25 t = tile25(threadIdx.x, C);
  const int lane16 = threadIdx.x & 15;
  // ---- own atom: pass-through block ...
  for (int idx = threadIdx.x; idx < kM * C; idx += blockDim.x) {
    const int lm = idx / C, cc = idx - lm * C, l = ell_of_lm(lm);
    sAcc[idx] = sDcat[L.offA[l] + (lm - l * l) * L.catA[l] + L.in_block[l] * C + cc];
  }
  __syncthreads();
  float2 dt[kA25T][kA25T];
  // ... and CG square: with dS[x][y] the scatter of the square blocks, dA_i[x] += conj(A_i[y]) dS[x][y], dA_i[y] += conj(A_i[x]) dS[x][y]
  {
    scatter_tile<true>(T, sDcat, t, lane16, dt);
    float2 w[kA25T], v[kA25T], ax[kA25T], ay[kA25T];
    MGB_UNROLL
    for (int q = 0; q < kA25T; ++q) {
      ax[q] = q < t.nx ? sAi[(t.x0 + q) * C + t.c] : make_float2(0.f, 0.f);
      ay[q] = q < t.ny ? sAi[(t.y0 + q) * C + t.c] : make_float2(0.f, 0.f);
      w[q] = make_float2(0.f, 0.f);
      v[q] = make_float2(0.f, 0.f);
    }
    MGB_UNROLL
    for (int xx = 0; xx < kA25T; ++xx)
      MGB_UNROLL
      for (int yy = 0; yy < kA25T; ++yy) {
        cfmacl(w[xx], ay[yy], dt[xx][yy]);
        cfmacl(v[yy], ax[xx], dt[xx][yy]);
      }
    MGB_UNROLL
    for (int q = 0; q < kA25T; ++q) {
      w[q].x += __shfl_xor_sync(0xffffffffu, w[q].x, 1); w[q].y += __shfl_xor_sync(0xffffffffu, w[q].y, 1);
      w[q].x += __shfl_xor_sync(0xffffffffu, w[q].x, 2); w[q].y += __shfl_xor_sync(0xffffffffu, w[q].y, 2);
      v[q].x += __shfl_xor_sync(0xffffffffu, v[q].x, 4); v[q].y += __shfl_xor_sync(0xffffffffu, v[q].y, 4);
      v[q].x += __shfl_xor_sync(0xffffffffu, v[q].x, 8); v[q].y += __shfl_xor_sync(0xffffffffu, v[q].y, 8);
    }
    if (t.owner) {
      MGB_UNROLL
      for (int q = 0; q < kA25T; ++q) {
        if ((q & 3) == t.yg && q < t.nx) smem_add2(sAcc + (t.x0 + q) * C + t.c, w[q]);
        if ((q & 3) == t.xg && q < t.ny) smem_add2(sAcc + (t.y0 + q) * C + t.c, v[q]);
      }
    }
  }
  // ---- the register tile of dT for the neighbour loop
  scatter_tile<false>(T, sDcat, t, lane16, dt);
  __syncthreads();   // sAcc complete; the cat cotangent is dead: its space becomes the staging area
  for (int idx = threadIdx.x; idx < kM * C; idx += blockDim.x) atomic_add2(dAb + (long long)i * kM * C + idx, sAcc[idx]);

  // ---- neighbour loop: row pass (-> dE_ij) and column pass (-> dA_j) on the same register tile
  float2* sU = smem + 2 * chunk;                  // [JC][25][C]  U_ij = E_ij * Y
  float2* sDE = sU + kA25JC * kM * C;             // [JC][25][C]  per-component dE contributions
  auto issue = [&](int j0, int buf) {
    const int nj = min(kA25JC, n - j0);
    float2* sE = smem + buf * chunk;
    copy_async_f2(sE, E_i + (long long)j0 * kNL * C, nj * kNL * C);
    copy_async_f2(sE + kA25JC * kNL * C, Ab + (long long)j0 * kM * C, nj * kM * C);
    cp_async_commit();
  };
  if (n > 0) issue(0, 0);
  int buf = 0;
  for (int j0 = 0; j0 < n; j0 += kA25JC, buf ^= 1) {
    const int nj = min(kA25JC, n - j0);
    cp_async_wait_all();
    __syncthreads();   // chunk visible; the other buffer, U and the dE parts are free
    if (j0 + kA25JC < n) issue(j0 + kA25JC, buf ^ 1);
    const float2* cE = smem + buf * chunk;
    const float2* cA = cE + kA25JC * kNL * C;
    build_u(cE, sYall + j0 * kM, C, nj, sU);
    __syncthreads();
    // software pipeline: the operands of neighbour jj + 1 are fetched from shared memory while jj is being multiplied (two CTAs of
    // five warps per SM leave little else to hide the load latency behind)
    float2 an[kA25T], un[kA25T];
    {
      const float2* a = cA + t.y0 * C + t.c;
      const float2* u = sU + t.x0 * C + t.c;
      MGB_UNROLL
      for (int q = 0; q < kA25T; ++q) {
        an[q] = q < t.ny ? a[q * C] : make_float2(0.f, 0.f);
        un[q] = q < t.nx ? u[q * C] : make_float2(0.f, 0.f);
      }
    }
    for (int jj = 0; jj < nj; ++jj) {
      float2 av[kA25T], uu[kA25T], v[kA25T];
      MGB_UNROLL
      for (int q = 0; q < kA25T; ++q) {
        av[q] = an[q];
        uu[q] = un[q];
        v[q] = make_float2(0.f, 0.f);
      }
      if (jj + 1 < nj) {
        const float2* a = cA + ((jj + 1) * kM + t.y0) * C + t.c;
        const float2* u = sU + ((jj + 1) * kM + t.x0) * C + t.c;
        MGB_UNROLL
        for (int q = 0; q < kA25T; ++q) {
          an[q] = q < t.ny ? a[q * C] : make_float2(0.f, 0.f);
          un[q] = q < t.nx ? u[q * C] : make_float2(0.f, 0.f);
        }
      }
      float2 w[kA25T + 1];
      MGB_UNROLL
      for (int xx = 0; xx < kA25T; ++xx) {
        const float2 uv = uu[xx];
        float2 wa = make_float2(0.f, 0.f), wb = make_float2(0.f, 0.f);   // two chains: the row sum is the longest dependency
        MGB_UNROLL
        for (int yy = 0; yy < kA25T; ++yy) {
          if (yy & 1) cfmacl(wb, av[yy], dt[xx][yy]); else cfmacl(wa, av[yy], dt[xx][yy]);   // row pass: w[x] += conj(A_j[y]) dT[x][y]
          cfmacl(v[yy], uv, dt[xx][yy]);                                                       // column pass: v[y] += conj(U_ij[x]) dT[x][y]
        }
        w[xx] = make_float2(wa.x + wb.x, wa.y + wb.y);
      }
      w[kA25T] = make_float2(0.f, 0.f);
      // The four yg lanes hold partial w, the four xg lanes partial v.  Reduce-scatter: exchange halves, then quarters — after
      // two rounds lane k of the quartet holds the complete sums of components 2k, 2k+1 (6 shuffles per float instead of 14).
      float2 wk[2], vk[2];
      reduce_scatter4(w, 1, t.yg, wk);
      float2 v8[kA25T + 1];
      MGB_UNROLL
      for (int q = 0; q < kA25T; ++q) v8[q] = v[q];
      v8[kA25T] = make_float2(0.f, 0.f);
      reduce_scatter4(v8, 4, t.xg, vk);
      if (t.owner) {
        const float2* yv = sYall + (j0 + jj) * kM + t.x0;
        float2* dst = dAb + ((long long)(j0 + jj) * kM + t.y0) * C + t.c;
        MGB_UNROLL
        for (int k = 0; k < 2; ++k) {
          const int qx = 2 * t.yg + k, qy = 2 * t.xg + k;
          if (qx < t.nx) {
            float2 de = make_float2(0.f, 0.f);
            cfmacl(de, yv[qx], wk[k]);             // conj(Y_x) w[x]: the component's share of dE_ij[l(x)]
            sDE[(jj * kM + t.x0 + qx) * C + t.c] = de;
          }
          if (qy < t.ny) atomic_add2(dst + qy * C, vk[k]);
        }
      }
    }
    __syncthreads();
    // sum the 2l+1 contributions of every (j, l, c) and write dE
    for (int idx = threadIdx.x; idx < nj * kNL * C; idx += blockDim.x) {
      const int jj = idx / (kNL * C), r = idx - jj * (kNL * C), l = r / C, cc = r - l * C;
      float2 acc = make_float2(0.f, 0.f);
      for (int m = 0; m < 2 * l + 1; ++m) { const float2 x = sDE[(jj * kM + l * l + m) * C + cc]; acc.x += x.x; acc.y += x.y; }
      float2* dst = dE_i + (long long)j0 * kNL * C + idx;
      if (accumulate_dE) { acc.x += dst->x; acc.y += dst->y; }
      *dst = acc;
    }
  }
}

}  // namespace mgb
