// EXPERIMENT, not part of the build: tensor-core (3xTF32 mma.sync) weight gradients of the MLP heads.  Parity-green on the emulator and
// on the GPU (33 GPU tests), but no step-time gain: the weight gradients run on the lowest-priority side stream in the gaps of the
// main chain (C3 b1024 step 6.11 ms with and without, C4 b1024 12.59 vs 12.48 ms, C5 b256 7.98 vs 7.93 ms), and it costs one more
// launch at C2.  Kept for the record; see tools/experiments/README.md.
// ------------------------------------------------------------------------------------------------------------
// Weight gradients of the MLPs on the tensor cores: dW[o][k] += sum_rows dY[r][o] X[r][k], db[o] += sum_rows dY[r][o] — a GEMM
// with the rows as the inner dimension (A = dY^T, B = [X | 1]).  grid = (row chunks, work items of k_dw_grouped's list: problem x
// 32 outputs); a CTA stages 32 rows of X (plus the ones column that yields the bias gradient) and of its dY slice at a time,
// warp w owns the n-tiles w, w + 8, ... of the K + 8 input columns for both m-tiles; the rows come from the compact active / valid
// lists (padding atoms cost nothing).  The partial sums leave with one atomic per entry and CTA.
// ------------------------------------------------------------------------------------------------------------
constexpr int kDwTcRows = 32;
constexpr int kDwTcMO = 32;      // outputs per CTA
constexpr int kDwTcNT = 5;       // n-tiles per warp: K + 8 <= 8 * 8 * kDwTcNT
__host__ __device__ inline bool dw_tc_ok(int K, int No) { return No % kDwTcMO == 0 && K % 8 == 0 && K + 8 <= 64 * kDwTcNT; }
__host__ __device__ inline size_t dw_tc_smem_bytes(int Kmax) {
  return sizeof(float) * ((size_t)kDwTcRows * tc_stride_kn(Kmax + 8) + (size_t)kDwTcRows * tc_stride_kn(kDwTcMO)) + sizeof(long long) * kDwTcRows;
}

__global__ void __launch_bounds__(kTcThreads, 2)
k_dw_tc(const MGB_GRID_CONSTANT DwProblemList list, int B, const int* __restrict__ act_off, const int* __restrict__ act_list,
        const int* __restrict__ atom_off, const int* __restrict__ atom_list, float* __restrict__ grad) {
  const DwWork wk = list.w[blockIdx.y];
  const DwProblem pr = list.p[wk.prob];
  const int* rlist = pr.mode == kRowsActive ? act_list : (pr.mode == kRowsValid ? atom_list : nullptr);
  const long long n_rows = pr.mode == kRowsActive ? act_off[B] : (pr.mode == kRowsValid ? atom_off[B] : pr.rows);
  long long per = (n_rows + gridDim.x - 1) / gridDim.x;
  per = (per + kDwTcRows - 1) / kDwTcRows * kDwTcRows;
  const long long r_begin = per * blockIdx.x, r_end = r_begin + per < n_rows ? r_begin + per : n_rows;
  if (r_begin >= r_end) return;
  const int K = pr.K, No = pr.No, o0 = wk.o0, n_tiles = (K + 8) / 8;
  const int sxs = tc_stride_kn(K + 8), sdo = tc_stride_kn(kDwTcMO);
  MGB_DYN_SMEM(float, sm);
  float* sx = sm;                                        // [32][sxs]: X rows, column K = 1 (bias), zero up to K + 8
  float* sdy = sx + (size_t)kDwTcRows * sxs;             // [32][sdo]: dY[r][o0 .. o0 + 32)
  long long* s_row = reinterpret_cast<long long*>(sdy + (size_t)kDwTcRows * sdo);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float acc[2][kDwTcNT][4];
  MGB_UNROLL
  for (int mt = 0; mt < 2; ++mt)
    MGB_UNROLL
    for (int j = 0; j < kDwTcNT; ++j) { acc[mt][j][0] = 0.f; acc[mt][j][1] = 0.f; acc[mt][j][2] = 0.f; acc[mt][j][3] = 0.f; }
  const int k4n = K / 4;
  for (long long rc = r_begin; rc < r_end; rc += kDwTcRows) {
    __syncthreads();   // previous rows consumed
    if ((int)threadIdx.x < kDwTcRows) {
      const long long r = rc + threadIdx.x;
      s_row[threadIdx.x] = r < r_end ? (rlist ? (long long)rlist[r] : r) : -1;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < kDwTcRows * k4n; idx += blockDim.x) {
      const int q = idx / k4n, k4 = idx - q * k4n;
      const long long row = s_row[q];
      *reinterpret_cast<float4*>(sx + q * sxs + 4 * k4) =
          row >= 0 ? reinterpret_cast<const float4*>(pr.X + row * K)[k4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int idx = threadIdx.x; idx < kDwTcRows * 8; idx += blockDim.x) {
      const int q = idx >> 3, j = idx & 7;
      sx[q * sxs + K + j] = (j == 0 && s_row[q] >= 0) ? 1.f : 0.f;
    }
    for (int idx = threadIdx.x; idx < kDwTcRows * (kDwTcMO / 4); idx += blockDim.x) {
      const int q = idx / (kDwTcMO / 4), o4 = idx - q * (kDwTcMO / 4);
      const long long row = s_row[q];
      *reinterpret_cast<float4*>(sdy + q * sdo + 4 * o4) =
          row >= 0 ? reinterpret_cast<const float4*>(pr.dY + row * No + o0)[o4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    MGB_UNROLL
    for (int ks = 0; ks < kDwTcRows / 8; ++ks) {
      const int r0 = 8 * ks;
      Tf32Pair a[2][4];
      MGB_UNROLL
      for (int mt = 0; mt < 2; ++mt) {
        a[mt][0] = tf32_split(sdy[(r0 + t) * sdo + 16 * mt + g]);
        a[mt][1] = tf32_split(sdy[(r0 + t) * sdo + 16 * mt + g + 8]);
        a[mt][2] = tf32_split(sdy[(r0 + t + 4) * sdo + 16 * mt + g]);
        a[mt][3] = tf32_split(sdy[(r0 + t + 4) * sdo + 16 * mt + g + 8]);
      }
      MGB_UNROLL
      for (int j = 0; j < kDwTcNT; ++j) {
        const int tile = warp + 8 * j;
        if (tile < n_tiles) {
          Tf32Pair b[2];
          b[0] = tf32_split(sx[(r0 + t) * sxs + 8 * tile + g]);
          b[1] = tf32_split(sx[(r0 + t + 4) * sxs + 8 * tile + g]);
          mma_3xtf32(acc[0][j], a[0], b);
          mma_3xtf32(acc[1][j], a[1], b);
        }
      }
    }
  }
  MGB_UNROLL
  for (int mt = 0; mt < 2; ++mt)
    MGB_UNROLL
    for (int j = 0; j < kDwTcNT; ++j) {
      const int tile = warp + 8 * j;
      if (tile < n_tiles) {
        MGB_UNROLL
        for (int e = 0; e < 4; ++e) {
          const int o = o0 + 16 * mt + g + ((e & 2) ? 8 : 0), k = 8 * tile + 2 * t + (e & 1);
          const float v = acc[mt][j][e];
          if (v != 0.f && o < No) {
            if (k < K) atomicAdd(grad + pr.dW + (long long)o * K + k, v);
            else if (k == K && pr.db >= 0) atomicAdd(grad + pr.db + o, v);
          }
        }
      }
    }
}

}  // namespace mgb
