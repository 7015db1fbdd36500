// common.cuh — shared device helpers: complex arithmetic, spherical harmonics, block reductions, descriptors.
#pragma once
#include "portable.h"
#include "../../include/molgym_b200.h"

namespace mgb {

constexpr int kL = 4;             // maxl supported by this build
constexpr int kNL = kL + 1;       // number of ells
constexpr int kM = 25;            // (kL+1)^2 spherical components, lm = l*l + l + m
constexpr int kRadFeat = 32;      // 8 trig x 4 inverse powers (basis_set=[3,3], covariant/agent.py:73)
constexpr int kTrig = 8;
constexpr float kTwoPi = 6.283185307179586f;
constexpr float kFourPi = 12.566370614359172f;

__host__ __device__ __forceinline__ int lm_index(int l, int m) { return l * l + l + m; }
__host__ __device__ __forceinline__ int ell_of_lm(int lm) { return lm >= 16 ? 4 : (lm >= 9 ? 3 : (lm >= 4 ? 2 : (lm >= 1 ? 1 : 0))); }

// ---- complex helpers (float2 = re, im) ---------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ void cfma(float2& acc, float2 a, float2 b) {  // acc += a*b
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfmac(float2& acc, float2 a, float2 b) {  // acc += a*conj(b)
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.y, b.x, acc.y); acc.y = fmaf(-a.x, b.y, acc.y);
}
__device__ __forceinline__ void cfmacl(float2& acc, float2 a, float2 b) {  // acc += conj(a)*b
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(-a.y, b.x, acc.y);
}

// ---- packed FP32 pairs (Blackwell FFMA2: fma.rn.f32x2, two IEEE fp32 FMAs per instruction) -----------------------
// A complex MAC acc += w * x costs four FFMAs; with two pair accumulators
//     P += (w.re, w.im) * (x.re, x.re)      Q += (w.re, w.im) * (x.im, x.im)
// it costs two FFMA2s and the product is recovered once at the end: w*x = (P.lo - Q.hi, P.hi + Q.lo),
// conj(w)*x = (P.lo + Q.hi, Q.lo - P.hi).  The pair (w.re, w.im) is the complex number as it sits in memory.
#ifndef MGB_CUSIM
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpack2(f32x2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ void fma2(f32x2& acc, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ f32x2 as_pair(float2 v) { return pack2(v.x, v.y); }
#else
struct f32x2 { float lo, hi; };
__device__ inline f32x2 pack2(float lo, float hi) { return f32x2{lo, hi}; }
__device__ inline float2 unpack2(f32x2 v) { return make_float2(v.lo, v.hi); }
__device__ inline void fma2(f32x2& acc, f32x2 a, f32x2 b) { acc.lo = fmaf(a.lo, b.lo, acc.lo); acc.hi = fmaf(a.hi, b.hi, acc.hi); }
__device__ inline f32x2 as_pair(float2 v) { return f32x2{v.x, v.y}; }
#endif
// w * x  and  conj(w) * x  from the pair accumulators above
__device__ __forceinline__ float2 cpair_mul(f32x2 P, f32x2 Q) { const float2 p = unpack2(P), q = unpack2(Q); return make_float2(p.x - q.y, p.y + q.x); }
__device__ __forceinline__ float2 cpair_mulc(f32x2 P, f32x2 Q) { const float2 p = unpack2(P), q = unpack2(Q); return make_float2(p.x + q.y, q.x - p.y); }

// ---- complex atomic add (one 8-byte RED on sm_90+; needs an 8-byte aligned destination) ------------------------------
#ifdef MGB_CUSIM
__device__ inline void atomic_add2(float2* p, float2 v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); }
#else
__device__ __forceinline__ void atomic_add2(float2* p, float2 v) { atomicAdd(p, v); }  // red.global.add.v2.f32
#endif

// ---- asynchronous 8-byte global -> shared copies (LDGSTS); the emulator build copies synchronously ----------------
__device__ __forceinline__ void cp_async8(float2* smem_dst, const float2* __restrict__ gmem_src) {
#ifndef MGB_CUSIM
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
#else
  *smem_dst = *gmem_src;
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef MGB_CUSIM
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifndef MGB_CUSIM
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// ---- spherical harmonics -------------------------------------------------------------------------------------
// Complex Y_lm(v) for l <= 4, Condon-Shortley phase, m = -l..l stored at lm = l*l+l+m, evaluated as solid
// harmonics (|v|^l Y_lm(v/|v|)) so that an un-normalised argument reproduces Cormorant's recursion
// Y_l ~ CG(Y_{l-1} x Y_1) (oracle/thirdparty/cormorant/cg_lib.py::spherical_harmonics).
//   unit_norm: multiply each l by sqrt(4 pi / (2l+1))   (sh_norm='unit', else 'qm')
//   conj:      complex conjugate                        (SphericalHarmonicsRel(conj=True), covariant/modules.py:52-56)
__device__ __forceinline__ void sph_harm_l4(float x, float y, float z, bool unit_norm, bool conj, float2* out) {
  const float r2 = x * x + y * y + z * z;
  const float z2 = z * z;
  // (x + i y)^m
  float2 e1 = make_float2(x, y);
  float2 e2 = cmul(e1, e1);
  float2 e3 = cmul(e2, e1);
  float2 e4 = cmul(e2, e2);
  // D_lm(z, r) = r^(l-m) d^m P_l / du^m (u = z/r)
  const float d00 = 1.f;
  const float d10 = z, d11 = 1.f;
  const float d20 = 0.5f * (3.f * z2 - r2), d21 = 3.f * z, d22 = 3.f;
  const float d30 = 0.5f * z * (5.f * z2 - 3.f * r2), d31 = 0.5f * (15.f * z2 - 3.f * r2), d32 = 15.f * z, d33 = 15.f;
  const float d40 = 0.125f * (35.f * z2 * z2 - 30.f * z2 * r2 + 3.f * r2 * r2);
  const float d41 = 0.5f * z * (35.f * z2 - 15.f * r2), d42 = 0.5f * (105.f * z2 - 15.f * r2), d43 = 105.f * z, d44 = 105.f;
  // N_lm = sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) for 'qm'; sqrt((l-m)!/(l+m)!) for 'unit'
  const float q0 = unit_norm ? 1.f : 0.28209479177387814f;   // sqrt(1/4pi)
  const float q1 = unit_norm ? 1.f : 0.4886025119029199f;    // sqrt(3/4pi)
  const float q2 = unit_norm ? 1.f : 0.6307831305050401f;    // sqrt(5/4pi)
  const float q3 = unit_norm ? 1.f : 0.7463526651802308f;    // sqrt(7/4pi)
  const float q4 = unit_norm ? 1.f : 0.8462843753216345f;    // sqrt(9/4pi)
  const float s = conj ? -1.f : 1.f;
#define MGB_SET(l, m, nrm, d, e)                                                   \
  {                                                                                \
    const float a_ = (nrm) * (d);                                                  \
    const float sg_ = ((m) & 1) ? -1.f : 1.f;                                      \
    out[(l) * (l) + (l) + (m)] = make_float2(sg_ * a_ * (e).x, s * sg_ * a_ * (e).y); \
    out[(l) * (l) + (l) - (m)] = make_float2(a_ * (e).x, -s * a_ * (e).y);          \
  }
  out[0] = make_float2(q0 * d00, 0.f);
  out[2] = make_float2(q1 * d10, 0.f);
  MGB_SET(1, 1, q1 * 0.7071067811865476f, d11, e1)
  out[6] = make_float2(q2 * d20, 0.f);
  MGB_SET(2, 1, q2 * 0.408248290463863f, d21, e1)      // sqrt(1/6)
  MGB_SET(2, 2, q2 * 0.2041241452319315f, d22, e2)     // sqrt(1/24)
  out[12] = make_float2(q3 * d30, 0.f);
  MGB_SET(3, 1, q3 * 0.2886751345948129f, d31, e1)     // sqrt(2!/4!) = sqrt(1/12)
  MGB_SET(3, 2, q3 * 0.09128709291752768f, d32, e2)    // sqrt(1/120)
  MGB_SET(3, 3, q3 * 0.03726779962499649f, d33, e3)    // sqrt(1/720)
  out[20] = make_float2(q4 * d40, 0.f);
  MGB_SET(4, 1, q4 * 0.22360679774997896f, d41, e1)    // sqrt(3!/5!) = sqrt(1/20)
  MGB_SET(4, 2, q4 * 0.05270462766947299f, d42, e2)    // sqrt(2!/6!) = sqrt(1/360)
  MGB_SET(4, 3, q4 * 0.014085904245475277f, d43, e3)   // sqrt(1/5040)
  MGB_SET(4, 4, q4 * 0.004980119205559973f, d44, e4)   // sqrt(1/40320)
#undef MGB_SET
}

// ---- block-wide reductions (blockDim.x multiple of 32, <= 1024) ----------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// scratch: >= 32 floats of shared memory. All threads must call. Result broadcast to all threads.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : -3.0e38f;
  r = warp_max(r);
  return r;
}

// ---- TMA bulk copies global -> shared memory (cp.async.bulk + mbarrier; SASS UBLKCP) ---------------------------
// Used to stage weight matrices: one thread issues the copy, the whole tile is in flight at once, and everybody waits on
// the mbarrier.  Source / destination must be 16-byte aligned and the size a multiple of 16 (smem_fill_ok); callers fall
// back to cooperative loads otherwise.  The CPU emulator build copies synchronously.
struct SmemBarrier { unsigned long long word; };
__device__ __forceinline__ bool smem_fill_ok(const void* src, long long bytes) {
  return ((reinterpret_cast<unsigned long long>(src) & 15ull) == 0ull) && (bytes % 16 == 0) && bytes > 0;
}
#ifndef MGB_CUSIM
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(SmemBarrier* bar, int arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // make the init visible to the async (TMA) proxy
}
__device__ __forceinline__ void mbar_expect(SmemBarrier* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, SmemBarrier* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(SmemBarrier* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MGB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MGB_DONE;\n"
      "bra MGB_WAIT;\n"
      "MGB_DONE:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
#else
__device__ inline void mbar_init(SmemBarrier*, int) {}
__device__ inline void mbar_expect(SmemBarrier*, unsigned) {}
__device__ inline void bulk_g2s(void* dst, const void* src, unsigned bytes, SmemBarrier*) { std::memcpy(dst, src, bytes); }
__device__ inline void mbar_wait(SmemBarrier*, unsigned) { __syncthreads(); }   // every thread of the CTA waits in our kernels
#endif
// Whole-CTA helper: start filling dst[0..n) (floats) from src.  Returns true when the copy is asynchronous (wait with
// mbar_wait(bar, parity) before reading); false when it was done with plain loads (a __syncthreads() is still needed).
// `bar` must have been initialised (mbar_init(bar, 1) + __syncthreads()) by the caller; one fill per barrier phase.
__device__ __forceinline__ bool smem_fill_begin(float* dst, const float* __restrict__ src, int n, SmemBarrier* bar) {
  const long long bytes = 4ll * n;
  if (smem_fill_ok(src, bytes) && smem_fill_ok(dst, bytes)) {
    if (threadIdx.x == 0) {
      mbar_expect(bar, (unsigned)bytes);
      bulk_g2s(dst, src, (unsigned)bytes, bar);
    }
    return true;
  }
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) dst[idx] = src[idx];
  return false;
}
__device__ __forceinline__ void smem_fill_end(bool async, SmemBarrier* bar, unsigned parity) {
  if (async) mbar_wait(bar, parity); else __syncthreads();
}

// ---- Clebsch-Gordan term tables (device pointers; built on the host in plan.cuh) --------------------------------
struct CgTable {
  int n_out;               // number of (path, m) outputs
  int n_pair;              // M1 * M2 input pairs
  int nlm2;                // number of lm components of the second factor (1 or 25)
  const int* out_l;        // [n_out] output ell
  const int* out_m;        // [n_out] output m index 0..2l
  const int* out_block;    // [n_out] channel-block index inside cat_l (slot = block*C + c)
  const int* term_start;   // [n_out+1]
  const int* term_lm1;     // [n_term]
  const int* term_lm2;     // [n_term]
  const float* term_coef;  // [n_term]
  const int* pair_start;   // [n_pair+1]   transposed table: pair = lm1*nlm2 + lm2
  const int* pair_out;     // [n_term]
  const float* pair_coef;  // [n_term]
  // resolved against one use site (cat layout, channel count): offsets in complex units for channel 0
  const int* out_dst;      // [n_out]  destination inside the cat vector
  const int2* term_src;    // [n_term] (a, b): product table -> a = (lm1*nlm2+lm2)*C ; square -> a = lm1*C, b = lm2*C
  const int2* pair_ent;    // [n_term] (destination inside the cat vector, float bits of the coefficient)
  // flat forms for the kernels
  const int4* flat;        // [n_term] output-major: (a, b, dst << 1 | last-term-of-output, float bits of the coefficient)
  const int* slot_start;   // [n_slots + 1] term ranges (whole outputs) of equal work
  int n_slots;
  const int2* pad_pair;    // [n_pair][kCgPad] (dst, coef bits), zero-padded: transposed table with a fixed trip count
  const int2* pad_sym;     // [n_pair][2 * kCgPad] entries of (x, y) and of (y, x)  (square tables only)
};
constexpr int kCgPad = 5;    // max number of l values a pair (l1 m1, l2 m2) couples to for l <= 4

// Forward CG gather tables of the levels whose input carries all ells (cov_forward.cuh::gather25): output-major, 8 bytes per term,
//   aggregate (a | last << 13 | dst << 14, coef), a = offset of the Kronecker sum T[lm1][lm2] (channel 0) in shared memory;
//   square    (a | b << 8 | last << 16 | dst << 17, coef), a / b = offsets of the two factors of A_i;
// dst = offset inside the atom's cat vector; `last` closes the run of one output.  ag_slot / sq_slot [kGatherSlots + 1] cut the
// lists into ranges of whole outputs of about equal length (one range per group of C threads of the 256-thread CTA).
constexpr int kGatherSlots = 25;
struct GatherTables {
  const int2* ag_flat8;
  const int2* sq_flat8;
  const int* ag_slot;
  const int* sq_slot;
};

// One Cormorant level (edge network + atom network), everything the kernels need by value.
struct LevelDesc {
  int nLin;          // ells present in the input atom reps (1 at level 0, else 5)
  int nlm_in;        // nLin^2
  int C;             // input / edge channels
  int Cout;          // atom-mix output channels
  int has_prev;      // previous-level edge scalars feed the edge mix
  int catE[kNL];     // edge cat sizes: [prev (C) | dot (nLin*C, only l < nLin) | radial (C)]
  int offE[kNL];     // complex offset of l block inside the packed edge weights (sum C*catE)
  int totE;          // total complex edge weights
  int sumCatE;       // sum_l catE[l]
  int catA[kNL];     // atom cat sizes: [ag paths | in (C, only l < nLin) | sq paths]
  int offA[kNL];     // complex offset of l block inside a per-atom CAT row: sum_{l'<l} catA[l']*(2l'+1)
  int totA;          // per-atom CAT size (complex)
  int offWA[kNL];    // complex offset of l block inside the packed atom weights (sum Cout*catA)
  int totWA;
  int mixCS;         // row stride (complex) of the transposed, padded copy of the atom weights in the scratch: [l][k][mixCS]
  int offWAt[kNL];   // complex offset of l block inside that copy (sum mixCS*catA)
  int in_block[kNL]; // block index of the pass-through input rep inside cat_l (-1 if absent)
  int sq_block[kNL]; // first block index of the CG-square paths inside cat_l
  long long p_scales, p_phases, p_radW, p_radb, p_edgeW, p_atomW;  // float offsets into the flat parameter buffer
  CgTable ag, sq;
  GatherTables gt;   // valid when nLin == 5
};

}  // namespace mgb
