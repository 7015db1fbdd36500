// cusim.h — TEST INFRASTRUCTURE: a tiny CPU emulator for CUDA kernels.
//
// The build container has nvcc but no GPU.  To debug kernel *logic* (indexing, barriers, reductions) before
// spending GPU time, the same .cu sources are compiled a second time with g++ -DMGB_CUSIM, this header standing
// in for <cuda_runtime.h>.  Every CUDA thread of a block runs as a ucontext fiber; __syncthreads / warp shuffles
// are cooperative yields; blocks run one after another (optionally spread over OS threads).  It is used ONLY by
// tests/ (built into tests/cusim/_build/) — the product library is the nvcc build and never falls back to this.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local
#define __constant__ static

struct uint3_ { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct int4 { int x, y, z, w; };
struct double2 { double x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline const char* cudaGetErrorString(cudaError_t) { return "cusim"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
typedef void* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return 0; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
template <class T> static inline cudaError_t cudaFuncSetAttribute(T, int, int) { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace cusim {

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
  unsigned tid = 0;
};

struct BlockState {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int current = -1;
  int alive = 0;
  int bar_count = 0;
  unsigned bar_gen = 0;
  // per-warp state
  std::vector<int> warp_count;
  std::vector<unsigned> warp_gen;
  std::vector<uint64_t> warp_buf;  // 32 slots per warp, 8 bytes each
  const std::function<void()>* body = nullptr;
  char* dyn_smem = nullptr;
};

inline thread_local BlockState* g_block = nullptr;
inline thread_local uint3_ g_threadIdx{0, 0, 0}, g_blockIdx{0, 0, 0};
inline thread_local dim3 g_blockDim, g_gridDim;

static constexpr size_t kStackBytes = 256 * 1024;

inline void yield_() {
  BlockState* b = g_block;
  swapcontext(&b->fibers[b->current].ctx, &b->sched);
}

inline void fiber_entry() {
  BlockState* b = g_block;
  (*b->body)();
  Fiber& f = b->fibers[b->current];
  f.done = true;
  b->alive--;
  // exited threads count as arrived at any pending barrier
  if (b->alive > 0 && b->bar_count >= b->alive) { b->bar_count = 0; b->bar_gen++; }
  swapcontext(&f.ctx, &b->sched);
}

inline void set_tid(unsigned t) {
  g_threadIdx.x = t % g_blockDim.x;
  g_threadIdx.y = (t / g_blockDim.x) % g_blockDim.y;
  g_threadIdx.z = t / (g_blockDim.x * g_blockDim.y);
}

inline void run_block(const std::function<void()>& body, dim3 grid, dim3 block, unsigned bx, unsigned by, unsigned bz,
                      size_t smem_bytes) {
  BlockState st;
  unsigned nt = block.x * block.y * block.z;
  st.fibers.resize(nt);
  st.alive = nt;
  unsigned nw = (nt + 31) / 32;
  st.warp_count.assign(nw, 0);
  st.warp_gen.assign(nw, 0);
  st.warp_buf.assign(nw * 32, 0);
  st.body = &body;
  st.dyn_smem = (char*)std::calloc(smem_bytes + 16, 1);
  g_block = &st;
  g_blockDim = block; g_gridDim = grid;
  g_blockIdx = uint3_{bx, by, bz};
  for (unsigned t = 0; t < nt; ++t) {
    Fiber& f = st.fibers[t];
    f.tid = t;
    f.stack = (char*)std::malloc(kStackBytes);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStackBytes;
    f.ctx.uc_link = &st.sched;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  int remaining = nt;
  while (remaining > 0) {
    remaining = 0;
    for (unsigned t = 0; t < nt; ++t) {
      Fiber& f = st.fibers[t];
      if (f.done) continue;
      st.current = t;
      set_tid(t);
      swapcontext(&st.sched, &f.ctx);
      if (!f.done) remaining++;
    }
  }
  for (auto& f : st.fibers) std::free(f.stack);
  std::free(st.dyn_smem);
  g_block = nullptr;
}

inline int num_workers() {
  const char* e = std::getenv("CUSIM_THREADS");
  int n = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
  return std::max(1, n);
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  if (nblocks == 0) return;
  int nworkers = (int)std::min<size_t>(num_workers(), nblocks);
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= nblocks) break;
      unsigned bx = i % grid.x, by = (i / grid.x) % grid.y, bz = i / ((size_t)grid.x * grid.y);
      run_block(body, grid, block, bx, by, bz, smem_bytes);
    }
  };
  if (nworkers == 1) { worker(); return; }
  std::vector<std::thread> ths;
  for (int w = 0; w < nworkers; ++w) ths.emplace_back(worker);
  for (auto& t : ths) t.join();
}

inline void syncthreads() {
  BlockState* b = g_block;
  unsigned gen = b->bar_gen;
  if (++b->bar_count >= b->alive) { b->bar_count = 0; b->bar_gen++; return; }
  while (b->bar_gen == gen) yield_();
}

inline void syncwarp_n(int expected) {
  BlockState* b = g_block;
  int w = b->current / 32;
  unsigned gen = b->warp_gen[w];
  if (++b->warp_count[w] >= expected) { b->warp_count[w] = 0; b->warp_gen[w]++; return; }
  while (b->warp_gen[w] == gen) yield_();
}

inline int warp_expected(unsigned mask) {
  BlockState* b = g_block;
  int w = b->current / 32;
  int nt = (int)b->fibers.size();
  int lanes_in_warp = std::min(32, nt - w * 32);
  unsigned valid = lanes_in_warp == 32 ? 0xffffffffu : ((1u << lanes_in_warp) - 1);
  return __builtin_popcount(mask & valid);
}

template <class T> inline T shfl_idx(unsigned mask, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl type too large");
  BlockState* b = g_block;
  int w = b->current / 32, lane = b->current % 32;
  int exp = warp_expected(mask);
  uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
  b->warp_buf[w * 32 + lane] = raw;
  syncwarp_n(exp);
  uint64_t got = b->warp_buf[w * 32 + (src_lane & 31)];
  syncwarp_n(exp);
  T out; std::memcpy(&out, &got, sizeof(T));
  return out;
}

}  // namespace cusim

#define threadIdx (cusim::g_threadIdx)
#define blockIdx (cusim::g_blockIdx)
#define blockDim (cusim::g_blockDim)
#define gridDim (cusim::g_gridDim)
static constexpr int warpSize = 32;

static inline void __syncthreads() { cusim::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { cusim::syncwarp_n(cusim::warp_expected(mask)); }
template <class T> static inline T __shfl_sync(unsigned m, T v, int src, int = 32) { return cusim::shfl_idx(m, v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int lm, int = 32) {
  return cusim::shfl_idx(m, v, (cusim::g_block->current % 32) ^ lm);
}
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int = 32) {
  int lane = cusim::g_block->current % 32;
  int src = lane + (int)d;
  if (src > 31) src = lane;
  return cusim::shfl_idx(m, v, src);
}

static inline unsigned __float_as_uint(float x) { unsigned u; std::memcpy(&u, &x, 4); return u; }
static inline float __uint_as_float(unsigned u) { float x; std::memcpy(&x, &u, 4); return x; }
static inline unsigned __ballot_sync(unsigned m, int pred) {
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) r |= (unsigned)(cusim::shfl_idx(m, pred ? 1 : 0, l) & 1) << l;
  return r;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int = 32) {
  int lane = cusim::g_block->current % 32;
  int src = lane - (int)d;
  if (src < 0) src = lane;
  return cusim::shfl_idx(m, v, src);
}

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) {
  uint32_t* ip = (uint32_t*)p;
  uint32_t cur = __atomic_load_n(ip, __ATOMIC_RELAXED);
  for (;;) {
    float f; std::memcpy(&f, &cur, 4);
    float nf = f + v; uint32_t ni; std::memcpy(&ni, &nf, 4);
    if (__atomic_compare_exchange_n(ip, &cur, ni, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return f;
  }
}
static inline double atomicAdd(double* p, double v) {
  uint64_t* ip = (uint64_t*)p;
  uint64_t cur = __atomic_load_n(ip, __ATOMIC_RELAXED);
  for (;;) {
    double f; std::memcpy(&f, &cur, 8);
    double nf = f + v; uint64_t ni; std::memcpy(&ni, &nf, 8);
    if (__atomic_compare_exchange_n(ip, &cur, ni, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return f;
  }
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

using std::min;
using std::max;
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline void sincosf_(float x, float* s, float* c) { *s = std::sin(x); *c = std::cos(x); }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }

// Launch + dynamic shared memory portability macros (mirrored for CUDA in csrc/portable.h)
namespace mgb { struct Profiler { long long launches = 0; }; inline Profiler g_prof; }
#define MGB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ++mgb::g_prof.launches;                                  \
  cusim::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#define MGB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(((uintptr_t)cusim::g_block->dyn_smem + 15) & ~(uintptr_t)15)
