// portable.h — one switch between the real CUDA toolchain (product) and the CPU kernel emulator used by tests.
#pragma once
#ifdef MGB_CUSIM
#include "cusim.h"  // tests/cusim/cusim.h — test infrastructure only
#define MGB_UNROLL
#define MGB_GRID_CONSTANT
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
namespace mgb {
// Launch accounting + optional per-kernel timing (CUDA events recorded on the launching stream around every launch whose
// stringified name contains the selected substring); read back through mgb_launch_count / mgb_profile_read.
struct Profiler {
  long long launches = 0;
  char pattern[64] = "";
  bool active = false, hit = false;
  std::vector<cudaEvent_t> ev;   // start/stop pairs not yet read
  std::vector<const char*> names;   // kernel name (string literal) of every pair
};
inline Profiler g_prof;
inline void prof_begin(const char* name, cudaStream_t st) {
  ++g_prof.launches;
  g_prof.hit = g_prof.active && std::strstr(name, g_prof.pattern) != nullptr;
  if (g_prof.hit) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
    g_prof.ev.push_back(a); g_prof.ev.push_back(b);
    g_prof.names.push_back(name);
  }
}
inline void prof_end(cudaStream_t st) {
  if (g_prof.hit) cudaEventRecord(g_prof.ev.back(), st);
}
}  // namespace mgb
#define MGB_LAUNCH(kernel, grid, block, smem, stream, ...)            \
  do {                                                                \
    mgb::prof_begin(#kernel, (stream));                               \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);       \
    mgb::prof_end((stream));                                          \
  } while (0)
#define MGB_DYN_SMEM(type, name)                                        \
  extern __shared__ __align__(16) unsigned char _mgb_dyn_smem[];        \
  type* name = reinterpret_cast<type*>(_mgb_dyn_smem)
#define MGB_UNROLL _Pragma("unroll")
#define MGB_GRID_CONSTANT __grid_constant__
#endif
