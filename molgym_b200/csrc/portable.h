// portable.h — one switch between the real CUDA toolchain (product) and the CPU kernel emulator used by tests.
#pragma once
#ifdef MGB_CUSIM
#include "cusim.h"  // tests/cusim/cusim.h — test infrastructure only
#define MGB_UNROLL
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#define MGB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MGB_DYN_SMEM(type, name)                                        \
  extern __shared__ __align__(16) unsigned char _mgb_dyn_smem[];        \
  type* name = reinterpret_cast<type*>(_mgb_dyn_smem)
#define MGB_UNROLL _Pragma("unroll")
#endif
