// cov_backward.cuh — backward kernels (filled in below).
#pragma once
#include "heads.cuh"
namespace mgb {}
