// Shared helpers for libfem_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/fem_b200.h"

namespace femb200 {

void set_error(const char* fmt, ...);
int check_device();

#define FEM_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      femb200::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                         cudaGetErrorString(_e));                                         \
      return FEM_ECUDA;                                                                   \
    }                                                                                     \
  } while (0)

#define FEM_LAUNCH_CHECK() FEM_CUDA_CHECK(cudaGetLastError())

#define FEM_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) {                                               \
      femb200::set_error("invalid argument: %s", msg);           \
      return FEM_EINVAL;                                         \
    }                                                            \
  } while (0)

constexpr int kNumSM = 148;   // B200

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block reduction of up to NV values; result valid in thread 0.
// Butterfly order inside a warp and a fixed warp order across warps => bit-reproducible.
template <int NV, int THREADS>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* NV*THREADS/32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) smem[k * (THREADS / 32) + warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      for (int w = 0; w < THREADS / 32; ++w) s += smem[k * (THREADS / 32) + w];
      v[k] = s;
    }
  }
}

}  // namespace femb200
