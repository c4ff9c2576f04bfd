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

// Packed storage of the symmetric element tangent: node-pair blocks (a, b >= a).
// For NN >= 8 the element kernel produces the row block in two column halves (b < NN/2, then b >= NN/2),
// so the pairs of the first half are stored first: [a <= b < NB] then [b >= NB].  Mirrored by
// jax_fem_b200/plan.py::pair_index.
template <int NN>
__host__ __device__ constexpr int pair_split() { return NN >= 8 ? NN / 2 : NN; }
template <int NN>
__host__ __device__ constexpr int pair_first_half() { return NN >= 8 ? (NN / 2) * (NN / 2 + 1) / 2 : 0; }
template <int NN>
__host__ __device__ __forceinline__ int pair_index(int a, int b) {   // requires a <= b
  constexpr int NB = pair_split<NN>();
  if (NB == NN) return a * NN - (a * (a - 1)) / 2 + (b - a);
  constexpr int P0 = pair_first_half<NN>(), W = NN - NB;
  if (b < NB) return a * NB - (a * (a - 1)) / 2 + (b - a);
  if (a < NB) return P0 + a * W + (b - NB);
  const int r = a - NB;
  return P0 + NB * W + r * W - (r * (r - 1)) / 2 + (b - a);
}

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
