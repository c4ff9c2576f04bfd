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

// The element kernels produce the row block of a corner in column chunks of pair_split<NN>() nodes.
template <int NN>
__host__ __device__ constexpr int pair_split() { return NN >= 8 ? NN / 2 : NN; }

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16; completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// shared -> global bulk copy (SASS UBLKCP.G.S): asynchronous, read by the TMA unit, no LSU issue slots.  The generic-proxy
// writes that filled the source must be fenced (fence_async_smem) before it is issued; the source may be reused once
// bulk_wait_read<N>() says that all but the N most recent groups of this thread have been read.
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// FP64 tensor-core tile: C(8x8) += A(8x4, row-major) B(4x8, col-major).  Lane l supplies A[l/4][l%4] and B[l%4][l/4] and
// holds C[l/4][2(l%4)], C[l/4][2(l%4)+1].  DMMA and DFMA share one FP64 pipe on B200 (tools/microbench.cu): this buys
// operand bandwidth (operands come from registers), not flops.
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
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
