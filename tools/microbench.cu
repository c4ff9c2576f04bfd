// FP64 pipe microbenchmarks for B200: DFMA peak, DMMA (mma.sync f64) peak, and both together.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define ITERS 4096

__global__ void __launch_bounds__(256) dfma_kernel(double* out, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma16_kernel(double* out, double a0, double b0) {
  double c[4][4], a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = a0 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = b0 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3;
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// both in every warp: 8 DMMA (256 FMA each per warp) + 16 DFMA per thread (=512 FMA per warp) per iteration
__global__ void __launch_bounds__(256) mixed_kernel(double* out, double a, double b) {
  double c[8][2], x[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3;
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dmma884(c[i][0], c[i][1], a, b);
      x[2 * i] = fma(x[2 * i], a, b);
      x[2 * i + 1] = fma(x[2 * i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
double time_kernel(K k, int blocks, double* out, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<<<blocks, 256>>>(out, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) k<<<blocks, 256>>>(out, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps * 1e-3;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * 256);
  const int reps = 20;
  const double threads = (double)blocks * 256, warps = threads / 32;
  double t;
  t = time_kernel(dfma_kernel, blocks, out, reps);
  printf("DFMA      : %.2f TFLOP/s (%.3f ms)\n", 2.0 * threads * ITERS * 16 / t / 1e12, t * 1e3);
  t = time_kernel(dmma_kernel, blocks, out, reps);
  printf("DMMA 884  : %.2f TFLOP/s (%.3f ms)\n", 2.0 * warps * ITERS * 8 * 256 / t / 1e12, t * 1e3);
  t = time_kernel(dmma16_kernel, blocks, out, reps);
  printf("DMMA 16816: %.2f TFLOP/s (%.3f ms)\n", 2.0 * warps * (ITERS / 4) * 4 * 2048 / t / 1e12, t * 1e3);
  t = time_kernel(mixed_kernel, blocks, out, reps);
  printf("mixed     : %.2f TFLOP/s total (DMMA %.2f + DFMA %.2f) (%.3f ms)\n",
         2.0 * (warps * ITERS * 8 * 256 + threads * ITERS * 16) / t / 1e12, 2.0 * warps * ITERS * 8 * 256 / t / 1e12,
         2.0 * threads * ITERS * 16 / t / 1e12, t * 1e3);
  cudaError_t e = cudaGetLastError();
  printf("status: %s, SMs %d\n", cudaGetErrorString(e), sms);
  return 0;
}
