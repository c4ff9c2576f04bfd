// HBM bandwidth microbenchmarks on B200: pure read, pure write, copy, and write with different store widths.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/membench tools/membench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) k_read(const double2* __restrict__ a, size_t n, double* out) {
  double s = 0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    double2 v = a[i];
    s += v.x + v.y;
  }
  if (s == 1.2345) out[0] = s;
}
__global__ void __launch_bounds__(256) k_write(double2* __restrict__ a, size_t n, double v) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) a[i] = make_double2(v, v);
}
__global__ void __launch_bounds__(256) k_write8(double* __restrict__ a, size_t n, double v) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) a[i] = v;
}
__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) b[i] = a[i];
}
// non-persistent: one element per thread
__global__ void __launch_bounds__(256) k_write_np(double2* __restrict__ a, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) a[i] = make_double2(v, v);
}

template <class F>
float timeit(F f, int reps = 10) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  const size_t bytes = (size_t)4 << 30;
  const size_t n2 = bytes / 16;
  double2 *a, *b; double* out;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&out, 8);
  cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
  for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
    float t;
    t = timeit([&] { k_read<<<g, 256>>>(a, n2, out); });
    printf("grid %5d read   %.0f GB/s\n", g, bytes / t / 1e6);
    t = timeit([&] { k_write<<<g, 256>>>(a, n2, 1.0); });
    printf("grid %5d write16 %.0f GB/s\n", g, bytes / t / 1e6);
    t = timeit([&] { k_write8<<<g, 256>>>((double*)a, n2 * 2, 1.0); });
    printf("grid %5d write8  %.0f GB/s\n", g, bytes / t / 1e6);
    t = timeit([&] { k_copy<<<g, 256>>>(a, b, n2); });
    printf("grid %5d copy   %.0f GB/s (read+write)\n", g, 2.0 * bytes / t / 1e6);
  }
  float t = timeit([&] { k_write_np<<<(unsigned)((n2 + 255) / 256), 256>>>(a, n2, 1.0); });
  printf("non-persistent write16 %.0f GB/s\n", bytes / t / 1e6);
  t = timeit([&] { cudaMemsetAsync(a, 0, bytes); });
  printf("cudaMemset %.0f GB/s\n", bytes / t / 1e6);
  t = timeit([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); });
  printf("cudaMemcpy D2D %.0f GB/s (read+write)\n", 2.0 * bytes / t / 1e6);
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
