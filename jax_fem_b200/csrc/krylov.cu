// Jacobi-preconditioned CG and BiCGSTAB, whole solve resident on the device.
//
// Replaces jax_solve (jax_fem/solver.py:63-92): CSR -> BCOO -> jax.scipy.sparse.linalg.bicgstab with
// M = 1/diag(A).  Same recurrences, same stopping rule (||r||^2 <= max(tol^2 ||b||^2, atol^2)), same
// breakdown codes as JAX's _bicgstab_solve / _cg_solve; every scalar (alpha, beta, omega, rho, the
// iteration counter and the convergence flag) lives in device memory and is produced by the last
// block of the kernel that completes the corresponding dot product, so an iteration is a fixed
// sequence of launches with no host round trip.  The host polls the flag every `check_every`
// iterations; kernels launched after convergence return immediately.
//
// Dot products are fused into the SpMV / vector-update kernels that produce their operands and are
// reduced in a fixed order (warp butterfly -> per-block partial -> last block sums partials by index),
// so results are bit-reproducible run to run.
#include "common.cuh"
#include "dist.cuh"

namespace femb200 {
int launch_spmv(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data, const double* x,
                double* y, cudaStream_t st);

namespace {

constexpr int kBlocks = kNumSM * 16;  // upper bound of any persistent grid (sizes the partial buffer)
constexpr int kThreads = 256;
constexpr int kScalars = 64;
constexpr int kMaxDots = 4;

// scalar slots (doubles) at the start of the workspace
enum {
  S_GAMMA = 0,   // CG: r.z ; BiCGSTAB: rho (previous)
  S_ALPHA = 1,
  S_BETA = 2,
  S_OMEGA = 3,
  S_RR = 4,      // ||r||^2
  S_ATOL2 = 5,
  S_K = 6,       // iteration counter (JAX breakdown codes -10 / -11)
  S_DONE = 7,
  S_MAXIT = 8,
  S_RHO_NEW = 9, // BiCGSTAB: rhat.r for the coming iteration
  S_EARLY = 10,  // BiCGSTAB: exit_early
  S_TOL2 = 11,
  S_ATOLIN2 = 12,
  S_TMP = 13,
  S_SUM0 = 16,   // distributed CG: rank-local partial sums, all-reduced by the caller between kernels
  S_SUM1 = 17,
  S_SUM2 = 18,
  S_SUM3 = 19,
  S_SUM4 = 20,
  S_SUM5 = 21,
  S_TICKET = 32  // unsigned counter (reinterpreted)
};

struct Ws {
  double* s;        // scalars
  double* partial;  // kBlocks * kMaxDots
  double* v[8];     // vectors
};

__host__ __device__ inline int64_t pad_n(int64_t n) { return (n + 31) / 32 * 32; }

inline Ws carve(double* w, int64_t n) {
  Ws o;
  o.s = w;
  o.partial = w + kScalars;
  double* base = o.partial + (int64_t)kBlocks * kMaxDots;
  for (int i = 0; i < 8; ++i) o.v[i] = base + i * pad_n(n);
  return o;
}

// Block partials -> global; the last block to arrive sums them in index order and runs `fin`.
// A product may be split over two launches (rows that need no ghost value first, the others after the halo has arrived):
// the launches write their partials side by side (`poff`, `ptotal` slots per dot) and only the last one (`finalize`) takes
// tickets and sums all of them -- stream order guarantees that the earlier launch has completed.
struct RowSplit {
  int64_t a0, a_cnt, b0, b_cnt;   // node ranges [a0, a0 + a_cnt) then [b0, b0 + b_cnt)
  int poff, ptotal, finalize;
};

template <int NV, class Fin>
__device__ __forceinline__ void reduce_and_finalize(double (&v)[NV], double* __restrict__ S,
                                                    double* __restrict__ partial, Fin fin, int poff, int ptotal,
                                                    bool finalize) {
  __shared__ double red[NV * (kThreads / 32)];
  __shared__ bool is_last;
  block_sum<NV, kThreads>(v, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partial[(int64_t)k * ptotal + poff + blockIdx.x] = v[k];
    is_last = false;
    if (finalize) {
      __threadfence();
      const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(S + S_TICKET), 1u);
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double tot[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double a = 0.0;
    for (int i = threadIdx.x; i < ptotal; i += kThreads) a += __ldcg(partial + (int64_t)k * ptotal + i);
    tot[k] = a;
  }
  __syncthreads();
  block_sum<NV, kThreads>(tot, red);
  if (threadIdx.x == 0) {
    *reinterpret_cast<unsigned*>(S + S_TICKET) = 0u;
    fin(tot);
    __threadfence();
  }
}

template <int NV, class Fin>
__device__ __forceinline__ void reduce_and_finalize(double (&v)[NV], double* __restrict__ S,
                                                    double* __restrict__ partial, Fin fin) {
  reduce_and_finalize<NV>(v, S, partial, fin, 0, (int)gridDim.x, true);
}

__device__ __forceinline__ bool solver_done(const double* S) { return __ldcg(S + S_DONE) != 0.0; }

// y = A x over a persistent grid, LPR lanes per row, optional fused dots with (row) entries.
// MODE 0: plain.  MODE 1 (CG): dot0 = d1[row]*y[row] -> alpha = gamma/dot0.
// MODE 2 (BiCGSTAB q = A phat): dot0 = rhat.q -> alpha = rho_new/dot0.
// MODE 3 (BiCGSTAB t = A shat): dot0 = t.s, dot1 = t.t -> omega = dot0/dot1.
// MODE 4 (distributed CG / BiCGSTAB): dot0 = d1.y over the rank's owned rows -> S[S_SUM0] (all-reduced by the caller).
// MODE 5 (distributed BiCGSTAB t = A shat): rank-local t.s -> S[S_SUM2], t.t -> S[S_SUM3].
template <int LPR, int MODE>
__global__ void __launch_bounds__(kThreads) spmv_fused_kernel(int64_t n, const int32_t* __restrict__ indptr,
                                                              const int32_t* __restrict__ indices,
                                                              const double* __restrict__ data,
                                                              const double* __restrict__ x, double* __restrict__ y,
                                                              const double* __restrict__ d1, double* __restrict__ S,
                                                              double* __restrict__ partial) {
  if (MODE != 0 && solver_done(S)) return;
  constexpr int RPB = kThreads / LPR;
  const int sub = threadIdx.x % LPR;
  double dots[2] = {0.0, 0.0};
  // the trip count is uniform across the block (loop on the block's first row) because the intra-row
  // shuffle below names every lane of the warp
  for (int64_t base = (int64_t)blockIdx.x * RPB; base < n; base += (int64_t)gridDim.x * RPB) {
    const int64_t row = base + threadIdx.x / LPR;
    const bool valid = row < n;
    double a0 = 0.0, a1 = 0.0;
    if (valid) {
      const int s = indptr[row], e = indptr[row + 1];
      int j = s + sub;
      for (; j + LPR < e; j += 2 * LPR) {
        const double v0 = data[j], v1 = data[j + LPR];
        const int c0 = indices[j], c1 = indices[j + LPR];
        a0 = fma(v0, __ldg(x + c0), a0);
        a1 = fma(v1, __ldg(x + c1), a1);
      }
      if (j < e) a0 = fma(data[j], __ldg(x + indices[j]), a0);
    }
    double acc = a0 + a1;
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (valid && sub == 0) {
      y[row] = acc;
      if (MODE == 1 || MODE == 2 || MODE == 4) dots[0] = fma(d1[row], acc, dots[0]);
      if (MODE == 3 || MODE == 5) {
        dots[0] = fma(acc, d1[row], dots[0]);
        dots[1] = fma(acc, acc, dots[1]);
      }
    }
  }
  if constexpr (MODE == 1) {
    double v[1] = {dots[0]};
    reduce_and_finalize<1>(v, S, partial, [&](double (&t)[1]) { S[S_ALPHA] = S[S_GAMMA] / t[0]; });
  } else if constexpr (MODE == 2) {
    double v[1] = {dots[0]};
    reduce_and_finalize<1>(v, S, partial, [&](double (&t)[1]) { S[S_ALPHA] = S[S_RHO_NEW] / t[0]; });
  } else if constexpr (MODE == 3) {
    reduce_and_finalize<2>(dots, S, partial, [&](double (&t)[2]) { S[S_OMEGA] = t[0] / t[1]; });
  } else if constexpr (MODE == 4) {   // distributed CG: rank-local p.Ap, all-reduced by the caller
    double v[1] = {dots[0]};
    reduce_and_finalize<1>(v, S, partial, [&](double (&t)[1]) { S[S_SUM0] = t[0]; });
  } else if constexpr (MODE == 5) {
    reduce_and_finalize<2>(dots, S, partial, [&](double (&t)[2]) { S[S_SUM2] = t[0]; S[S_SUM3] = t[1]; });
  }
}


// Same contract on the NODE-BLOCK structure of the FE matrix: the VEC scalar rows of a mesh node share one column
// pattern (VEC consecutive columns per neighbour node), so LPN lanes walk the node's neighbour list once, read ONE
// column index per VEC x VEC block (4/9 instead of 4 index bytes per nonzero for VEC = 3), gather VEC consecutive
// entries of x and stream the VEC row segments at full width.  `indices` is not read at all.
template <int VEC, int LPN, int MODE>
__global__ void __launch_bounds__(kThreads) spmv_block_fused_kernel(const RowSplit R, const int32_t* __restrict__ brow_ptr,
                                                                    const int32_t* __restrict__ bcol,
                                                                    const double* __restrict__ data,
                                                                    const double* __restrict__ x, double* __restrict__ y,
                                                                    const double* __restrict__ d1, double* __restrict__ S,
                                                                    double* __restrict__ partial) {
  if (MODE != 0 && solver_done(S)) return;
  constexpr int NPB = kThreads / LPN;
  constexpr int VV = VEC * VEC;
  const int sub = threadIdx.x % LPN;
  const int64_t n_nodes = R.a_cnt + R.b_cnt;
  const int ptotal = R.ptotal ? R.ptotal : (int)gridDim.x;
  double dots[2] = {0.0, 0.0};
  for (int64_t base = (int64_t)blockIdx.x * NPB; base < n_nodes; base += (int64_t)gridDim.x * NPB) {
    const int64_t li = base + threadIdx.x / LPN;
    const bool valid = li < n_nodes;
    const int64_t nd = li < R.a_cnt ? R.a0 + li : R.b0 + (li - R.a_cnt);
    double acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.0;
    if (valid) {
      const int e0 = brow_ptr[nd], len = brow_ptr[nd + 1] - e0;
      const double* rowp = data + (int64_t)VV * e0;
      const int32_t* colp = bcol + e0;
#pragma unroll 2
      for (int s = sub; s < len; s += LPN) {
        const int64_t m = colp[s];
        double xv[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) xv[k] = __ldg(x + VEC * m + k);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const double* d = rowp + (int64_t)i * VEC * len + VEC * s;
#pragma unroll
          for (int k = 0; k < VEC; ++k) acc[i] = fma(d[k], xv[k], acc[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i)
#pragma unroll
      for (int o = LPN / 2; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    if (valid && sub == 0) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int64_t row = VEC * nd + i;
        y[row] = acc[i];
        if (MODE == 1 || MODE == 2 || MODE == 4) dots[0] = fma(d1[row], acc[i], dots[0]);
        if (MODE == 3 || MODE == 5) {
          dots[0] = fma(acc[i], d1[row], dots[0]);
          dots[1] = fma(acc[i], acc[i], dots[1]);
        }
      }
    }
  }
  if constexpr (MODE == 1) {
    double v[1] = {dots[0]};
    reduce_and_finalize<1>(v, S, partial, [&](double (&t)[1]) { S[S_ALPHA] = S[S_GAMMA] / t[0]; }, R.poff, ptotal, R.finalize);
  } else if constexpr (MODE == 2) {
    double v[1] = {dots[0]};
    reduce_and_finalize<1>(v, S, partial, [&](double (&t)[1]) { S[S_ALPHA] = S[S_RHO_NEW] / t[0]; }, R.poff, ptotal, R.finalize);
  } else if constexpr (MODE == 3) {
    reduce_and_finalize<2>(dots, S, partial, [&](double (&t)[2]) { S[S_OMEGA] = t[0] / t[1]; }, R.poff, ptotal, R.finalize);
  } else if constexpr (MODE == 4) {
    double v[1] = {dots[0]};
    reduce_and_finalize<1>(v, S, partial, [&](double (&t)[1]) { S[S_SUM0] = t[0]; }, R.poff, ptotal, R.finalize);
  } else if constexpr (MODE == 5) {
    reduce_and_finalize<2>(dots, S, partial, [&](double (&t)[2]) { S[S_SUM2] = t[0]; S[S_SUM3] = t[1]; }, R.poff, ptotal, R.finalize);
  }
}

__device__ __forceinline__ double precond(const double* __restrict__ diag, int64_t i, double v) {
  return diag ? v * (1.0 / diag[i]) : v;      // pc = x * (1. / jacobi), solver.py:69
}

// ---------------------------------------------------------------- CG ------------------------------
// after q = A x0:  r = b - q ; z = M r ; p = z ; gamma = r.z ; rr = r.r ; bb = b.b
__global__ void __launch_bounds__(kThreads) cg_init_kernel(int64_t n, const double* __restrict__ b,
                                                           const double* __restrict__ diag, const double* __restrict__ q,
                                                           double* __restrict__ r, double* __restrict__ p,
                                                           double* __restrict__ S, double* __restrict__ partial) {
  double v[3] = {0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double bi = b[i], ri = bi - q[i];
    const double zi = precond(diag, i, ri);
    r[i] = ri;
    p[i] = zi;
    v[0] = fma(ri, zi, v[0]);
    v[1] = fma(ri, ri, v[1]);
    v[2] = fma(bi, bi, v[2]);
  }
  reduce_and_finalize<3>(v, S, partial, [&](double (&t)[3]) {
    S[S_GAMMA] = t[0];
    S[S_RR] = t[1];
    const double atol2 = fmax(S[S_TOL2] * t[2], S[S_ATOLIN2]);
    S[S_ATOL2] = atol2;
    S[S_K] = 0.0;
    S[S_DONE] = (t[1] > atol2 && 0.0 < S[S_MAXIT]) ? 0.0 : 1.0;
  });
}

// x += alpha p ; r -= alpha q ; gamma' = r.(M r) ; rr = r.r ; beta = gamma'/gamma
__global__ void __launch_bounds__(kThreads) cg_update_kernel(int64_t n, const double* __restrict__ diag,
                                                             const double* __restrict__ p, const double* __restrict__ q,
                                                             double* __restrict__ x, double* __restrict__ r,
                                                             double* __restrict__ S, double* __restrict__ partial) {
  if (solver_done(S)) return;
  const double alpha = __ldcg(S + S_ALPHA);
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    v[0] = fma(ri, precond(diag, i, ri), v[0]);
    v[1] = fma(ri, ri, v[1]);
  }
  reduce_and_finalize<2>(v, S, partial, [&](double (&t)[2]) {
    S[S_BETA] = t[0] / S[S_GAMMA];
    S[S_GAMMA] = t[0];
    S[S_RR] = t[1];
    const double k = S[S_K] + 1.0;
    S[S_K] = k;
    S[S_DONE] = (t[1] > S[S_ATOL2] && k < S[S_MAXIT]) ? 0.0 : 1.0;
  });
}

// p = M r + beta p   (skipped once converged: p is dead after the last update of x)
__global__ void __launch_bounds__(kThreads) cg_direction_kernel(int64_t n, const double* __restrict__ diag,
                                                                const double* __restrict__ r, double* __restrict__ p,
                                                                double* __restrict__ S) {
  if (solver_done(S)) return;
  const double beta = __ldcg(S + S_BETA);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    p[i] = fma(beta, p[i], precond(diag, i, r[i]));
}

// ------------------------------------------------ distributed CG (one rank's share) ---------------
// Vectors hold the rank's owned dofs first, then its ghosts; kernels touch the owned range only.  Each kernel
// leaves rank-local partial sums in S[S_SUM*]; the caller all-reduces that 4-double slice (NCCL, same stream)
// before launching the next kernel, which derives alpha / beta from the reduced values on the device.
__global__ void __launch_bounds__(kThreads) dcg_init_kernel(int64_t n, const double* __restrict__ b,
                                                            const double* __restrict__ diag, const double* __restrict__ q,
                                                            double* __restrict__ r, double* __restrict__ p,
                                                            double* __restrict__ S, double* __restrict__ partial) {
  double v[3] = {0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double bi = b[i], ri = bi - q[i];
    const double zi = precond(diag, i, ri);
    r[i] = ri;
    p[i] = zi;
    v[0] = fma(ri, zi, v[0]);
    v[1] = fma(ri, ri, v[1]);
    v[2] = fma(bi, bi, v[2]);
  }
  reduce_and_finalize<3>(v, S, partial, [&](double (&t)[3]) {
    S[S_SUM0] = t[0];
    S[S_SUM1] = t[1];
    S[S_SUM2] = t[2];
    S[S_SUM3] = 0.0;
  });
}

__global__ void dcg_scalars_init_kernel(double* __restrict__ S) {
  S[S_GAMMA] = S[S_SUM0];
  S[S_RR] = S[S_SUM1];
  const double atol2 = fmax(S[S_TOL2] * S[S_SUM2], S[S_ATOLIN2]);
  S[S_ATOL2] = atol2;
  S[S_K] = 0.0;
  S[S_DONE] = (S[S_SUM1] > atol2 && 0.0 < S[S_MAXIT]) ? 0.0 : 1.0;
}

__global__ void __launch_bounds__(kThreads) dcg_update_kernel(int64_t n, const double* __restrict__ diag,
                                                              const double* __restrict__ p, const double* __restrict__ q,
                                                              double* __restrict__ x, double* __restrict__ r,
                                                              double* __restrict__ S, double* __restrict__ partial) {
  if (solver_done(S)) return;
  const double alpha = __ldcg(S + S_GAMMA) / __ldcg(S + S_SUM0);
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    v[0] = fma(ri, precond(diag, i, ri), v[0]);
    v[1] = fma(ri, ri, v[1]);
  }
  reduce_and_finalize<2>(v, S, partial, [&](double (&t)[2]) {
    S[S_SUM1] = t[0];
    S[S_SUM2] = t[1];
  });
}

__global__ void __launch_bounds__(kThreads) dcg_direction_kernel(int64_t n, const double* __restrict__ diag,
                                                                 const double* __restrict__ r, double* __restrict__ p,
                                                                 const double* __restrict__ S) {
  if (solver_done(S)) return;
  const double beta = __ldcg(S + S_SUM1) / __ldcg(S + S_GAMMA);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    p[i] = fma(beta, p[i], precond(diag, i, r[i]));
}

__global__ void dcg_scalars_step_kernel(double* __restrict__ S) {
  if (S[S_DONE] != 0.0) return;
  S[S_GAMMA] = S[S_SUM1];
  S[S_RR] = S[S_SUM2];
  const double k = S[S_K] + 1.0;
  S[S_K] = k;
  S[S_DONE] = (S[S_SUM2] > S[S_ATOL2] && k < S[S_MAXIT]) ? 0.0 : 1.0;
}

// ------------------------------------------------------------ BiCGSTAB ----------------------------
// after q0 = A x0:  r = b - q0 ; rhat = p = q = r ; rho = alpha = omega = 1
__global__ void __launch_bounds__(kThreads) bicg_init_kernel(int64_t n, const double* __restrict__ b,
                                                             const double* __restrict__ ax, double* __restrict__ r,
                                                             double* __restrict__ rhat, double* __restrict__ p,
                                                             double* __restrict__ q, double* __restrict__ S,
                                                             double* __restrict__ partial) {
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double bi = b[i], ri = bi - ax[i];
    r[i] = ri;
    rhat[i] = ri;
    p[i] = ri;
    q[i] = ri;
    v[0] = fma(ri, ri, v[0]);
    v[1] = fma(bi, bi, v[1]);
  }
  reduce_and_finalize<2>(v, S, partial, [&](double (&t)[2]) {
    S[S_GAMMA] = 1.0;   // rho
    S[S_ALPHA] = 1.0;
    S[S_OMEGA] = 1.0;
    S[S_RR] = t[0];
    S[S_RHO_NEW] = t[0];                       // rhat.r with rhat = r
    S[S_BETA] = t[0] / 1.0 * 1.0 / 1.0;        // rho_/rho * alpha/omega
    const double atol2 = fmax(S[S_TOL2] * t[1], S[S_ATOLIN2]);
    S[S_ATOL2] = atol2;
    S[S_K] = 0.0;
    S[S_DONE] = (t[0] > atol2 && 0.0 < S[S_MAXIT]) ? 0.0 : 1.0;
  });
}

// p = r + beta (p - omega q) ; phat = M p
__global__ void __launch_bounds__(kThreads) bicg_p_kernel(int64_t n, const double* __restrict__ diag,
                                                          const double* __restrict__ r, const double* __restrict__ q,
                                                          double* __restrict__ p, double* __restrict__ phat,
                                                          const double* __restrict__ S) {
  if (solver_done(S)) return;
  const double beta = __ldcg(S + S_BETA), omega = __ldcg(S + S_OMEGA);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double pi = r[i] + beta * (p[i] - omega * q[i]);
    p[i] = pi;
    phat[i] = precond(diag, i, pi);
  }
}

// s = r - alpha q ; shat = M s ; ss = s.s -> exit_early
__global__ void __launch_bounds__(kThreads) bicg_s_kernel(int64_t n, const double* __restrict__ diag,
                                                          const double* __restrict__ r, const double* __restrict__ q,
                                                          double* __restrict__ s, double* __restrict__ shat,
                                                          double* __restrict__ S, double* __restrict__ partial) {
  if (solver_done(S)) return;
  const double alpha = __ldcg(S + S_ALPHA);
  double v[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double si = r[i] - alpha * q[i];
    s[i] = si;
    shat[i] = precond(diag, i, si);
    v[0] = fma(si, si, v[0]);
  }
  reduce_and_finalize<1>(v, S, partial, [&](double (&t)[1]) { S[S_EARLY] = (t[0] < S[S_ATOL2]) ? 1.0 : 0.0; });
}

// x += alpha phat (+ omega shat) ; r = s (- omega t) ; rr, rhat.r ; bookkeeping of the JAX body
__global__ void __launch_bounds__(kThreads) bicg_x_kernel(int64_t n, const double* __restrict__ phat,
                                                          const double* __restrict__ shat, const double* __restrict__ s,
                                                          const double* __restrict__ t, const double* __restrict__ rhat,
                                                          double* __restrict__ x, double* __restrict__ r,
                                                          double* __restrict__ S, double* __restrict__ partial) {
  if (solver_done(S)) return;
  const double alpha = __ldcg(S + S_ALPHA), omega = __ldcg(S + S_OMEGA);
  const bool early = __ldcg(S + S_EARLY) != 0.0;
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    double xi, ri;
    if (early) {
      xi = x[i] + alpha * phat[i];
      ri = s[i];
    } else {
      xi = x[i] + (alpha * phat[i] + omega * shat[i]);
      ri = s[i] - omega * t[i];
    }
    x[i] = xi;
    r[i] = ri;
    v[0] = fma(ri, ri, v[0]);
    v[1] = fma(rhat[i], ri, v[1]);
  }
  reduce_and_finalize<2>(v, S, partial, [&](double (&tt)[2]) {
    const double rho_ = S[S_RHO_NEW], al = S[S_ALPHA], om = S[S_OMEGA];
    double k = (om == 0.0 || al == 0.0) ? -11.0 : S[S_K] + 1.0;
    if (rho_ == 0.0) k = -10.0;
    S[S_K] = k;
    S[S_GAMMA] = rho_;
    S[S_RR] = tt[0];
    S[S_RHO_NEW] = tt[1];
    S[S_BETA] = tt[1] / rho_ * al / om;
    S[S_DONE] = (tt[0] > S[S_ATOL2] && k < S[S_MAXIT] && k >= 0.0) ? 0.0 : 1.0;
  });
}

// ------------------------------------------- distributed BiCGSTAB (one rank's share) -------------
// Same recurrences as above (jax.scipy.sparse.linalg._bicgstab_solve) on the rank's owned dofs.  Every kernel leaves
// rank-local partial sums in S[S_SUM*]; the driver all-reduces the slots of that phase (ncclAllReduce, same stream) and
// the next kernel derives alpha / omega from the reduced values on the device:
//   SUM0 = rhat.q, SUM1 = s.s, SUM2 = t.s, SUM3 = t.t, SUM4 = r.r, SUM5 = rhat.r (b.b at start-up).
__global__ void __launch_bounds__(kThreads) dbicg_init_kernel(int64_t n, const double* __restrict__ b,
                                                              const double* __restrict__ ax, double* __restrict__ r,
                                                              double* __restrict__ rhat, double* __restrict__ p,
                                                              double* __restrict__ q, double* __restrict__ S,
                                                              double* __restrict__ partial) {
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double bi = b[i], ri = bi - ax[i];
    r[i] = ri;
    rhat[i] = ri;
    p[i] = ri;
    q[i] = ri;
    v[0] = fma(ri, ri, v[0]);
    v[1] = fma(bi, bi, v[1]);
  }
  reduce_and_finalize<2>(v, S, partial, [&](double (&t)[2]) { S[S_SUM4] = t[0]; S[S_SUM5] = t[1]; });
}

__global__ void dbicg_scalars_init_kernel(double* __restrict__ S) {
  const double rr = S[S_SUM4];
  S[S_GAMMA] = 1.0;
  S[S_ALPHA] = 1.0;
  S[S_OMEGA] = 1.0;
  S[S_RR] = rr;
  S[S_RHO_NEW] = rr;
  S[S_BETA] = rr;
  const double atol2 = fmax(S[S_TOL2] * S[S_SUM5], S[S_ATOLIN2]);
  S[S_ATOL2] = atol2;
  S[S_K] = 0.0;
  S[S_DONE] = (rr > atol2 && 0.0 < S[S_MAXIT]) ? 0.0 : 1.0;
}

__global__ void __launch_bounds__(kThreads) dbicg_s_kernel(int64_t n, const double* __restrict__ diag,
                                                           const double* __restrict__ r, const double* __restrict__ q,
                                                           double* __restrict__ s, double* __restrict__ shat,
                                                           double* __restrict__ S, double* __restrict__ partial) {
  if (solver_done(S)) return;
  const double alpha = __ldcg(S + S_RHO_NEW) / __ldcg(S + S_SUM0);
  double v[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double si = r[i] - alpha * q[i];
    s[i] = si;
    shat[i] = precond(diag, i, si);
    v[0] = fma(si, si, v[0]);
  }
  reduce_and_finalize<1>(v, S, partial, [&](double (&t)[1]) { S[S_SUM1] = t[0]; });
}

__global__ void __launch_bounds__(kThreads) dbicg_x_kernel(int64_t n, const double* __restrict__ phat,
                                                           const double* __restrict__ shat, const double* __restrict__ s,
                                                           const double* __restrict__ t, const double* __restrict__ rhat,
                                                           double* __restrict__ x, double* __restrict__ r,
                                                           double* __restrict__ S, double* __restrict__ partial) {
  if (solver_done(S)) return;
  const double alpha = __ldcg(S + S_RHO_NEW) / __ldcg(S + S_SUM0);
  const double omega = __ldcg(S + S_SUM2) / __ldcg(S + S_SUM3);
  const bool early = __ldcg(S + S_SUM1) < __ldcg(S + S_ATOL2);
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    double xi, ri;
    if (early) {
      xi = x[i] + alpha * phat[i];
      ri = s[i];
    } else {
      xi = x[i] + (alpha * phat[i] + omega * shat[i]);
      ri = s[i] - omega * t[i];
    }
    x[i] = xi;
    r[i] = ri;
    v[0] = fma(ri, ri, v[0]);
    v[1] = fma(rhat[i], ri, v[1]);
  }
  reduce_and_finalize<2>(v, S, partial, [&](double (&tt)[2]) { S[S_SUM4] = tt[0]; S[S_SUM5] = tt[1]; });
}

__global__ void dbicg_scalars_step_kernel(double* __restrict__ S) {
  if (S[S_DONE] != 0.0) return;
  const double rho_ = S[S_RHO_NEW], al = rho_ / S[S_SUM0], om = S[S_SUM2] / S[S_SUM3];
  double k = (om == 0.0 || al == 0.0) ? -11.0 : S[S_K] + 1.0;
  if (rho_ == 0.0) k = -10.0;
  S[S_K] = k;
  S[S_GAMMA] = rho_;
  S[S_ALPHA] = al;
  S[S_OMEGA] = om;
  S[S_RR] = S[S_SUM4];
  S[S_RHO_NEW] = S[S_SUM5];
  S[S_BETA] = S[S_SUM5] / rho_ * al / om;
  S[S_DONE] = (S[S_SUM4] > S[S_ATOL2] && k < S[S_MAXIT] && k >= 0.0) ? 0.0 : 1.0;
}

// rank-local sum of (A x - b)^2 over the owned rows -> S[S_SUM0]
__global__ void __launch_bounds__(kThreads) dresnorm_kernel(int64_t n, const double* __restrict__ ax,
                                                            const double* __restrict__ b, double* __restrict__ S,
                                                            double* __restrict__ partial) {
  double v[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double d = ax[i] - b[i];
    v[0] = fma(d, d, v[0]);
  }
  reduce_and_finalize<1>(v, S, partial, [&](double (&t)[1]) { S[S_SUM0] = t[0]; });
}

// ||A x - b|| for the reference's post-solve check (solver.py:87)
__global__ void __launch_bounds__(kThreads) resnorm_kernel(int64_t n, const double* __restrict__ ax,
                                                           const double* __restrict__ b, double* __restrict__ S,
                                                           double* __restrict__ partial) {
  double v[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double d = ax[i] - b[i];
    v[0] = fma(d, d, v[0]);
  }
  reduce_and_finalize<1>(v, S, partial, [&](double (&t)[1]) { S[S_TMP] = sqrt(t[0]); });
}

constexpr int kLPR = 8;

// Persistent grid = resident CTAs per SM (from the occupancy API) x SM count: exactly one wave.
template <class K>
int persistent_grid(K kernel) {
  static int cached[64] = {0};   // per device (a process may drive several GPUs)
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = cached[dev & 63];
  if (!c) {
    int per_sm = 0, sms = kNumSM;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    c = per_sm * sms;
    if (c > kBlocks) c = kBlocks;
  }
  return c;
}
#define FEM_PGRID(kernel) persistent_grid(kernel), kThreads, 0, st

// CSR operand; with vec > 1 and the node-block graph (brow_ptr, bcol) the block kernel is used and `indices` is unread.
struct Mat {
  int64_t n;
  const int32_t* indptr;
  const int32_t* indices;
  const double* data;
  int vec;
  const int32_t* brow_ptr;
  const int32_t* bcol;
};

inline RowSplit whole_rows(int64_t n_nodes) { return RowSplit{0, n_nodes, 0, 0, 0, 0, 1}; }

template <int MODE>
void spmv_fused(const Mat& A, const double* x, double* y, const double* d1, const Ws& w, cudaStream_t st) {
  constexpr int LPN = 8;
  if (A.brow_ptr && A.bcol && A.vec == 3) {
    spmv_block_fused_kernel<3, LPN, MODE><<<FEM_PGRID((spmv_block_fused_kernel<3, LPN, MODE>))>>>(
        whole_rows(A.n / 3), A.brow_ptr, A.bcol, A.data, x, y, d1, w.s, w.partial);
  } else if (A.brow_ptr && A.bcol && A.vec == 2) {
    spmv_block_fused_kernel<2, LPN, MODE><<<FEM_PGRID((spmv_block_fused_kernel<2, LPN, MODE>))>>>(
        whole_rows(A.n / 2), A.brow_ptr, A.bcol, A.data, x, y, d1, w.s, w.partial);
  } else {
    spmv_fused_kernel<kLPR, MODE><<<FEM_PGRID((spmv_fused_kernel<kLPR, MODE>))>>>(A.n, A.indptr, A.indices, A.data, x, y,
                                                                                 d1, w.s, w.partial);
  }
}

// y[owned] = (A x)[owned] with x's ghosts refreshed first.  When the plan knows a range of owned nodes without ghost
// neighbours (fem_halo_set_interior) and the matrix has its node-block structure, those rows are multiplied while the sends
// and receives are in flight on the plan's stream; the remaining rows follow once the ghosts have arrived.
template <int MODE, int VEC>
int dist_spmv_split(const HaloPlan* h, const Mat& A, double* x, double* y, const double* d1, const Ws& w, cudaStream_t st) {
  constexpr int LPN = 8, NPB = kThreads / LPN;
  auto k = spmv_block_fused_kernel<VEC, LPN, MODE>;
  const int64_t n_nodes = A.n / VEC, lo = h->int_lo, hi = h->int_hi;
  const int cap = persistent_grid(k);
  auto blocks = [&](int64_t cnt) { return (int)std::max<int64_t>(1, std::min<int64_t>(cap, (cnt + NPB - 1) / NPB)); };
  const int g1 = blocks(hi - lo), g2 = blocks(n_nodes - (hi - lo));
  // peer-memory exchange when the plan has it (stores into the neighbours' mailboxes, flags, wait + unpack: no NCCL call, one
  // stream); otherwise ncclSend / ncclRecv on the plan's second stream
  if (int e = h->p2p_ready ? halo_p2p_send(h, x, st) : halo_begin(h, x, st)) return e;
  k<<<g1, kThreads, 0, st>>>(RowSplit{lo, hi - lo, 0, 0, 0, g1 + g2, 0}, A.brow_ptr, A.bcol, A.data, x, y, d1, w.s, w.partial);
  if (int e = h->p2p_ready ? halo_p2p_receive(h, x, st) : halo_end(h, st)) return e;
  k<<<g2, kThreads, 0, st>>>(RowSplit{0, lo, hi, n_nodes - hi, g1, g1 + g2, 1}, A.brow_ptr, A.bcol, A.data, x, y, d1, w.s, w.partial);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

template <int MODE>
int dist_spmv(const HaloPlan* h, const Mat& A, double* x, double* y, const double* d1, const Ws& w, cudaStream_t st) {
  const int64_t n_nodes = A.vec > 0 ? A.n / A.vec : 0;
  const bool split = h->n_nb > 0 && A.brow_ptr && A.bcol && h->int_hi > h->int_lo && h->int_hi <= n_nodes &&
                     (h->int_hi - h->int_lo) * 2 > n_nodes;          // worth two launches only if most rows are interior
  if (split && A.vec == 3) return dist_spmv_split<MODE, 3>(h, A, x, y, d1, w, st);
  if (split && A.vec == 2) return dist_spmv_split<MODE, 2>(h, A, x, y, d1, w, st);
  if (h->p2p_ready) {
    if (int e = halo_p2p_send(h, x, st)) return e;
    if (int e = halo_p2p_receive(h, x, st)) return e;
  } else if (int e = halo_exchange(h, x, st)) {
    return e;
  }
  spmv_fused<MODE>(A, x, y, d1, w, st);
  return FEM_OK;
}

int init_scalars(const Ws& w, double tol, double atol, int maxiter, cudaStream_t st) {
  double h[kScalars] = {0};
  h[S_TOL2] = tol * tol;
  h[S_ATOLIN2] = atol * atol;
  h[S_MAXIT] = (double)maxiter;
  FEM_CUDA_CHECK(cudaMemcpyAsync(w.s, h, sizeof(h), cudaMemcpyHostToDevice, st));
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));   // h is a stack buffer
  return FEM_OK;
}

int finish(const Mat& A, const double* b, const double* x, const Ws& w, double* scratch, double* info_host,
           cudaStream_t st) {
  const int64_t n = A.n;
  double h[kScalars];
  FEM_CUDA_CHECK(cudaMemcpyAsync(h, w.s, sizeof(h), cudaMemcpyDeviceToHost, st));
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  info_host[0] = h[S_K];
  info_host[1] = h[S_RR];
  spmv_fused<0>(A, x, scratch, nullptr, w, st);
  resnorm_kernel<<<FEM_PGRID(resnorm_kernel)>>>(n, scratch, b, w.s, w.partial);
  FEM_LAUNCH_CHECK();
  FEM_CUDA_CHECK(cudaMemcpyAsync(h, w.s, sizeof(h), cudaMemcpyDeviceToHost, st));
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  info_host[2] = h[S_TMP];
  return FEM_OK;
}

int poll_done(const Ws& w, bool* done, cudaStream_t st) {
  double flag = 0.0;
  FEM_CUDA_CHECK(cudaMemcpyAsync(&flag, w.s + S_DONE, sizeof(double), cudaMemcpyDeviceToHost, st));
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  *done = flag != 0.0;
  return FEM_OK;
}

}  // namespace
}  // namespace femb200

namespace femb200 {
// plain y = A x on the node-block structure (non-persistent: one node row block per LPN lanes)
int launch_spmv_block(int64_t n, int vec, const int32_t* brow_ptr, const int32_t* bcol, const double* data,
                      const double* x, double* y, cudaStream_t st) {
  constexpr int LPN = 8;
  const int64_t n_nodes = n / vec;
  const unsigned grid = (unsigned)((n_nodes * LPN + kThreads - 1) / kThreads);
  if (vec == 3)
    spmv_block_fused_kernel<3, LPN, 0><<<grid, kThreads, 0, st>>>(whole_rows(n_nodes), brow_ptr, bcol, data, x, y, nullptr, nullptr, nullptr);
  else
    spmv_block_fused_kernel<2, LPN, 0><<<grid, kThreads, 0, st>>>(whole_rows(n_nodes), brow_ptr, bcol, data, x, y, nullptr, nullptr, nullptr);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}
}  // namespace femb200

using namespace femb200;

extern "C" int64_t fem_krylov_workspace(int64_t n) {
  return kScalars + (int64_t)kBlocks * kMaxDots + 8 * pad_n(n);
}

extern "C" int fem_pcg(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data, int vec,
                       const int32_t* brow_ptr, const int32_t* bcol, const double* diag, const double* b, double* x,
                       double tol, double atol, int maxiter, int check_every, double* workspace, double* info_host,
                       void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(indptr && indices && data && b && x && workspace && info_host, "null pointer");
  FEM_REQUIRE(n > 0 && maxiter >= 0, "bad size");
  cudaStream_t st = (cudaStream_t)stream;
  if (check_every <= 0) check_every = 25;
  const Ws w = carve(workspace, n);
  const Mat A{n, indptr, indices, data, vec, brow_ptr, bcol};
  double *r = w.v[0], *p = w.v[1], *q = w.v[2];
  if (int e = init_scalars(w, tol, atol, maxiter, st)) return e;
  spmv_fused<0>(A, x, q, nullptr, w, st);
  cg_init_kernel<<<FEM_PGRID(cg_init_kernel)>>>(n, b, diag, q, r, p, w.s, w.partial);
  FEM_LAUNCH_CHECK();
  bool done = false;
  if (int e = poll_done(w, &done, st)) return e;
  for (int it = 0; !done && it < maxiter; it += check_every) {
    for (int j = 0; j < check_every; ++j) {
      spmv_fused<1>(A, p, q, p, w, st);                                        // q = A p ; alpha
      cg_update_kernel<<<FEM_PGRID(cg_update_kernel)>>>(n, diag, p, q, x, r, w.s, w.partial);
      cg_direction_kernel<<<FEM_PGRID(cg_direction_kernel)>>>(n, diag, r, p, w.s);
    }
    FEM_LAUNCH_CHECK();
    if (int e = poll_done(w, &done, st)) return e;
  }
  return finish(A, b, x, w, q, info_host, st);
}

extern "C" int fem_pbicgstab(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data, int vec,
                             const int32_t* brow_ptr, const int32_t* bcol, const double* diag, const double* b, double* x,
                             double tol, double atol, int maxiter, int check_every, double* workspace,
                             double* info_host, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(indptr && indices && data && b && x && workspace && info_host, "null pointer");
  FEM_REQUIRE(n > 0 && maxiter >= 0, "bad size");
  cudaStream_t st = (cudaStream_t)stream;
  if (check_every <= 0) check_every = 25;
  const Ws w = carve(workspace, n);
  const Mat A{n, indptr, indices, data, vec, brow_ptr, bcol};
  double *r = w.v[0], *rhat = w.v[1], *p = w.v[2], *q = w.v[3], *s = w.v[4], *t = w.v[5], *phat = w.v[6],
         *shat = w.v[7];
  if (int e = init_scalars(w, tol, atol, maxiter, st)) return e;
  spmv_fused<0>(A, x, t, nullptr, w, st);
  bicg_init_kernel<<<FEM_PGRID(bicg_init_kernel)>>>(n, b, t, r, rhat, p, q, w.s, w.partial);
  FEM_LAUNCH_CHECK();
  bool done = false;
  if (int e = poll_done(w, &done, st)) return e;
  for (int it = 0; !done && it < maxiter; it += check_every) {
    for (int j = 0; j < check_every; ++j) {
      bicg_p_kernel<<<FEM_PGRID(bicg_p_kernel)>>>(n, diag, r, q, p, phat, w.s);
      spmv_fused<2>(A, phat, q, rhat, w, st);                                  // q = A phat ; alpha
      bicg_s_kernel<<<FEM_PGRID(bicg_s_kernel)>>>(n, diag, r, q, s, shat, w.s, w.partial);
      spmv_fused<3>(A, shat, t, s, w, st);                                     // t = A shat ; omega
      bicg_x_kernel<<<FEM_PGRID(bicg_x_kernel)>>>(n, phat, shat, s, t, rhat, x, r, w.s, w.partial);
    }
    FEM_LAUNCH_CHECK();
    if (int e = poll_done(w, &done, st)) return e;
  }
  return finish(A, b, x, w, t, info_host, st);
}

// ---- distributed CG building blocks (the caller all-reduces workspace[16..20) between them) -------------
extern "C" int fem_dcg_begin(double* workspace, double tol, double atol, int maxiter, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(workspace, "null pointer");
  Ws w;
  w.s = workspace;
  return init_scalars(w, tol, atol, maxiter, (cudaStream_t)stream);
}

extern "C" int fem_dcg_spmv_dot(int64_t n_owned, int64_t n_local, const int32_t* indptr, const int32_t* indices,
                                const double* data, int vec, const int32_t* brow_ptr, const int32_t* bcol,
                                const double* p, double* q, int with_dot, double* workspace, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(indptr && indices && data && p && q && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const Ws w = carve(workspace, n_local);
  const Mat A{n_owned, indptr, indices, data, vec, brow_ptr, bcol};
  if (with_dot) spmv_fused<4>(A, p, q, p, w, st);
  else spmv_fused<0>(A, p, q, nullptr, w, st);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

extern "C" int fem_dcg_init(int64_t n_owned, int64_t n_local, const double* b, const double* diag, const double* q,
                            double* r, double* p, double* workspace, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(b && q && r && p && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const Ws w = carve(workspace, n_local);
  dcg_init_kernel<<<FEM_PGRID(dcg_init_kernel)>>>(n_owned, b, diag, q, r, p, w.s, w.partial);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

extern "C" int fem_dcg_scalars(int phase, double* workspace, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (phase == 0) dcg_scalars_init_kernel<<<1, 1, 0, st>>>(workspace);
  else dcg_scalars_step_kernel<<<1, 1, 0, st>>>(workspace);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

extern "C" int fem_dcg_update(int64_t n_owned, int64_t n_local, const double* diag, const double* p, const double* q,
                              double* x, double* r, double* workspace, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(p && q && x && r && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const Ws w = carve(workspace, n_local);
  dcg_update_kernel<<<FEM_PGRID(dcg_update_kernel)>>>(n_owned, diag, p, q, x, r, w.s, w.partial);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

extern "C" int fem_dcg_direction(int64_t n_owned, const double* diag, const double* r, double* p, double* workspace,
                                 void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(r && p && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  dcg_direction_kernel<<<FEM_PGRID(dcg_direction_kernel)>>>(n_owned, diag, r, p, workspace);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

// ---- distributed Krylov drivers: the whole loop (kernels, halo exchange, all-reduces) issued from C on one stream ----
namespace femb200 {
namespace {
int dist_finish(const HaloPlan* h, const Mat& A, const double* b, double* x, const Ws& w, double* scratch,
                double* info_host, cudaStream_t st) {
  double hs[kScalars];
  FEM_CUDA_CHECK(cudaMemcpyAsync(hs, w.s, sizeof(hs), cudaMemcpyDeviceToHost, st));
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  info_host[0] = hs[S_K];
  info_host[1] = hs[S_RR];
  if (int e = halo_p2p_status(h, st)) return e;
  if (int e = halo_exchange(h, x, st)) return e;                 // the caller gets up-to-date ghosts
  spmv_fused<0>(A, x, scratch, nullptr, w, st);
  dresnorm_kernel<<<FEM_PGRID(dresnorm_kernel)>>>(A.n, scratch, b, w.s, w.partial);
  FEM_LAUNCH_CHECK();
  if (int e = allreduce_sum(h, w.s + S_SUM0, 1, st)) return e;
  FEM_CUDA_CHECK(cudaMemcpyAsync(hs, w.s, sizeof(hs), cudaMemcpyDeviceToHost, st));
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  info_host[2] = sqrt(hs[S_SUM0]);
  return FEM_OK;
}
}  // namespace
}  // namespace femb200

extern "C" int fem_dist_pcg(void* halo, int64_t n_owned, int64_t n_local, const int32_t* indptr, const int32_t* indices,
                            const double* data, int vec, const int32_t* brow_ptr, const int32_t* bcol, const double* diag,
                            const double* b, double* x, double tol, double atol, int maxiter, int check_every,
                            double* workspace, double* info_host, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(halo && indptr && indices && data && b && x && workspace && info_host, "null pointer");
  FEM_REQUIRE(n_owned > 0 && n_local >= n_owned && maxiter >= 0, "bad size");
  const HaloPlan* h = reinterpret_cast<const HaloPlan*>(halo);
  cudaStream_t st = (cudaStream_t)stream;
  if (check_every <= 0) check_every = 25;
  const Ws w = carve(workspace, n_local);
  const Mat A{n_owned, indptr, indices, data, vec, brow_ptr, bcol};
  double *r = w.v[0], *p = w.v[1], *q = w.v[2];
  if (int e = init_scalars(w, tol, atol, maxiter, st)) return e;
  if (int e = halo_exchange(h, x, st)) return e;
  spmv_fused<0>(A, x, q, nullptr, w, st);
  dcg_init_kernel<<<FEM_PGRID(dcg_init_kernel)>>>(n_owned, b, diag, q, r, p, w.s, w.partial);
  FEM_LAUNCH_CHECK();
  if (int e = allreduce_sum(h, w.s + S_SUM0, 3, st)) return e;
  dcg_scalars_init_kernel<<<1, 1, 0, st>>>(w.s);
  bool done = false;
  if (int e = poll_done(w, &done, st)) return e;
  for (int it = 0; !done && it < maxiter; it += check_every) {
    for (int j = 0; j < check_every; ++j) {
      if (int e = dist_spmv<4>(h, A, p, q, p, w, st)) return e;                // ghosts of p ; q = A p ; rank-local p.Ap
      if (int e = allreduce_sum(h, w.s + S_SUM0, 1, st)) return e;
      dcg_update_kernel<<<FEM_PGRID(dcg_update_kernel)>>>(n_owned, diag, p, q, x, r, w.s, w.partial);
      if (int e = allreduce_sum(h, w.s + S_SUM1, 2, st)) return e;
      dcg_direction_kernel<<<FEM_PGRID(dcg_direction_kernel)>>>(n_owned, diag, r, p, w.s);
      dcg_scalars_step_kernel<<<1, 1, 0, st>>>(w.s);
    }
    FEM_LAUNCH_CHECK();
    if (int e = poll_done(w, &done, st)) return e;
  }
  return dist_finish(h, A, b, x, w, q, info_host, st);
}

extern "C" int fem_dist_pbicgstab(void* halo, int64_t n_owned, int64_t n_local, const int32_t* indptr,
                                  const int32_t* indices, const double* data, int vec, const int32_t* brow_ptr,
                                  const int32_t* bcol, const double* diag, const double* b, double* x, double tol,
                                  double atol, int maxiter, int check_every, double* workspace, double* info_host,
                                  void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(halo && indptr && indices && data && b && x && workspace && info_host, "null pointer");
  FEM_REQUIRE(n_owned > 0 && n_local >= n_owned && maxiter >= 0, "bad size");
  const HaloPlan* h = reinterpret_cast<const HaloPlan*>(halo);
  cudaStream_t st = (cudaStream_t)stream;
  if (check_every <= 0) check_every = 25;
  const Ws w = carve(workspace, n_local);
  const Mat A{n_owned, indptr, indices, data, vec, brow_ptr, bcol};
  double *r = w.v[0], *rhat = w.v[1], *p = w.v[2], *q = w.v[3], *s = w.v[4], *t = w.v[5], *phat = w.v[6],
         *shat = w.v[7];
  if (int e = init_scalars(w, tol, atol, maxiter, st)) return e;
  if (int e = halo_exchange(h, x, st)) return e;
  spmv_fused<0>(A, x, t, nullptr, w, st);
  dbicg_init_kernel<<<FEM_PGRID(dbicg_init_kernel)>>>(n_owned, b, t, r, rhat, p, q, w.s, w.partial);
  FEM_LAUNCH_CHECK();
  if (int e = allreduce_sum(h, w.s + S_SUM4, 2, st)) return e;
  dbicg_scalars_init_kernel<<<1, 1, 0, st>>>(w.s);
  bool done = false;
  if (int e = poll_done(w, &done, st)) return e;
  for (int it = 0; !done && it < maxiter; it += check_every) {
    for (int j = 0; j < check_every; ++j) {
      bicg_p_kernel<<<FEM_PGRID(bicg_p_kernel)>>>(n_owned, diag, r, q, p, phat, w.s);
      if (int e = dist_spmv<4>(h, A, phat, q, rhat, w, st)) return e;          // ghosts of phat ; q = A phat ; rank-local rhat.q
      if (int e = allreduce_sum(h, w.s + S_SUM0, 1, st)) return e;
      dbicg_s_kernel<<<FEM_PGRID(dbicg_s_kernel)>>>(n_owned, diag, r, q, s, shat, w.s, w.partial);
      if (int e = allreduce_sum(h, w.s + S_SUM1, 1, st)) return e;
      if (int e = dist_spmv<5>(h, A, shat, t, s, w, st)) return e;             // ghosts of shat ; t = A shat ; rank-local t.s, t.t
      if (int e = allreduce_sum(h, w.s + S_SUM2, 2, st)) return e;
      dbicg_x_kernel<<<FEM_PGRID(dbicg_x_kernel)>>>(n_owned, phat, shat, s, t, rhat, x, r, w.s, w.partial);
      if (int e = allreduce_sum(h, w.s + S_SUM4, 2, st)) return e;
      dbicg_scalars_step_kernel<<<1, 1, 0, st>>>(w.s);
    }
    FEM_LAUNCH_CHECK();
    if (int e = poll_done(w, &done, st)) return e;
  }
  return dist_finish(h, A, b, x, w, t, info_host, st);
}
