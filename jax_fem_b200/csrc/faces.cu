// Solution-dependent surface maps (Robin-type boundary terms): face residual and face tangent.
//
// Replaces get_surface_kernel (jax_fem/problem.py:238-259) and the face part of pre_jit_fns / value_and_jacfwd
// (problem.py:289-325) for the REGISTERED surface law  val_i(u) = coef_i (u_i - uref_i)^power  (power 1: convection / spring
// foundation; power 2, coef 5: the map of applications/robin_bc/example.py:59-67).  u-independent maps stay a constant load
// vector (jax_fem_b200/problem.py::_assemble_loads).  Boundary faces are O(N^2) against O(N^3) cells, so these kernels are
// small post-passes on the nodal residual and on the assembled CSR values; they keep the assembly's guarantees: every
// output is written by exactly one thread, contributions are added in ascending face order (bit-reproducible, no atomics).
//
// Geometry (Nanson scale x face quadrature weight, fe.py:180-182) does not depend on u and is precomputed per face and
// quadrature point on the host.  Tables: face_vals (local faces, FQ, NN) = shape values of the cell's nodes on its local faces
// (basis.py:178-250); per boundary node: its incident (face, local node) pairs in ascending face order.
#include "common.cuh"

namespace femb200 {
namespace {

struct FaceArgs {
  int nn, fq;
  int64_t n_bnodes;
  const int32_t* bnode;      // boundary nodes of the set
  const int32_t* bf_ptr;     // (n_bnodes + 1)
  const int32_t* bf_face;    // incident faces (index into the set), ascending
  const int32_t* bf_local;   // local node of the boundary node in that face's cell
  const int32_t* face_cell;  // (F)
  const int32_t* face_lid;   // (F) local face of the cell
  const double* nanson;      // (F, FQ)
  const double* fvals;       // (local faces, FQ, NN)
  const int32_t* cells;      // (C, NN)
  const double* sol;         // (nodes, VEC)
  double coef[3], uref[3], power;
};

__device__ __forceinline__ double powf64(double x, double p, bool derivative) {
  if (!derivative) {
    if (p == 1.0) return x;
    if (p == 2.0) return x * x;
    return pow(x, p);
  }
  if (p == 1.0) return 1.0;
  if (p == 2.0) return 2.0 * x;
  return p * pow(x, p - 1.0);
}

// u at quadrature point q of face f: sum_b N_b(q) u_b over the cell's nodes
template <int VEC>
__device__ __forceinline__ void face_u(const FaceArgs& A, int c, const double* N, double (&u)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) u[i] = 0.0;
  for (int b = 0; b < A.nn; ++b) {
    const double Nb = N[b];
    if (Nb == 0.0) continue;
    const int64_t node = A.cells[(int64_t)c * A.nn + b];
#pragma unroll
    for (int i = 0; i < VEC; ++i) u[i] = fma(Nb, A.sol[node * VEC + i], u[i]);
  }
}

template <int VEC>
__global__ void face_residual_kernel(const FaceArgs A, double* __restrict__ res) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.n_bnodes) return;
  double acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.0;
  for (int p = A.bf_ptr[t]; p < A.bf_ptr[t + 1]; ++p) {
    const int f = A.bf_face[p], a = A.bf_local[p], c = A.face_cell[f];
    const double* Nf = A.fvals + (int64_t)A.face_lid[f] * A.fq * A.nn;
    for (int q = 0; q < A.fq; ++q) {
      const double* N = Nf + q * A.nn;
      double u[VEC];
      face_u<VEC>(A, c, N, u);
      const double w = N[a] * A.nanson[(int64_t)f * A.fq + q];
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = fma(A.coef[i] * powf64(u[i] - A.uref[i], A.power, false), w, acc[i]);   // problem.py:254-256
    }
  }
  const int64_t n = A.bnode[t];
#pragma unroll
  for (int i = 0; i < VEC; ++i) res[n * VEC + i] += acc[i];
}

// one warp per boundary node; lanes = block entries (n, m) of the node's CSR rows
template <int VEC>
__global__ void face_tangent_kernel(const FaceArgs A, const int32_t* __restrict__ brow_ptr, const int32_t* __restrict__ bcol,
                                    const uint8_t* __restrict__ bc_flag, double* __restrict__ data) {
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= A.n_bnodes) return;
  const int l = threadIdx.x & 31;
  const int64_t n = A.bnode[t];
  const int e0 = brow_ptr[n], len = brow_ptr[n + 1] - e0;
  for (int s = l; s < len; s += 32) {
    const int m = bcol[e0 + s];
    double K[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) K[i] = 0.0;
    for (int p = A.bf_ptr[t]; p < A.bf_ptr[t + 1]; ++p) {
      const int f = A.bf_face[p], a = A.bf_local[p], c = A.face_cell[f];
      int b = -1;
      for (int j = 0; j < A.nn; ++j)
        if (A.cells[(int64_t)c * A.nn + j] == m) b = j;
      if (b < 0) continue;
      const double* Nf = A.fvals + (int64_t)A.face_lid[f] * A.fq * A.nn;
      for (int q = 0; q < A.fq; ++q) {
        const double* N = Nf + q * A.nn;
        const double w = N[a] * N[b] * A.nanson[(int64_t)f * A.fq + q];
        if (w == 0.0) continue;
        double u[VEC];
        face_u<VEC>(A, c, N, u);
#pragma unroll
        for (int i = 0; i < VEC; ++i) K[i] = fma(A.coef[i] * powf64(u[i] - A.uref[i], A.power, true), w, K[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      if (bc_flag && bc_flag[n * VEC + i]) continue;          // Dirichlet rows stay unit rows (zeroRows, solver.py:527-528)
      data[(int64_t)VEC * VEC * e0 + (int64_t)i * VEC * len + VEC * s + i] += K[i];
    }
  }
}

int fill(FaceArgs& A, int nn, int fq, int64_t n_bnodes, const int32_t* bnode, const int32_t* bf_ptr, const int32_t* bf_face,
         const int32_t* bf_local, const int32_t* face_cell, const int32_t* face_lid, const double* nanson, const double* fvals,
         const int32_t* cells, const double* sol, int vec, const double* law_host) {
  FEM_REQUIRE(bnode && bf_ptr && bf_face && bf_local && face_cell && face_lid && nanson && fvals && cells && sol && law_host,
              "null pointer");
  FEM_REQUIRE(vec >= 1 && vec <= 3 && nn > 0 && fq > 0, "bad sizes");
  A.nn = nn; A.fq = fq; A.n_bnodes = n_bnodes; A.bnode = bnode; A.bf_ptr = bf_ptr; A.bf_face = bf_face; A.bf_local = bf_local;
  A.face_cell = face_cell; A.face_lid = face_lid; A.nanson = nanson; A.fvals = fvals; A.cells = cells; A.sol = sol;
  for (int i = 0; i < 3; ++i) {
    A.coef[i] = i < vec ? law_host[i] : 0.0;
    A.uref[i] = i < vec ? law_host[3 + i] : 0.0;
  }
  A.power = law_host[6];
  return FEM_OK;
}

}  // namespace
}  // namespace femb200

using namespace femb200;

extern "C" int fem_face_residual(int vec, int nn, int fq, int64_t n_bnodes, const int32_t* bnode, const int32_t* bf_ptr,
                                 const int32_t* bf_face, const int32_t* bf_local, const int32_t* face_cell,
                                 const int32_t* face_lid, const double* nanson, const double* face_vals, const int32_t* cells,
                                 const double* sol, const double* law_host, double* res, void* stream) {
  if (int e = check_device()) return e;
  if (n_bnodes == 0) return FEM_OK;
  FEM_REQUIRE(res, "null pointer");
  FaceArgs A{};
  if (int e = fill(A, nn, fq, n_bnodes, bnode, bf_ptr, bf_face, bf_local, face_cell, face_lid, nanson, face_vals, cells, sol, vec, law_host)) return e;
  const unsigned grid = (unsigned)((n_bnodes + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec == 1) face_residual_kernel<1><<<grid, 128, 0, st>>>(A, res);
  else if (vec == 2) face_residual_kernel<2><<<grid, 128, 0, st>>>(A, res);
  else face_residual_kernel<3><<<grid, 128, 0, st>>>(A, res);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

extern "C" int fem_face_tangent(int vec, int nn, int fq, int64_t n_bnodes, const int32_t* bnode, const int32_t* bf_ptr,
                                const int32_t* bf_face, const int32_t* bf_local, const int32_t* face_cell,
                                const int32_t* face_lid, const double* nanson, const double* face_vals, const int32_t* cells,
                                const double* sol, const double* law_host, const int32_t* brow_ptr, const int32_t* bcol,
                                const uint8_t* bc_flag, double* data, void* stream) {
  if (int e = check_device()) return e;
  if (n_bnodes == 0) return FEM_OK;
  FEM_REQUIRE(brow_ptr && bcol && data, "null pointer");
  FaceArgs A{};
  if (int e = fill(A, nn, fq, n_bnodes, bnode, bf_ptr, bf_face, bf_local, face_cell, face_lid, nanson, face_vals, cells, sol, vec, law_host)) return e;
  const unsigned grid = (unsigned)((n_bnodes + 3) / 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec == 1) face_tangent_kernel<1><<<grid, 128, 0, st>>>(A, brow_ptr, bcol, bc_flag, data);
  else if (vec == 2) face_tangent_kernel<2><<<grid, 128, 0, st>>>(A, brow_ptr, bcol, bc_flag, data);
  else face_tangent_kernel<3><<<grid, 128, 0, st>>>(A, brow_ptr, bcol, bc_flag, data);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}
