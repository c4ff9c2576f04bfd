// Solution-dependent mass maps: the volumetric term  int m(u, x) . v  and its tangent.
//
// Replaces get_mass_kernel (jax_fem/problem.py:216-236) and its share of value_and_jacfwd (problem.py:262-266) for the
// REGISTERED mass law  m_i(u) = a u_i + b_i  with per-quadrature-point fields a (cells, quads) and b (cells, quads, vec) or
// constants: backward-Euler heat capacity rho Cp (T - T_old) / dt (applications/thermal_mechanical/example.py:40-43,
// quiet_element/example.py:50-56), the phase-field driving term (G_c / l + 2 H) d - 2 H (phase_field_fracture/example.py:54-57),
// elastic foundations / penalty terms (a u).  u-independent maps stay a constant load vector
// (jax_fem_b200/problem.py::_assemble_loads).
//
// The kernel runs between the element kernel and the gathers and ADDS to the element kernel's outputs in place -- the element
// residuals Re and the staged row blocks Ke (reference block layout, NOT the tile-major rows of fem_element_tiles):
//   Re[c][a][i]      += sum_q JxW_q N_a(q) (a_q u_i(q) + b_qi)                       (problem.py:229-234)
//   Ke[(c,a)][b][i][i] += sum_q JxW_q a_q N_a(q) N_b(q)
// so the CSR gather, the Dirichlet rows and problem.V see one tangent.  One thread per (cell, a) owns its row block and its
// residual entries: no atomics, fixed summation order.
#include "common.cuh"
#include "element_math.cuh"

namespace femb200 {
namespace {

constexpr int kMassMaxQ = 27;      // HEX8 / QUAD4 (up to the 27-point rule)
constexpr int kMassMaxQ27 = 216;   // HEX27 (basis.py:64: default degree 10 = 6 x 6 x 6 points)

struct MassArgs {
  int64_t C;
  const int32_t* cells;
  const double* points;
  const double* sol;
  const double* ref;          // [nq*NN*DIM] dN, [nq] w
  const double* vals;         // [nq*NN] shape values (basis.py:141-175)
  const double* coef_field;   // (C, nq) or nullptr
  const double* const_field;  // (C, nq, VEC) or nullptr
  double coef, cst[3];
  const int32_t* corner_pos;
  double* Ke;                 // (C*NN, row_block) or nullptr
  double* Re;                 // (C, NN*VEC)
  int nq;
};

template <int NN, int DIM, int VEC, int CPB, int MAXQ>
__global__ void __launch_bounds__(CPB* NN) mass_term_kernel(const MassArgs A) {
  constexpr int VV = VEC * VEC, ROW = (NN * VV + 1) / 2 * 2;
  __shared__ double X[CPB][NN * DIM], U[CPB][NN * VEC], AQ[CPB][MAXQ], BQ[CPB][MAXQ * VEC];
  const int lc = threadIdx.x / NN, a = threadIdx.x % NN;
  const int64_t c = (int64_t)blockIdx.x * CPB + lc;
  const bool act = c < A.C;
  const int nq = A.nq;
  if (act) {
    const int64_t node = A.cells[c * NN + a];
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[lc][a * DIM + d] = A.points[node * DIM + d];
#pragma unroll
    for (int i = 0; i < VEC; ++i) U[lc][a * VEC + i] = A.sol[node * VEC + i];
  }
  __syncthreads();
  if (act) {
    for (int q = a; q < nq; q += NN) {
      const double* dN = A.ref + (int64_t)q * NN * DIM;
      const double* N = A.vals + (int64_t)q * NN;
      double J[DIM][DIM], inv[DIM][DIM], u[VEC];
#pragma unroll
      for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int e = 0; e < DIM; ++e) J[d][e] = 0.0;
#pragma unroll
      for (int i = 0; i < VEC; ++i) u[i] = 0.0;
#pragma unroll
      for (int n = 0; n < NN; ++n) {
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
          for (int e = 0; e < DIM; ++e) J[d][e] = fma(X[lc][n * DIM + d], __ldg(dN + n * DIM + e), J[d][e]);   // fe.py:132
        const double Nn = __ldg(N + n);
#pragma unroll
        for (int i = 0; i < VEC; ++i) u[i] = fma(Nn, U[lc][n * VEC + i], u[i]);                                 // problem.py:229
      }
      const double jw = det_inv<DIM>(J, inv) * __ldg(A.ref + (int64_t)nq * NN * DIM + q);                        // fe.py:140
      const double al = A.coef_field ? A.coef_field[c * nq + q] : A.coef;
      AQ[lc][q] = jw * al;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const double be = A.const_field ? A.const_field[(c * nq + q) * VEC + i] : A.cst[i];
        BQ[lc][q * VEC + i] = jw * fma(al, u[i], be);
      }
    }
  }
  __syncthreads();
  if (!act) return;
  double r[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) r[i] = 0.0;
  for (int q = 0; q < nq; ++q) {
    const double Na = __ldg(A.vals + (int64_t)q * NN + a);
#pragma unroll
    for (int i = 0; i < VEC; ++i) r[i] = fma(Na, BQ[lc][q * VEC + i], r[i]);
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) A.Re[c * (NN * VEC) + a * VEC + i] += r[i];
  if (A.Ke == nullptr) return;
  double* row = A.Ke + (int64_t)(A.corner_pos ? A.corner_pos[c * NN + a] : (int)(c * NN + a)) * ROW;
  for (int b = 0; b < NN; ++b) {
    double m = 0.0;
    for (int q = 0; q < nq; ++q) m = fma(AQ[lc][q] * __ldg(A.vals + (int64_t)q * NN + a), __ldg(A.vals + (int64_t)q * NN + b), m);
#pragma unroll
    for (int i = 0; i < VEC; ++i) row[b * VV + i * VEC + i] += m;
  }
}

template <int NN, int DIM, int VEC>
int launch_mass(const MassArgs& A, cudaStream_t st) {
  constexpr int CPB = 128 / NN, MAXQ = NN == 27 ? kMassMaxQ27 : kMassMaxQ;
  mass_term_kernel<NN, DIM, VEC, CPB, MAXQ><<<(unsigned)((A.C + CPB - 1) / CPB), CPB * NN, 0, st>>>(A);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

}  // namespace
}  // namespace femb200

using namespace femb200;

extern "C" int fem_mass_term(int ele_type, int vec, const double* points, const int32_t* cells, int64_t n_cells,
                             const double* sol, const double* ref_tables, const double* shape_vals, int n_quad, double coef,
                             const double* coef_field, const double* const_host, const double* const_field,
                             const int32_t* corner_pos, double* Ke, double* Re, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(points && cells && sol && ref_tables && shape_vals && const_host && Re, "null pointer");
  FEM_REQUIRE(n_quad > 0 && n_quad <= (ele_type == FEM_ELE_HEX27 ? kMassMaxQ27 : kMassMaxQ), "unsupported number of quadrature points");
  if (n_cells == 0) return FEM_OK;
  MassArgs A{};
  A.C = n_cells; A.cells = cells; A.points = points; A.sol = sol; A.ref = ref_tables; A.vals = shape_vals;
  A.coef_field = coef_field; A.const_field = const_field; A.coef = coef;
  for (int i = 0; i < 3; ++i) A.cst[i] = i < vec ? const_host[i] : 0.0;
  A.corner_pos = corner_pos; A.Ke = Ke; A.Re = Re; A.nq = n_quad;
  cudaStream_t st = (cudaStream_t)stream;
#define FEM_M(E, NN, DIM, V) \
  if (ele_type == E && vec == V) return launch_mass<NN, DIM, V>(A, st);
  FEM_M(FEM_ELE_HEX8, 8, 3, 1) FEM_M(FEM_ELE_HEX8, 8, 3, 3) FEM_M(FEM_ELE_QUAD4, 4, 2, 1) FEM_M(FEM_ELE_QUAD4, 4, 2, 2)
  FEM_M(FEM_ELE_HEX27, 27, 3, 3)
#undef FEM_M
  set_error("fem_mass_term: unregistered (ele_type=%d, vec=%d)", ele_type, vec);
  return FEM_EINVAL;
}
