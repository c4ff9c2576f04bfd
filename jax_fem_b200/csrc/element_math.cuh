// Per-quadrature-point building blocks shared by the element kernels (element.cu) and the fused
// owner-computes assembly (fused.cu): geometry of FiniteElement.get_shape_grads (jax_fem/fe.py:112-141),
// grad u of get_laplace_kernel (jax_fem/problem.py:204-205) and the registered constitutive laws.
#pragma once
#include "common.cuh"

namespace femb200 {

template <int DIM>
__device__ __forceinline__ double det_inv(const double (&J)[DIM][DIM], double (&inv)[DIM][DIM]);

template <>
__device__ __forceinline__ double det_inv<2>(const double (&J)[2][2], double (&inv)[2][2]) {
  const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  const double r = 1.0 / det;
  inv[0][0] = J[1][1] * r;
  inv[0][1] = -J[0][1] * r;
  inv[1][0] = -J[1][0] * r;
  inv[1][1] = J[0][0] * r;
  return det;
}

template <>
__device__ __forceinline__ double det_inv<3>(const double (&J)[3][3], double (&inv)[3][3]) {
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  const double r = 1.0 / det;
  inv[0][0] = c00 * r;
  inv[1][0] = c01 * r;
  inv[2][0] = c02 * r;
  inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
  inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
  inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
  inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
  inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
  inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
  return det;
}

// ---- phase-1 building blocks -------------------------------------------------------------------
template <int NN, int DIM>
__device__ __forceinline__ double qp_geometry(const double* __restrict__ X, const double* __restrict__ tabq,
                                              double wq, double (&g)[NN][DIM]) {
  double J[DIM][DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d)
#pragma unroll
    for (int e = 0; e < DIM; ++e) J[d][e] = 0.0;
#pragma unroll
  for (int n = 0; n < NN; ++n)
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int e = 0; e < DIM; ++e) J[d][e] = fma(X[n * DIM + d], tabq[n * DIM + e], J[d][e]);   // fe.py:132
  double inv[DIM][DIM];
  const double det = det_inv<DIM>(J, inv);                                                      // fe.py:134-135
#pragma unroll
  for (int n = 0; n < NN; ++n)
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double s = 0.0;
#pragma unroll
      for (int e = 0; e < DIM; ++e) s = fma(tabq[n * DIM + e], inv[e][d], s);                   // fe.py:138-139
      g[n][d] = s;
    }
  return det * wq;                                                                              // fe.py:140
}

template <int NN, int DIM, int VEC>
__device__ __forceinline__ void qp_grad_u(const double* __restrict__ U, const double (&g)[NN][DIM],
                                          double (&ug)[VEC][DIM]) {
#pragma unroll
  for (int i = 0; i < VEC; ++i)
#pragma unroll
    for (int d = 0; d < DIM; ++d) ug[i][d] = 0.0;
#pragma unroll
  for (int n = 0; n < NN; ++n)
#pragma unroll
    for (int i = 0; i < VEC; ++i)
#pragma unroll
      for (int d = 0; d < DIM; ++d) ug[i][d] = fma(U[n * VEC + i], g[n][d], ug[i][d]);          // problem.py:204-205
}

// Young's modulus at a quadrature point for the isotropic laws, and its derivative wrt theta.
template <int LAW>
__device__ __forceinline__ double iso_modulus(const double* p, const double* ivq, bool derivative) {
  if constexpr (LAW == FEM_LAW_SIMP) {
    const double theta = *ivq;
    if (derivative) return (p[0] - p[1]) * p[3] * pow(theta, p[3] - 1.0);
    return p[1] + (p[0] - p[1]) * pow(theta, p[3]);
  } else {
    return p[0];
  }
}
// lambda / E of the isotropic laws.  Plane stress (2-D only; parameter slot 2 of LINEAR_ELASTIC, slot 4 of SIMP) is the law
// of docs/source/learn/topology_optimization/example.ipynb cell 9: lambda* = E nu / ((1+nu)(1-nu)).
template <int LAW, int DIM>
__device__ __forceinline__ double iso_lam1(const double* p, double nu) {
  const bool plane_stress = DIM == 2 && (LAW == FEM_LAW_SIMP ? p[4] : p[2]) != 0.0;
  return plane_stress ? nu / ((1.0 + nu) * (1.0 - nu)) : nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
}

template <int LAW>
__device__ __forceinline__ double iso_nu(const double* p) {
  return LAW == FEM_LAW_SIMP ? p[2] : p[1];
}

template <int DIM>
__device__ __forceinline__ void iso_stress(double lam, double mu, const double (&ug)[DIM][DIM], double (&sig)[DIM][DIM]) {
  double tr = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) tr += ug[d][d];
#pragma unroll
  for (int i = 0; i < DIM; ++i)
#pragma unroll
    for (int d = 0; d < DIM; ++d) sig[i][d] = mu * (ug[i][d] + ug[d][i]) + (i == d ? lam * tr : 0.0);
}

struct NHPoint {
  double F[3][3], H[3][3];   // H = F^-T
  double J, I1, m;           // m = mu J^-2/3
};

__device__ __forceinline__ void nh_kinematics(const double (&ug)[3][3], double mu, bool clamp, NHPoint& k) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) k.F[i][j] = ug[i][j] + (i == j ? 1.0 : 0.0);
  double Finv[3][3];
  double J = det_inv<3>(k.F, Finv);
  if (clamp) J = fmax(J, 1e-14);
  k.J = J;
  k.I1 = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      k.H[i][j] = Finv[j][i];
      k.I1 = fma(k.F[i][j], k.F[i][j], k.I1);
    }
  k.m = mu * pow(J, -2.0 / 3.0);
}

__device__ __forceinline__ void nh_stress(const NHPoint& k, double kappa, double (&P)[3][3]) {
  const double a = k.I1 / 3.0, b = kappa * (k.J - 1.0) * k.J;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) P[i][j] = k.m * (k.F[i][j] - a * k.H[i][j]) + b * k.H[i][j];
}

}  // namespace femb200
