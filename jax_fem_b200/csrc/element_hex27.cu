// HEX27 (27-node, second-order hexahedron) element kernel: the dense 81x81 element tangent on the FP64 tensor
// cores (DMMA, mma.sync.m8n8k4.f64), isotropic elasticity (linear / SIMP).
//
// Replaces, for ele_type='HEX27' (jax_fem/basis.py:58-65: degree 2, default quadrature degree 10 = 6x6x6 points),
// the same reference functions as element.cu: get_laplace_kernel + value_and_jacfwd (jax_fem/problem.py:189-214,
// 262-266) and FiniteElement.get_shape_grads (jax_fem/fe.py:112-141).  The reference would materialise
// shape_grads (C,216,27,3) = 30 GB at 60^3; here geometry is recomputed per cell.
//
// General kernel (hex27_kernel): persistent CTAs of 10 warps, one cell at a time per CTA (all cells, or the list the affine-cell
// pass at the end of this file leaves):
//   phase 1  thread = quadrature point: J, J^-1, JxW, grad u, stress  ->  shared {J^-1, E w, S = sigma JxW}
//   phase 2  for chunks of 32 quadrature points:
//              all threads: g[q][n][:] = dN[q][n] J^-1(q)                     -> shared Gq[q][3n+d]
//              DMMA: G(81x81) += sum_q (E_q w_q g(q)) g(q)^T as 8x8 tiles; only the 66 upper tiles of the 11x11
//                    grid are computed (G is symmetric).  The tile rows/columns form 4 groups (3, 3, 3, 2) and every
//                    warp owns one block of the 4x4 upper block grid (<= 9 tiles): per k-step it loads the 2-3 row and
//                    2-3 column fragments of its block ONCE and issues up to 9 DMMA with them (0.7 shared loads per
//                    DMMA instead of 3: the kernel was bound by shared-memory bandwidth), accumulators in registers
//              threads 0..80: r_(a,i) += sum_q S_q[i][:] . g_a(q)
//   phase 3  tiles -> shared G (mirrored), in-place 3x3 conversion K_ab = lam' G_ab + mu' G_ab^T + mu' tr(G_ab) I,
//            coalesced copy of the 27 row blocks to their node-sorted positions (row block = 27 blocks of 3x3,
//            padded to 244 doubles so that every row block is a 16-byte multiple for the TMA gather).
#include "common.cuh"

namespace femb200 {
namespace {

constexpr int H27_NN = 27, H27_ND = 81, H27_T = 11;        // 11 x 11 tiles of 8 x 8 cover 81 x 81 (padded to 88)
constexpr int H27_QC = 32;                                 // quadrature points per chunk (8 DMMA k-steps)
constexpr int H27_GS = 100;                                // Gq row stride: 4 (mod 16) => conflict-free fragments
constexpr int H27_GSS = 89;                                // G row stride (odd)
constexpr int H27_QP = 19;                                 // per point: J^-1 (9), E w (1), S (9)
constexpr int H27_THREADS = 320, H27_WARPS = 10;              // one warp per block of the upper 4x4 block grid
constexpr int H27_TPW = 9;                                 // tiles per warp (3 x 3 block)

struct Hex27Args {
  const double* points;
  const int32_t* cells;
  const double* sol;
  const double* iv;          // (C, nq) or nullptr
  const double* ref;         // [nq*27*3] dN, [nq] w
  const double* ref_t;       // [27*3][nq] the same dN with the point index fastest (coalesced reads when thread = point) or nullptr
  const int32_t* corner_pos;
  const double* affine;      // [27][3] reference node coordinates, then [3][3][27][27] reference Gram tables; or nullptr
  double* rec;               // per cell: J^-1, E detJ, affine flag (written by the affine pass's first kernel) or nullptr
  int32_t* list;             // [0] = number of cells left to the general kernel, [1..] their ids (written by the affine pass) or nullptr
  double* Ke;                // (C*27, 244)
  double* Re;                // (C, 81)
  int64_t C;
  int nq;
  int law;
  double p[8];
};


__device__ __forceinline__ double det_inv3(const double (&J)[3][3], double (&inv)[3][3]) {
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

// One 32-point chunk (8 k-steps) of a block of NR x NC tiles starting at tile (I0, J0); DIAG: only tiles c >= r, and the
// row fragments are the column fragments scaled by E w.  Lane (row = l/4, k = l%4) holds A[row][k] = E w g[q][8I + l/4]
// and B[k][col] = g[q][8J + l/4] with q = 4s + l%4.
template <int NR, int NC, bool DIAG>
__device__ __forceinline__ void h27_block_chunk(const double* __restrict__ Gq, const double* __restrict__ ew, int I0, int J0,
                                                int l, double (&C)[H27_TPW][2]) {
  const double* g = Gq + (l & 3) * H27_GS + (l >> 2);
#pragma unroll
  for (int s = 0; s < H27_QC / 4; ++s) {
    const double e = ew[4 * s * H27_QP];
    double a[NR], b[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) b[c] = g[4 * s * H27_GS + 8 * (J0 + c)];
#pragma unroll
    for (int r = 0; r < NR; ++r) a[r] = e * (DIAG ? b[r] : g[4 * s * H27_GS + 8 * (I0 + r)]);
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (!DIAG || c >= r) dmma884(C[r * 3 + c], a[r], b[c]);
  }
}

__device__ __forceinline__ void hex27_cell(const Hex27Args& A, const int64_t c, double* sm) {
  const int nq = A.nq;
  const int nq_pad = (nq + H27_QC - 1) / H27_QC * H27_QC;
  double* X = sm;                          // [27][3]
  double* U = X + H27_ND;                  // [27][3]
  double* R = U + H27_ND;                  // [81] residual
  double* work = R + H27_ND + 1;           // even offset (244) => 16-byte aligned
  double* QP = work;                       // [nq_pad][19]
  double* Gq = QP + nq_pad * H27_QP + (nq_pad * H27_QP) % 2;   // 2 x [32][100]
  double* G = work;                        // [88][89] overlays QP + Gq after the main loop
  __shared__ int pos[H27_NN];
  const int tid = threadIdx.x, warp = tid >> 5, l = tid & 31;

  if (tid < H27_NN) {
    const int64_t node = A.cells[c * H27_NN + tid];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      X[tid * 3 + d] = A.points[node * 3 + d];
      U[tid * 3 + d] = A.sol[node * 3 + d];
    }
    pos[tid] = A.corner_pos ? A.corner_pos[c * H27_NN + tid] : (int)(c * H27_NN + tid);
  }
  __syncthreads();

  const double nu = A.law == FEM_LAW_SIMP ? A.p[2] : A.p[1];
  const double mu1 = 1.0 / (2.0 * (1.0 + nu)), lam1 = nu / ((1.0 + nu) * (1.0 - 2.0 * nu));

  // ---- phase 1: thread = quadrature point ----
  for (int q = tid; q < nq_pad; q += H27_THREADS) {
    double* rec = QP + q * H27_QP;
    if (q >= nq) {                                        // padding points carry zero weight
      for (int j = 0; j < H27_QP; ++j) rec[j] = 0.0;
      continue;
    }
    // dN of point q: consecutive threads are consecutive points, so the table is read with q fastest when the caller supplies
    // the transposed copy (one coalesced 256-byte request per warp instead of 32 different cache lines)
    const double* dN = A.ref + (int64_t)q * H27_ND;
    const double* dT = A.ref_t ? A.ref_t + q : nullptr;
    auto dn = [&](int n, int d) { return dT ? __ldg(dT + (int64_t)(n * 3 + d) * nq) : __ldg(dN + n * 3 + d); };
    // One pass over the 27 nodes gives both J = sum_n x_n (x) dN_n (fe.py:132) and the reference gradient of u,
    // Gu = sum_n u_n (x) dN_n: 18 independent accumulation chains.  grad u = Gu J^-1 then costs 27 FMA per point instead of
    // forming every physical shape gradient here (problem.py:204-205 with fe.py:138-139 substituted).
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Gu[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll 3
    for (int n = 0; n < H27_NN; ++n) {
      const double d0 = dn(n, 0), d1 = dn(n, 1), d2 = dn(n, 2);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double x = X[n * 3 + d], u = U[n * 3 + d];
        J[d][0] = fma(x, d0, J[d][0]);
        J[d][1] = fma(x, d1, J[d][1]);
        J[d][2] = fma(x, d2, J[d][2]);
        Gu[d][0] = fma(u, d0, Gu[d][0]);
        Gu[d][1] = fma(u, d1, Gu[d][1]);
        Gu[d][2] = fma(u, d2, Gu[d][2]);
      }
    }
    double inv[3][3];
    const double det = det_inv3(J, inv);                  // fe.py:134-135
    const double w = det * __ldg(A.ref + (int64_t)nq * H27_ND + q);        // fe.py:140
    double ug[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int d = 0; d < 3; ++d) ug[i][d] = Gu[i][0] * inv[0][d] + Gu[i][1] * inv[1][d] + Gu[i][2] * inv[2][d];
    double E;
    if (A.law == FEM_LAW_SIMP) E = A.p[1] + (A.p[0] - A.p[1]) * pow(A.iv[c * nq + q], A.p[3]);
    else E = A.p[0];
    const double mu = E * mu1, lam = E * lam1;
    const double tr = ug[0][0] + ug[1][1] + ug[2][2];
#pragma unroll
    for (int e = 0; e < 3; ++e)
#pragma unroll
      for (int d = 0; d < 3; ++d) rec[e * 3 + d] = inv[e][d];
    rec[9] = E * w;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int d = 0; d < 3; ++d) rec[10 + i * 3 + d] = (mu * (ug[i][d] + ug[d][i]) + (i == d ? lam * tr : 0.0)) * w;
  }
  // zero the padding columns 81..99 of Gq once
  for (int j = tid; j < 2 * H27_QC * (H27_GS - H27_ND); j += H27_THREADS)
    Gq[(j / (H27_GS - H27_ND)) * H27_GS + H27_ND + j % (H27_GS - H27_ND)] = 0.0;
  __syncthreads();

  // ---- phase 2: chunks of 32 points ----
  // this warp's block of the upper 4 x 4 block grid: tile groups {0,1,2}, {3,4,5}, {6,7,8}, {9,10}
  //   warps 0-2: (0,1) (0,2) (1,2)  3x3 tiles;  warps 3-5: (0,3) (1,3) (2,3)  3x2;  warps 6-8: diagonal 3x3 (6 tiles);  warp 9: diagonal 2x2 (3)
  const int bgi = warp < 3 ? (warp == 2 ? 1 : 0) : (warp < 6 ? warp - 3 : warp - 6);
  const int bgj = warp < 3 ? (warp == 0 ? 1 : 2) : (warp < 6 ? 3 : warp - 6);
  const int I0 = 3 * bgi, J0 = 3 * bgj;
  const int bnr = bgi == 3 ? 2 : 3, bnc = bgj == 3 ? 2 : 3;
  const bool bdiag = bgi == bgj;
  double Cacc[H27_TPW][2];
#pragma unroll
  for (int k = 0; k < H27_TPW; ++k) Cacc[k][0] = Cacc[k][1] = 0.0;
  double racc = 0.0;

  // g[q][n][:] = dN[q][n] J^-1(q) of one chunk -> Gq buffer; done by the warps with the fewest tiles (3..9: the three warps that
  // own 9 tiles each are the critical path of the DMMA step and get none of it)
  constexpr int G_FIRST = 96, G_THREADS = H27_THREADS - G_FIRST;
  constexpr int G_ITEMS = (H27_QC * H27_NN + G_THREADS - 1) / G_THREADS;   // 4 (point, node) pairs per thread
  auto fill_chunk = [&](int q0, double* Gb) {
    if (tid < G_FIRST) return;
    // all loads of the thread's pairs are issued before the first use: one L2 round trip per chunk instead of four
    double d[G_ITEMS][3];
#pragma unroll
    for (int k = 0; k < G_ITEMS; ++k) {
      const int j = tid - G_FIRST + k * G_THREADS, ql = j / H27_NN, n = j % H27_NN, q = q0 + ql;
      const bool on = j < H27_QC * H27_NN && q < nq;
      const double* dN = A.ref + ((int64_t)(on ? q : 0) * H27_NN + (on ? n : 0)) * 3;
#pragma unroll
      for (int e = 0; e < 3; ++e) d[k][e] = on ? __ldg(dN + e) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < G_ITEMS; ++k) {
      const int j = tid - G_FIRST + k * G_THREADS, ql = j / H27_NN, n = j % H27_NN, q = q0 + ql;
      if (j >= H27_QC * H27_NN) break;
      const double* inv = QP + (q < nq ? q : 0) * H27_QP;
#pragma unroll
      for (int e = 0; e < 3; ++e) Gb[ql * H27_GS + n * 3 + e] = d[k][0] * inv[e] + d[k][1] * inv[3 + e] + d[k][2] * inv[6 + e];
    }
  };
  // Two Gq buffers: while the slower warps still multiply chunk k, the others already fill chunk k+1 -- ONE barrier per chunk.
  fill_chunk(0, Gq);
  __syncthreads();
  const int rt = tid - (H27_THREADS - 96);                   // residual rows on the three lightest warps (7, 8, 9)
  for (int q0 = 0, kc = 0; q0 < nq_pad; q0 += H27_QC, ++kc) {
    const double* Gc = Gq + (kc & 1) * (H27_QC * H27_GS);
    {
      const double* ew = QP + (q0 + (l & 3)) * H27_QP + 9;
      if (!bdiag) {                                        // warp-uniform
        if (bnc == 3) h27_block_chunk<3, 3, false>(Gc, ew, I0, J0, l, Cacc);
        else h27_block_chunk<3, 2, false>(Gc, ew, I0, J0, l, Cacc);
      } else {
        if (bnc == 3) h27_block_chunk<3, 3, true>(Gc, ew, I0, J0, l, Cacc);
        else h27_block_chunk<2, 2, true>(Gc, ew, I0, J0, l, Cacc);
      }
    }
    // residual r_(a,i) += sum_q S_q[i][:] . g_a(q).  Tried and reverted: from the reference table instead of the shared
    // gradients (dependent global loads: 38.4 -> 43.7 ms); 4 lanes per row with 11 warps x 6 tiles (more fragment loads: 40.8 ms)
    if (rt >= 0 && rt < H27_ND) {
      const int a = rt / 3, i = rt % 3;
      double r4[4] = {0.0, 0.0, 0.0, 0.0};                 // four independent chains instead of one of 96 dependent FMAs
#pragma unroll 2
      for (int ql = 0; ql < H27_QC; ql += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double* S = QP + (q0 + ql + u) * H27_QP + 10 + i * 3;
          const double* g = Gc + (ql + u) * H27_GS + a * 3;
          r4[u] = fma(S[0], g[0], fma(S[1], g[1], fma(S[2], g[2], r4[u])));                 // problem.py:210
        }
      }
      racc += (r4[0] + r4[1]) + (r4[2] + r4[3]);
    }
    if (q0 + H27_QC < nq_pad) fill_chunk(q0 + H27_QC, Gq + ((kc + 1) & 1) * (H27_QC * H27_GS));
    __syncthreads();
  }
  if (rt >= 0 && rt < H27_ND) A.Re[c * H27_ND + rt] = racc;
  if (A.Ke == nullptr) return;                             // residual only (block-uniform)

  // ---- phase 3: tiles -> G (mirrored); fragment: row = l/4, cols = 2(l%4), 2(l%4)+1 ----
#pragma unroll
  for (int br = 0; br < 3; ++br)
#pragma unroll
    for (int bc = 0; bc < 3; ++bc) {
      if (br >= bnr || bc >= bnc || (bdiag && bc < br)) continue;
      const int tI = I0 + br, tJ = J0 + bc;
      const int r = 8 * tI + (l >> 2), c0 = 8 * tJ + 2 * (l & 3);
      G[r * H27_GSS + c0] = Cacc[br * 3 + bc][0];
      G[r * H27_GSS + c0 + 1] = Cacc[br * 3 + bc][1];
      if (tI != tJ) {
        G[c0 * H27_GSS + r] = Cacc[br * 3 + bc][0];
        G[(c0 + 1) * H27_GSS + r] = Cacc[br * 3 + bc][1];
      }
    }
  __syncthreads();
  // in-place conversion of every 3x3 node-pair block
  for (int j = tid; j < H27_NN * H27_NN; j += H27_THREADS) {
    const int a = j / H27_NN, b = j % H27_NN;
    double g[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) g[i][k] = G[(3 * a + i) * H27_GSS + 3 * b + k];
    const double tr = g[0][0] + g[1][1] + g[2][2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        G[(3 * a + i) * H27_GSS + 3 * b + k] = lam1 * g[i][k] + mu1 * g[k][i] + (i == k ? mu1 * tr : 0.0);
  }
  __syncthreads();
  // coalesced copy-out: row block a = [b][i][k], 243 doubles (+1 pad) at position pos[a]
  for (int j = tid; j < H27_NN * 244; j += H27_THREADS) {
    const int a = j / 244, rem = j % 244;
    double v = 0.0;
    if (rem < 243) {
      const int b = rem / 9, i = (rem % 9) / 3, k = rem % 3;
      v = G[(3 * a + i) * H27_GSS + 3 * b + k];
    }
    A.Ke[(int64_t)pos[a] * 244 + rem] = v;
  }
}

// Persistent CTAs over the cells: all of them, or the list the affine pass left (its order does not matter: cells are independent).
__global__ void __launch_bounds__(H27_THREADS, 2) hex27_kernel(const Hex27Args A) {
  extern __shared__ __align__(16) double sm[];
  const int64_t n = A.list ? A.list[0] : A.C;
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    hex27_cell(A, A.list ? A.list[1 + i] : i, sm);
    __syncthreads();
  }
}


// ---- affine cells ------------------------------------------------------------------------------------------
// If the map x(xi) = sum_n X_n N_n(xi) of a cell is affine (every cell of a box / sheared / stretched mesh), J is constant and
//   G_ab[I][J] = sum_q E w_q detJ (dN_a J^-1)_I (dN_b J^-1)_J = E detJ (J^-T Ghat_ab J^-1)[I][J],   Ghat_ab[e][f] = sum_q w_q dN_a^e dN_b^f,
// with the 9 x 27 x 27 reference Gram tables Ghat computed once: 59 k FMA per cell instead of 912 k, exact up to the rounding of
// the reassociated sum.  The test is exact and made per cell on every call: the Lagrange basis reproduces affine functions, so the
// map is affine iff every node sits where the affine map through the corners xi = 0, e_1, e_2, e_3 puts its reference point,
// X_n = X_0 + J xi_n with J = [X_1 - X_0, X_2 - X_0, X_3 - X_0]; with SIMP the density must also be constant over the cell's
// points.  Everything else (curved cells, graded density) is listed for hex27_kernel.  The stress is linear in grad u for the
// registered laws, so the element residual is K_e u_e (problem.py:204-210 with a constant tangent).
constexpr int H27A_WARPS = 9, H27A_THREADS = 32 * (H27A_WARPS + 1);   // compute warp w owns the row blocks a = w, w + 9, w + 18 (lane = column node b); warp 9 feeds them
constexpr int H27A_NODE_TAB = H27_ND;                        // xi_n [27][3]
constexpr int H27A_GRAM = 9 * H27_NN * H27_NN;               // Ghat[e][f][a][b]
constexpr int H27A_GRAM6 = 6 * H27_NN * H27_NN;              // ... of which the kernel keeps e <= f
constexpr int H27A_ROW = 244;

constexpr int H27A_REC = 12;                                 // per cell: J^-1 (9), E detJ, affine flag, pad

// Pass 1, one warp per cell (lane n = node n): the affinity test, J^-1, E detJ -> the cell's record; cells that fail are listed.
__global__ void __launch_bounds__(256) hex27_affine_prep_kernel(const Hex27Args A) {
  const int l = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= A.C) return;                                      // warp-uniform
  const bool in = l < H27_NN;
  double xi[3] = {0, 0, 0}, x[3] = {0, 0, 0};
  if (in) {
    const int64_t node = A.cells[c * H27_NN + l];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      xi[d] = A.affine[l * 3 + d];
      x[d] = A.points[node * 3 + d];
    }
  }
  // which lanes hold the corners xi = 0, e_1, e_2, e_3
  const double sum = xi[0] + xi[1] + xi[2];
  int corner[4];
  corner[0] = __ffs(__ballot_sync(0xffffffffu, in && sum == 0.0)) - 1;
#pragma unroll
  for (int e = 0; e < 3; ++e) corner[e + 1] = __ffs(__ballot_sync(0xffffffffu, in && xi[e] == 1.0 && sum == 1.0)) - 1;
  double J[3][3], x0[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) x0[d] = __shfl_sync(0xffffffffu, x[d], corner[0]);
#pragma unroll
  for (int e = 0; e < 3; ++e)
#pragma unroll
    for (int d = 0; d < 3; ++d) J[d][e] = __shfl_sync(0xffffffffu, x[d], corner[e + 1]) - x0[d];
  double scale = 0.0, dev = 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double pred = x0[d] + J[d][0] * xi[0] + J[d][1] * xi[1] + J[d][2] * xi[2];
    dev = fmax(dev, fabs(x[d] - pred));
    scale = fmax(scale, fmax(fabs(J[d][0]), fmax(fabs(J[d][1]), fabs(J[d][2]))));
  }
  bool ok = !in || dev <= 1e-13 * scale;
  double E = A.p[0];
  if (A.law == FEM_LAW_SIMP) {
    const double* iv = A.iv + c * A.nq;
    const double r0 = iv[0];
    for (int q = l; q < A.nq; q += 32) ok = ok && iv[q] == r0;
    E = A.p[1] + (A.p[0] - A.p[1]) * pow(r0, A.p[3]);
  }
  double inv[3][3];
  const double det = det_inv3(J, inv);
  ok = __all_sync(0xffffffffu, ok) && det > 0.0;             // inverted cell: let the general kernel reproduce the reference
  if (l == 0) {
    double* rec = A.rec + c * H27A_REC;
#pragma unroll
    for (int e = 0; e < 3; ++e)
#pragma unroll
      for (int d = 0; d < 3; ++d) rec[e * 3 + d] = inv[e][d];
    rec[9] = E * det;
    rec[10] = ok ? 1.0 : 0.0;
    rec[11] = 0.0;
    if (!ok) A.list[1 + atomicAdd(A.list, 1)] = (int)c;
  }
}

// Pass 2, persistent CTAs of 9 warps; warp w owns the row blocks a = w, w + 9, w + 18 of the cell, lane = column node b.
__global__ void __launch_bounds__(H27A_THREADS, 2) hex27_affine_kernel(const Hex27Args A) {
  extern __shared__ __align__(16) double sm[];
  double* gram = sm;                                         // the 6 tables e <= f ([e][f][a][b] = [f][e][b][a]): 4374 doubles
  double* out = gram + H27A_GRAM6;                           // [27][244] row blocks of the cell
  double* Y = out + H27_NN * H27A_ROW;                       // [27][27][3]: K_ab u_b, summed over b into the residual
  double* U = Y + H27_NN * H27_ND + 1;                       // 2 x [27][3] (double-buffered by cell parity)
  double* JI = U + 2 * H27_ND;                               // 2 x the cell's record
  __shared__ int pos[2][H27_NN];
  const int tid = threadIdx.x, warp = tid >> 5, l = tid & 31;
  const double nu = A.law == FEM_LAW_SIMP ? A.p[2] : A.p[1];
  const double mu1 = 1.0 / (2.0 * (1.0 + nu)), lam1 = nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
  for (int i = tid; i < H27A_GRAM6; i += H27A_THREADS) {
    const int t = i / (H27_NN * H27_NN), e = t < 3 ? 0 : (t < 5 ? 1 : 2), f = t < 3 ? t : (t < 5 ? t - 2 : 2);
    gram[i] = A.affine[H27A_NODE_TAB + (e * 3 + f) * (H27_NN * H27_NN) + i % (H27_NN * H27_NN)];
  }

  // Warp 9 is the producer: while the nine compute warps work on cell c it stages the solution / row positions / record of the
  // CTA's next cell in the other half of the double buffers and has the loads of the cell after that (and the node ids of one
  // further) in flight, so that no global-load latency sits between the barriers of the compute warps.
  const int64_t stride = gridDim.x;
  int node_next = 0, ppos = 0;
  double pu[3] = {0, 0, 0}, prec = 0.0;
  auto load_node = [&](int64_t c) { return (c < A.C && l < H27_NN) ? A.cells[c * H27_NN + l] : 0; };
  auto load_data = [&](int64_t c, int node) {
    if (c >= A.C) return;
    if (l < H27_NN) {
#pragma unroll
      for (int d = 0; d < 3; ++d) pu[d] = A.sol[(int64_t)node * 3 + d];
      ppos = A.corner_pos ? A.corner_pos[c * H27_NN + l] : (int)(c * H27_NN + l);
    }
    if (l < H27A_REC) prec = A.rec[c * H27A_REC + l];
  };
  auto stage = [&](int half) {
    if (l < H27_NN) {
#pragma unroll
      for (int d = 0; d < 3; ++d) U[half * H27_ND + l * 3 + d] = pu[d];
      pos[half][l] = ppos;
    }
    if (l < H27A_REC) JI[half * H27A_REC + l] = prec;
  };
  if (warp == H27A_WARPS) {
    load_data(blockIdx.x, load_node(blockIdx.x));
    stage(0);
    load_data(blockIdx.x + stride, load_node(blockIdx.x + stride));
    node_next = load_node(blockIdx.x + 2 * stride);
  }
  __syncthreads();

  int par = 0;
  for (int64_t c = blockIdx.x; c < A.C; c += stride, par ^= 1) {
    const bool affine = JI[par * H27A_REC + 10] != 0.0;      // block-uniform; otherwise the cell is listed for hex27_kernel
    if (warp == H27A_WARPS) {
      if (c + stride < A.C) stage(par ^ 1);
      load_data(c + 2 * stride, node_next);
      node_next = load_node(c + 3 * stride);
    } else if (affine) {
    double inv[3][3];
#pragma unroll
    for (int e = 0; e < 3; ++e)
#pragma unroll
      for (int d = 0; d < 3; ++d) inv[e][d] = JI[par * H27A_REC + e * 3 + d];
    const double cdet = JI[par * H27A_REC + 9];
    double ub[3] = {0, 0, 0};
    if (l < H27_NN) {
#pragma unroll
      for (int d = 0; d < 3; ++d) ub[d] = U[par * H27_ND + l * 3 + d];
    }
#pragma unroll
    for (int ai = 0; ai < H27_NN / H27A_WARPS; ++ai) {       // three independent items per lane: instruction-level parallelism
      const int a = warp + ai * H27A_WARPS;
      if (l < H27_NN) {
        const int j = a * H27_NN + l, jt = l * H27_NN + a;
        constexpr int S = H27_NN * H27_NN;
        // Ghat_ab[e][f]: upper tables directly, lower ones from the transposed entry of the mirrored table
        const double g00 = gram[j], g01 = gram[S + j], g02 = gram[2 * S + j], g11 = gram[3 * S + j], g12 = gram[4 * S + j],
                     g22 = gram[5 * S + j], g10 = gram[S + jt], g20 = gram[2 * S + jt], g21 = gram[4 * S + jt];
        const double gh[3][3] = {{g00, g01, g02}, {g10, g11, g12}, {g20, g21, g22}};
        double T[3][3], G[3][3];
#pragma unroll
        for (int e = 0; e < 3; ++e)
#pragma unroll
          for (int k = 0; k < 3; ++k) T[e][k] = gh[e][0] * inv[0][k] + gh[e][1] * inv[1][k] + gh[e][2] * inv[2][k];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int k = 0; k < 3; ++k) G[i][k] = cdet * (inv[0][i] * T[0][k] + inv[1][i] * T[1][k] + inv[2][i] * T[2][k]);
        const double tr = G[0][0] + G[1][1] + G[2][2];
        double* o = out + a * H27A_ROW + l * 9;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double y = 0.0;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const double v = lam1 * G[i][k] + mu1 * G[k][i] + (i == k ? mu1 * tr : 0.0);
            o[i * 3 + k] = v;
            y = fma(v, ub[k], y);
          }
          Y[j * 3 + i] = y;
        }
      } else if (l == 31) {
        out[a * H27A_ROW + 243] = 0.0;
      }
    }
    fence_async_smem();
    }
    __syncthreads();
    if (affine) {
      if (warp < 3) {
        // residual row r = (a, i): sum of K_ab u_b over the column nodes in ascending b
        const int r = tid;
        if (r < H27_ND) {
          const double* y = Y + (r / 3) * H27_ND + r % 3;
          double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int b = 0; b < H27_NN; b += 3) {
            s0 += y[b * 3];
            s1 += y[b * 3 + 3];
            s2 += y[b * 3 + 6];
          }
          A.Re[c * H27_ND + r] = (s0 + s1) + s2;
        }
      } else if (warp == 3 && A.Ke && l < H27_NN) {
        // the 27 row blocks go to their node-sorted positions as bulk copies (no LDS / STG instructions); the next cell may
        // overwrite them once they have been read
        bulk_s2g(A.Ke + (int64_t)pos[par][l] * H27A_ROW, out + l * H27A_ROW, H27A_ROW * sizeof(double));
        bulk_commit();
        bulk_wait_read<0>();
      }
    }
    __syncthreads();
  }
}


// ---- adjoint: -lambda^T dc/dtheta per quadrature point (SIMP), one CTA per cell, thread = point --------------------------
// d r_a / d theta_q = dE/dtheta (lam' tr(eps) I + 2 mu' eps)(q) grad N_a(q) JxW_q, so the contraction with lambda is the double
// dot of that stress derivative with grad(lambda)(q) = GL J^-1, GL = sum_n lambda_n (x) dN_n: no per-node gradient is formed
// (solver.py:1362-1418 with problem.py:204-210 differentiated in the density).
__global__ void __launch_bounds__(224) hex27_param_grad_kernel(const Hex27Args A, const double* __restrict__ lam,
                                                               double* __restrict__ grad) {
  __shared__ double X[H27_ND], U[H27_ND], L[H27_ND];
  const int64_t c = blockIdx.x;
  const int tid = threadIdx.x, nq = A.nq;
  if (tid < H27_NN) {
    const int64_t node = A.cells[c * H27_NN + tid];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      X[tid * 3 + d] = A.points[node * 3 + d];
      U[tid * 3 + d] = A.sol[node * 3 + d];
      L[tid * 3 + d] = lam[node * 3 + d];
    }
  }
  __syncthreads();
  const double nu = A.p[2], mu1 = 1.0 / (2.0 * (1.0 + nu)), lam1 = nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
  for (int q = tid; q < nq; q += blockDim.x) {
    const double* dN = A.ref + (int64_t)q * H27_ND;
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Gu[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}},
           GL[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll 3
    for (int n = 0; n < H27_NN; ++n) {
      const double d0 = __ldg(dN + n * 3), d1 = __ldg(dN + n * 3 + 1), d2 = __ldg(dN + n * 3 + 2);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double x = X[n * 3 + i], u = U[n * 3 + i], l = L[n * 3 + i];
        J[i][0] = fma(x, d0, J[i][0]);  J[i][1] = fma(x, d1, J[i][1]);  J[i][2] = fma(x, d2, J[i][2]);
        Gu[i][0] = fma(u, d0, Gu[i][0]); Gu[i][1] = fma(u, d1, Gu[i][1]); Gu[i][2] = fma(u, d2, Gu[i][2]);
        GL[i][0] = fma(l, d0, GL[i][0]); GL[i][1] = fma(l, d1, GL[i][1]); GL[i][2] = fma(l, d2, GL[i][2]);
      }
    }
    double inv[3][3];
    const double w = det_inv3(J, inv) * __ldg(A.ref + (int64_t)nq * H27_ND + q);
    double ug[3][3], lg[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        ug[i][d] = Gu[i][0] * inv[0][d] + Gu[i][1] * inv[1][d] + Gu[i][2] * inv[2][d];
        lg[i][d] = GL[i][0] * inv[0][d] + GL[i][1] * inv[1][d] + GL[i][2] * inv[2][d];
      }
    const double theta = A.iv[c * nq + q];
    const double dE = (A.p[0] - A.p[1]) * A.p[3] * pow(theta, A.p[3] - 1.0);       // E = Emin + (Emax - Emin) theta^p
    const double tr = ug[0][0] + ug[1][1] + ug[2][2];
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int d = 0; d < 3; ++d)
        acc = fma(dE * (mu1 * (ug[i][d] + ug[d][i]) + (i == d ? lam1 * tr : 0.0)), lg[i][d], acc);
    grad[c * nq + q] = -acc * w;
  }
}

}  // namespace
}  // namespace femb200

using namespace femb200;

extern "C" int fem_hex27_residual_jacobian(int law_id, const double* law_params_host, const double* points,
                                           const int32_t* cells, int64_t n_cells, const double* sol,
                                           const double* internal_var, const double* ref_tables,
                                           const double* ref_tables_t, int n_quad, const double* affine_tables,
                                           void* affine_work, const int32_t* corner_pos, double* Ke, double* Re,
                                           void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(points && cells && sol && ref_tables && Re && law_params_host, "null pointer");
  FEM_REQUIRE(law_id == FEM_LAW_LINEAR_ELASTIC || law_id == FEM_LAW_SIMP,
              "HEX27 is registered for isotropic elasticity (linear, SIMP) only");
  FEM_REQUIRE(!(law_id == FEM_LAW_SIMP && !internal_var), "SIMP needs the per-quadrature-point density");
  FEM_REQUIRE(n_quad > 0 && n_quad <= 512, "unsupported number of quadrature points");
  FEM_REQUIRE(!affine_tables == !affine_work, "the affine pass needs both its tables and its workspace");
  FEM_REQUIRE((reinterpret_cast<uintptr_t>(affine_work) & 15) == 0, "affine_work must be 16-byte aligned");
  if (n_cells == 0) return FEM_OK;
  Hex27Args A{};
  A.points = points; A.cells = cells; A.sol = sol; A.iv = internal_var; A.ref = ref_tables; A.ref_t = ref_tables_t;
  A.corner_pos = corner_pos; A.Ke = Ke; A.Re = Re; A.C = n_cells; A.nq = n_quad; A.law = law_id;
  for (int i = 0; i < 8; ++i) A.p[i] = law_params_host[i];
  A.affine = affine_tables;
  A.rec = static_cast<double*>(affine_work);                                 // [n_cells][12] doubles, then the int32 list
  A.list = affine_work ? reinterpret_cast<int32_t*>(A.rec + n_cells * H27A_REC) : nullptr;
  static int grid_of[64] = {0};
  int dev = 0;
  FEM_CUDA_CHECK(cudaGetDevice(&dev));
  FEM_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  const int nq_pad = (n_quad + H27_QC - 1) / H27_QC * H27_QC;
  const size_t work = (size_t)nq_pad * H27_QP + (nq_pad * H27_QP) % 2 + 2 * H27_QC * H27_GS;
  const size_t gsz = (size_t)88 * H27_GSS;
  const size_t smem = sizeof(double) * (3 * H27_ND + 1 + (work > gsz ? work : gsz));
  const size_t asmem = sizeof(double) * (H27A_GRAM6 + H27_NN * H27A_ROW + H27_NN * H27_ND + 1 + 2 * H27_ND + 2 * H27A_REC);
  if (grid_of[dev] == 0) {
    int sms = 0;
    FEM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FEM_CUDA_CHECK(cudaFuncSetAttribute(hex27_affine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asmem));
    FEM_CUDA_CHECK(cudaFuncSetAttribute(hex27_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    grid_of[dev] = 2 * sms;
  }
  FEM_REQUIRE(smem <= 200 * 1024, "quadrature rule too large for the HEX27 kernel's shared memory");
  const unsigned grid = (unsigned)(n_cells < grid_of[dev] ? n_cells : grid_of[dev]);
  if (affine_tables) {
    FEM_CUDA_CHECK(cudaMemsetAsync(A.list, 0, sizeof(int32_t), (cudaStream_t)stream));
    hex27_affine_prep_kernel<<<(unsigned)((n_cells + 7) / 8), 256, 0, (cudaStream_t)stream>>>(A);
    FEM_LAUNCH_CHECK();
    hex27_affine_kernel<<<grid, H27A_THREADS, asmem, (cudaStream_t)stream>>>(A);
    FEM_LAUNCH_CHECK();
  }
  hex27_kernel<<<grid, H27_THREADS, smem, (cudaStream_t)stream>>>(A);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

extern "C" int fem_hex27_adjoint_param_grad(int law_id, const double* law_params_host, const double* points,
                                            const int32_t* cells, int64_t n_cells, const double* sol,
                                            const double* internal_var, const double* lam, const double* ref_tables,
                                            int n_quad, double* grad, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(points && cells && sol && internal_var && lam && ref_tables && grad && law_params_host, "null pointer");
  FEM_REQUIRE(law_id == FEM_LAW_SIMP, "the HEX27 parameter gradient is registered for SIMP (per-point density) only");
  FEM_REQUIRE(n_quad > 0 && n_quad <= 512, "unsupported number of quadrature points");
  if (n_cells == 0) return FEM_OK;
  Hex27Args A{};
  A.points = points; A.cells = cells; A.sol = sol; A.iv = internal_var; A.ref = ref_tables; A.C = n_cells; A.nq = n_quad;
  A.law = law_id;
  for (int i = 0; i < 8; ++i) A.p[i] = law_params_host[i];
  hex27_param_grad_kernel<<<(unsigned)n_cells, 224, 0, (cudaStream_t)stream>>>(A, lam, grad);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}
