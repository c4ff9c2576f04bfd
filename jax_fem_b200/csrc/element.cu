// Element kernels: per-cell residual and tangent for the registered constitutive laws.
//
// Replaces Problem.compute_residual_vars / compute_newton_vars (jax_fem/problem.py:439-460), i.e.
// get_laplace_kernel (:189-214) + value_and_jacfwd (:262-266), with the geometry of
// FiniteElement.get_shape_grads (jax_fem/fe.py:112-141) recomputed per cell instead of being
// materialised as (C,Q,N,dim) arrays.
//
// Thread mapping: NN consecutive lanes own one cell (NN == NQ for HEX8 and QUAD4).
//   phase 1: lane = quadrature point q.  J, det, J^-1, physical gradients g[n][:], JxW,
//            grad u, stress and the law's tangent data -> shared-memory record of (cell, q).
//   phase 2: lane = local node a.  r_a = sum_q S_q g_a(q);  K row block (a, :) accumulated in
//            registers from broadcast shared-memory reads of g_b(q) (all NN lanes of a cell read
//            the same address), closed-form tangents -- no AD, no per-cell (24x24) temporaries.
//   output : one contiguous row block (NN blocks of VEC x VEC) per corner (cell, a), placed at the
//            corner's position in the plan's node-sorted order.
#include <stdlib.h>
#include "common.cuh"
#include "element_math.cuh"

namespace femb200 {
namespace {

constexpr int pad_odd(int x) { return (x % 2) ? x : x + 1; }

// Law-specific extra doubles in a (cell, q) record.
template <int LAW, int NN, int DIM>
struct LawExtra {
  static constexpr int value = 1;   // Poisson: k_q * w ; isotropic elasticity: E_q * w
};
template <int NN, int DIM>
struct LawExtra<FEM_LAW_NEO_HOOKEAN, NN, DIM> {
  static constexpr int value = 2 * NN * DIM + 4;   // f_n = F g_n, h_n = F^-T g_n, c1..c4
};

template <int NN, int DIM, int VEC, int LAW>
struct Layout {
  static constexpr int NQ = NN;
  static constexpr int ND = NN * VEC;
  static constexpr int TAB_STRIDE = pad_odd(NN * DIM);      // per-q stride of the dN table (bank spread)
  static constexpr int TAB_SIZE = NQ * TAB_STRIDE + NQ;     // + quadrature weights
  static constexpr int OFF_G = 0;                           // g[NN][DIM]
  static constexpr int OFF_W = NN * DIM;                    // JxW
  static constexpr int OFF_S = OFF_W + 1;                   // S[VEC][DIM] = sigma * JxW
  static constexpr int OFF_X = OFF_S + VEC * DIM;           // law extras
  static constexpr int REC = pad_odd(OFF_X + LawExtra<LAW, NN, DIM>::value);
  static constexpr int OFF_XN = 0;                          // X[NN][DIM] node coordinates
  static constexpr int OFF_UN = NN * DIM;                   // U[NN][VEC] nodal solution
  static constexpr int OFF_REC = NN * DIM + NN * VEC;
  // Output staging: the row block is produced in HALVES column chunks of NB nodes.  The last chunk overlays
  // the head of the cell area (dead once the warp is past phase 2); with two chunks the first one needs its
  // own zone because the records are still live while the second chunk is computed.
  static constexpr int NB = pair_split<NN>();
  static constexpr int HALVES = NN / NB;
  static constexpr int CHUNK = NN * NB * VEC * VEC;          // doubles of one column chunk of a cell (all rows)
  static constexpr int OFF_STAGE = ((OFF_REC + NQ * REC + 1) / 2) * 2;
  static constexpr int CELL_RAW = OFF_STAGE + (HALVES == 2 ? CHUNK : 0);
  // cell stride == 8 (mod 16) doubles: the two cells of a half-warp hit disjoint bank sets
  static constexpr int CELL = CELL_RAW + ((8 - CELL_RAW % 16) + 16) % 16;
};

struct ElemArgs {
  const double* points;
  const int32_t* cells;
  const double* sol;
  const double* iv;      // (C, NQ) or nullptr
  const double* lam;     // adjoint vector (nodes, vec) for the parameter-gradient kernel
  const double* ref;     // [NQ*NN*DIM] dN, [NQ] w
  const int32_t* corner_pos;   // (C*NN) output row-block position of corner (c,a); nullptr = c*NN + a
  double* Ke;
  double* Re;
  double* grad;          // (C, NQ)
  int64_t C;
  double p[8];
};

// ---- the element kernel ----------------------------------------------------------------------------
template <int NN, int DIM, int VEC, int LAW, int CPB, bool JAC>
__global__ void __launch_bounds__(CPB* NN, (LAW == FEM_LAW_NEO_HOOKEAN ? 256 / (CPB * NN) : 512 / (CPB * NN))) element_kernel(const ElemArgs A) {
  using L = Layout<NN, DIM, VEC, LAW>;
  constexpr int NQ = L::NQ, ND = L::ND;
  extern __shared__ __align__(16) double sm[];
  double* tab = sm;
  const int lc = threadIdx.x / NN, lane = threadIdx.x % NN;
  double* cb = sm + L::TAB_SIZE + lc * L::CELL;
  const int64_t c = (int64_t)blockIdx.x * CPB + lc;
  const bool active = c < A.C;

  for (int i = threadIdx.x; i < NQ * NN * DIM; i += CPB * NN)
    tab[(i / (NN * DIM)) * L::TAB_STRIDE + i % (NN * DIM)] = A.ref[i];
  if (threadIdx.x < NQ) tab[NQ * L::TAB_STRIDE + threadIdx.x] = A.ref[NQ * NN * DIM + threadIdx.x];
  if (active) {
    const int64_t node = A.cells[c * NN + lane];
#pragma unroll
    for (int d = 0; d < DIM; ++d) cb[L::OFF_XN + lane * DIM + d] = A.points[node * DIM + d];
#pragma unroll
    for (int i = 0; i < VEC; ++i) cb[L::OFF_UN + lane * VEC + i] = A.sol[node * VEC + i];
  }
  __syncthreads();

  // ---------------- phase 1: lane = quadrature point ----------------
  if (active) {
    const int q = lane;
    double g[NN][DIM];
    const double w = qp_geometry<NN, DIM>(cb + L::OFF_XN, tab + q * L::TAB_STRIDE, tab[NQ * L::TAB_STRIDE + q], g);
    double ug[VEC][DIM];
    qp_grad_u<NN, DIM, VEC>(cb + L::OFF_UN, g, ug);
    double* rec = cb + L::OFF_REC + q * L::REC;
#pragma unroll
    for (int n = 0; n < NN; ++n)
#pragma unroll
      for (int d = 0; d < DIM; ++d) rec[L::OFF_G + n * DIM + d] = g[n][d];
    rec[L::OFF_W] = w;
    const double* ivq = A.iv ? A.iv + c * NQ + q : nullptr;

    if constexpr (LAW == FEM_LAW_POISSON) {
      const double kq = A.p[0] * (ivq ? *ivq : 1.0);
#pragma unroll
      for (int i = 0; i < VEC; ++i)
#pragma unroll
        for (int d = 0; d < DIM; ++d) rec[L::OFF_S + i * DIM + d] = kq * ug[i][d] * w;
      rec[L::OFF_X] = kq * w;
    } else if constexpr (LAW == FEM_LAW_LINEAR_ELASTIC || LAW == FEM_LAW_SIMP) {
      static_assert(VEC == DIM, "elasticity needs vec == dim");
      const double E = iso_modulus<LAW>(A.p, ivq, false), nu = iso_nu<LAW>(A.p);
      const double mu = E / (2.0 * (1.0 + nu)), lam = E * iso_lam1<LAW, DIM>(A.p, nu);
      double sig[DIM][DIM];
      iso_stress<DIM>(lam, mu, ug, sig);
#pragma unroll
      for (int i = 0; i < VEC; ++i)
#pragma unroll
        for (int d = 0; d < DIM; ++d) rec[L::OFF_S + i * DIM + d] = sig[i][d] * w;
      rec[L::OFF_X] = E * w;
    } else {   // Neo-Hookean
      static_assert(VEC == 3 && DIM == 3, "Neo-Hookean is 3-D");
      const double E = A.p[0] * (ivq ? *ivq : 1.0), nu = A.p[1];
      const double mu = E / (2.0 * (1.0 + nu)), kappa = E / (3.0 * (1.0 - 2.0 * nu));
      NHPoint k;
      nh_kinematics(ug, mu, A.p[2] != 0.0, k);
      double P[3][3];
      nh_stress(k, kappa, P);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int d = 0; d < 3; ++d) rec[L::OFF_S + i * 3 + d] = P[i][d] * w;
      if constexpr (JAC) {
        double* f = rec + L::OFF_X;
        double* h = f + NN * 3;
#pragma unroll
        for (int n = 0; n < NN; ++n)
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            double sf = 0.0, sh = 0.0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              sf = fma(k.F[i][j], g[n][j], sf);
              sh = fma(k.H[i][j], g[n][j], sh);
            }
            f[n * 3 + i] = sf;
            h[n * 3 + i] = sh;
          }
        double* cc = h + NN * 3;
        cc[0] = k.m * w;                                                             // delta_ik (g_a . g_b)
        cc[1] = -(2.0 / 3.0) * k.m * w;                                              // f_a h_b + h_a f_b
        cc[2] = ((2.0 / 9.0) * k.m * k.I1 + kappa * (2.0 * k.J - 1.0) * k.J) * w;    // h_a h_b
        cc[3] = (k.m * k.I1 / 3.0 - kappa * (k.J - 1.0) * k.J) * w;                  // h_b h_a (swapped)
      }
    }
  }
  __syncthreads();

  // ---------------- phase 2: lane = local node a ----------------
  // (no early return: every lane of the warp must reach the __syncwarp()s below)
  const int a = lane;
  const double* recs = cb + L::OFF_REC;
  if (active) {
    double r[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) r[i] = 0.0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double* rec = recs + q * L::REC;
#pragma unroll
      for (int i = 0; i < VEC; ++i)
#pragma unroll
        for (int d = 0; d < DIM; ++d) r[i] = fma(rec[L::OFF_S + i * DIM + d], rec[L::OFF_G + a * DIM + d], r[i]);   // problem.py:210
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) A.Re[c * ND + a * VEC + i] = r[i];
  }
  if constexpr (JAC) {
    // Row block (a, :) of the element tangent, NB column nodes at a time in registers
    // (K[j][i][k] = d r_(a,i) / d u_(b0+j,k)); NB = NN/2 keeps the kernel under 128 registers so that
    // several CTAs share an SM and hide each other's global-load and phase-change latencies.
    // Output: the row block of corner (c,a) is NN*VEC*VEC contiguous doubles at position corner_pos[c*NN+a];
    // with corner_pos = the node-sorted corner order of the plan, all row blocks of a mesh node are adjacent
    // in memory ("COO sorted by row"), which is what the CSR gather streams.
    constexpr int VV = VEC * VEC;
    constexpr int NB = L::NB, HALVES = L::HALVES;
    static_assert(L::CHUNK <= L::OFF_STAGE && (NB * VV) % 2 == 0, "output staging layout");
    const int opos = active ? (A.corner_pos ? A.corner_pos[c * NN + a] : (int)(c * NN + a)) : 0;
#pragma unroll 1
    for (int b0 = 0; b0 < NN; b0 += NB) {
      double K[NB][VEC][VEC];
#pragma unroll
      for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
          for (int k = 0; k < VEC; ++k) K[j][i][k] = 0.0;
      if (active) {
        if constexpr (LAW == FEM_LAW_POISSON) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const double* rec = recs + q * L::REC;
            double ga[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) ga[d] = rec[L::OFF_G + a * DIM + d] * rec[L::OFF_X];
#pragma unroll
            for (int j = 0; j < NB; ++j)
#pragma unroll
              for (int d = 0; d < DIM; ++d) K[j][0][0] = fma(ga[d], rec[L::OFF_G + (b0 + j) * DIM + d], K[j][0][0]);
          }
#pragma unroll
          for (int j = 0; j < NB; ++j)
#pragma unroll
            for (int i = 1; i < VEC; ++i) K[j][i][i] = K[j][0][0];
        } else if constexpr (LAW == FEM_LAW_LINEAR_ELASTIC || LAW == FEM_LAW_SIMP) {
          // K_ab[i][k] = lam' G[i][k] + mu' G[k][i] + mu' tr(G) delta_ik,  G = sum_q E_q w_q g_a (x) g_b
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const double* rec = recs + q * L::REC;
            double ga[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) ga[d] = rec[L::OFF_G + a * DIM + d] * rec[L::OFF_X];
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              double gb[DIM];
#pragma unroll
              for (int d = 0; d < DIM; ++d) gb[d] = rec[L::OFF_G + (b0 + j) * DIM + d];
#pragma unroll
              for (int i = 0; i < DIM; ++i)
#pragma unroll
                for (int k = 0; k < DIM; ++k) K[j][i][k] = fma(ga[i], gb[k], K[j][i][k]);
            }
          }
          const double nu = iso_nu<LAW>(A.p);
          const double mu1 = 1.0 / (2.0 * (1.0 + nu)), lam1 = iso_lam1<LAW, DIM>(A.p, nu);
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            double G[DIM][DIM];
            double tr = 0.0;
#pragma unroll
            for (int i = 0; i < DIM; ++i) {
              tr += K[j][i][i];
#pragma unroll
              for (int k = 0; k < DIM; ++k) G[i][k] = K[j][i][k];
            }
#pragma unroll
            for (int i = 0; i < DIM; ++i)
#pragma unroll
              for (int k = 0; k < DIM; ++k) K[j][i][k] = lam1 * G[i][k] + mu1 * G[k][i] + (i == k ? mu1 * tr : 0.0);
          }
        } else {
          // Neo-Hookean: K_ab[i][k] = c1 (g_a.g_b) d_ik + (c2 f_a + c3 h_a)_i h_b[k] + c2 h_a[i] f_b[k] + c4 h_b[i] h_a[k]
#pragma unroll 1
          for (int q = 0; q < NQ; ++q) {
            const double* rec = recs + q * L::REC;
            const double* f = rec + L::OFF_X;
            const double* h = f + NN * 3;
            const double* cc = h + NN * 3;
            const double c1 = cc[0], c2 = cc[1], c3 = cc[2], c4 = cc[3];
            double ga[3], pa[3], qa[3], ra[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const double fa = f[a * 3 + i], ha = h[a * 3 + i];
              ga[i] = c1 * rec[L::OFF_G + a * 3 + i];
              pa[i] = c2 * fa + c3 * ha;
              qa[i] = c2 * ha;
              ra[i] = c4 * ha;
            }
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              const int b = b0 + j;
              double gb[3], fb[3], hb[3];
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                gb[i] = rec[L::OFF_G + b * 3 + i];
                fb[i] = f[b * 3 + i];
                hb[i] = h[b * 3 + i];
              }
              const double dgg = ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2];
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                  double v = K[j][i][k];
                  v = fma(pa[i], hb[k], v);
                  v = fma(qa[i], fb[k], v);
                  v = fma(hb[i], ra[k], v);
                  K[j][i][k] = v;
                }
#pragma unroll
              for (int i = 0; i < 3; ++i) K[j][i][i] += dgg;
            }
          }
        }
      }
      // stage this column chunk in shared memory: [row a][NB blocks][VEC][VEC]
      const bool last = (b0 + NB >= NN);
      if (last) __syncwarp();                        // every lane of the warp is done reading X,U and the records
      if (active) {
        double* dst = cb + ((HALVES == 2 && !last) ? L::OFF_STAGE : 0) + a * (NB * VV);
#pragma unroll
        for (int t = 0; t < NB * VV / 2; ++t) {
          const int e0 = 2 * t, e1 = 2 * t + 1;
          reinterpret_cast<double2*>(dst)[t] =
              make_double2(K[e0 / VV][(e0 % VV) / VEC][e0 % VEC], K[e1 / VV][(e1 % VV) / VEC][e1 % VEC]);
        }
      }
    }
    __syncwarp();
    // coalesced copy-out: per cell, NN rows x HALVES chunks of NB*VV contiguous doubles each; the row block of
    // corner (c, a) goes to position corner_pos[c*NN + a] (held by lane a of the cell -> warp shuffle)
    constexpr int CPW = 32 / NN;                        // cells per warp
    constexpr int P2 = NB * VV / 2;                     // double2 pieces of one (row, chunk)
    const int wl = threadIdx.x & 31;
    const int lc0 = (threadIdx.x >> 5) * CPW;
#pragma unroll 1
    for (int j = 0; j < CPW; ++j) {
      const int64_t cj = (int64_t)blockIdx.x * CPB + lc0 + j;
      const bool live = cj < A.C;                       // warp-uniform
      const double* cellp = sm + L::TAB_SIZE + (lc0 + j) * L::CELL;
#pragma unroll
      for (int it = 0; it < (NN * HALVES * P2 + 31) / 32; ++it) {
        const int p = it * 32 + wl;
        const int row = (p / P2) / HALVES, h = (p / P2) % HALVES, w2 = p % P2;
        const int pos = __shfl_sync(0xffffffffu, opos, (j * NN + row) & 31);
        if (live && p < NN * HALVES * P2) {
          const double* srcp = cellp + ((HALVES == 2 && h == 0) ? L::OFF_STAGE : 0) + row * (NB * VV);
          double* dstp = A.Ke + (int64_t)pos * (NN * VV) + h * (NB * VV);
          reinterpret_cast<double2*>(dstp)[w2] = reinterpret_cast<const double2*>(srcp)[w2];
        }
      }
    }
  }
}

// ---- HEX8 isotropic elasticity on the FP64 tensor cores (DMMA) -------------------------------------
// The DFMA kernel above is bounded by shared->register bandwidth: G_ab += (E w g_a) (x) g_b needs one loaded
// double per 2.25 FMA while the SM sustains one per 4.  mma.sync.m8n8k4.f64 contracts over the quadrature points
// with the operands DISTRIBUTED over the warp's registers: with the 24 element dofs ordered (component, node) the
// 24x24 matrix  G = sum_q (E_q w_q g(q)) g(q)^T  is a 3x3 grid of 8x8 tiles T_IJ[a][b] = G_ab[I][J], every tile is
// two k-steps (8 quadrature points), and lane l = (node n = l/4, t = l%4) supplies g_n(q=t) and g_n(q=t+4) as BOTH
// the A row and the B column -- 8 shared loads per lane and cell instead of ~350.  DMMA and DFMA share one pipe on
// B200 (tools/microbench.cu), so this saves operand bandwidth, not flops.  The element residual
// r_a[i] = sum_{q,d} g_a(q)[d] S_q[i][d] reuses the same A fragments against S as the B operand (6 more DMMA).
// One warp owns 4 cells: phase 1 (lane = (cell, q)) is the same geometry/stress code as above; then the warp
// walks its cells one at a time: 18 + 6 DMMA, G -> K in registers (lane (n,t) holds K_{n,2t} and K_{n,2t+1}:
// 18 contiguous doubles of row block n), staging in shared memory and a coalesced copy to the row block's
// node-sorted position.

struct DmmaLayout {
  static constexpr int NN = 8, NQ = 8, DIM = 3, VEC = 3;
  static constexpr int TAB_STRIDE = 25, TAB_SIZE = NQ * TAB_STRIDE + NQ;      // 208
  static constexpr int GS = 28;                    // per-q stride of g[n][d]: 12 (mod 16) => conflict-free fragments
  static constexpr int OFF_X = 0, OFF_U = 24;      // per cell: X[8][3], U[8][3]
  static constexpr int OFF_G = 48;                 // g[q][n][d]
  static constexpr int OFF_S = OFF_G + NQ * GS;    // S[q][i][d] = sigma JxW        (272)
  static constexpr int OFF_E = OFF_S + NQ * 9;     // E_q JxW                       (344)
  static constexpr int CELL = 354;                 // 352 + 2: cell stride 2 (mod 16) spreads phase-1 stores
  static constexpr int OFF_SPARE = 4 * CELL + 16;  // after the 32 ints of corner positions
  static constexpr int WARP = OFF_SPARE + 4 * 72;  // the row blocks of a cell leave through two staging areas of 4 rows:
                                                   // rows 0-3 in the cell's own (by then dead) area, rows 4-7 in the spare
  static constexpr int WARPS = 4;
};

// TILES = true: the accumulator fragments go straight from registers to the corner's row in TILE-MAJOR layout -- row (c, a) =
// 9 tiles (I, J) of 8 doubles (b = 0..7) holding G_ab[I][J] = sum_q E_q w_q g_a[I] g_b[J]; lane (a, t) writes 16 bytes of
// every tile, so each warp-wide store is 8 full 64-byte pieces: no shared-memory staging, no bulk copies, and the isotropic map
// K = lam' G + mu' G^T + mu' tr(G) I is applied by the CSR gather AFTER the sum over the cells (it is linear; 2.35x fewer
// blocks to transform).  TILES = false: K blocks in the reference's V layout through the staged bulk copies described above.
template <int LAW, bool TILES>
__global__ void __launch_bounds__(DmmaLayout::WARPS * 32, 4) element_dmma_kernel(const ElemArgs A) {
  using L = DmmaLayout;
  constexpr int NN = 8, NQ = 8, DIM = 3, VEC = 3, ND = 24;
  extern __shared__ __align__(16) double sm[];
  double* tab = sm;
  const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
  double* wb = sm + L::TAB_SIZE + warp * L::WARP;
  int* pos = reinterpret_cast<int*>(wb + 4 * L::CELL);
  for (int i = threadIdx.x; i < NQ * NN * DIM; i += L::WARPS * 32)
    tab[(i / (NN * DIM)) * L::TAB_STRIDE + i % (NN * DIM)] = A.ref[i];
  if (threadIdx.x < NQ) tab[NQ * L::TAB_STRIDE + threadIdx.x] = A.ref[NQ * NN * DIM + threadIdx.x];

  // ---- phase 0/1: lane = (cell j, q) ----
  const int j1 = l >> 3, q = l & 7;
  const int64_t c1 = ((int64_t)blockIdx.x * L::WARPS + warp) * 4 + j1;
  const bool act1 = c1 < A.C;
  double* cb = wb + j1 * L::CELL;
  if (act1) {
    const int64_t node = A.cells[c1 * NN + q];
#pragma unroll
    for (int d = 0; d < DIM; ++d) cb[L::OFF_X + q * DIM + d] = A.points[node * DIM + d];
#pragma unroll
    for (int i = 0; i < VEC; ++i) cb[L::OFF_U + q * VEC + i] = A.sol[node * VEC + i];
    pos[l] = A.corner_pos ? A.corner_pos[c1 * NN + q] : (int)(c1 * NN + q);
  }
  __syncthreads();
  if (act1) {
    double g[NN][DIM];
    const double w = qp_geometry<NN, DIM>(cb + L::OFF_X, tab + q * L::TAB_STRIDE, tab[NQ * L::TAB_STRIDE + q], g);
    double ug[VEC][DIM];
    qp_grad_u<NN, DIM, VEC>(cb + L::OFF_U, g, ug);
    const double* ivq = A.iv ? A.iv + c1 * NQ + q : nullptr;
    const double E = iso_modulus<LAW>(A.p, ivq, false), nu = iso_nu<LAW>(A.p);
    const double mu = E / (2.0 * (1.0 + nu)), lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
    double sig[DIM][DIM];
    iso_stress<DIM>(lam, mu, ug, sig);
#pragma unroll
    for (int n = 0; n < NN; ++n)
#pragma unroll
      for (int d = 0; d < DIM; ++d) cb[L::OFF_G + q * L::GS + n * DIM + d] = g[n][d];
#pragma unroll
    for (int i = 0; i < VEC; ++i)
#pragma unroll
      for (int d = 0; d < DIM; ++d) cb[L::OFF_S + q * 9 + i * DIM + d] = sig[i][d] * w;
    cb[L::OFF_E + q] = E * w;
  }
  __syncwarp();

  // ---- phase 2: the warp walks its 4 cells; lane = (node n, t) ----
  const int n = l >> 2, t = l & 3;
  const double nu = iso_nu<LAW>(A.p);
  const double mu1 = 1.0 / (2.0 * (1.0 + nu)), lam1 = nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
    const int64_t c = ((int64_t)blockIdx.x * L::WARPS + warp) * 4 + j;
    if (c >= A.C) break;                                   // warp-uniform
    const double* cj = wb + j * L::CELL;
    double g0[3], g1[3], a0[3], a1[3];
    const double e0 = cj[L::OFF_E + t], e1 = cj[L::OFF_E + t + 4];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      g0[d] = cj[L::OFF_G + t * L::GS + n * 3 + d];
      g1[d] = cj[L::OFF_G + (t + 4) * L::GS + n * 3 + d];
      a0[d] = e0 * g0[d];
      a1[d] = e1 * g1[d];
    }
    double C[3][3][2];
#pragma unroll
    for (int I = 0; I < 3; ++I)
#pragma unroll
      for (int J = 0; J < 3; ++J) {
        C[I][J][0] = C[I][J][1] = 0.0;
        dmma884(C[I][J], a0[I], g0[J]);
        dmma884(C[I][J], a1[I], g1[J]);
        if constexpr (TILES)
          *reinterpret_cast<double2*>(A.Ke + (int64_t)pos[j * 8 + n] * 72 + (I * 3 + J) * 8 + 2 * t) = make_double2(C[I][J][0], C[I][J][1]);
      }
    // residual: B[q][col] = S_q[i = col][d] for col < 3 (col = l/4), zero otherwise
    double R[2] = {0.0, 0.0};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double b0 = (n < 3) ? cj[L::OFF_S + t * 9 + n * 3 + d] : 0.0;
      const double b1 = (n < 3) ? cj[L::OFF_S + (t + 4) * 9 + n * 3 + d] : 0.0;
      dmma884(R, g0[d], b0);
      dmma884(R, g1[d], b1);
    }
    // R = r_n[2t], r_n[2t+1]: components 0,1 live in lanes t == 0, component 2 in lanes t == 1
    if (t == 0) {
      A.Re[c * ND + n * 3 + 0] = R[0];
      A.Re[c * ND + n * 3 + 1] = R[1];
    } else if (t == 1) {
      A.Re[c * ND + n * 3 + 2] = R[0];
    }
    if constexpr (TILES) continue;
    // G -> K for the lane's two column nodes b = 2t, 2t+1: 18 contiguous doubles of row block n
    double K[2][9];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const double tr = C[0][0][e] + C[1][1][e] + C[2][2][e];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) K[e][i * 3 + k] = lam1 * C[i][k][e] + mu1 * C[k][i][e] + (i == k ? mu1 * tr : 0.0);
    }
    // stage 4 row blocks at a time -- rows 0-3 in the cell's own area (its g / S / E are in registers by now), rows 4-7
    // in the warp's spare area -- and hand every 576-byte row block to the TMA unit (cp.async.bulk shared -> global) for
    // its node-sorted position: no LDS / STG for the copy-out.  The spare area was last used one cell earlier: all but
    // the most recent bulk group of the issuing lanes must have been read before it is overwritten.
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      double* out = half == 0 ? wb + j * L::CELL : wb + L::OFF_SPARE;
      if (half == 1 && l < 4) bulk_wait_read<1>();
      __syncwarp();
      if ((n >> 2) == half) {
        double* dst = out + (n & 3) * 72 + t * 18;
#pragma unroll
        for (int m = 0; m < 9; ++m) reinterpret_cast<double2*>(dst)[m] = make_double2(K[(2 * m) / 9][(2 * m) % 9], K[(2 * m + 1) / 9][(2 * m + 1) % 9]);
        fence_async_smem();
      }
      __syncwarp();
      if (l < 4) {
        bulk_s2g(A.Ke + (int64_t)pos[j * 8 + half * 4 + l] * 72, out + l * 72, 72 * sizeof(double));
        bulk_commit();
      }
    }
  }
  if (!TILES && l < 4) bulk_wait_read<0>();          // shared memory must outlive the outstanding bulk reads
}

template <int LAW, bool TILES = false>
int launch_element_dmma(const ElemArgs& A, cudaStream_t st) {
  using L = DmmaLayout;
  const size_t smem = sizeof(double) * (L::TAB_SIZE + (size_t)L::WARPS * L::WARP);
  const unsigned grid = (unsigned)((A.C + L::WARPS * 4 - 1) / (L::WARPS * 4));
  auto k = element_dmma_kernel<LAW, TILES>;
  FEM_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, L::WARPS * 32, smem, st>>>(A);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

// ---- HEX8 Neo-Hookean on the FP64 tensor cores ------------------------------------------------------------
// K_ab[i][k] = sum_q  p_a[i] h_b[k] + q_a[i] f_b[k] + h_b[i] r_a[k] + delta_ik (c1 g_a . g_b)   (closed form of section 8a,
// p = c2 f + c3 h, q = c2 h, r = c4 h, f = F g, h = F^-T g): every term is a sum over the quadrature points of an
// outer product of per-node values, i.e. the same 8x8 (a, b) tiles contracted over q as in element_dmma_kernel --
// 3 x 18 + 6 mma.sync.m8n8k4.f64 per cell for the tangent and 6 for the residual, operands in registers (26 loads per
// lane and cell instead of ~1000 shared loads in the CUDA-core formulation).  Phase 1 (lane = (cell, q)) is the
// CUDA-core code of element_kernel; it also stores f and h.  One warp owns 4 cells; 2 CTAs of 4 warps per SM.
struct NhDmmaLayout {
  static constexpr int NN = 8, NQ = 8, DIM = 3, VEC = 3;
  static constexpr int TAB_STRIDE = 25, TAB_SIZE = NQ * TAB_STRIDE + NQ;
  static constexpr int GS = 28;                        // per-q stride of g[n][d] (conflict-free fragments)
  static constexpr int OFF_X = 0, OFF_U = 24;          // X[8][3], U[8][3]: dead once every lane of the cell holds g and grad u,
  static constexpr int OFF_G = 0;                      // g[q][n][d] is written over them
  static constexpr int OFF_FH = OFF_G + NQ * GS;       // F[q][3][3], H[q][3][3] = F^-T (f = F g and h = H g are formed in phase 2)
  static constexpr int OFF_C = OFF_FH + NQ * 18;       // c1..c4 per q
  static constexpr int CELL = OFF_C + NQ * 4 + 10;     // 410 = 10 (mod 16); >= 4 * 72 doubles of output staging; 4 CTAs per SM
  static constexpr int WARP = 4 * CELL + 16;           // + 32 ints of corner positions
  static constexpr int WARPS = 4;
};

template <bool TILES>
__global__ void __launch_bounds__(NhDmmaLayout::WARPS * 32, 4) element_nh_dmma_kernel(const ElemArgs A) {
  using L = NhDmmaLayout;
  constexpr int NN = 8, NQ = 8, DIM = 3, VEC = 3, ND = 24;
  extern __shared__ __align__(16) double sm[];
  double* tab = sm;
  const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
  double* wb = sm + L::TAB_SIZE + warp * L::WARP;
  int* pos = reinterpret_cast<int*>(wb + 4 * L::CELL);
  for (int i = threadIdx.x; i < NQ * NN * DIM; i += L::WARPS * 32)
    tab[(i / (NN * DIM)) * L::TAB_STRIDE + i % (NN * DIM)] = A.ref[i];
  if (threadIdx.x < NQ) tab[NQ * L::TAB_STRIDE + threadIdx.x] = A.ref[NQ * NN * DIM + threadIdx.x];

  // ---- phase 0/1: lane = (cell j, q) ----
  const int j1 = l >> 3, q = l & 7;
  const int64_t c1 = ((int64_t)blockIdx.x * L::WARPS + warp) * 4 + j1;
  const bool act1 = c1 < A.C;
  double* cb = wb + j1 * L::CELL;
  if (act1) {
    const int64_t node = A.cells[c1 * NN + q];
#pragma unroll
    for (int d = 0; d < DIM; ++d) cb[L::OFF_X + q * DIM + d] = A.points[node * DIM + d];
#pragma unroll
    for (int i = 0; i < VEC; ++i) cb[L::OFF_U + q * VEC + i] = A.sol[node * VEC + i];
    pos[l] = A.corner_pos ? A.corner_pos[c1 * NN + q] : (int)(c1 * NN + q);
  }
  __syncthreads();
  double g[NN][DIM], ug[VEC][DIM], w = 0.0;
  if (act1) {
    w = qp_geometry<NN, DIM>(cb + L::OFF_X, tab + q * L::TAB_STRIDE, tab[NQ * L::TAB_STRIDE + q], g);
    qp_grad_u<NN, DIM, VEC>(cb + L::OFF_U, g, ug);
  }
  __syncwarp();                                            // X / U of the warp's cells are consumed: g goes on top of them
  if (act1) {
    const double* ivq = A.iv ? A.iv + c1 * NQ + q : nullptr;
    const double E = A.p[0] * (ivq ? *ivq : 1.0), nu = A.p[1];
    const double mu = E / (2.0 * (1.0 + nu)), kappa = E / (3.0 * (1.0 - 2.0 * nu));
    NHPoint k;
    nh_kinematics(ug, mu, A.p[2] != 0.0, k);
#pragma unroll
    for (int n = 0; n < NN; ++n)
#pragma unroll
      for (int i = 0; i < 3; ++i) cb[L::OFF_G + q * L::GS + n * 3 + i] = g[n][i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        cb[L::OFF_FH + q * 18 + i * 3 + j] = k.F[i][j];
        cb[L::OFF_FH + q * 18 + 9 + i * 3 + j] = k.H[i][j];
      }
    double* cc = cb + L::OFF_C + q * 4;
    cc[0] = k.m * w;                                                             // delta_ik (g_a . g_b); also P JxW = c1 F - c4 H
    cc[1] = -(2.0 / 3.0) * k.m * w;                                              // f_a h_b + h_a f_b
    cc[2] = ((2.0 / 9.0) * k.m * k.I1 + kappa * (2.0 * k.J - 1.0) * k.J) * w;    // h_a h_b
    cc[3] = (k.m * k.I1 / 3.0 - kappa * (k.J - 1.0) * k.J) * w;                  // h_b h_a (swapped)
  }
  __syncwarp();

  // ---- phase 2: the warp walks its 4 cells; lane = (node n, t) ----
  const int n = l >> 2, t = l & 3;
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
    const int64_t c = ((int64_t)blockIdx.x * L::WARPS + warp) * 4 + j;
    if (c >= A.C) break;                                   // warp-uniform
    const double* cj = wb + j * L::CELL;
    double gq[2][3], fq[2][3], hq[2][3], ga[2][3], pa[2][3], qa[2][3], ra[2][3];
    double R[3] = {0.0, 0.0, 0.0};                          // P JxW g_a = c1 f_a - c4 h_a, summed over the lane's two points
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int qq = t + 4 * s;
      const double* cc = cj + L::OFF_C + qq * 4;
      const double k1 = cc[0], k2 = cc[1], k3 = cc[2], k4 = cc[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) gq[s][d] = cj[L::OFF_G + qq * L::GS + n * 3 + d];
      const double* FH = cj + L::OFF_FH + qq * 18;        // broadcast: the 8 lanes of a q read the same 18 doubles
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        fq[s][i] = FH[i * 3] * gq[s][0] + FH[i * 3 + 1] * gq[s][1] + FH[i * 3 + 2] * gq[s][2];
        hq[s][i] = FH[9 + i * 3] * gq[s][0] + FH[9 + i * 3 + 1] * gq[s][1] + FH[9 + i * 3 + 2] * gq[s][2];
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        ga[s][d] = k1 * gq[s][d];
        pa[s][d] = k2 * fq[s][d] + k3 * hq[s][d];
        qa[s][d] = k2 * hq[s][d];
        ra[s][d] = k4 * hq[s][d];
        R[d] += fma(k1, fq[s][d], -ra[s][d]);
      }
    }
    double D[2] = {0.0, 0.0};
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int d = 0; d < 3; ++d) dmma884(D, ga[s][d], gq[s][d]);
    // The tangent is symmetric, K_ab[I][J] = K_ba[J][I]: with tile-major staging only the 6 tiles I <= J are computed (36 instead
    // of 54 DMMA -- this kernel is FP64-bound); tile (J, I) is the transpose of tile (I, J) and is written by the lanes that hold
    // it: lane (r, t) stores element [r][2t] to row 2t, column r and [r][2t+1] to row 2t+1, column r (8 lanes = 64 contiguous bytes).
    double C[3][3][2];
#pragma unroll
    for (int I = 0; I < 3; ++I)
#pragma unroll
      for (int J = 0; J < 3; ++J) {
        if (TILES && J < I) continue;
        C[I][J][0] = (I == J) ? D[0] : 0.0;
        C[I][J][1] = (I == J) ? D[1] : 0.0;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          dmma884(C[I][J], pa[s][I], hq[s][J]);
          dmma884(C[I][J], qa[s][I], fq[s][J]);
          dmma884(C[I][J], ra[s][J], hq[s][I]);
        }
      }
    // residual r_n[i] = sum over the 8 points: the 4 lanes of node n hold two points each
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      R[d] += __shfl_xor_sync(0xffffffffu, R[d], 1);
      R[d] += __shfl_xor_sync(0xffffffffu, R[d], 2);
    }
    if (t == 0) {
      A.Re[c * ND + n * 3 + 0] = R[0];
      A.Re[c * ND + n * 3 + 1] = R[1];
    } else if (t == 1) {
      A.Re[c * ND + n * 3 + 2] = R[2];
    }
    if constexpr (TILES) {
      // tile-major row (9 tiles of 8 doubles): the fragments ARE the tiles, 16 bytes per lane and tile, no staging
      double* row = A.Ke + (int64_t)pos[j * 8 + n] * 72 + 2 * t;
      double* rowT0 = A.Ke + (int64_t)pos[j * 8 + 2 * t] * 72 + n;          // transposed element [n][2t]   -> row 2t,   column n
      double* rowT1 = A.Ke + (int64_t)pos[j * 8 + 2 * t + 1] * 72 + n;      // transposed element [n][2t+1] -> row 2t+1, column n
#pragma unroll
      for (int I = 0; I < 3; ++I)
#pragma unroll
        for (int J = I; J < 3; ++J) {
          *reinterpret_cast<double2*>(row + (I * 3 + J) * 8) = make_double2(C[I][J][0], C[I][J][1]);
          if (J > I) {
            rowT0[(J * 3 + I) * 8] = C[I][J][0];
            rowT1[(J * 3 + I) * 8] = C[I][J][1];
          }
        }
      continue;
    }
    // lane (n, t) holds K_{n,2t} and K_{n,2t+1}: 18 contiguous doubles of row block n; stage 4 row blocks at a time in
    // the cell's own (dead) area and copy them to their node-sorted positions with coalesced 16-byte stores
    double* out = wb + j * L::CELL;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      __syncwarp();
      if ((n >> 2) == half) {
        double* dst = out + (n & 3) * 72 + t * 18;
#pragma unroll
        for (int m = 0; m < 9; ++m) {
          const int e0 = 2 * m, e1 = 2 * m + 1;          // element e of the 18: block e / 9, (i, k) = (e % 9) / 3, e % 3
          reinterpret_cast<double2*>(dst)[m] = make_double2(C[(e0 % 9) / 3][e0 % 3][e0 / 9], C[(e1 % 9) / 3][e1 % 3][e1 / 9]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 5; ++it) {
        const int p2 = it * 32 + l;                     // 4 rows x 36 double2
        if (p2 < 144) {
          const int row = p2 / 36, w2 = p2 % 36;
          reinterpret_cast<double2*>(A.Ke + (int64_t)pos[j * 8 + half * 4 + row] * 72)[w2] =
              reinterpret_cast<const double2*>(out + row * 72)[w2];
        }
      }
    }
  }
}

template <bool TILES = false>
int launch_element_nh_dmma(const ElemArgs& A, cudaStream_t st) {
  using L = NhDmmaLayout;
  const size_t smem = sizeof(double) * (L::TAB_SIZE + (size_t)L::WARPS * L::WARP);
  const unsigned grid = (unsigned)((A.C + L::WARPS * 4 - 1) / (L::WARPS * 4));
  FEM_CUDA_CHECK(cudaFuncSetAttribute(element_nh_dmma_kernel<TILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  element_nh_dmma_kernel<TILES><<<grid, L::WARPS * 32, smem, st>>>(A);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

// ---- adjoint: -lambda^T dc/dtheta per quadrature point (thread per (cell, q)) -------------------
template <int NN, int DIM, int VEC, int LAW, int CPB>
__global__ void __launch_bounds__(CPB* NN) param_grad_kernel(const ElemArgs A) {
  using L = Layout<NN, DIM, VEC, LAW>;
  constexpr int NQ = L::NQ;
  extern __shared__ double sm[];
  double* tab = sm;
  const int lc = threadIdx.x / NN, q = threadIdx.x % NN;
  // per cell: X, U, LAM
  double* cb = sm + L::TAB_SIZE + lc * (NN * DIM + 2 * NN * VEC);
  const int64_t c = (int64_t)blockIdx.x * CPB + lc;
  const bool active = c < A.C;
  for (int i = threadIdx.x; i < NQ * NN * DIM; i += CPB * NN)
    tab[(i / (NN * DIM)) * L::TAB_STRIDE + i % (NN * DIM)] = A.ref[i];
  if (threadIdx.x < NQ) tab[NQ * L::TAB_STRIDE + threadIdx.x] = A.ref[NQ * NN * DIM + threadIdx.x];
  double* Xs = cb;
  double* Us = cb + NN * DIM;
  double* Ls = Us + NN * VEC;
  if (active) {
    const int64_t node = A.cells[c * NN + q];
#pragma unroll
    for (int d = 0; d < DIM; ++d) Xs[q * DIM + d] = A.points[node * DIM + d];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      Us[q * VEC + i] = A.sol[node * VEC + i];
      Ls[q * VEC + i] = A.lam[node * VEC + i];
    }
  }
  __syncthreads();
  if (!active) return;
  double g[NN][DIM];
  const double w = qp_geometry<NN, DIM>(Xs, tab + q * L::TAB_STRIDE, tab[NQ * L::TAB_STRIDE + q], g);
  double ug[VEC][DIM];
  qp_grad_u<NN, DIM, VEC>(Us, g, ug);
  const double* ivq = A.iv + c * NQ + q;
  double ds[VEC][DIM];
  if constexpr (LAW == FEM_LAW_POISSON) {
#pragma unroll
    for (int i = 0; i < VEC; ++i)
#pragma unroll
      for (int d = 0; d < DIM; ++d) ds[i][d] = A.p[0] * ug[i][d];
  } else if constexpr (LAW == FEM_LAW_LINEAR_ELASTIC || LAW == FEM_LAW_SIMP) {
    const double dE = iso_modulus<LAW>(A.p, ivq, true), nu = iso_nu<LAW>(A.p);
    iso_stress<DIM>(dE * iso_lam1<LAW, DIM>(A.p, nu), dE / (2.0 * (1.0 + nu)), ug, ds);
  } else {
    const double E = A.p[0], nu = A.p[1];      // P is linear in E: dP/drho = P(rho = 1)
    NHPoint k;
    nh_kinematics(ug, E / (2.0 * (1.0 + nu)), A.p[2] != 0.0, k);
    nh_stress(k, E / (3.0 * (1.0 - 2.0 * nu)), ds);
  }
  double acc = 0.0;
#pragma unroll
  for (int n = 0; n < NN; ++n)
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      double t = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) t = fma(ds[i][d], g[n][d], t);
      acc = fma(Ls[n * VEC + i], t, acc);
    }
  A.grad[c * NQ + q] = -acc * w;
}

template <int NN, int DIM, int VEC, int LAW, int CPB>
int launch_element(const ElemArgs& A, cudaStream_t st) {
  using L = Layout<NN, DIM, VEC, LAW>;
  const size_t smem = sizeof(double) * (L::TAB_SIZE + (size_t)CPB * L::CELL);
  const unsigned grid = (unsigned)((A.C + CPB - 1) / CPB);
  if (A.Ke) {
    auto k = element_kernel<NN, DIM, VEC, LAW, CPB, true>;
    FEM_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, CPB * NN, smem, st>>>(A);
  } else {
    auto k = element_kernel<NN, DIM, VEC, LAW, CPB, false>;
    FEM_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, CPB * NN, smem, st>>>(A);
  }
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

template <int NN, int DIM, int VEC, int LAW, int CPB>
int launch_param_grad(const ElemArgs& A, cudaStream_t st) {
  using L = Layout<NN, DIM, VEC, LAW>;
  const size_t smem = sizeof(double) * (L::TAB_SIZE + (size_t)CPB * (NN * DIM + 2 * NN * VEC));
  const unsigned grid = (unsigned)((A.C + CPB - 1) / CPB);
  auto k = param_grad_kernel<NN, DIM, VEC, LAW, CPB>;
  FEM_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, CPB * NN, smem, st>>>(A);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

// Registry of (element, vec, law) combinations.  Anything else is an error, never a fallback.
template <bool GRAD>
int dispatch(int ele, int vec, int law, const ElemArgs& A, cudaStream_t st) {
  if constexpr (!GRAD) {
    // tensor-core path for the headline configuration (FEM_ELEMENT_PATH=dfma in the environment selects the
    // CUDA-core kernel instead, for A/B measurements; both are parity-tested)
    static const bool use_dmma = []() { const char* e = getenv("FEM_ELEMENT_PATH"); return !(e && e[0] == 'd' && e[1] == 'f'); }();
    if (use_dmma && A.Ke && ele == FEM_ELE_HEX8 && vec == 3) {
      if (law == FEM_LAW_LINEAR_ELASTIC) return launch_element_dmma<FEM_LAW_LINEAR_ELASTIC>(A, st);
      if (law == FEM_LAW_SIMP) return launch_element_dmma<FEM_LAW_SIMP>(A, st);
      if (law == FEM_LAW_NEO_HOOKEAN) return launch_element_nh_dmma<false>(A, st);
    }
  }
#define FEM_CASE(ELE, NN, DIM, VEC, LAW, CPB)                                     \
  if (ele == ELE && vec == VEC && law == LAW) {                                   \
    if constexpr (GRAD) return launch_param_grad<NN, DIM, VEC, LAW, CPB>(A, st);  \
    else return launch_element<NN, DIM, VEC, LAW, CPB>(A, st);                    \
  }
  FEM_CASE(FEM_ELE_HEX8, 8, 3, 1, FEM_LAW_POISSON, 16)
  FEM_CASE(FEM_ELE_HEX8, 8, 3, 3, FEM_LAW_LINEAR_ELASTIC, 8)
  FEM_CASE(FEM_ELE_HEX8, 8, 3, 3, FEM_LAW_SIMP, 8)
  FEM_CASE(FEM_ELE_HEX8, 8, 3, 3, FEM_LAW_NEO_HOOKEAN, 8)
  FEM_CASE(FEM_ELE_QUAD4, 4, 2, 1, FEM_LAW_POISSON, 32)
  FEM_CASE(FEM_ELE_QUAD4, 4, 2, 2, FEM_LAW_LINEAR_ELASTIC, 32)
  FEM_CASE(FEM_ELE_QUAD4, 4, 2, 2, FEM_LAW_SIMP, 32)
#undef FEM_CASE
  set_error("unregistered (element=%d, vec=%d, law=%d) combination: no kernel, no fallback", ele, vec, law);
  return FEM_EINVAL;
}

}  // namespace
}  // namespace femb200

using namespace femb200;

extern "C" int fem_element_residual_jacobian(int ele_type, int vec, int law_id, const double* law_params_host,
                                             const double* points, const int32_t* cells, int64_t n_cells,
                                             const double* sol, const double* internal_var,
                                             const double* ref_tables, const int32_t* corner_pos, double* Ke, double* Re,
                                             void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(points && cells && sol && ref_tables && Re && law_params_host, "null pointer");
  FEM_REQUIRE(n_cells >= 0, "n_cells < 0");
  FEM_REQUIRE(!(law_id == FEM_LAW_SIMP && !internal_var), "SIMP needs the per-quadrature-point density");
  if (n_cells == 0) return FEM_OK;
  ElemArgs A{};
  A.points = points; A.cells = cells; A.sol = sol; A.iv = internal_var; A.ref = ref_tables;
  A.Ke = Ke; A.Re = Re; A.C = n_cells; A.corner_pos = corner_pos;
  for (int i = 0; i < 8; ++i) A.p[i] = law_params_host[i];
  return dispatch<false>(ele_type, vec, law_id, A, (cudaStream_t)stream);
}

extern "C" int fem_element_tiles(int law_id, const double* law_params_host, const double* points, const int32_t* cells,
                                 int64_t n_cells, const double* sol, const double* internal_var, const double* ref_tables,
                                 const int32_t* corner_pos, double* Ke_tiles, double* Re, double* post_host, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(points && cells && sol && ref_tables && Ke_tiles && Re && law_params_host && post_host, "null pointer");
  FEM_REQUIRE((reinterpret_cast<uintptr_t>(Ke_tiles) & 15) == 0, "Ke_tiles must be 16-byte aligned");
  FEM_REQUIRE(!(law_id == FEM_LAW_SIMP && !internal_var), "SIMP needs the per-quadrature-point density");
  post_host[0] = post_host[1] = post_host[2] = 0.0;
  if (n_cells <= 0) return FEM_OK;
  ElemArgs A{};
  A.points = points; A.cells = cells; A.sol = sol; A.iv = internal_var; A.ref = ref_tables;
  A.Ke = Ke_tiles; A.Re = Re; A.C = n_cells; A.corner_pos = corner_pos;
  for (int i = 0; i < 8; ++i) A.p[i] = law_params_host[i];
  cudaStream_t st = (cudaStream_t)stream;
  if (law_id == FEM_LAW_LINEAR_ELASTIC || law_id == FEM_LAW_SIMP) {
    const double nu = law_id == FEM_LAW_SIMP ? A.p[2] : A.p[1];
    post_host[0] = 1.0;                                          // the gather applies K = lam' G + mu' G^T + mu' tr(G) I
    post_host[1] = nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
    post_host[2] = 1.0 / (2.0 * (1.0 + nu));
    return law_id == FEM_LAW_SIMP ? launch_element_dmma<FEM_LAW_SIMP, true>(A, st) : launch_element_dmma<FEM_LAW_LINEAR_ELASTIC, true>(A, st);
  }
  if (law_id == FEM_LAW_NEO_HOOKEAN) return launch_element_nh_dmma<true>(A, st);     // tiles hold K itself
  set_error("fem_element_tiles: unregistered law %d (registered on HEX8 / vec 3: linear elasticity, SIMP, Neo-Hookean)", law_id);
  return FEM_EINVAL;
}

extern "C" int fem_adjoint_param_grad(int ele_type, int vec, int law_id, const double* law_params_host,
                                      const double* points, const int32_t* cells, int64_t n_cells,
                                      const double* sol, const double* internal_var, const double* lam,
                                      const double* ref_tables, double* grad, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(points && cells && sol && ref_tables && lam && grad && internal_var && law_params_host, "null pointer");
  if (n_cells == 0) return FEM_OK;
  ElemArgs A{};
  A.points = points; A.cells = cells; A.sol = sol; A.iv = internal_var; A.lam = lam; A.ref = ref_tables;
  A.grad = grad; A.C = n_cells;
  for (int i = 0; i < 8; ++i) A.p[i] = law_params_host[i];
  return dispatch<true>(ele_type, vec, law_id, A, (cudaStream_t)stream);
}
