// Fused owner-computes assembly: element evaluation + global CSR assembly + Dirichlet rows + nodal residual in ONE
// kernel, without the element tangents ever touching HBM.
//
// Replaces, for HEX8 isotropic elasticity (linear, SIMP), the chain
//   Problem.compute_newton_vars (jax_fem/problem.py:447-460: get_laplace_kernel :189-214 + value_and_jacfwd :262-266)
//   -> problem.V -> _PetscTangentCache.update / get_A (jax_fem/solver.py:469-553: setValuesCOO + zeroRows)
//   and compute_residual_vars_helper (jax_fem/problem.py:426-437)
// which the two-kernel path (element.cu + sparse.cu::gather_csr_kernel) runs through a C*8 x 8 x 3 x 3 staging
// buffer (4.6 GB written and read again at 100^3).
//
// One persistent CTA per SM walks over *patches* (jax_fem_b200/patch_plan.py): <= 64 mesh nodes owned by the CTA
// (4x4x4 on a structured grid) together with every cell touching them.  Per patch:
//   prologue : coordinates and solution of the patch's local nodes (owned + halo) -> shared memory
//   per chunk of 32 cells
//     phase 1: thread = (cell, quadrature point): J, J^-1, physical gradients g, JxW, grad u, stress -> record
//              (same arithmetic as element.cu, fe.py:112-141 / problem.py:204-210)
//     phase 2: thread = (owned corner (cell, a), column half): G_ab = sum_q E_q w_q g_a (x) g_b for 4 column nodes
//              from 16-byte broadcast shared loads, K_ab = lam' G + mu' G^T + mu' tr(G) I, element residual r_a
//     rounds : the row blocks are added into the patch's shared-memory accumulator, which is laid out exactly like
//              the CSR rows of the owned nodes; lanes that hit the same row are serialised by their plan-computed
//              round (= rank of the cell among the node's cells) => ascending cell order, bit-reproducible, no atomics
//   epilogue : Dirichlet rows -> unit rows (pattern kept, Mat.zeroRows), rows copied to `data` with coalesced
//              stores, residual (+ constant load vector) written; every CSR value is written exactly once.
// Cells on a patch surface are evaluated by every patch they touch (phase 1 only: 125 cells per 64 owned nodes on
// a structured grid); phase 2 is done once per (cell, corner) by the corner's owner.
#include "common.cuh"
#include "element_math.cuh"

namespace femb200 {
namespace {

// MAX_OWNED / CHUNK / RMAX mirror jax_fem_b200/patch_plan.py::CONFIGS; SPLIT = tasks per owned corner in phase 2
// (8 / SPLIT column nodes each); CTAS = resident CTAs per SM the shared-memory footprint is sized for.
template <int MAX_OWNED_, int CHUNK_, int THREADS_, int SPLIT_, int CTAS_>
struct FusedCfg {
  static constexpr int THREADS = THREADS_, CHUNK = CHUNK_, MAX_OWNED = MAX_OWNED_, SPLIT = SPLIT_, CTAS = CTAS_;
  static constexpr int MAX_LOCAL = MAX_OWNED == 64 ? 256 : 160;   // (patch_plan max_local + 1)
  static constexpr int ACC = MAX_OWNED * 27 * 9;                  // accumulator doubles (patch_plan acc_doubles)
  static constexpr int NN = 8, NQ = 8, DIM = 3, VEC = 3;
  static constexpr int CPT = NN / SPLIT;                          // column nodes per task
  static constexpr int TAB_STRIDE = 25, TAB_SIZE = NQ * TAB_STRIDE + NQ;
  // record of (cell, q): g[8][3] | E w | pad | S[3][3] = sigma JxW | pad  -- 16-byte aligned pieces
  static constexpr int OFF_G = 0, OFF_E = 24, OFF_S = 26, QREC = 36;
  static constexpr int CELLREC = NQ * QREC + 2;   // +16 B: neighbouring cells broadcast from different banks
  static constexpr int OFF_TAB = 0;
  static constexpr int OFF_ACC = OFF_TAB + TAB_SIZE;
  static constexpr int OFF_RACC = OFF_ACC + ACC;
  static constexpr int OFF_FEXT = OFF_RACC + MAX_OWNED * VEC;
  static constexpr int OFF_XU = OFF_FEXT + MAX_OWNED * VEC;
  static constexpr int OFF_REC = OFF_XU + MAX_LOCAL * (DIM + VEC);
  static constexpr int OFF_INT = OFF_REC + CHUNK * CELLREC;
  static constexpr int SMEM_DOUBLES = OFF_INT + (5 * MAX_OWNED) / 2;
  static_assert(CHUNK * NQ <= THREADS, "phase 1 needs one thread per (cell, q)");
  static_assert(SPLIT == 2 || SPLIT == 4, "column split");
};

struct FusedArgs {
  const double* points;
  const double* sol;
  const double* iv;
  const double* ref;
  const int32_t *phdr, *pn_node, *pn_out, *pn_acc, *pn_info, *lnodes, *pc_cell, *pc_ln, *pc_lm, *ck_cell, *ck_lane, *ck_rnd, *ln_desc, *ln_slot;
  const uint8_t* bc_flag;
  const double* f_ext;
  double* data;
  double* res;
  int n_patches;
  double p[8];
};


// ---- pieces shared by the two kernels ------------------------------------------------------------------------------
struct PatchView {       // per-patch shared-memory state
  double *acc, *racc, *fext, *xu;
  int *s_node, *s_out, *s_acc, *s_info;      // s_info: len | diagonal slot << 8 | Dirichlet flags << 16
};

struct PatchRange {
  int node0, lnode0, chunk0, n_owned, n_local, n_chunks;
};

__device__ __forceinline__ PatchRange patch_range(const FusedArgs& A, int patch) {
  const int* h0 = A.phdr + (int64_t)patch * 8;
  PatchRange r;
  r.node0 = h0[0]; r.lnode0 = h0[1]; r.chunk0 = h0[3];
  r.n_owned = h0[8] - r.node0; r.n_local = h0[9] - r.lnode0; r.n_chunks = h0[11] - r.chunk0;
  return r;
}

// prologue: owned-node tables (with the Dirichlet flags and the constant loads) and the coordinates / solution of the
// patch's local nodes -> shared memory
template <int THREADS>
__device__ __forceinline__ void patch_prologue(const FusedArgs& A, const PatchRange& r, const PatchView& v) {
  const int tid = threadIdx.x;
  __syncthreads();            // previous patch: epilogue done with s_*, xu, fext
  if (tid < r.n_owned) {
    const int nd = A.pn_node[r.node0 + tid];
    const uint8_t* f = A.bc_flag + 3 * (int64_t)nd;
    v.s_node[tid] = nd;
    v.s_out[tid] = A.pn_out[r.node0 + tid];
    v.s_acc[tid] = A.pn_acc[r.node0 + tid];
    v.s_info[tid] = A.pn_info[r.node0 + tid] | ((f[0] ? 1 : 0) << 16) | ((f[1] ? 1 : 0) << 17) | ((f[2] ? 1 : 0) << 18);
#pragma unroll
    for (int d = 0; d < 3; ++d) v.fext[tid * 3 + d] = A.f_ext ? A.f_ext[3 * (int64_t)nd + d] : 0.0;
  }
  for (int i = tid; i < r.n_local; i += THREADS) {
    const int64_t node = A.lnodes[r.lnode0 + i];
#pragma unroll
    for (int d = 0; d < 3; ++d) v.xu[i * 6 + d] = A.points[node * 3 + d];
#pragma unroll
    for (int d = 0; d < 3; ++d) v.xu[i * 6 + 3 + d] = A.sol[node * 3 + d];
  }
  __syncthreads();
}

// phase 1 of one (cell, q): geometry, grad u, stress -> g[8][3], E w, S = sigma JxW (fe.py:112-141, problem.py:204-210)
template <int LAW>
__device__ __forceinline__ void cell_point(const FusedArgs& A, const double* xu, const double* tabq, double wq, int pc, int q,
                                           double nu, double (&g)[8][3], double& Ew, double (&S)[3][3]) {
  const int2 lw = reinterpret_cast<const int2*>(A.pc_ln)[pc];
  const double* ivq = A.iv ? A.iv + (int64_t)A.pc_cell[pc] * 8 + q : nullptr;
  const double E = iso_modulus<LAW>(A.p, ivq, false);
  double X[24], U[24];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int ln = ((m < 4 ? lw.x : lw.y) >> (8 * (m & 3))) & 255;
    const double2* src = reinterpret_cast<const double2*>(xu + ln * 6);
    const double2 v0 = src[0], v1 = src[1], v2 = src[2];
    X[m * 3 + 0] = v0.x; X[m * 3 + 1] = v0.y; X[m * 3 + 2] = v1.x;
    U[m * 3 + 0] = v1.y; U[m * 3 + 1] = v2.x; U[m * 3 + 2] = v2.y;
  }
  const double w = qp_geometry<8, 3>(X, tabq, wq, g);
  double ug[3][3];
  qp_grad_u<8, 3, 3>(U, g, ug);
  const double mu = E / (2.0 * (1.0 + nu)), lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
  iso_stress<3>(lam, mu, ug, S);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d) S[i][d] *= w;
  Ew = E * w;
}

// epilogue: Dirichlet rows -> unit rows (pattern kept), coalesced copy of the finished CSR rows, residual + constant
// loads; the accumulators are zeroed for the next patch on the way out
template <int THREADS>
__device__ __forceinline__ void patch_epilogue(const FusedArgs& A, const PatchRange& r, const PatchView& v) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < r.n_owned; i += THREADS / 32) {
    const int nd = v.s_node[i], info = v.s_info[i];
    const int len = info & 255, dg = (info >> 8) & 255, fl = info >> 16;
    const int tot = 9 * len, len3 = 3 * len;
    double* src = v.acc + v.s_acc[i];
    double* dst = A.data + v.s_out[i];
    for (int e = lane; e < tot; e += 32) {
      double val = src[e];
      src[e] = 0.0;
      if (fl) {
        const int row = e / len3, col = e - row * len3;
        if ((fl >> row) & 1) val = (col == 3 * dg + row) ? 1.0 : 0.0;
      }
      dst[e] = val;
    }
    if (lane < 3) {
      A.res[3 * (int64_t)nd + lane] = v.racc[i * 3 + lane] + v.fext[i * 3 + lane];
      v.racc[i * 3 + lane] = 0.0;
    }
  }
}

template <int LAW, class L>
__global__ void __launch_bounds__(L::THREADS, L::CTAS) fused_assembly_kernel(const FusedArgs A) {
  constexpr int NN = L::NN, NQ = L::NQ, DIM = L::DIM, VEC = L::VEC, SPLIT = L::SPLIT, CPT = L::CPT;
  extern __shared__ __align__(16) double sm[];
  double* tab = sm + L::OFF_TAB;
  double* acc = sm + L::OFF_ACC;
  double* racc = sm + L::OFF_RACC;
  double* fext = sm + L::OFF_FEXT;
  double* xu = sm + L::OFF_XU;
  double* recs = sm + L::OFF_REC;
  int* s_node = reinterpret_cast<int*>(sm + L::OFF_INT);
  int* s_out = s_node + L::MAX_OWNED;
  int* s_acc = s_out + L::MAX_OWNED;
  int* s_info = s_acc + L::MAX_OWNED;      // len | diagonal slot << 8 | Dirichlet flags << 16
  const PatchView pv{acc, racc, fext, xu, s_node, s_out, s_acc, s_info};
  const int tid = threadIdx.x;

  for (int i = tid; i < NQ * NN * DIM; i += L::THREADS) tab[(i / (NN * DIM)) * L::TAB_STRIDE + i % (NN * DIM)] = A.ref[i];
  if (tid < NQ) tab[NQ * L::TAB_STRIDE + tid] = A.ref[NQ * NN * DIM + tid];
  for (int i = tid; i < L::ACC + L::MAX_OWNED * VEC; i += L::THREADS) acc[i] = 0.0;   // acc and racc are adjacent

  const double nu = iso_nu<LAW>(A.p);
  const double mu1 = 1.0 / (2.0 * (1.0 + nu)), lam1 = nu / ((1.0 + nu) * (1.0 - 2.0 * nu));

#pragma unroll 1
  for (int patch = blockIdx.x; patch < A.n_patches; patch += gridDim.x) {
    const PatchRange pr = patch_range(A, patch);
    const int chunk0 = pr.chunk0, n_chunks = pr.n_chunks;
    patch_prologue<L::THREADS>(A, pr, pv);

#pragma unroll 1
    for (int k = 0; k < n_chunks; ++k) {
      const int c0 = A.ck_cell[chunk0 + k], ncell = A.ck_cell[chunk0 + k + 1] - c0;
      // phase-2 metadata of this chunk: issued before phase 1 so that the loads overlap it
      const int l0 = A.ck_lane[chunk0 + k], l1 = A.ck_lane[chunk0 + k + 1];
      const int rounds = A.ck_rnd[chunk0 + k];
      const int ntask = SPLIT * (l1 - l0);
      int desc0 = 0, slot0 = 0;
      if (tid < ntask) {
        desc0 = A.ln_desc[l0 + tid / SPLIT];
        slot0 = A.ln_slot[2 * (int64_t)(l0 + tid / SPLIT) + ((tid % SPLIT) * CPT) / 4];
      }
      // ---------------- phase 1: thread = (cell, q) ----------------
      if (tid < ncell * NQ) {
        const int cl = tid >> 3, q = tid & 7;
        double g[NN][DIM], sig[DIM][DIM], Ew;
        cell_point<LAW>(A, xu, tab + q * L::TAB_STRIDE, tab[NQ * L::TAB_STRIDE + q], c0 + cl, q, nu, g, Ew, sig);
        double* rq = recs + cl * L::CELLREC + q * L::QREC;
#pragma unroll
        for (int t = 0; t < 12; ++t)
          reinterpret_cast<double2*>(rq + L::OFF_G)[t] = make_double2(g[(2 * t) / 3][(2 * t) % 3], g[(2 * t + 1) / 3][(2 * t + 1) % 3]);
        rq[L::OFF_E] = Ew;
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
          for (int d = 0; d < DIM; ++d) rq[L::OFF_S + i * DIM + d] = sig[i][d];
      }
      __syncthreads();

      // ---------------- phase 2: thread = (owned corner, column part) ----------------
#pragma unroll 1
      for (int t0 = 0; t0 < ntask; t0 += L::THREADS) {
        const int t = t0 + tid;
        const bool valid = t < ntask;
        const int h = t % SPLIT;
        int desc = desc0, slots = slot0;
        if (t0 > 0 && valid) {
          desc = A.ln_desc[l0 + t / SPLIT];
          slots = A.ln_slot[2 * (int64_t)(l0 + t / SPLIT) + (h * CPT) / 4];
        }
        slots >>= 8 * ((h * CPT) % 4);
        const int cl = desc & 31, a = (desc >> 5) & 7, nl = (desc >> 8) & 255, rk = (desc >> 16) & 255;
        double K[CPT][3][3];
        double r[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < CPT; ++j)
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) K[j][i][c] = 0.0;
        if (valid) {
          const double* rec = recs + cl * L::CELLREC;
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const double* rq = rec + q * L::QREC;
            const double ew = rq[L::OFF_E];
            double ga[3], gw[3], gb[CPT * 3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              ga[d] = rq[L::OFF_G + a * 3 + d];
              gw[d] = ga[d] * ew;
            }
#pragma unroll
            for (int u = 0; u < CPT * 3 / 2; ++u) {
              const double2 v = reinterpret_cast<const double2*>(rq + L::OFF_G + h * (CPT * 3))[u];
              gb[2 * u] = v.x;
              gb[2 * u + 1] = v.y;
            }
#pragma unroll
            for (int j = 0; j < CPT; ++j)
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int c = 0; c < 3; ++c) K[j][i][c] = fma(gw[i], gb[j * 3 + c], K[j][i][c]);
            // element residual r_a = sum_q S_q g_a(q): part h takes the quadrature points q = h NQ/SPLIT ...
            if (q / (NQ / SPLIT) == h) {
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int d = 0; d < 3; ++d) r[i] = fma(rq[L::OFF_S + i * 3 + d], ga[d], r[i]);   // problem.py:210
            }
          }
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            double G[3][3];
            double tr = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              tr += K[j][i][i];
#pragma unroll
              for (int c = 0; c < 3; ++c) G[i][c] = K[j][i][c];
            }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int c = 0; c < 3; ++c) K[j][i][c] = lam1 * G[i][c] + mu1 * G[c][i] + (i == c ? mu1 * tr : 0.0);
          }
        }
        // the SPLIT parts of a corner sit in adjacent lanes: butterfly sum of the partial residuals (commutative =>
        // every lane of the group holds the same bits); the part-0 lane adds it to the node
#pragma unroll
        for (int o = 1; o < SPLIT; o <<= 1)
#pragma unroll
          for (int i = 0; i < 3; ++i) r[i] += __shfl_xor_sync(0xffffffffu, r[i], o);
        // ---- ordered accumulation into the patch's CSR-shaped rows ----
        const int len3 = valid ? 3 * (s_info[nl] & 255) : 0;
        double* rowp = acc + (valid ? s_acc[nl] : 0);
#pragma unroll 1
        for (int rd = 0; rd < rounds; ++rd) {
          if (valid && rk == rd) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
              double* p = rowp + 3 * ((slots >> (8 * j)) & 255);
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int c = 0; c < 3; ++c) p[i * len3 + c] += K[j][i][c];
            }
            if (h == 0) {
#pragma unroll
              for (int i = 0; i < 3; ++i) racc[nl * 3 + i] += r[i];
            }
          }
          __syncthreads();
        }
      }
      if (ntask == 0) __syncthreads();     // records may not be overwritten before every thread left phase 2
    }

    patch_epilogue<L::THREADS>(A, pr, pv);
  }
}

template <int LAW, class L>
int launch_fused(const FusedArgs& A, cudaStream_t st) {
  const size_t smem = sizeof(double) * L::SMEM_DOUBLES;
  auto k = fused_assembly_kernel<LAW, L>;
  FEM_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int want = kNumSM * L::CTAS;
  const int grid = A.n_patches < want ? A.n_patches : want;
  k<<<grid, L::THREADS, smem, st>>>(A);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

// ---- DMMA variant ----------------------------------------------------------------------------------------------
// The CUDA-core phase 2 above is bound by shared-memory bandwidth (one wavefront per operand double and lane, plus
// the read-modify-write of the accumulator).  Here the element tangent of a cell is formed by ONE warp on the FP64
// tensor cores exactly as in element.cu::element_dmma_kernel (G = sum_q (E_q w_q g(q)) g(q)^T as a 3x3 grid of 8x8
// tiles, 18 + 6 mma.sync.m8n8k4.f64, 8 + 6 operand loads per lane and cell), lane (n, t) ends up with the blocks
// K[n][2t], K[n][2t+1] of row corner n and adds them into the patch accumulator if the patch owns that corner.
// Inside a cell no two lanes touch the same accumulator word; the plan's chunks hold <= 8 cells (one per warp) that
// share no owned node (rmax = 1), so a barrier between chunks is the only ordering needed.  Cells on the patch
// surface pay the full tile for the few rows that are kept (125 cells per 64 owned nodes).

struct FusedDmmaCfg {
  static constexpr int THREADS = 256, MAX_OWNED = 64, MAX_LOCAL = 256;
  static constexpr int SUB = 8;                    // cells per chunk (patch_plan CONFIGS[4].chunk) = warps
  static constexpr int NSUB = 4;                   // chunks per phase-1 batch
  static constexpr int BATCH = SUB * NSUB;
  static constexpr int ACC = MAX_OWNED * 27 * 9;
  static constexpr int NN = 8, NQ = 8, DIM = 3, VEC = 3;
  static constexpr int TAB_STRIDE = 25, TAB_SIZE = NQ * TAB_STRIDE + NQ;
  static constexpr int GS = 28;                    // per-q stride of g[n][d]: 12 (mod 16) => conflict-free fragments
  static constexpr int OFF_G = 0, OFF_S = NQ * GS, OFF_E = OFF_S + NQ * 9;
  static constexpr int CELL = OFF_E + NQ + 2;      // 306: cell stride 2 (mod 16) spreads the phase-1 stores
  static constexpr int OFF_TAB = 0;
  static constexpr int OFF_ACC = OFF_TAB + TAB_SIZE;
  static constexpr int OFF_RACC = OFF_ACC + ACC;
  static constexpr int OFF_FEXT = OFF_RACC + MAX_OWNED * VEC;
  static constexpr int OFF_XU = OFF_FEXT + MAX_OWNED * VEC;
  static constexpr int OFF_CELLS = OFF_XU + MAX_LOCAL * (DIM + VEC);
  static constexpr int OFF_INT = OFF_CELLS + BATCH * CELL;
  static constexpr int SMEM_DOUBLES = OFF_INT + (5 * MAX_OWNED) / 2;
};

template <int LAW>
__global__ void __launch_bounds__(FusedDmmaCfg::THREADS, 1) fused_dmma_kernel(const FusedArgs A) {
  using L = FusedDmmaCfg;
  constexpr int NN = L::NN, NQ = L::NQ, DIM = L::DIM, VEC = L::VEC;
  extern __shared__ __align__(16) double sm[];
  double* tab = sm + L::OFF_TAB;
  double* acc = sm + L::OFF_ACC;
  double* racc = sm + L::OFF_RACC;
  double* fext = sm + L::OFF_FEXT;
  double* xu = sm + L::OFF_XU;
  double* cellsm = sm + L::OFF_CELLS;
  int* s_node = reinterpret_cast<int*>(sm + L::OFF_INT);
  int* s_out = s_node + L::MAX_OWNED;
  int* s_acc = s_out + L::MAX_OWNED;
  int* s_info = s_acc + L::MAX_OWNED;      // len | diagonal slot << 8 | Dirichlet flags << 16
  const PatchView pv{acc, racc, fext, xu, s_node, s_out, s_acc, s_info};
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < NQ * NN * DIM; i += L::THREADS) tab[(i / (NN * DIM)) * L::TAB_STRIDE + i % (NN * DIM)] = A.ref[i];
  if (tid < NQ) tab[NQ * L::TAB_STRIDE + tid] = A.ref[NQ * NN * DIM + tid];
  for (int i = tid; i < L::ACC + L::MAX_OWNED * VEC; i += L::THREADS) acc[i] = 0.0;   // acc and racc are adjacent

  const double nu = iso_nu<LAW>(A.p);
  const double mu1 = 1.0 / (2.0 * (1.0 + nu)), lam1 = nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
  const int n = lane >> 2, t = lane & 3;     // phase 2: row corner n, column pair t

#pragma unroll 1
  for (int patch = blockIdx.x; patch < A.n_patches; patch += gridDim.x) {
    const PatchRange pr = patch_range(A, patch);
    const int chunk0 = pr.chunk0, n_chunks = pr.n_chunks;
    patch_prologue<L::THREADS>(A, pr, pv);

#pragma unroll 1
    for (int k0 = 0; k0 < n_chunks; k0 += L::NSUB) {
      const int nsub = min(L::NSUB, n_chunks - k0);
      const int cbase = A.ck_cell[chunk0 + k0];
      const int ncell = A.ck_cell[chunk0 + k0 + nsub] - cbase;
      // phase-2 metadata of the warp's cells (one per chunk of the batch), fetched ahead of phase 1:
      // cw[s] = cell index inside the batch or -1, desc/slots of the lane's corner if the patch owns it
      int cw[L::NSUB], desc[L::NSUB], slots[L::NSUB];
#pragma unroll
      for (int s = 0; s < L::NSUB; ++s) {
        cw[s] = -1;
        desc[s] = -1;
        slots[s] = 0;
        if (s < nsub) {
          const int c0 = A.ck_cell[chunk0 + k0 + s], c1 = A.ck_cell[chunk0 + k0 + s + 1];
          if (c0 + warp < c1) {
            cw[s] = c0 + warp - cbase;
            const int2 lm = reinterpret_cast<const int2*>(A.pc_lm)[c0 + warp];      // first lane, owned-corner mask
            if ((lm.y >> n) & 1) {
              const int li = lm.x + __popc(lm.y & ((1 << n) - 1));
              desc[s] = A.ln_desc[li];
              slots[s] = A.ln_slot[2 * (int64_t)li + (t >> 1)] >> (16 * (t & 1));
            }
          }
        }
      }
      // ---------------- phase 1: thread = (cell, q) ----------------
      if (tid < ncell * NQ) {
        const int cl = tid >> 3, q = tid & 7;
        double g[NN][DIM], sig[DIM][DIM], Ew;
        cell_point<LAW>(A, xu, tab + q * L::TAB_STRIDE, tab[NQ * L::TAB_STRIDE + q], cbase + cl, q, nu, g, Ew, sig);
        double* cb = cellsm + cl * L::CELL;
#pragma unroll
        for (int u = 0; u < 12; ++u)
          reinterpret_cast<double2*>(cb + L::OFF_G + q * L::GS)[u] = make_double2(g[(2 * u) / 3][(2 * u) % 3], g[(2 * u + 1) / 3][(2 * u + 1) % 3]);
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
          for (int d = 0; d < DIM; ++d) cb[L::OFF_S + q * 9 + i * DIM + d] = sig[i][d];
        cb[L::OFF_E + q] = Ew;
      }
      __syncthreads();

      // ---------------- phase 2: one warp per cell of a chunk; lane = (row corner n, column pair t) ----------------
#pragma unroll
      for (int s = 0; s < L::NSUB; ++s) {
        if (s < nsub) {
          if (cw[s] >= 0) {                                       // warp-uniform
            const double* cj = cellsm + cw[s] * L::CELL;
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
              }
            // residual: B[q][col] = S_q[i = col][d] for col < 3 (col = lane / 4), zero otherwise
            double R[2] = {0.0, 0.0};
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const double b0 = (n < 3) ? cj[L::OFF_S + t * 9 + n * 3 + d] : 0.0;
              const double b1 = (n < 3) ? cj[L::OFF_S + (t + 4) * 9 + n * 3 + d] : 0.0;
              dmma884(R, g0[d], b0);
              dmma884(R, g1[d], b1);
            }
            if (desc[s] >= 0) {                                   // the patch owns corner n of this cell
              const int nl = (desc[s] >> 8) & 255;
              const int len3 = 3 * (s_info[nl] & 255);
              double* rowp = acc + s_acc[nl];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const double tr = C[0][0][e] + C[1][1][e] + C[2][2][e];
                double* p = rowp + 3 * ((slots[s] >> (8 * e)) & 255);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                  for (int c = 0; c < 3; ++c)
                    p[i * len3 + c] += lam1 * C[i][c][e] + mu1 * C[c][i][e] + (i == c ? mu1 * tr : 0.0);
              }
              // R = r_n[2t], r_n[2t+1]: components 0,1 live in lanes t == 0, component 2 in lanes t == 1
              if (t == 0) {
                racc[nl * 3 + 0] += R[0];
                racc[nl * 3 + 1] += R[1];
              } else if (t == 1) {
                racc[nl * 3 + 2] += R[0];
              }
            }
          }
          __syncthreads();      // the next chunk may touch the same owned nodes; after the last one: cell areas free
        }
      }
    }

    patch_epilogue<L::THREADS>(A, pr, pv);
  }
}

template <int LAW>
int launch_fused_dmma(const FusedArgs& A, cudaStream_t st) {
  using L = FusedDmmaCfg;
  const size_t smem = sizeof(double) * L::SMEM_DOUBLES;
  auto k = fused_dmma_kernel<LAW>;
  FEM_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = A.n_patches < kNumSM ? A.n_patches : kNumSM;
  k<<<grid, L::THREADS, smem, st>>>(A);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

template <class L>
int launch_fused_law(int law_id, const FusedArgs& A, cudaStream_t st) {
  if (law_id == FEM_LAW_SIMP) return launch_fused<FEM_LAW_SIMP, L>(A, st);
  return launch_fused<FEM_LAW_LINEAR_ELASTIC, L>(A, st);
}

}  // namespace
}  // namespace femb200

using namespace femb200;

extern "C" int fem_assemble_fused(int ele_type, int vec, int law_id, const double* law_params_host,
                                  const double* points, const double* sol, const double* internal_var,
                                  const double* ref_tables, int64_t n_patches, const int32_t* phdr,
                                  const int32_t* pn_node, const int32_t* pn_out, const int32_t* pn_acc,
                                  const int32_t* pn_info, const int32_t* lnodes, const int32_t* pc_cell,
                                  const int32_t* pc_ln, const int32_t* pc_lm, const int32_t* ck_cell, const int32_t* ck_lane, const int32_t* ck_rnd,
                                  const int32_t* ln_desc, const int32_t* ln_slot, const uint8_t* bc_flag,
                                  const double* f_ext, double* data, double* res, int config, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(points && sol && ref_tables && law_params_host && phdr && pn_node && pn_out && pn_acc && pn_info &&
                  lnodes && pc_cell && pc_ln && pc_lm && ck_cell && ck_lane && ck_rnd && ln_desc && ln_slot && bc_flag && data && res,
              "null pointer");
  FEM_REQUIRE(n_patches >= 0 && n_patches < (1ll << 31), "n_patches out of range");
  if (!(ele_type == FEM_ELE_HEX8 && vec == 3 && (law_id == FEM_LAW_LINEAR_ELASTIC || law_id == FEM_LAW_SIMP))) {
    set_error("fused assembly is registered for HEX8 / vec 3 / isotropic elasticity only (element=%d, vec=%d, law=%d)",
              ele_type, vec, law_id);
    return FEM_EINVAL;
  }
  FEM_REQUIRE(!(law_id == FEM_LAW_SIMP && !internal_var), "SIMP needs the per-quadrature-point density");
  if (n_patches == 0) return FEM_OK;
  FusedArgs A{};
  A.points = points; A.sol = sol; A.iv = internal_var; A.ref = ref_tables;
  A.phdr = phdr; A.pn_node = pn_node; A.pn_out = pn_out; A.pn_acc = pn_acc; A.pn_info = pn_info; A.lnodes = lnodes;
  A.pc_cell = pc_cell; A.pc_ln = pc_ln; A.pc_lm = pc_lm; A.ck_cell = ck_cell; A.ck_lane = ck_lane; A.ck_rnd = ck_rnd; A.ln_desc = ln_desc; A.ln_slot = ln_slot;
  A.bc_flag = bc_flag; A.f_ext = f_ext; A.data = data; A.res = res; A.n_patches = (int)n_patches;
  for (int i = 0; i < 8; ++i) A.p[i] = law_params_host[i];
  // config == index into jax_fem_b200/patch_plan.py::CONFIGS (the tables must have been built for it)
  switch (config) {
    case 0: return launch_fused_law<FusedCfg<64, 32, 256, 2, 1>>(law_id, A, (cudaStream_t)stream);
    case 1: return launch_fused_law<FusedCfg<32, 16, 256, 4, 2>>(law_id, A, (cudaStream_t)stream);
    case 2: return launch_fused_law<FusedCfg<32, 16, 128, 2, 2>>(law_id, A, (cudaStream_t)stream);
    case 3: return launch_fused_law<FusedCfg<32, 16, 256, 2, 2>>(law_id, A, (cudaStream_t)stream);
    case 4:
      if (law_id == FEM_LAW_SIMP) return launch_fused_dmma<FEM_LAW_SIMP>(A, (cudaStream_t)stream);
      return launch_fused_dmma<FEM_LAW_LINEAR_ELASTIC>(A, (cudaStream_t)stream);
  }
  set_error("unknown fused-assembly configuration %d", config);
  return FEM_EINVAL;
}
