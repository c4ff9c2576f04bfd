// Staged assembly, warp-item version: element tangents -> CSR in ONE persistent kernel whose work items are processed by
// single WARPS, the staging rows kept in L2 (ring + discard).  Same replacement as staged.cu -- Problem.compute_newton_vars
// (jax_fem/problem.py:447-460) -> problem.V -> _PetscTangentCache.update / get_A (jax_fem/solver.py:469-553) for HEX8 / vec 3 /
// isotropic elasticity -- and the same schedule tables (jax_fem_b200/stage_plan.py with cells_per_item = 4), but no CTA-wide
// barrier anywhere after start-up: staged.cu spends 37 % of its stall samples at __syncthreads (profiles/r02_ring_assembly.md).
//
//   grid = one CTA per SM, W warps each; every warp owns 11.5 KB of shared memory.  Tickets are dealt round-robin to the
//   W * #SM resident warps (no atomics); a warp processes its tickets in order and prefetches the next descriptor.
//   E item : 4 cells.  lane = (cell, q): geometry / grad u / stress (fe.py:112-141, problem.py:204-210) -> the warp's shared
//            memory; then per cell 18 + 6 mma.sync.m8n8k4.f64, accumulator fragments stored straight to the corner's staging
//            row in tile-major layout (see element.cu, TILES).  Only __syncwarp.
//   G item : the CSR rows of ~4 consecutive mesh nodes, one node at a time: wait for the E items of the node's cells (one flag
//            per lane), ONE TMA bulk copy of the node's <= 16 staging rows into the warp's buffer (the next node's copy is in
//            flight meanwhile when both fit), lane = block entry of the node: fixed-order sum of its sources, isotropic map
//            K = lam' G + mu' G^T + mu' tr(G) I, Dirichlet rows -> unit rows, and the entry goes straight from registers to
//            its CSR position (a node's scalar row = 3 doubles per lane, contiguous across the warp).  No output staging.
//            When the item is done its staging lines are dropped from L2 (discard.L2) and its flag is released.
//
// Waits are on EARLIER tickets only (stage_plan.py checks it), which are held by resident warps: no deadlock as long as the
// whole grid is resident (one CTA per SM, checked at launch).  Fixed summation order => bit-reproducible, no atomics on data.
#include "common.cuh"
#include "element_math.cuh"

namespace femb200 {
namespace {

constexpr int kCtrlInts = 16;          // ctrl[3] error; then the done flags of the E items and of the G items
constexpr unsigned kSpinLimit = 1u << 20;
constexpr int kRow = 72;               // doubles per staging row
constexpr int kDescInts = 32;
enum { D_CODE = 0, D_ROW0 = 1, D_C0 = 2, D_C1 = 6, D_N0 = 30, D_N1 = 31 };

struct WarpArgs {
  const double* points;
  const double* sol;
  const double* iv;
  const double* ref;
  const int32_t* cells_p;       // (C, 8) connectivity in processing order
  const int32_t* corder;        // (C) processing slot -> cell id
  const int32_t* dest_row;      // (C*8) staging row of corner (slot, a)
  const int32_t* prev_g;        // (C*8) G item that read the previous occupant of that row, or -1
  double* Re;
  const int32_t* tdesc;         // (n_tickets, 32): [0] E: 0x80000000 | item, G: item; G: [1] first staging row, [2] / [6] first / last
                                // corner, [30] / [31] first / last node
  const int32_t* corner_eitem;  // (C*8) E item of every corner in node-sorted order
  const int32_t* nc_ptr;        // (nodes + 1) corners per node
  const int32_t* brow_ptr;      // (nodes + 1) block entries per node
  const uint4* esrc;            // (nnzb) 16 source bytes per entry (relative block inside the node's rows; bit 7 of byte 0: diagonal)
  const uint8_t* bc_flag;       // (3 nodes) 1 = Dirichlet row, or nullptr
  double* data;
  double* stage;
  int* ctrl;
  int n_tickets, n_e, n_g;
  int64_t C;
  double p[8];
  double lam1, mu1;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {     // no L1 invalidation (ld.acquire costs a CCTL.IVALL per load)
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// legal waits are short (the awaited ticket is earlier than ours and its warp is resident); exceeding the limit is a plan bug:
// flag it and go on (wrong numbers, reported by the host) instead of hanging the GPU
__device__ __forceinline__ void wait_flag(const int* flag, int* err) {
  unsigned spins = 0;
  while (ld_relaxed(flag) == 0) {                    // the caller fences once after all its flags are set
    __nanosleep(32);
    if ((++spins & 255u) == 0 && (spins > kSpinLimit || ld_acquire(err) != 0)) {
      atomicExch(err, 1);
      return;
    }
  }
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity, int* err) {
  unsigned spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++spins & 1023u) == 0 && (spins > (kSpinLimit << 4) || ld_acquire(err) != 0)) {
      atomicExch(err, 2);
      return;
    }
  }
}

struct WLayout {
  static constexpr int NQ = 8;
  static constexpr int TAB_STRIDE = 25, TAB_SIZE = NQ * TAB_STRIDE + NQ;      // 208
  static constexpr int GS = 28;                    // per-q stride of g[n][d]: 12 (mod 16) => conflict-free fragments
  static constexpr int OFF_X = 0, OFF_U = 24;      // per cell: X[8][3], U[8][3]
  static constexpr int OFF_G = 48;                 // g[q][n][d]
  static constexpr int OFF_S = OFF_G + NQ * GS;    // S[q][i][d] = sigma JxW
  static constexpr int OFF_E = OFF_S + NQ * 9;     // E_q JxW
  static constexpr int CELL = 354;
  static constexpr int WARP = 4 * CELL + 18;       // E: 4 cells + 32 staging rows + 4 cell ids (ints); G: 16 rows of 72 doubles
  static_assert(WARP >= 16 * kRow, "the warp's region must hold the staging rows of one node");
};

// ---- E item: 4 cells, one warp ------------------------------------------------------------------------------------------
template <int LAW>
__device__ __forceinline__ void run_element_warp(const WarpArgs& A, int item, const double* tab, double* wb, int l) {
  using L = WLayout;
  constexpr int NN = 8, NQ = 8, DIM = 3, VEC = 3, ND = 24;
  int* pos = reinterpret_cast<int*>(wb + 4 * L::CELL);      // [0,32): staging rows, [32,36): cell ids
  const int j1 = l >> 3, q = l & 7;
  const int64_t s1 = (int64_t)item * 4 + j1;
  const bool act1 = s1 < A.C;
  double* cb = wb + j1 * L::CELL;
  int c1 = 0, pg = -1;
  if (act1) {
    const int64_t node = A.cells_p[s1 * NN + q];
    c1 = A.corder[s1];
    pg = A.prev_g[s1 * NN + q];
#pragma unroll
    for (int d = 0; d < DIM; ++d) cb[L::OFF_X + q * DIM + d] = A.points[node * DIM + d];
#pragma unroll
    for (int i = 0; i < VEC; ++i) cb[L::OFF_U + q * VEC + i] = A.sol[node * VEC + i];
    pos[l] = A.dest_row[s1 * NN + q];
    if (q == 0) pos[32 + j1] = c1;
  }
  __syncwarp();
  if (act1) {
    double g[NN][DIM];
    const double w = qp_geometry<NN, DIM>(cb + L::OFF_X, tab + q * L::TAB_STRIDE, tab[NQ * L::TAB_STRIDE + q], g);
    double ug[VEC][DIM];
    qp_grad_u<NN, DIM, VEC>(cb + L::OFF_U, g, ug);
    const double* ivq = A.iv ? A.iv + (int64_t)c1 * NQ + q : nullptr;
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
    // the corner's staging row may still be read by the G item of its previous occupant
    if (pg >= 0) wait_flag(A.ctrl + kCtrlInts + A.n_e + pg, A.ctrl + 3);
  }
  __syncwarp();
  fence_acq_rel();

  const int n = l >> 2, t = l & 3;
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
    const int64_t s = (int64_t)item * 4 + j;
    if (s >= A.C) break;                                   // warp-uniform
    const int64_t c = pos[32 + j];
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
    double* row = A.stage + (int64_t)pos[j * 8 + n] * kRow + 2 * t;
#pragma unroll
    for (int I = 0; I < 3; ++I)
#pragma unroll
      for (int J = 0; J < 3; ++J) {
        double Cc[2] = {0.0, 0.0};
        dmma884(Cc, a0[I], g0[J]);
        dmma884(Cc, a1[I], g1[J]);
        *reinterpret_cast<double2*>(row + (I * 3 + J) * 8) = make_double2(Cc[0], Cc[1]);
      }
    double R[2] = {0.0, 0.0};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double b0 = (n < 3) ? cj[L::OFF_S + t * 9 + n * 3 + d] : 0.0;
      const double b1 = (n < 3) ? cj[L::OFF_S + (t + 4) * 9 + n * 3 + d] : 0.0;
      dmma884(R, g0[d], b0);
      dmma884(R, g1[d], b1);
    }
    if (t == 0) {
      A.Re[c * ND + n * 3 + 0] = R[0];
      A.Re[c * ND + n * 3 + 1] = R[1];
    } else if (t == 1) {
      A.Re[c * ND + n * 3 + 2] = R[0];
    }
  }
  __syncwarp();
  if (l == 0) st_release(A.ctrl + kCtrlInts + item, 1);     // release: cumulative over the warp's stores (ordered by __syncwarp)
  __syncwarp();
}

// ---- G item: the CSR rows of the nodes [N0, N1), one warp -------------------------------------------------------------
struct GState {
  uint32_t uses[2];     // completed phases of the warp's two mbarriers
};

__device__ __forceinline__ void node_issue(const WarpArgs& A, int64_t row, int nr, double* buf, uint64_t* bar, int l) {
  if (l == 0) {
    const uint32_t bytes = (uint32_t)nr * kRow * sizeof(double);
    mbar_expect_tx(bar, bytes);
    if (bytes) bulk_g2s(buf, A.stage + row * kRow, bytes, bar);
  }
}

__device__ __forceinline__ void run_gather_warp(const WarpArgs& A, int item, int row0, int C0, int C1, int N0, int N1, double* wb,
                                                uint64_t* bars, GState& st, int l) {
  int* const err = A.ctrl + 3;
  // per-node offsets of the item, one node per lane (items rarely have more than a handful of nodes)
  const int nn = N1 - N0;
  const int ncp = A.nc_ptr[N0 + (l < nn ? l : nn)];
  int maxr = 0;
  {
    const int nxt = __shfl_down_sync(0xffffffffu, ncp, 1);
    maxr = (l < nn && l < 31) ? nxt - ncp : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxr = max(maxr, __shfl_xor_sync(0xffffffffu, maxr, o));
  }
  const bool halves = maxr <= 8 && nn <= 31;       // two nodes fit the buffer side by side: the next copy runs ahead
  auto corner0 = [&](int k) { return (k < 32 && nn <= 31) ? __shfl_sync(0xffffffffu, ncp, k) : A.nc_ptr[N0 + k]; };

  // every E item holding a cell around the item's nodes must be done: one flag per corner, two corners per lane; then ONE
  // fence for the whole item (generic-proxy writes of the E items -> the async-proxy reads of the bulk copies below)
  for (int c = C0 + l; c < C1; c += 32) wait_flag(A.ctrl + kCtrlInts + A.corner_eitem[c], err);
  __syncwarp();
  fence_acq_rel();
  fence_proxy_async_all();
  {
    const int r0 = corner0(0), nr = corner0(1) - r0;
    node_issue(A, (int64_t)row0 + (r0 - C0), nr, wb, &bars[0], l);
  }
  for (int k = 0; k < nn; ++k) {
    const int n = N0 + k;
    const int h = halves ? (k & 1) : 0;
    const int r0 = corner0(k), r1 = corner0(k + 1), nr = r1 - r0;
    const double* buf = wb + h * (8 * kRow);
    const int e0 = A.brow_ptr[n], ne = A.brow_ptr[n + 1] - e0;
    // the next node's rows go into the other half right away
    bool next_now = false;
    if (k + 1 < nn && halves) {
      node_issue(A, (int64_t)row0 + (r1 - C0), corner0(k + 2) - r1, wb + (h ^ 1) * (8 * kRow), &bars[h ^ 1], l);
      next_now = true;
    }
    uint4 src = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (l < ne) src = __ldcs(A.esrc + e0 + l);
    int bc = 0;
    if (A.bc_flag) bc = A.bc_flag[3 * (int64_t)n] | (A.bc_flag[3 * (int64_t)n + 1] << 1) | (A.bc_flag[3 * (int64_t)n + 2] << 2);
    mbar_wait_bounded(&bars[h], st.uses[h] & 1, err);
    ++st.uses[h];
    const int rowlen = 3 * ne;
    double* out = A.data + 9 * (int64_t)e0;
    for (int eb = 0; eb < ne; eb += 32) {
      const int ei = eb + l;
      if (eb > 0) {
        src = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        if (ei < ne) src = __ldcs(A.esrc + e0 + ei);
      }
      if (ei < ne) {
        double G[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) G[j] = 0.0;
        const bool diag = (src.x & 0x80u) != 0;
#pragma unroll 1
        for (int s = 0; s < 16; ++s) {
          const uint32_t word = s < 4 ? src.x : (s < 8 ? src.y : (s < 12 ? src.z : src.w));
          uint32_t b = (word >> (8 * (s & 3))) & 0xffu;
          if (s == 0) b &= 0x7fu;
          else if (b == 0xffu) break;
          const double* blk = buf + (b >> 3) * kRow + (b & 7);
#pragma unroll
          for (int j = 0; j < 9; ++j) G[j] += blk[j * 8];
        }
        const double tr = A.mu1 * (G[0] + G[4] + G[8]);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double* o = out + (int64_t)i * rowlen + 3 * ei;
          if ((bc >> i) & 1) {
            o[0] = (diag && i == 0) ? 1.0 : 0.0;
            o[1] = (diag && i == 1) ? 1.0 : 0.0;
            o[2] = (diag && i == 2) ? 1.0 : 0.0;
          } else {
#pragma unroll
            for (int kk = 0; kk < 3; ++kk)
              __stcs(o + kk, A.lam1 * G[i * 3 + kk] + A.mu1 * G[kk * 3 + i] + (i == kk ? tr : 0.0));
          }
        }
      }
    }
    __syncwarp();                                   // every lane is done reading this half before it is refilled
    if (k + 1 < nn && !next_now) {
      const int nr2 = corner0(k + 2) - r1;
      const int h2 = halves ? (h ^ 1) : 0;
      node_issue(A, (int64_t)row0 + (r1 - C0), nr2, wb + h2 * (8 * kRow), &bars[h2], l);
    }
  }
  // the item's staging rows are dead: drop the lines that lie entirely inside them from L2 (no write-back), then release them
  {
    const double* first = A.stage + (int64_t)row0 * kRow;
    const uintptr_t lo = (reinterpret_cast<uintptr_t>(first) + 127) & ~(uintptr_t)127;
    const uintptr_t hi = (reinterpret_cast<uintptr_t>(first) + (size_t)(C1 - C0) * kRow * sizeof(double)) & ~(uintptr_t)127;
    for (uintptr_t a = lo + 128 * (uintptr_t)l; a < hi; a += 128 * 32) asm volatile("discard.global.L2 [%0], 128;" ::"l"(a) : "memory");
  }
  __syncwarp();
  if (l == 0) st_release(A.ctrl + kCtrlInts + A.n_e + item, 1);
  __syncwarp();
}

template <int LAW, int W>
__global__ void __launch_bounds__(32 * W, 1) staged_warp_kernel(const WarpArgs A) {
  extern __shared__ __align__(128) double wsm[];
  __shared__ __align__(8) uint64_t bars[W][2];
  double* tab = wsm;
  const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
  double* wb = wsm + WLayout::TAB_SIZE + warp * WLayout::WARP;
  for (int i = threadIdx.x; i < 8 * 8 * 3; i += 32 * W) tab[(i / 24) * WLayout::TAB_STRIDE + i % 24] = A.ref[i];
  if (threadIdx.x < 8) tab[8 * WLayout::TAB_STRIDE + threadIdx.x] = A.ref[8 * 8 * 3 + threadIdx.x];
  if (l == 0) {
    mbar_init(&bars[warp][0], 1);
    mbar_init(&bars[warp][1], 1);
  }
  __syncthreads();                                  // the only CTA-wide barrier

  GState st;
  st.uses[0] = st.uses[1] = 0;
  const int n_warps = gridDim.x * W;
  int t = warp * gridDim.x + blockIdx.x;            // consecutive tickets go to different SMs
  int desc = t < A.n_tickets ? __ldcs(A.tdesc + (int64_t)t * kDescInts + l) : 0;
  while (t < A.n_tickets) {
    const int t_next = t + n_warps;
    const int desc_next = t_next < A.n_tickets ? __ldcs(A.tdesc + (int64_t)t_next * kDescInts + l) : 0;
    const int code = __shfl_sync(0xffffffffu, desc, D_CODE);
    if (code < 0) {
      run_element_warp<LAW>(A, code & 0x7fffffff, tab, wb, l);
    } else {
      run_gather_warp(A, code, __shfl_sync(0xffffffffu, desc, D_ROW0), __shfl_sync(0xffffffffu, desc, D_C0),
                      __shfl_sync(0xffffffffu, desc, D_C1), __shfl_sync(0xffffffffu, desc, D_N0),
                      __shfl_sync(0xffffffffu, desc, D_N1), wb, bars[warp], st, l);
    }
    t = t_next;
    desc = desc_next;
  }
}

template <int LAW, int W>
int launch_warp(const WarpArgs& A, cudaStream_t st) {
  const size_t smem = sizeof(double) * (WLayout::TAB_SIZE + (size_t)W * WLayout::WARP);
  auto k = staged_warp_kernel<LAW, W>;
  FEM_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0, dev = 0, sms = kNumSM;
  FEM_CUDA_CHECK(cudaGetDevice(&dev));
  FEM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  FEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 32 * W, smem));
  if (per_sm < 1) {
    set_error("fem_assemble_staged_warp: a CTA of %d warps does not fit an SM", W);
    return FEM_ECUDA;
  }
  FEM_CUDA_CHECK(cudaMemsetAsync(A.ctrl, 0, sizeof(int) * (kCtrlInts + (size_t)A.n_e + A.n_g), st));
  k<<<sms, 32 * W, smem, st>>>(A);                   // one CTA per SM: the whole grid is resident (the waits rely on it)
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

}  // namespace
}  // namespace femb200

using namespace femb200;

extern "C" int fem_staged_warp_count(int warps_per_sm) {
  int dev = 0, sms = kNumSM;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    sms = kNumSM;
  }
  return sms * warps_per_sm;
}

extern "C" int fem_assemble_staged_warp(int law_id, const double* law_params_host, const double* points, const double* sol,
                                        const double* internal_var, const double* ref_tables, int64_t n_cells,
                                        const int32_t* cells_p, const int32_t* corder, const int32_t* dest_row,
                                        const int32_t* prev_g, int64_t n_gather, const int32_t* tdesc,
                                        const int32_t* corner_eitem, const int32_t* nc_ptr, const int32_t* brow_ptr,
                                        const uint8_t* esrc, const uint8_t* bc_flag, double* stage, int32_t* ctrl, double* Re,
                                        double* data, int warps_per_sm, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(law_params_host && points && sol && ref_tables && cells_p && corder && dest_row && prev_g && tdesc && corner_eitem &&
                  nc_ptr && brow_ptr && esrc && stage && ctrl && Re && data, "null pointer");
  FEM_REQUIRE((reinterpret_cast<uintptr_t>(esrc) & 15) == 0 && (reinterpret_cast<uintptr_t>(stage) & 127) == 0, "esrc / stage alignment");
  FEM_REQUIRE(n_cells > 0 && n_gather > 0, "empty mesh");
  FEM_REQUIRE(!(law_id == FEM_LAW_SIMP && !internal_var), "SIMP needs the per-quadrature-point density");
  FEM_REQUIRE(law_id == FEM_LAW_LINEAR_ELASTIC || law_id == FEM_LAW_SIMP,
              "fem_assemble_staged_warp is registered for linear elasticity and SIMP on HEX8 / vec 3");
  WarpArgs A{};
  A.points = points; A.sol = sol; A.iv = internal_var; A.ref = ref_tables;
  A.cells_p = cells_p; A.corder = corder; A.dest_row = dest_row; A.prev_g = prev_g; A.Re = Re;
  A.tdesc = tdesc; A.corner_eitem = corner_eitem; A.nc_ptr = nc_ptr; A.brow_ptr = brow_ptr;
  A.esrc = reinterpret_cast<const uint4*>(esrc); A.bc_flag = bc_flag; A.data = data; A.stage = stage; A.ctrl = ctrl;
  A.C = n_cells;
  A.n_e = (int)((n_cells + 3) / 4);
  A.n_g = (int)n_gather;
  A.n_tickets = A.n_e + A.n_g;
  for (int i = 0; i < 8; ++i) A.p[i] = law_params_host[i];
  const double nu = law_id == FEM_LAW_SIMP ? A.p[2] : A.p[1];
  A.lam1 = nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
  A.mu1 = 1.0 / (2.0 * (1.0 + nu));
  cudaStream_t st = (cudaStream_t)stream;
#define FEM_W(WW)                                                                                                        \
  if (warps_per_sm == WW)                                                                                                \
    return law_id == FEM_LAW_SIMP ? launch_warp<FEM_LAW_SIMP, WW>(A, st) : launch_warp<FEM_LAW_LINEAR_ELASTIC, WW>(A, st);
  FEM_W(12) FEM_W(16) FEM_W(19)
#undef FEM_W
  set_error("fem_assemble_staged_warp: warps_per_sm must be 12, 16 or 19");
  return FEM_EINVAL;
}
