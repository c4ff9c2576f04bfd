// Staged assembly: element tangents -> CSR in ONE persistent kernel, the staging buffer kept in L2.
//
// Replaces, for HEX8 / vec 3 / {linear elasticity, SIMP, Neo-Hookean}, the chain
//   Problem.compute_newton_vars (jax_fem/problem.py:447-460) -> problem.V -> _PetscTangentCache.update / get_A
//   (jax_fem/solver.py:469-553: setValuesCOO + zeroRows)
// which the two-kernel path (element.cu + sparse.cu::gather_csr_kernel) runs through a C*8*8*3*3 buffer that is
// written to HBM and read back (4.6 GB each way at 100^3).  Here both steps are work items of one persistent grid:
//
//   E item i : 16 cells (in the plan's processing order).  Phase 1 as in element.cu (geometry of fe.py:112-141, grad u
//              and stress of problem.py:204-210), then per cell 18 + 6 mma.sync.m8n8k4.f64 and the accumulator
//              fragments go STRAIGHT from registers to the staging row of the corner: row (cell, a) = 9 tiles (I, J) of
//              8 doubles (b = 0..7), lane (a, t) writes 16 bytes of every tile -> each warp-wide store is 8 full 64-byte
//              pieces, no shared-memory staging and no G -> K transform here.
//   G item g : the CSR rows of ~4 mesh nodes (sparse.cu's work item): ONE TMA bulk copy of the nodes' staging rows into
//              shared memory, fixed-order sums per block entry, the isotropic map K = lam' G + mu' G^T + mu' tr(G) I
//              applied AFTER the sum (it is linear), Dirichlet rows -> unit rows, coalesced copy to `data`.
//
// Items are handed out by one atomic ticket over a plan-computed sequence in which every G item follows the E items of
// all cells around its nodes (plus a slack, so that they have normally finished), and staging rows are recycled: an E
// item may only overwrite rows whose previous G item is done.  Every item sets a done flag; a G item polls the flags of
// the (few) E items holding cells around its nodes, an E item the flags of the previous occupants of its 128 rows, one
// flag per thread in parallel.  The plan guarantees that these are tickets EARLIER than the waiting one, which are
// held by running CTAs: no deadlock, whatever the number of resident CTAs.  The ring
// holds a few thousand cells (tens of MB), is rewritten in place and therefore stays in the 126 MB L2: the element
// tangents never reach HBM.  Rows with a long life (nodes between two sweeps of the processing order) live in a
// spill area behind the ring.  Summation order is fixed by the plan => bit-reproducible, no atomics on data.
#include "common.cuh"
#include "element_math.cuh"

namespace femb200 {
namespace {

constexpr int kSThreads = 128;
constexpr int kCellsPerItem = 16;
constexpr int kCtrlInts = 16;          // ctrl[0] ticket, [3] error; then the done flags of the E and of the G items
constexpr unsigned kSpinLimit = 1u << 20;

struct StagedArgs {
  const double* points;
  const double* sol;
  const double* iv;
  const double* ref;
  const int32_t* cells_p;    // (C, 8) connectivity in processing order
  const int32_t* corder;     // (C) processing slot -> cell id
  const int32_t* dest_row;   // (C*8) staging row of corner (slot, a)
  const int32_t* prev_g;     // (C*8) G item that read the previous occupant of the corner's staging row, or -1
  double* Re;
  const int32_t* tdesc;      // (n_tickets, 32) ticket descriptors, see below
  const int32_t* gdep;       // overflow lists of tdesc
  const int4* emeta;
  const int32_t* src;
  double* data;
  double* stage;
  int* ctrl;
  int n_tickets, n_e, n_g;
  int64_t C;
  double p[8];
  double lam1, mu1;
  int transform;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {     // polling: ld.acquire would cost an L1 invalidation (CCTL.IVALL) per load
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_sc() { asm volatile("fence.sc.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Poll a done flag.  Legal waits are short (the awaited ticket is earlier than ours, so it is running); a wait that
// exceeds the limit is a plan bug: flag it and go on (wrong numbers, reported by the host) instead of hanging the GPU.
__device__ __forceinline__ void wait_flag(const int* flag, int* err) {
  unsigned spins = 0;
  while (ld_relaxed(flag) == 0) {        // relaxed: the caller issues one acquire fence after all of its flags are set
    __nanosleep(64);
    if ((++spins & 255u) == 0 && (spins > kSpinLimit || ld_acquire(err) != 0)) {
      atomicExch(err, 1);
      return;
    }
  }
}

// mbarrier wait with the same safety limit (a copy that was never issued must not hang the GPU)
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

// ---- E item -----------------------------------------------------------------------------------------------------------
struct ELayout {
  static constexpr int NN = 8, NQ = 8, DIM = 3, VEC = 3;
  static constexpr int TAB_STRIDE = 25, TAB_SIZE = NQ * TAB_STRIDE + NQ;      // 208
  static constexpr int GS = 28;                    // per-q stride of g[n][d]: 12 (mod 16) => conflict-free fragments
  static constexpr int OFF_X = 0, OFF_U = 24;      // per cell: X[8][3], U[8][3]
  static constexpr int OFF_G = 48;                 // g[q][n][d]
  static constexpr int OFF_S = OFF_G + NQ * GS;    // S[q][i][d] = sigma JxW
  static constexpr int OFF_E = OFF_S + NQ * 9;     // E_q JxW  (Neo-Hookean: the extra records start here)
  static constexpr int CELL = 354;                 // 352 + 2: cell stride 2 (mod 16) spreads phase-1 stores
  static constexpr int WARP = 4 * CELL + 18;       // + 32 staging rows and 4 cell ids (ints)
  static constexpr int WARPS = 4;
  static constexpr int DOUBLES = TAB_SIZE + WARPS * WARP;
};

__device__ __forceinline__ void fetch_desc(const StagedArgs& A, int* slot, int ticket);

template <int LAW>
__device__ __forceinline__ void run_element_item(const StagedArgs& A, int item, double* sm, int* next_slot, int next_ticket) {
  using L = ELayout;
  constexpr int NN = 8, NQ = 8, DIM = 3, VEC = 3, ND = 24;
  double* tab = sm;
  const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
  double* wb = sm + L::TAB_SIZE + warp * L::WARP;
  int* pos = reinterpret_cast<int*>(wb + 4 * L::CELL);      // [0,32): staging rows, [32,36): cell ids
  for (int i = threadIdx.x; i < NQ * NN * DIM; i += kSThreads)
    tab[(i / (NN * DIM)) * L::TAB_STRIDE + i % (NN * DIM)] = A.ref[i];
  if (threadIdx.x < NQ) tab[NQ * L::TAB_STRIDE + threadIdx.x] = A.ref[NQ * NN * DIM + threadIdx.x];

  // ---- phase 0/1: lane = (cell j, q) ----
  const int j1 = l >> 3, q = l & 7;
  const int64_t s1 = (int64_t)item * kCellsPerItem + warp * 4 + j1;
  const bool act1 = s1 < A.C;
  double* cb = wb + j1 * L::CELL;
  int c1 = 0;
  if (act1) {
    const int64_t node = A.cells_p[s1 * NN + q];
    c1 = A.corder[s1];
#pragma unroll
    for (int d = 0; d < DIM; ++d) cb[L::OFF_X + q * DIM + d] = A.points[node * DIM + d];
#pragma unroll
    for (int i = 0; i < VEC; ++i) cb[L::OFF_U + q * VEC + i] = A.sol[node * VEC + i];
    pos[l] = A.dest_row[s1 * NN + q];
    if (q == 0) pos[32 + j1] = c1;
  }
  __syncthreads();
  if (threadIdx.x == 0) fetch_desc(A, next_slot, next_ticket);   // descriptor of the item after next (lands during this item)
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
  }
  // the staging rows of this item may still be read by the G items of their previous occupants: one row per thread
  if (act1) {
    const int pg = A.prev_g[s1 * NN + q];
    if (pg >= 0) wait_flag(A.ctrl + kCtrlInts + A.n_e + pg, A.ctrl + 3);
  }
  __syncthreads();
  fence_acq_rel();                       // the stores below are ordered after the flag reads

  // ---- phase 2: the warp walks its 4 cells; lane = (node n, t) ----
  const int n = l >> 2, t = l & 3;
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
    const int64_t s = (int64_t)item * kCellsPerItem + warp * 4 + j;
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
    double* row = A.stage + (int64_t)pos[j * 8 + n] * 72 + 2 * t;
#pragma unroll
    for (int I = 0; I < 3; ++I)
#pragma unroll
      for (int J = 0; J < 3; ++J) {
        double Cc[2] = {0.0, 0.0};
        dmma884(Cc, a0[I], g0[J]);
        dmma884(Cc, a1[I], g1[J]);
        *reinterpret_cast<double2*>(row + (I * 3 + J) * 8) = make_double2(Cc[0], Cc[1]);
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
    if (t == 0) {
      A.Re[c * ND + n * 3 + 0] = R[0];
      A.Re[c * ND + n * 3 + 1] = R[1];
    } else if (t == 1) {
      A.Re[c * ND + n * 3 + 2] = R[0];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) st_release(A.ctrl + kCtrlInts + item, 1);   // cumulative over the CTA's stores (ordered by the barrier)
}

// ---- G item -----------------------------------------------------------------------------------------------------------
constexpr int kMaxC = 48;                         // corners of a G item: 32 + the tail of its last node (<= 16)
constexpr int kRow = 72;                          // doubles per staging row
constexpr int kMaxS = kMaxC * 8;
constexpr int kEPT = (kMaxS + kSThreads - 1) / kSThreads;   // 3

// Ticket descriptor (32 ints, one 128-byte line per ticket, in ticket order): [0] E: 0x80000000 | item, G: item;
// G only: [1] first staging row, [2..5] first corner / entry / source / emeta row, [6..9] their ends, [10] number of E
// items holding cells around the item's nodes, [11] -1 or their offset in `gdep` when they do not fit, [12..31] those E items.
constexpr int kDescInts = 32, kDescInline = 20, kDescSlots = 3;
enum { D_CODE = 0, D_ROW0, D_C0, D_E0, D_S0, D_M0, D_C1, D_E1, D_S1, D_M1, D_NDEP, D_OVF, D_DEPS };

struct GMeta {   // per-thread share of an item's metadata, prefetched one item ahead
  int soff[kEPT], sb[kEPT], se[kEPT], dst[kEPT], info[kEPT];
};

__device__ __forceinline__ void g_load_meta(const StagedArgs& A, const int* d, GMeta& m) {
  const int S0 = d[D_S0], nS = d[D_S1] - S0, M0 = d[D_M0], nM = d[D_M1] - M0;
#pragma unroll
  for (int r = 0; r < kEPT; ++r) {
    int t = threadIdx.x + r * kSThreads;
    m.soff[r] = (t < nS) ? __ldcs(A.src + S0 + t) : 0;
    m.sb[r] = m.se[r] = 0;
    m.dst[r] = -1;
    m.info[r] = 0;
    // rows come sorted by descending source count: the second round is dealt out in reverse thread order, so that the
    // threads with the lightest first row get the extra one (the warps reach the barrier after the sums together)
    if (r == 1) t = 2 * kSThreads - 1 - threadIdx.x;
    if (t < nM) {
      const int4 em = __ldcs(A.emeta + M0 + t);
      m.sb[r] = em.x;
      m.se[r] = em.y;
      m.dst[r] = em.z;
      m.info[r] = em.w;
    }
  }
}

// done flag of the thread's dependency of G item `d` (one E item per thread), or nullptr
__device__ __forceinline__ const int* g_dep_flag(const StagedArgs& A, const int* d) {
  const int nd = d[D_NDEP], t = threadIdx.x;
  if (t >= nd) return nullptr;
  const int e = d[D_OVF] >= 0 ? A.gdep[d[D_OVF] + t] : d[D_DEPS + t];
  return A.ctrl + kCtrlInts + e;
}

// one elected thread, once the item's E items are done: start the bulk copy of its staging rows
__device__ __forceinline__ void g_issue(const StagedArgs& A, const int* d, double* stage_smem, uint64_t* bar) {
  fence_acq_rel();                                 // acquire side of the E items' release stores (the flags were polled relaxed)
  fence_proxy_async_all();                         // generic-proxy writes of the E items -> async-proxy read below
  const uint32_t bytes = (uint32_t)(d[D_C1] - d[D_C0]) * kRow * sizeof(double);
  mbar_expect_tx(bar, bytes);
  if (bytes) bulk_g2s(stage_smem, A.stage + (int64_t)d[D_ROW0] * kRow, bytes, bar);
}

// ticket -> descriptor slot: asynchronous 128-byte copy issued by ONE thread (tickets past the end read as the END code)
__device__ __forceinline__ void fetch_desc(const StagedArgs& A, int* slot, int ticket) {
  if (ticket < A.n_tickets) {
    const int32_t* srcp = A.tdesc + (int64_t)ticket * kDescInts;
#pragma unroll
    for (int i = 0; i < kDescInts / 4; ++i)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(slot + 4 * i)), "l"(srcp + 4 * i) : "memory");
  } else {
    slot[D_CODE] = -1;
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void fetch_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Every staging row is read exactly once: after the bulk copy has landed its lines are dead.  Dropping them from L2
// (discard.L2: no write-back) keeps the dirty ring lines from ever being evicted to HBM.  Only lines that lie entirely
// inside the item's rows are dropped (rows are 4.5 lines long, the first and last line may be shared with neighbours).
__device__ __forceinline__ void discard_rows(const double* first, int n_rows) {
  const uintptr_t lo = (reinterpret_cast<uintptr_t>(first) + 127) & ~(uintptr_t)127;
  const uintptr_t hi = (reinterpret_cast<uintptr_t>(first) + (size_t)n_rows * kRow * sizeof(double)) & ~(uintptr_t)127;
  for (uintptr_t a = lo + 128 * (uintptr_t)threadIdx.x; a < hi; a += 128 * kSThreads)
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(a) : "memory");
}

template <int LAW>
__global__ void __launch_bounds__(kSThreads, 4) staged_assembly_kernel(const StagedArgs A) {
  constexpr int VEC = 3, VV = 9, NN = 8;
  extern __shared__ __align__(128) double ssm[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ __align__(16) int tq[kDescSlots][kDescInts];
  unsigned short* s_off = reinterpret_cast<unsigned short*>(ssm + 2 * kMaxC * kRow);   // source block inside the stage (< 384)
  int* const err = A.ctrl + 3;
  int* const flag_e = A.ctrl + kCtrlInts;
  int* const flag_g = flag_e + A.n_e;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    const int base = atomicAdd(A.ctrl, 2);
    fetch_desc(A, tq[0], base);
    fetch_desc(A, tq[1], base + 1);
    fetch_wait();
  }
  __syncthreads();

  GMeta cm, nm;
  int gcount = 0;                 // G items processed by this CTA: stage = gcount & 1, mbarrier parity = (gcount >> 1) & 1
  bool cur_issued = false;        // the bulk copy of the current G item is already in flight
  if (tq[0][D_CODE] >= 0) g_load_meta(A, tq[0], cm);

  for (int k = 0;; ++k) {
    const int* dc = tq[k % 3];
    const int* dn = tq[(k + 1) % 3];
    int* dnn = tq[(k + 2) % 3];
    const int code = dc[D_CODE];
    if (code == -1) break;
    int t2 = 0;
    if (threadIdx.x == 0) t2 = atomicAdd(A.ctrl, 1);          // ticket of item k + 2; used mid-way through this item
    const bool next_is_g = dn[D_CODE] >= 0;
    // the item after this one: its metadata and the state of its dependencies travel while this one is processed
    int nflag = 1;
    if (next_is_g) {
      g_load_meta(A, dn, nm);
      const int* f = g_dep_flag(A, dn);
      if (f) nflag = ld_relaxed(f);
    }
    bool nxt_issued = false;
    if (code < 0) {
      run_element_item<LAW>(A, code & 0x7fffffff, ssm, dnn, t2);
    } else {
      const int stage = gcount & 1;
      double* sh = ssm + stage * (kMaxC * kRow);
      if (!cur_issued) {
        const int* f = g_dep_flag(A, dc);
        if (f) wait_flag(f, err);
        __syncthreads();
        if (threadIdx.x == 0) g_issue(A, dc, sh, &bar[stage]);
      }
      const int C0 = dc[D_C0], E0 = dc[D_E0], S0 = dc[D_S0], nE = dc[D_E1] - E0, nS = dc[D_S1] - S0;
#pragma unroll
      for (int r = 0; r < kEPT; ++r) {
        const int t = threadIdx.x + r * kSThreads;
        if (t < nS) s_off[t] = (unsigned short)(cm.soff[r] - C0 * NN);
      }
      // barrier after s_off; at the same time: are the next item's E items done?  Then its rows go into the other stage now
      nxt_issued = __syncthreads_and(nflag != 0) && next_is_g;
      if (threadIdx.x == 0) {
        if (nxt_issued) g_issue(A, dn, ssm + (stage ^ 1) * (kMaxC * kRow), &bar[stage ^ 1]);
        fetch_desc(A, dnn, t2);
      }
      mbar_wait_bounded(&bar[stage], (gcount >> 1) & 1, err);
      // the staging rows have landed in shared memory: they are dead in L2 and may be overwritten from now on
      discard_rows(A.stage + (int64_t)dc[D_ROW0] * kRow, dc[D_C1] - C0);
      __syncthreads();
      if (threadIdx.x == 0) st_release(flag_g + code, 1);

      // phase B: per entry, add its sources in the plan's fixed order; source block (row r, b) = sh[r*72 + IJ*8 + b]
      double res[kEPT][VV];
#pragma unroll
      for (int r = 0; r < kEPT; ++r) {
#pragma unroll
        for (int j = 0; j < VV; ++j) res[r][j] = 0.0;
        for (int sidx = cm.sb[r] - S0; sidx < cm.se[r] - S0; ++sidx) {
          const int so = s_off[sidx];
          const double* blk = sh + (so >> 3) * kRow + (so & 7);
#pragma unroll
          for (int j = 0; j < VV; ++j) res[r][j] += blk[j * 8];
        }
        if (A.transform) {   // K = lam' G + mu' G^T + mu' tr(G) I  (linear: applied to the sum, or to each half of a split entry)
          double G[VV];
#pragma unroll
          for (int j = 0; j < VV; ++j) G[j] = res[r][j];
          const double tr = A.mu1 * (G[0] + G[4] + G[8]);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int kk = 0; kk < 3; ++kk)
              res[r][i * 3 + kk] = A.lam1 * G[i * 3 + kk] + A.mu1 * G[kk * 3 + i] + (i == kk ? tr : 0.0);
        }
      }
      __syncthreads();

      // phase C: results -> shared memory in CSR order, then a coalesced copy (see sparse.cu::gather_csr_kernel)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int r = 0; r < kEPT; ++r) {
          if (cm.dst[r] >= 0 && ((cm.info[r] >> 20) & 1) == half) {
            const int rowlen = cm.info[r] & 0xffff;
            const int dst = cm.dst[r] - VV * E0;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              const bool bc = (cm.info[r] >> (17 + i)) & 1;
#pragma unroll
              for (int kk = 0; kk < VEC; ++kk) {
                double* o = sh + dst + i * rowlen + kk;
                if (half == 0) *o = bc ? ((((cm.info[r] >> 16) & 1) && i == kk) ? 1.0 : 0.0) : res[r][i * VEC + kk];
                else if (!bc) *o += res[r][i * VEC + kk];
              }
            }
          }
        }
        __syncthreads();
      }
      double* __restrict__ out = A.data + (int64_t)VV * E0;
      for (int t = threadIdx.x; t < nE * VV; t += kSThreads) __stcs(out + t, sh[t]);   // streaming: do not displace the ring
      ++gcount;
    }
    if (threadIdx.x == 0) fetch_wait();     // descriptor of item k + 2
    __syncthreads();
    cm = nm;
    cur_issued = nxt_issued;
  }
}

template <int LAW>
int launch_staged(const StagedArgs& A, cudaStream_t st) {
  const size_t smem_g = sizeof(double) * 2 * kMaxC * kRow + sizeof(unsigned short) * kMaxS;
  const size_t smem_e = sizeof(double) * ELayout::DOUBLES;
  const size_t smem = smem_g > smem_e ? smem_g : smem_e;
  auto k = staged_assembly_kernel<LAW>;
  FEM_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0, dev = 0, sms = kNumSM;
  FEM_CUDA_CHECK(cudaGetDevice(&dev));
  FEM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  FEM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kSThreads, smem));
  if (per_sm < 1) per_sm = 1;
  int grid = per_sm * sms;
  if (grid > A.n_tickets) grid = A.n_tickets;
  FEM_CUDA_CHECK(cudaMemsetAsync(A.ctrl, 0, sizeof(int) * (kCtrlInts + (size_t)A.n_e + A.n_g), st));
  k<<<grid, kSThreads, smem, st>>>(A);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

}  // namespace
}  // namespace femb200

using namespace femb200;

extern "C" int64_t fem_staged_ctrl_ints(int64_t n_e, int64_t n_g) { return kCtrlInts + n_e + n_g; }

extern "C" int fem_assemble_staged(int law_id, const double* law_params_host, const double* points, const double* sol,
                                   const double* internal_var, const double* ref_tables, int64_t n_cells,
                                   const int32_t* cells_p, const int32_t* corder, const int32_t* dest_row,
                                   const int32_t* prev_g, int64_t n_gather, const int32_t* tdesc, const int32_t* gdep,
                                   const int32_t* emeta, const int32_t* src, double* stage,
                                   int32_t* ctrl, double* Re, double* data, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(law_params_host && points && sol && ref_tables && cells_p && corder && dest_row && prev_g && tdesc && gdep &&
                  emeta && src && stage && ctrl && Re && data, "null pointer");
  FEM_REQUIRE((reinterpret_cast<uintptr_t>(emeta) & 15) == 0 && (reinterpret_cast<uintptr_t>(tdesc) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(stage) & 127) == 0, "emeta / tdesc / stage alignment");
  FEM_REQUIRE(n_cells > 0 && n_gather > 0, "empty mesh");
  FEM_REQUIRE(!(law_id == FEM_LAW_SIMP && !internal_var), "SIMP needs the per-quadrature-point density");
  StagedArgs A{};
  A.points = points; A.sol = sol; A.iv = internal_var; A.ref = ref_tables;
  A.cells_p = cells_p; A.corder = corder; A.dest_row = dest_row; A.prev_g = prev_g; A.Re = Re;
  A.tdesc = tdesc; A.gdep = gdep; A.emeta = reinterpret_cast<const int4*>(emeta); A.src = src; A.data = data;
  A.stage = stage; A.ctrl = ctrl;
  A.C = n_cells;
  A.n_e = (int)((n_cells + kCellsPerItem - 1) / kCellsPerItem);
  A.n_g = (int)n_gather;
  A.n_tickets = A.n_e + A.n_g;
  for (int i = 0; i < 8; ++i) A.p[i] = law_params_host[i];
  cudaStream_t st = (cudaStream_t)stream;
  if (law_id == FEM_LAW_LINEAR_ELASTIC || law_id == FEM_LAW_SIMP) {
    const double nu = law_id == FEM_LAW_SIMP ? A.p[2] : A.p[1];
    A.lam1 = nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
    A.mu1 = 1.0 / (2.0 * (1.0 + nu));
    A.transform = 1;
    return law_id == FEM_LAW_SIMP ? launch_staged<FEM_LAW_SIMP>(A, st) : launch_staged<FEM_LAW_LINEAR_ELASTIC>(A, st);
  }
  set_error("fem_assemble_staged: unregistered law %d (registered: linear elasticity, SIMP on HEX8 / vec 3)", law_id);
  return FEM_EINVAL;
}

// ctrl[3] != 0 after a launch: a wait exceeded its limit (plan bug); the values of that assembly are invalid
extern "C" int fem_staged_status(const int32_t* ctrl, int32_t* status_host, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(ctrl && status_host, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FEM_CUDA_CHECK(cudaMemcpyAsync(status_host, ctrl + 3, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  return FEM_OK;
}
