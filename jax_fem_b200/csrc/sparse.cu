// Global assembly (deterministic gather), Dirichlet operations, CSR SpMV and small vector kernels.
#include <stdarg.h>
#include "common.cuh"

namespace femb200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_device() {
  static int cached = -1;
  if (cached < 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
      set_error("no CUDA device (%s): libfem_b200 has no CPU fallback", cudaGetErrorString(e));
      cudaGetLastError();
      return FEM_ENODEV;
    }
    cached = n;
  }
  return FEM_OK;
}

namespace {

// ---- get_A: deterministic segmented sum of the element row blocks into CSR -------------------------
// The element kernel stores the row block of every corner (cell, a) in node-sorted order, so the row
// blocks of a mesh node are one contiguous segment ("COO sorted by row").  CTA b owns the nodes whose
// first corner lies in [kGatherCorners*b, kGatherCorners*(b+1)):
//   phase A: its segment is brought into shared memory by ONE TMA bulk copy (cp.async.bulk + mbarrier: no
//            registers, no indirection, the whole segment in flight at once) while the threads fetch the
//            per-source block offsets and the per-entry metadata;
//   phase B: one thread per output scalar, ordered (row i, entry, column k) so that consecutive threads
//            write consecutive addresses of a CSR row; the sources of an entry are added in ascending
//            (cell, a, b) order -- the same fixed order on every run => bit-reproducible, no atomics.
// Dirichlet rows become unit rows with the pattern kept (zeroRows, solver.py:527-528).
constexpr int kGatherThreads = 128;
// corners per work item / max corners of one node (checked when the plan is built; mirrored by plan.py::gather_config)
template <int NN> struct GatherCfg { static constexpr int CORNERS = NN <= 8 ? 32 : 8, TAIL = NN <= 8 ? 16 : 8; };

// Per-item metadata a thread keeps in registers (prefetched one work item ahead).
template <int EPT, int SPT>
struct GatherMeta {
  int C0, C1, E0, E1, S0, S1, M0, M1;   // corner / entry / source / gather-row ranges of the item
  int soff[SPT];                   // raw block indices of the thread's sources
  int sb[EPT], se[EPT], dst[EPT], info[EPT];
};

// TILES: the rows come from fem_element_tiles (tile-major: block (row r, b) = 9 doubles at r*72 + IJ*8 + b) and, with
// post.x != 0, hold G: the isotropic map K = lam' G + mu' G^T + mu' tr(G) I (post.y = lam', post.z = mu') is applied to the sum.
template <int VEC, int NN, bool TILES>
__global__ void __launch_bounds__(kGatherThreads) gather_csr_kernel(
    int n_items, const int32_t* __restrict__ gdesc, const int4* __restrict__ emeta, const int32_t* __restrict__ src,
    const double* __restrict__ Ke, double* __restrict__ data, const double3 post) {
  constexpr int VV = VEC * VEC;
  constexpr int ROW = (NN * VV + 1) / 2 * 2;                      // doubles per corner row block (16-byte multiple)
  constexpr int MAXC = GatherCfg<NN>::CORNERS + GatherCfg<NN>::TAIL;
  constexpr int MAXS = MAXC * NN;                                 // sources (blocks) of an item
  constexpr int EPT = (MAXS + kGatherThreads - 1) / kGatherThreads;   // entries / sources per thread (worst case)
  static_assert(ROW % 2 == 0, "row blocks must be 16-byte multiples");
  extern __shared__ __align__(128) double gsm[];
  auto stage_ptr = [&](int stage) { return gsm + stage * (MAXC * ROW); };   // two TMA stages
  int* s_off = reinterpret_cast<int*>(gsm + 2 * MAXC * ROW);      // block offset (in blocks) inside the stage
  __shared__ __align__(8) uint64_t bar[2];
  using Meta = GatherMeta<EPT, EPT>;

  auto load_desc = [&](int it, Meta& m) {
    m.C0 = gdesc[it * 4 + 0]; m.C1 = gdesc[it * 4 + 4];
    m.E0 = gdesc[it * 4 + 1]; m.E1 = gdesc[it * 4 + 5];
    m.S0 = gdesc[it * 4 + 2]; m.S1 = gdesc[it * 4 + 6];
    m.M0 = gdesc[it * 4 + 3]; m.M1 = gdesc[it * 4 + 7];
  };
  auto load_meta = [&](Meta& m) {          // global loads only; consumed one iteration later
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
      const int t = threadIdx.x + r * kGatherThreads;
      m.soff[r] = (t < m.S1 - m.S0) ? src[m.S0 + t] : 0;
      m.sb[r] = m.se[r] = 0;
      m.dst[r] = -1;   // raw (absolute) values; an entry slot beyond the item keeps dst < 0
      m.info[r] = 0;
      if (t < m.M1 - m.M0) {
        // gather rows in processing order: one independent 16-byte load each; the raw values are kept (nothing is
        // computed from them here) so that the load stays in flight until the item is processed
        const int4 em = emeta[m.M0 + t];
        m.sb[r] = em.x;
        m.se[r] = em.y;
        m.dst[r] = em.z;
        m.info[r] = em.w;
      }
    }
  };
  auto issue_tma = [&](const Meta& m, int stage) {   // one elected thread; the segment is contiguous and 16B-aligned
    const uint32_t bytes = (uint32_t)(m.C1 - m.C0) * ROW * sizeof(double);
    mbar_expect_tx(&bar[stage], bytes);            // also the (single) arrival: an empty item completes the phase at once
    if (bytes) bulk_g2s(stage_ptr(stage), Ke + (int64_t)m.C0 * ROW, bytes, &bar[stage]);
  };

  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
  }
  __syncthreads();
  int it = blockIdx.x;
  if (it >= n_items) return;
  Meta cur, nxt;
  load_desc(it, cur);
  if (threadIdx.x == 0) issue_tma(cur, 0);
  load_meta(cur);
  const int it1 = it + gridDim.x;
  if (it1 < n_items) load_desc(it1, nxt);

  for (int k = 0; it < n_items; ++k, it += gridDim.x) {
    const int stage = k & 1;
    const bool has_next = it + (int)gridDim.x < n_items;
    // prefetch the next item: its row blocks by TMA into the other stage (free since the end of iteration k-1),
    // its metadata into registers; both are consumed in iteration k+1
    if (has_next) {
      if (threadIdx.x == 0) issue_tma(nxt, stage ^ 1);
      load_meta(nxt);
    }
    Meta nn2;
    const bool has_next2 = it + 2 * (int)gridDim.x < n_items;
    if (has_next2) load_desc(it + 2 * gridDim.x, nn2);

    const int nE = cur.E1 - cur.E0, nS = cur.S1 - cur.S0;
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
      const int t = threadIdx.x + r * kGatherThreads;
      if (t < nS) s_off[t] = cur.soff[r] - cur.C0 * NN;
    }
    __syncthreads();
    mbar_wait(&bar[stage], (k >> 1) & 1);
    const double* sh = stage_ptr(stage);

    // phase B: per entry, add its sources in ascending (cell, a, b) order -- fixed order => bit-reproducible
    double res[EPT][VV];
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
#pragma unroll
      for (int j = 0; j < VV; ++j) res[r][j] = 0.0;
      for (int sidx = cur.sb[r] - cur.S0; sidx < cur.se[r] - cur.S0; ++sidx) {
        const int so = s_off[sidx];
        if constexpr (TILES) {
          const double* blk = sh + (so / NN) * ROW + (so % NN);
#pragma unroll
          for (int j = 0; j < VV; ++j) res[r][j] += blk[j * NN];
        } else {
          const double* blk = (ROW == NN * VV) ? sh + so * VV : sh + (so / NN) * ROW + (so % NN) * VV;
#pragma unroll
          for (int j = 0; j < VV; ++j) res[r][j] += blk[j];
        }
      }
      if constexpr (TILES && VEC == 3) {
        if (post.x != 0.0) {     // linear: applied to the sum (or to each half of a split entry)
          double G[VV];
#pragma unroll
          for (int j = 0; j < VV; ++j) G[j] = res[r][j];
          const double tr = post.z * (G[0] + G[4] + G[8]);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) res[r][i * 3 + kk] = post.y * G[i * 3 + kk] + post.z * G[kk * 3 + i] + (i == kk ? tr : 0.0);
        }
      }
    }
    __syncthreads();

    // phase C: results -> shared memory in CSR order (the item's rows are one contiguous range of `data`), then a
    // coalesced copy.  einfo: bits 0..15 = VEC*len(row node), bit 16 = diagonal block, bits 17.. = Dirichlet rows.
    // A row with bit 20 is the second half of a split entry: it adds to what the first half stored (fixed order).
    double* so = stage_ptr(stage);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int r = 0; r < EPT; ++r) {
        if (cur.dst[r] >= 0 && ((cur.info[r] >> 20) & 1) == half) {
          const int rowlen = cur.info[r] & 0xffff;
          const int dst = cur.dst[r] - VV * cur.E0;
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const bool bc = (cur.info[r] >> (17 + i)) & 1;
#pragma unroll
            for (int kk = 0; kk < VEC; ++kk) {
              double* o = so + dst + i * rowlen + kk;
              if (half == 0) *o = bc ? ((((cur.info[r] >> 16) & 1) && i == kk) ? 1.0 : 0.0) : res[r][i * VEC + kk];
              else if (!bc) *o += res[r][i * VEC + kk];
            }
          }
        }
      }
      __syncthreads();
    }
    double* __restrict__ out = data + (int64_t)VV * cur.E0;
    for (int t = threadIdx.x; t < nE * VV; t += kGatherThreads) out[t] = so[t];
    __syncthreads();                       // stage is free for the TMA issued in the next iteration
    cur = nxt;
    nxt = nn2;
  }
}

template <int VEC, int NN>
__global__ void gather_residual_kernel(int64_t n_nodes, const int32_t* __restrict__ nc_ptr,
                                       const int32_t* __restrict__ nc, const double* __restrict__ Re,
                                       const double* __restrict__ f_ext, double* __restrict__ res) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  double acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.0;
  for (int p = nc_ptr[n]; p < nc_ptr[n + 1]; ++p) {
    const double* r = Re + (int64_t)nc[p] * VEC;      // code = c*NN + a  ->  Re[(c*NN + a)*VEC]
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] += r[i];
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) res[n * VEC + i] = acc[i] + (f_ext ? f_ext[n * VEC + i] : 0.0);
}

__global__ void apply_bc_vec_kernel(int64_t n_bc, const int32_t* __restrict__ rows, const double* __restrict__ vals,
                                    double scale, const double* __restrict__ sol, double* __restrict__ res) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_bc) {
    const int r = rows[i];
    res[r] = sol[r] - vals[i] * scale;                 // solver.py:299-301
  }
}

__global__ void bc_x0_kernel(int64_t n_bc, const int32_t* __restrict__ rows, const double* __restrict__ vals,
                             const double* __restrict__ dofs, double* __restrict__ x0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_bc) {
    const int r = rows[i];
    x0[r] = vals[i] - dofs[r];                         // solver.py:403-409
  }
}

// ---- CSR SpMV: LPR lanes per row, 64-bit values + 32-bit columns streamed at full sector width --
template <int LPR>
__global__ void __launch_bounds__(256) spmv_kernel(int64_t n, const int32_t* __restrict__ indptr,
                                                   const int32_t* __restrict__ indices,
                                                   const double* __restrict__ data, const double* __restrict__ x,
                                                   double* __restrict__ y) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int sub = threadIdx.x % LPR;
  double acc = 0.0;
  if (row < n) {
    const int s = indptr[row], e = indptr[row + 1];
    for (int j = s + sub; j < e; j += LPR) acc = fma(data[j], __ldg(x + indices[j]), acc);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (row < n && sub == 0) y[row] = acc;
}

__global__ void csr_diag_kernel(int64_t n, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                const double* __restrict__ data, double* __restrict__ diag) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  int lo = indptr[row], hi = indptr[row + 1] - 1;
  double d = 0.0;
  while (lo <= hi) {
    const int mid = lo + ((hi - lo) >> 1);      // lo + hi overflows int32 once nnz > 2^30 (200^3 HEX8: nnz = 1.95e9)
    const int c = indices[mid];
    if (c == row) { d = data[mid]; break; }
    if (c < row) lo = mid + 1; else hi = mid - 1;
  }
  diag[row] = d;
}

template <int VEC>
__global__ void transpose_values_kernel(int64_t n_nodes, const int32_t* __restrict__ brow_ptr,
                                        const int32_t* __restrict__ bcol, const int32_t* __restrict__ tperm, const double* __restrict__ data,
                                        double* __restrict__ data_t) {
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= n_nodes) return;
  const int e0 = brow_ptr[n], len = brow_ptr[n + 1] - e0;
  const int64_t row0 = (int64_t)VEC * VEC * e0;
  for (int s = (threadIdx.x & 31); s < len; s += 32) {
    // entry (n, m) at slot s; its transpose (m, n) is entry te of row block m
    const int te = tperm[e0 + s];
    const int64_t m = bcol[e0 + s];
    const int f0 = brow_ptr[m], lenm = brow_ptr[m + 1] - f0, sm = te - f0;
    const int64_t rowm = (int64_t)VEC * VEC * f0;
#pragma unroll
    for (int i = 0; i < VEC; ++i)
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        data_t[row0 + (int64_t)i * VEC * len + (int64_t)VEC * s + k] =
            data[rowm + (int64_t)k * VEC * lenm + (int64_t)VEC * sm + i];
  }
}

__global__ void axpy_kernel(int64_t n, double alpha, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fma(alpha, x[i], y[i]);
}

__global__ void __launch_bounds__(256) dot_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                                                  double* __restrict__ partial) {
  __shared__ double red[8];
  double v[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[0] = fma(x[i], y[i], v[0]);
  block_sum<1, 256>(v, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = v[0];
}

__global__ void final_sum_kernel(int nblocks, const double* __restrict__ partial, double* __restrict__ out) {
  __shared__ double red[8];
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < nblocks; i += 256) v[0] += partial[i];
  block_sum<1, 256>(v, red);
  if (threadIdx.x == 0) *out = v[0];
}

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace
}  // namespace femb200

using namespace femb200;

extern "C" const char* fem_last_error(void) { return g_err; }
extern "C" int fem_version(void) { return 100; }
extern "C" int fem_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device (%s)", cudaGetErrorString(e));
    return FEM_ENODEV;
  }
  return n;
}

namespace femb200 {
namespace {
template <int VEC, int NN, bool TILES = false>
int launch_gather(int n_items, const int32_t* gdesc, const int32_t* emeta, const int32_t* src, const double* Ke,
                  double* data, cudaStream_t st, double3 post = make_double3(0., 0., 0.)) {
  constexpr int MAXC = GatherCfg<NN>::CORNERS + GatherCfg<NN>::TAIL;
  constexpr int ROW = (NN * VEC * VEC + 1) / 2 * 2;
  const size_t smem = sizeof(double) * 2 * MAXC * ROW + sizeof(int) * MAXC * NN;
  auto k = gather_csr_kernel<VEC, NN, TILES>;
  // persistent grid: resident CTAs x SMs, cached per device (the shared-memory attribute is per device too)
  static int grids[64] = {0};
  int dev = 0;
  FEM_CUDA_CHECK(cudaGetDevice(&dev));
  int& grid = grids[dev & 63];
  if (!grid) {
    FEM_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, sms = kNumSM;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kGatherThreads, smem);
    grid = (per_sm < 1 ? 1 : per_sm) * sms;
  }
  k<<<grid < n_items ? grid : n_items, kGatherThreads, smem, st>>>(n_items, gdesc, reinterpret_cast<const int4*>(emeta), src, Ke, data, post);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}
}  // namespace
}  // namespace femb200

extern "C" int fem_gather_csr(int vec, int nn, int64_t n_blocks, const int32_t* gdesc, const int32_t* emeta,
                              const int32_t* src, const double* Ke, double* data, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(gdesc && emeta && src && Ke && data, "null pointer");
  FEM_REQUIRE((reinterpret_cast<uintptr_t>(emeta) & 15) == 0, "emeta must be 16-byte aligned");
  if (n_blocks == 0) return FEM_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define FEM_G(V, N)                                                                                                  \
  if (vec == V && nn == N) {                                                                                         \
    return launch_gather<V, N>((int)n_blocks, gdesc, emeta, src, Ke, data, st);              \
  }
  FEM_G(3, 8) FEM_G(1, 8) FEM_G(1, 4) FEM_G(2, 4) FEM_G(3, 27)
#undef FEM_G
  set_error("fem_gather_csr: unregistered (vec=%d, nodes/cell=%d)", vec, nn);
  return FEM_EINVAL;
}

extern "C" int fem_gather_csr_tiles(int64_t n_blocks, const int32_t* gdesc, const int32_t* emeta, const int32_t* src,
                                    const double* Ke_tiles, double* data, const double* post_host, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(gdesc && emeta && src && Ke_tiles && data && post_host, "null pointer");
  FEM_REQUIRE((reinterpret_cast<uintptr_t>(emeta) & 15) == 0, "emeta must be 16-byte aligned");
  if (n_blocks == 0) return FEM_OK;
  return launch_gather<3, 8, true>((int)n_blocks, gdesc, emeta, src, Ke_tiles, data, (cudaStream_t)stream,
                                   make_double3(post_host[0], post_host[1], post_host[2]));
}

extern "C" int fem_gather_residual(int vec, int nn, int64_t n_nodes, const int32_t* nc_ptr, const int32_t* nc,
                                   const double* Re, const double* f_ext, double* res, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(nc_ptr && nc && Re && res, "null pointer");
  if (n_nodes == 0) return FEM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = blocks_for(n_nodes, 256);
#define FEM_G(V, N)                                                                                  \
  if (vec == V && nn == N) {                                                                         \
    gather_residual_kernel<V, N><<<grid, 256, 0, st>>>(n_nodes, nc_ptr, nc, Re, f_ext, res);         \
    FEM_LAUNCH_CHECK();                                                                              \
    return FEM_OK;                                                                                   \
  }
  FEM_G(3, 8) FEM_G(1, 8) FEM_G(1, 4) FEM_G(2, 4) FEM_G(3, 27) FEM_G(1, 27)
#undef FEM_G
  set_error("fem_gather_residual: unregistered (vec=%d, nodes/cell=%d)", vec, nn);
  return FEM_EINVAL;
}

extern "C" int fem_apply_bc_vec(int64_t n_bc, const int32_t* bc_rows, const double* bc_vals, double scale,
                                const double* sol, double* res, void* stream) {
  if (int e = check_device()) return e;
  if (n_bc == 0) return FEM_OK;
  FEM_REQUIRE(bc_rows && bc_vals && sol && res, "null pointer");
  apply_bc_vec_kernel<<<blocks_for(n_bc, 256), 256, 0, (cudaStream_t)stream>>>(n_bc, bc_rows, bc_vals, scale, sol, res);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

extern "C" int fem_bc_initial_guess(int64_t n, int64_t n_bc, const int32_t* bc_rows, const double* bc_vals,
                                    const double* dofs, double* x0, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(dofs && x0, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FEM_CUDA_CHECK(cudaMemsetAsync(x0, 0, sizeof(double) * n, st));
  if (n_bc == 0) return FEM_OK;
  bc_x0_kernel<<<blocks_for(n_bc, 256), 256, 0, st>>>(n_bc, bc_rows, bc_vals, dofs, x0);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

namespace femb200 {
int launch_spmv(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data, const double* x,
                double* y, cudaStream_t st) {
  constexpr int LPR = 8;
  spmv_kernel<LPR><<<blocks_for(n * LPR, 256), 256, 0, st>>>(n, indptr, indices, data, x, y);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}
}  // namespace femb200

namespace femb200 {
int launch_spmv_block(int64_t n, int vec, const int32_t* brow_ptr, const int32_t* bcol, const double* data,
                      const double* x, double* y, cudaStream_t st);
}

extern "C" int fem_spmv(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data, int vec,
                        const int32_t* brow_ptr, const int32_t* bcol, const double* x, double* y, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(indptr && indices && data && x && y, "null pointer");
  if (n == 0) return FEM_OK;
  if (brow_ptr && bcol && (vec == 2 || vec == 3)) return launch_spmv_block(n, vec, brow_ptr, bcol, data, x, y, (cudaStream_t)stream);
  return launch_spmv(n, indptr, indices, data, x, y, (cudaStream_t)stream);
}

extern "C" int fem_csr_diagonal(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data,
                                double* diag, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(indptr && indices && data && diag, "null pointer");
  if (n == 0) return FEM_OK;
  csr_diag_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, indptr, indices, data, diag);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

extern "C" int fem_csr_transpose_values(int vec, int64_t n_nodes, const int32_t* brow_ptr, const int32_t* bcol,
                                        const int32_t* tperm,
                                        const double* data, double* data_t, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(brow_ptr && bcol && tperm && data && data_t, "null pointer");
  if (n_nodes == 0) return FEM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = blocks_for(n_nodes, 8);
  if (vec == 1) transpose_values_kernel<1><<<grid, 256, 0, st>>>(n_nodes, brow_ptr, bcol, tperm, data, data_t);
  else if (vec == 2) transpose_values_kernel<2><<<grid, 256, 0, st>>>(n_nodes, brow_ptr, bcol, tperm, data, data_t);
  else if (vec == 3) transpose_values_kernel<3><<<grid, 256, 0, st>>>(n_nodes, brow_ptr, bcol, tperm, data, data_t);
  else {
    set_error("fem_csr_transpose_values: unregistered vec=%d", vec);
    return FEM_EINVAL;
  }
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

extern "C" int fem_dot(int64_t n, const double* x, const double* y, double* result_host, double* workspace,
                       void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(x && y && result_host && workspace, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (int)((n + 255) / 256 < 1184 ? ((n + 255) / 256 > 0 ? (n + 255) / 256 : 1) : 1184);
  dot_kernel<<<nb, 256, 0, st>>>(n, x, y, workspace + 1);
  final_sum_kernel<<<1, 256, 0, st>>>(nb, workspace + 1, workspace);
  FEM_LAUNCH_CHECK();
  FEM_CUDA_CHECK(cudaMemcpyAsync(result_host, workspace, sizeof(double), cudaMemcpyDeviceToHost, st));
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  return FEM_OK;
}

extern "C" int fem_axpy(int64_t n, double alpha, const double* x, double* y, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(x && y, "null pointer");
  if (n == 0) return FEM_OK;
  axpy_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, alpha, x, y);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}
