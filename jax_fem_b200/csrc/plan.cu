// fem_plan_create: the assembly plan (sparsity pattern + cell -> CSR-slot maps) built on the device from raw buffers.
//
// Replaces the (I, J) COO index arrays of Problem.__post_init__ (jax_fem/problem.py:86-107) and PETSc's
// setPreallocationCOO (jax_fem/solver.py:476-478) behind the C ABI: a jax.ffi / ctypes caller hands over the connectivity
// and gets every table the element, gather, SpMV and transpose kernels need -- no Python, no torch.  The reference
// materialises C*ndof^2 integers twice (37-74 GB at 200^3); here only the node-block graph is built (sort / run-length /
// scan / binary search with Thrust, a few small kernels).  jax_fem_b200/plan.py is the same construction in torch (it runs
// on the CPU for the host-logic tests); tests/test_gpu_parity.py checks that both produce identical tables.
#include <thrust/binary_search.h>
#include <thrust/count.h>
#include <thrust/device_ptr.h>
#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/extrema.h>
#include <thrust/gather.h>
#include <thrust/transform.h>
#include <thrust/iterator/constant_iterator.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <new>
#include "common.cuh"

namespace femb200 {
namespace {

constexpr int kSplitSources = 4;     // plan.py::SPLIT_SOURCES
constexpr int kNumTables = 17;

struct Plan {
  int64_t C = 0, num_nodes = 0, nnzb = 0, nnz = 0, n_src = 0, n_rows = 0, n_items = 0;
  int N = 0, vec = 0;
  int32_t* t[kNumTables] = {nullptr};
  int64_t count[kNumTables] = {0};
};

template <class T>
int dev_alloc(T** p, int64_t n) {
  FEM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * (size_t)(n > 0 ? n : 1)));
  return FEM_OK;
}

__global__ void pair_keys_kernel(int64_t C, int N, int64_t num_nodes, const int32_t* __restrict__ cells, int64_t* __restrict__ keys) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= C * N * N) return;
  const int64_t c = k / (N * N);
  const int a = (int)(k / N % N), b = (int)(k % N);
  keys[k] = (int64_t)cells[c * N + a] * num_nodes + cells[c * N + b];
}

__global__ void split_keys_kernel(int64_t n, int64_t num_nodes, const int64_t* __restrict__ ukeys, int32_t* __restrict__ brow,
                                  int32_t* __restrict__ bcol) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  brow[e] = (int32_t)(ukeys[e] / num_nodes);
  bcol[e] = (int32_t)(ukeys[e] % num_nodes);
}

__global__ void scatter_pos_kernel(int64_t n, const int32_t* __restrict__ nc, int32_t* __restrict__ corner_pos) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) corner_pos[nc[i]] = (int32_t)i;
}

__global__ void src_kernel(int64_t n, int N, const int32_t* __restrict__ codes, const int32_t* __restrict__ corner_pos,
                           int32_t* __restrict__ src) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) src[k] = corner_pos[codes[k] / N] * N + codes[k] % N;
}

__global__ void indptr_kernel(int64_t num_nodes, int vec, const int32_t* __restrict__ brow_ptr, int32_t* __restrict__ indptr) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n > num_nodes) return;
  if (n == num_nodes) {
    indptr[num_nodes * vec] = vec * vec * brow_ptr[num_nodes];
    return;
  }
  const int len = brow_ptr[n + 1] - brow_ptr[n];
  for (int i = 0; i < vec; ++i) indptr[n * vec + i] = vec * vec * brow_ptr[n] + i * vec * len;
}

// per block entry: scalar column indices, CSR destination of its (0,0) element
__global__ void entries_kernel(int64_t nnzb, int vec, const int32_t* __restrict__ brow_ptr, const int32_t* __restrict__ erow,
                               const int32_t* __restrict__ bcol, int32_t* __restrict__ indices, int32_t* __restrict__ edst) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnzb) return;
  const int n = erow[e];
  const int p0 = brow_ptr[n], len = brow_ptr[n + 1] - p0, slot = (int)(e - p0);
  const int base = vec * vec * p0 + vec * slot;
  edst[e] = base;
  for (int i = 0; i < vec; ++i)
    for (int k = 0; k < vec; ++k) indices[base + i * vec * len + k] = vec * bcol[e] + k;
}

// gather rows before sorting: one per entry, a second one for an entry with more than kSplitSources sources
__global__ void rows_kernel(int64_t nnzb, const int32_t* __restrict__ src_ptr, const int32_t* __restrict__ extra_rank,
                            const int32_t* __restrict__ cta_of_entry, int32_t* __restrict__ r_ent, int32_t* __restrict__ r_sb,
                            int32_t* __restrict__ r_se, int32_t* __restrict__ r_add, int64_t* __restrict__ r_key) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnzb) return;
  const int s0 = src_ptr[e], s1 = src_ptr[e + 1], cnt = s1 - s0;
  const bool split = cnt > kSplitSources;
  const int half = (cnt + 1) / 2;
  const int se0 = split ? s0 + half : s1;
  r_ent[e] = (int32_t)e;
  r_sb[e] = s0;
  r_se[e] = se0;
  r_add[e] = 0;
  auto key = [&](int len) { return (int64_t)cta_of_entry[e] * 64 + (63 - (len < 63 ? len : 63)); };
  r_key[e] = key(se0 - s0);
  if (split) {
    const int64_t r = nnzb + extra_rank[e];
    r_ent[r] = (int32_t)e;
    r_sb[r] = s0 + half;
    r_se[r] = s1;
    r_add[r] = 1;
    r_key[r] = key(s1 - s0 - half);
  }
}

__global__ void gdesc_kernel(int64_t n_items, const int32_t* __restrict__ node0, const int32_t* __restrict__ nc_ptr,
                             const int32_t* __restrict__ brow_ptr, const int32_t* __restrict__ src_ptr,
                             const int32_t* __restrict__ row0, int32_t* __restrict__ gdesc) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_items) return;
  const int n0 = node0[b], e0 = brow_ptr[n0];
  gdesc[4 * b + 0] = nc_ptr[n0];
  gdesc[4 * b + 1] = e0;
  gdesc[4 * b + 2] = src_ptr[e0];
  gdesc[4 * b + 3] = row0[b];
}

__global__ void emeta_kernel(int64_t n_rows, int vec, const int32_t* __restrict__ m_sb, const int32_t* __restrict__ m_se,
                             const int32_t* __restrict__ m_ent, const int32_t* __restrict__ m_add, const int32_t* __restrict__ edst,
                             const int32_t* __restrict__ erow, const int32_t* __restrict__ bcol, const int32_t* __restrict__ brow_ptr,
                             const uint8_t* __restrict__ bc_flag, int4* __restrict__ emeta) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int e = m_ent[r], n = erow[e];
  int info = vec * (brow_ptr[n + 1] - brow_ptr[n]);
  if (bcol[e] == n) info |= 1 << 16;
  if (bc_flag)
    for (int i = 0; i < vec; ++i)
      if (bc_flag[(int64_t)n * vec + i]) info |= 1 << (17 + i);
  info |= m_add[r] << 20;
  emeta[r] = make_int4(m_sb[r], m_se[r], edst[e], info);
}

struct Degree {
  const int32_t* nc_ptr;
  __host__ __device__ int operator()(int64_t n) const { return nc_ptr[n + 1] - nc_ptr[n]; }
};
struct ItemStart {
  int width;
  __host__ __device__ int32_t operator()(int64_t b) const { return (int32_t)(b * width); }
};
struct Lookup {
  const int32_t* table;
  __host__ __device__ int32_t operator()(int32_t i) const { return table[i]; }
};
struct TransposeMismatch {
  const int64_t* ukeys;
  const int32_t* tperm;
  const int32_t *brow, *bcol;
  int64_t num_nodes, nnzb;
  __host__ __device__ int operator()(int64_t e) const {
    const int64_t want = (int64_t)bcol[e] * num_nodes + brow[e];
    return (tperm[e] >= nnzb || ukeys[tperm[e]] != want) ? 1 : 0;
  }
};
struct DivBy64 {
  __host__ __device__ int64_t operator()(int64_t k) const { return k >> 6; }
};
struct IsSplit {
  const int32_t* src_ptr;
  __host__ __device__ int operator()(int64_t e) const { return (src_ptr[e + 1] - src_ptr[e]) > kSplitSources ? 1 : 0; }
};
struct TransposedKey {
  const int32_t *brow, *bcol;
  int64_t num_nodes;
  __host__ __device__ int64_t operator()(int64_t e) const { return (int64_t)bcol[e] * num_nodes + brow[e]; }
};

inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256 > 0 ? (n + 255) / 256 : 1); }

int build(Plan& P, const int32_t* cells, cudaStream_t st) {
  const int64_t C = P.C, num_nodes = P.num_nodes;
  const int N = P.N, vec = P.vec;
  auto pol = thrust::cuda::par.on(st);
  const int64_t n_src = C * N * N, n_corners = C * N;
  P.n_src = n_src;

  // 1. node-pair keys of every (cell, a, b), sorted: run lengths = sources per block entry
  thrust::device_vector<int64_t> keys(n_src);
  int32_t* codes = nullptr;
  if (int e = dev_alloc(&codes, n_src)) return e;
  pair_keys_kernel<<<nblk(n_src), 256, 0, st>>>(C, N, num_nodes, cells, thrust::raw_pointer_cast(keys.data()));
  FEM_LAUNCH_CHECK();
  thrust::sequence(pol, thrust::device_pointer_cast(codes), thrust::device_pointer_cast(codes) + n_src, 0);
  thrust::stable_sort_by_key(pol, keys.begin(), keys.end(), thrust::device_pointer_cast(codes));
  thrust::device_vector<int64_t> ukeys(n_src);
  thrust::device_vector<int32_t> counts(n_src);
  auto ends = thrust::reduce_by_key(pol, keys.begin(), keys.end(), thrust::constant_iterator<int32_t>(1), ukeys.begin(), counts.begin());
  const int64_t nnzb = ends.first - ukeys.begin();
  keys.clear();
  keys.shrink_to_fit();
  P.nnzb = nnzb;
  P.nnz = nnzb * vec * vec;
  if (P.nnz > 2147483647LL) {
    cudaFree(codes);
    set_error("nnz = %lld exceeds int32 (the reference's PETSc.IntType); shard the mesh", (long long)P.nnz);
    return FEM_EINVAL;
  }
  enum { BROW_PTR, BCOL, INDPTR, INDICES, CORNER_POS, NC_PTR, NC, GDESC, SRC, SRC_PTR, TPERM, M_SB, M_SE, M_ENT, M_ADD, EDST, EROW };
  auto alloc_table = [&](int which, int64_t n) -> int {
    P.count[which] = n;
    return dev_alloc(&P.t[which], n);
  };
  if (int e = alloc_table(SRC_PTR, nnzb + 1)) return e;
  if (int e = alloc_table(BCOL, nnzb)) return e;
  if (int e = alloc_table(EROW, nnzb)) return e;
  if (int e = alloc_table(BROW_PTR, num_nodes + 1)) return e;
  FEM_CUDA_CHECK(cudaMemsetAsync(P.t[SRC_PTR], 0, sizeof(int32_t), st));
  thrust::inclusive_scan(pol, counts.begin(), counts.begin() + nnzb, thrust::device_pointer_cast(P.t[SRC_PTR]) + 1);
  split_keys_kernel<<<nblk(nnzb), 256, 0, st>>>(nnzb, num_nodes, thrust::raw_pointer_cast(ukeys.data()), P.t[EROW], P.t[BCOL]);
  FEM_LAUNCH_CHECK();
  thrust::lower_bound(pol, thrust::device_pointer_cast(P.t[EROW]), thrust::device_pointer_cast(P.t[EROW]) + nnzb,
                      thrust::counting_iterator<int32_t>(0), thrust::counting_iterator<int32_t>((int32_t)num_nodes + 1),
                      thrust::device_pointer_cast(P.t[BROW_PTR]));
  counts.clear();
  counts.shrink_to_fit();

  // 2. corners sorted by node: nc / nc_ptr / corner_pos, then the source block of every (cell, a, b)
  if (int e = alloc_table(NC, n_corners)) return e;
  if (int e = alloc_table(NC_PTR, num_nodes + 1)) return e;
  if (int e = alloc_table(CORNER_POS, n_corners)) return e;
  if (int e = alloc_table(SRC, n_src)) return e;
  {
    thrust::device_vector<int32_t> flat(thrust::device_pointer_cast(cells), thrust::device_pointer_cast(cells) + n_corners);
    thrust::sequence(pol, thrust::device_pointer_cast(P.t[NC]), thrust::device_pointer_cast(P.t[NC]) + n_corners, 0);
    thrust::stable_sort_by_key(pol, flat.begin(), flat.end(), thrust::device_pointer_cast(P.t[NC]));
    thrust::lower_bound(pol, flat.begin(), flat.end(), thrust::counting_iterator<int32_t>(0),
                        thrust::counting_iterator<int32_t>((int32_t)num_nodes + 1), thrust::device_pointer_cast(P.t[NC_PTR]));
  }
  scatter_pos_kernel<<<nblk(n_corners), 256, 0, st>>>(n_corners, P.t[NC], P.t[CORNER_POS]);
  src_kernel<<<nblk(n_src), 256, 0, st>>>(n_src, N, codes, P.t[CORNER_POS], P.t[SRC]);
  FEM_LAUNCH_CHECK();
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  FEM_CUDA_CHECK(cudaFree(codes));

  // 3. scalar CSR pattern, CSR destinations
  if (int e = alloc_table(INDPTR, num_nodes * vec + 1)) return e;
  if (int e = alloc_table(INDICES, P.nnz)) return e;
  if (int e = alloc_table(EDST, nnzb)) return e;
  indptr_kernel<<<nblk(num_nodes + 1), 256, 0, st>>>(num_nodes, vec, P.t[BROW_PTR], P.t[INDPTR]);
  entries_kernel<<<nblk(nnzb), 256, 0, st>>>(nnzb, vec, P.t[BROW_PTR], P.t[EROW], P.t[BCOL], P.t[INDICES], P.t[EDST]);
  FEM_LAUNCH_CHECK();

  // 4. work split of the CSR gather (csrc/sparse.cu::gather_csr_kernel)
  const int width = N <= 8 ? 32 : 8, tail = N <= 8 ? 16 : 8;
  {
    auto it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), Degree{P.t[NC_PTR]});
    const int maxdeg = num_nodes ? thrust::reduce(pol, it, it + num_nodes, 0, thrust::maximum<int>()) : 0;
    if (maxdeg > tail) {
      set_error("a node belongs to %d cells (> %d): mesh valence too high", maxdeg, tail);
      return FEM_EINVAL;
    }
  }
  const int64_t n_items = (n_corners + width - 1) / width;
  P.n_items = n_items;
  thrust::device_vector<int32_t> node0(n_items + 1), ent0(n_items + 1), cta_of_entry(nnzb), extra_rank(nnzb);
  {
    auto starts = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), ItemStart{width});
    thrust::lower_bound(pol, thrust::device_pointer_cast(P.t[NC_PTR]), thrust::device_pointer_cast(P.t[NC_PTR]) + num_nodes,
                        starts, starts + (n_items + 1), node0.begin());
    thrust::transform(pol, node0.begin(), node0.end(), ent0.begin(), Lookup{P.t[BROW_PTR]});
    thrust::upper_bound(pol, ent0.begin() + 1, ent0.end(), thrust::counting_iterator<int32_t>(0),
                        thrust::counting_iterator<int32_t>((int32_t)nnzb), cta_of_entry.begin());
  }
  IsSplit is_split{P.t[SRC_PTR]};
  auto split_it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), is_split);
  thrust::exclusive_scan(pol, split_it, split_it + nnzb, extra_rank.begin());
  const int64_t n_split = nnzb ? thrust::reduce(pol, split_it, split_it + nnzb, 0) : 0;
  const int64_t n_rows = nnzb + n_split;
  P.n_rows = n_rows;
  {
    thrust::device_vector<int32_t> r_ent(n_rows), r_sb(n_rows), r_se(n_rows), r_add(n_rows), order(n_rows);
    thrust::device_vector<int64_t> r_key(n_rows);
    rows_kernel<<<nblk(nnzb), 256, 0, st>>>(nnzb, P.t[SRC_PTR], thrust::raw_pointer_cast(extra_rank.data()),
                                            thrust::raw_pointer_cast(cta_of_entry.data()), thrust::raw_pointer_cast(r_ent.data()),
                                            thrust::raw_pointer_cast(r_sb.data()), thrust::raw_pointer_cast(r_se.data()),
                                            thrust::raw_pointer_cast(r_add.data()), thrust::raw_pointer_cast(r_key.data()));
    FEM_LAUNCH_CHECK();
    thrust::sequence(pol, order.begin(), order.end(), 0);
    thrust::stable_sort_by_key(pol, r_key.begin(), r_key.end(), order.begin());
    if (int e = alloc_table(M_SB, n_rows)) return e;
    if (int e = alloc_table(M_SE, n_rows)) return e;
    if (int e = alloc_table(M_ENT, n_rows)) return e;
    if (int e = alloc_table(M_ADD, n_rows)) return e;
    thrust::gather(pol, order.begin(), order.end(), r_sb.begin(), thrust::device_pointer_cast(P.t[M_SB]));
    thrust::gather(pol, order.begin(), order.end(), r_se.begin(), thrust::device_pointer_cast(P.t[M_SE]));
    thrust::gather(pol, order.begin(), order.end(), r_ent.begin(), thrust::device_pointer_cast(P.t[M_ENT]));
    thrust::gather(pol, order.begin(), order.end(), r_add.begin(), thrust::device_pointer_cast(P.t[M_ADD]));
    // first gather row of every item: the sorted keys are (item, ...), so a binary search on key >> 6
    thrust::device_vector<int32_t> row0(n_items + 1);
    auto cta_sorted = thrust::make_transform_iterator(r_key.begin(), DivBy64());
    thrust::lower_bound(pol, cta_sorted, cta_sorted + n_rows, thrust::counting_iterator<int64_t>(0),
                        thrust::counting_iterator<int64_t>(n_items + 1), row0.begin());
    if (int e = alloc_table(GDESC, 4 * (n_items + 1))) return e;
    gdesc_kernel<<<nblk(n_items + 1), 256, 0, st>>>(n_items, thrust::raw_pointer_cast(node0.data()), P.t[NC_PTR], P.t[BROW_PTR],
                                                    P.t[SRC_PTR], thrust::raw_pointer_cast(row0.data()), P.t[GDESC]);
    FEM_LAUNCH_CHECK();
    FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  }

  // 5. block transpose map (the graph is structurally symmetric)
  if (int e = alloc_table(TPERM, nnzb)) return e;
  {
    TransposedKey tk{P.t[EROW], P.t[BCOL], num_nodes};
    auto tkeys = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), tk);
    thrust::lower_bound(pol, ukeys.begin(), ukeys.begin() + nnzb, tkeys, tkeys + nnzb, thrust::device_pointer_cast(P.t[TPERM]));
    auto bad = thrust::make_transform_iterator(
        thrust::counting_iterator<int64_t>(0),
        TransposeMismatch{thrust::raw_pointer_cast(ukeys.data()), P.t[TPERM], P.t[EROW], P.t[BCOL], num_nodes, nnzb});
    if (nnzb && thrust::reduce(pol, bad, bad + nnzb, 0) != 0) {
      set_error("pattern is not structurally symmetric");
      return FEM_EINVAL;
    }
  }
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  return FEM_OK;
}

void destroy(Plan* P) {
  if (!P) return;
  for (int i = 0; i < kNumTables; ++i)
    if (P->t[i]) cudaFree(P->t[i]);
  delete P;
}

}  // namespace
}  // namespace femb200

using namespace femb200;

extern "C" int fem_plan_create(const int32_t* cells, int64_t n_cells, int64_t n_nodes, int nodes_per_cell, int vec,
                               void* stream, void** plan_out) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(cells && plan_out, "null pointer");
  FEM_REQUIRE(n_cells > 0 && n_nodes > 0 && nodes_per_cell > 0 && vec >= 1 && vec <= 3, "bad sizes");
  FEM_REQUIRE(n_cells * nodes_per_cell * nodes_per_cell <= 2147483647LL, "C*N*N exceeds int32 source codes; shard the mesh across GPUs");
  Plan* P = new (std::nothrow) Plan();
  FEM_REQUIRE(P, "out of host memory");
  P->C = n_cells;
  P->num_nodes = n_nodes;
  P->N = nodes_per_cell;
  P->vec = vec;
  int rc;
  try {
    rc = build(*P, cells, (cudaStream_t)stream);
  } catch (const std::exception& ex) {     // Thrust reports allocation / launch failures by exception
    set_error("fem_plan_create: %s", ex.what());
    cudaGetLastError();
    rc = FEM_ECUDA;
  }
  if (rc != FEM_OK) {
    destroy(P);
    return rc;
  }
  *plan_out = P;
  return FEM_OK;
}

extern "C" int fem_plan_destroy(void* plan) {
  destroy(reinterpret_cast<Plan*>(plan));
  return FEM_OK;
}

extern "C" int fem_plan_sizes(const void* plan, int64_t* sizes_host) {
  FEM_REQUIRE(plan && sizes_host, "null pointer");
  const Plan* P = reinterpret_cast<const Plan*>(plan);
  sizes_host[0] = P->nnzb;
  sizes_host[1] = P->nnz;
  sizes_host[2] = P->n_items;
  sizes_host[3] = P->n_rows;
  sizes_host[4] = P->n_src;
  sizes_host[5] = P->num_nodes * P->vec;
  sizes_host[6] = (P->N * P->vec * P->vec + 1) / 2 * 2;      // doubles per corner row block of the element-tangent buffer
  sizes_host[7] = P->C * P->N;                               // row blocks of that buffer
  return FEM_OK;
}

extern "C" int fem_plan_table(const void* plan, int which, const int32_t** table_out, int64_t* count_out) {
  FEM_REQUIRE(plan && table_out, "null pointer");
  FEM_REQUIRE(which >= 0 && which < kNumTables, "unknown table");
  const Plan* P = reinterpret_cast<const Plan*>(plan);
  *table_out = P->t[which];
  if (count_out) *count_out = P->count[which];
  return FEM_OK;
}

extern "C" int fem_plan_entry_meta(const void* plan, const uint8_t* bc_flag, int32_t* emeta, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(plan && emeta, "null pointer");
  FEM_REQUIRE((reinterpret_cast<uintptr_t>(emeta) & 15) == 0, "emeta must be 16-byte aligned");
  const Plan* P = reinterpret_cast<const Plan*>(plan);
  enum { BROW_PTR, BCOL, INDPTR, INDICES, CORNER_POS, NC_PTR, NC, GDESC, SRC, SRC_PTR, TPERM, M_SB, M_SE, M_ENT, M_ADD, EDST, EROW };
  if (P->n_rows == 0) return FEM_OK;
  emeta_kernel<<<nblk(P->n_rows), 256, 0, (cudaStream_t)stream>>>(P->n_rows, P->vec, P->t[M_SB], P->t[M_SE], P->t[M_ENT], P->t[M_ADD],
                                                                  P->t[EDST], P->t[EROW], P->t[BCOL], P->t[BROW_PTR], bc_flag,
                                                                  reinterpret_cast<int4*>(emeta));
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}
