// Multi-GPU pieces of the hot path: NCCL plumbing, halo exchange of the interface dofs, all-reduced scalars.
//
// The reference has no multi-GPU path in jax_fem/; its MPI demo (applications/parallel/poisson_mpi.py) ships off-rank
// matrix contributions through PETSc (:152-164) and all-reduces the whole solution every Newton step (:123-132).  Here a
// rank owns a node range plus one layer of ghost cells, assembles its rows without communication, and the Krylov loop
// exchanges only the interface values of the search direction with its neighbours (ncclSend / ncclRecv over NVLink) and
// a few scalars per iteration (ncclAllReduce) -- SURVEY.md 8(e).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include "dist.cuh"

namespace femb200 {

const NcclApi* nccl_api() {
  static NcclApi api;
  static int state = 0;   // 0 = not tried, 1 = ok, -1 = failed
  if (state == 0) {
    const char* names[] = {getenv("FEM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
      if (n && !h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      set_error("libnccl.so.2 not found (%s): set FEM_NCCL_LIB or import torch first", dlerror());
      state = -1;
      return nullptr;
    }
    bool ok = true;
    auto sym = [&](const char* n) {
      void* p = dlsym(h, n);
      if (!p) ok = false;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!ok) {
      set_error("libnccl.so.2 lacks a required symbol");
      state = -1;
      return nullptr;
    }
    state = 1;
  }
  return state == 1 ? &api : nullptr;
}

namespace {
template <int VEC>
__global__ void halo_pack_kernel(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ x,
                                 double* __restrict__ buf) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t node = idx[k];
#pragma unroll
  for (int i = 0; i < VEC; ++i) buf[k * VEC + i] = x[node * VEC + i];
}
}  // namespace

namespace {
int halo_pack(const HaloPlan* h, const double* x, cudaStream_t st) {
  const int64_t n_send = h->send_ptr[h->n_nb];
  if (n_send > 0) {
    const unsigned grid = (unsigned)((n_send + 255) / 256);
    if (h->vec == 1) halo_pack_kernel<1><<<grid, 256, 0, st>>>(n_send, h->send_idx, x, h->sendbuf);
    else if (h->vec == 2) halo_pack_kernel<2><<<grid, 256, 0, st>>>(n_send, h->send_idx, x, h->sendbuf);
    else halo_pack_kernel<3><<<grid, 256, 0, st>>>(n_send, h->send_idx, x, h->sendbuf);
    FEM_LAUNCH_CHECK();
  }
  return FEM_OK;
}

int halo_transfer(const HaloPlan* h, double* x, cudaStream_t st) {
  const NcclApi* nc = nccl_api();
  if (!nc) return FEM_ECUDA;
  FEM_NCCL_CHECK(nc->GroupStart());
  for (int k = 0; k < h->n_nb; ++k) {
    const int64_t ns = h->send_ptr[k + 1] - h->send_ptr[k];
    if (ns > 0)
      FEM_NCCL_CHECK(nc->Send(h->sendbuf + h->send_ptr[k] * h->vec, (size_t)ns * h->vec, ncclDouble, h->peer[k], h->comm, st));
    if (h->recv_count[k] > 0)
      FEM_NCCL_CHECK(nc->Recv(x + h->recv_start[k] * h->vec, (size_t)h->recv_count[k] * h->vec, ncclDouble, h->peer[k], h->comm, st));
  }
  FEM_NCCL_CHECK(nc->GroupEnd());
  return FEM_OK;
}
}  // namespace

int halo_exchange(const HaloPlan* h, double* x, cudaStream_t st) {
  if (int e = halo_pack(h, x, st)) return e;
  if (h->n_nb == 0) return FEM_OK;
  return halo_transfer(h, x, st);
}

int halo_begin(const HaloPlan* h, double* x, cudaStream_t st) {
  if (int e = halo_pack(h, x, st)) return e;
  if (h->n_nb == 0) return FEM_OK;
  // everything queued on `st` so far (the producer of x, the previous readers of its ghosts) precedes the transfer
  FEM_CUDA_CHECK(cudaEventRecord(h->ev_packed, st));
  FEM_CUDA_CHECK(cudaStreamWaitEvent(h->comm_stream, h->ev_packed, 0));
  if (int e = halo_transfer(h, x, h->comm_stream)) return e;
  FEM_CUDA_CHECK(cudaEventRecord(h->ev_arrived, h->comm_stream));
  return FEM_OK;
}

int halo_end(const HaloPlan* h, cudaStream_t st) {
  if (h->n_nb == 0) return FEM_OK;
  FEM_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_arrived, 0));
  return FEM_OK;
}

// ---- peer-memory halo exchange -----------------------------------------------------------------------------------------
namespace {
struct P2PSend {
  int64_t n_send;
  const int32_t* idx;
  int n_nb;
  int64_t send_ptr[17];
  double* dst[16];                       // neighbour k's receive buffer (already offset to my block and to the parity)
  unsigned long long* flag[16];
  unsigned long long seq;
  unsigned int* ticket;
};

// Every interface value goes straight to its place in the neighbour's mailbox (NVLink store); the last block to finish raises
// the flags.  Each thread fences its own stores at system scope before its block takes a ticket.
template <int VEC>
__global__ void halo_p2p_send_kernel(const P2PSend a, const double* __restrict__ x) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < a.n_send) {
    int j = 0;
    while (j + 1 < a.n_nb && k >= a.send_ptr[j + 1]) ++j;
    const int64_t node = a.idx[k], o = (k - a.send_ptr[j]) * VEC;
#pragma unroll
    for (int i = 0; i < VEC; ++i) a.dst[j][o + i] = x[node * VEC + i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(a.ticket, 1u);
    if (t == gridDim.x - 1) {
      *a.ticket = 0u;
      __threadfence_system();
      for (int j = 0; j < a.n_nb; ++j)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[j]), "l"(a.seq) : "memory");
    }
  }
}

struct P2PRecv {
  const unsigned long long* flags;       // this rank's mailbox
  unsigned long long* error;
  unsigned long long seq;
  const double* buf;                     // receive buffer of this parity
  int64_t n_owned;                       // nodes
  int64_t recv_start[16], recv_count[16];
};

// blockIdx.y = neighbour: wait for its flag (bounded: a lost peer raises the error word instead of hanging the GPU), then copy
// its ghost block from the mailbox into x.  The mailbox is written by another GPU: read it past L1.
template <int VEC>
__global__ void halo_p2p_receive_kernel(const P2PRecv a, double* __restrict__ x) {
  const int nb = blockIdx.y;
  if (threadIdx.x == 0) {
    unsigned long long v = 0;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.flags + nb) : "memory");
      if (v >= a.seq) break;
      if (++spins > (1ll << 28)) {       // several seconds
        atomicExch(a.error, 1ull);
        break;
      }
    } while (true);
  }
  __syncthreads();
  const int64_t n = a.recv_count[nb] * VEC, src0 = (a.recv_start[nb] - a.n_owned) * VEC, dst0 = a.recv_start[nb] * VEC;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
  if (((reinterpret_cast<uintptr_t>(a.buf + src0) | reinterpret_cast<uintptr_t>(x + dst0)) & 15) == 0) {   // two doubles per load
    const double2* s2 = reinterpret_cast<const double2*>(a.buf + src0);
    double2* d2 = reinterpret_cast<double2*>(x + dst0);
    for (int64_t i = tid; i < n / 2; i += nthreads) d2[i] = __ldcg(s2 + i);
    if (tid == 0 && (n & 1)) x[dst0 + n - 1] = __ldcg(a.buf + src0 + n - 1);
  } else {
    for (int64_t i = tid; i < n; i += nthreads) x[dst0 + i] = __ldcg(a.buf + src0 + i);
  }
}
}  // namespace

int halo_p2p_send(const HaloPlan* h, const double* x, cudaStream_t st) {
  const unsigned long long seq = ++h->p2p_seq;
  if (h->n_nb == 0) return FEM_OK;
  P2PSend a{};
  a.n_send = h->send_ptr[h->n_nb];
  a.idx = h->send_idx;
  a.n_nb = h->n_nb;
  for (int k = 0; k <= h->n_nb; ++k) a.send_ptr[k] = h->send_ptr[k];
  for (int k = 0; k < h->n_nb; ++k) {
    a.dst[k] = h->peer_buf[k] + (seq & 1) * h->peer_stride[k];
    a.flag[k] = h->peer_flag[k];
  }
  a.seq = seq;
  a.ticket = reinterpret_cast<unsigned int*>(h->mailbox + 17);
  const unsigned grid = (unsigned)std::max<int64_t>(1, (a.n_send + 255) / 256);
  if (h->vec == 1) halo_p2p_send_kernel<1><<<grid, 256, 0, st>>>(a, x);
  else if (h->vec == 2) halo_p2p_send_kernel<2><<<grid, 256, 0, st>>>(a, x);
  else halo_p2p_send_kernel<3><<<grid, 256, 0, st>>>(a, x);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

int halo_p2p_receive(const HaloPlan* h, double* x, cudaStream_t st) {
  if (h->n_nb == 0) return FEM_OK;
  const unsigned long long seq = h->p2p_seq;
  P2PRecv a{};
  a.flags = h->mailbox;
  a.error = h->mailbox + 16;
  a.seq = seq;
  a.buf = reinterpret_cast<const double*>(h->mailbox + kMailboxHeaderWords) + (seq & 1) * h->n_ghost * h->vec;
  a.n_owned = h->n_owned_nodes;
  for (int k = 0; k < h->n_nb; ++k) {
    a.recv_start[k] = h->recv_start[k];
    a.recv_count[k] = h->recv_count[k];
  }
  int64_t most = 1;
  for (int k = 0; k < h->n_nb; ++k) most = std::max(most, h->recv_count[k] * h->vec);
  const dim3 grid((unsigned)std::min<int64_t>(148, (most / 2 + 255) / 256 + 1), (unsigned)h->n_nb);   // ~one 16-byte load per thread
  if (h->vec == 1) halo_p2p_receive_kernel<1><<<grid, 256, 0, st>>>(a, x);
  else if (h->vec == 2) halo_p2p_receive_kernel<2><<<grid, 256, 0, st>>>(a, x);
  else halo_p2p_receive_kernel<3><<<grid, 256, 0, st>>>(a, x);
  FEM_LAUNCH_CHECK();
  return FEM_OK;
}

int halo_p2p_status(const HaloPlan* h, cudaStream_t st) {
  if (!h->p2p_ready || h->n_nb == 0) return FEM_OK;
  unsigned long long err = 0;
  FEM_CUDA_CHECK(cudaMemcpyAsync(&err, h->mailbox + 16, sizeof(err), cudaMemcpyDeviceToHost, st));
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  if (err) {
    set_error("peer-memory halo exchange: a neighbour's values never arrived (flag wait gave up)");
    return FEM_ECUDA;
  }
  return FEM_OK;
}

int allreduce_sum(const HaloPlan* h, double* buf, int count, cudaStream_t st) {
  const NcclApi* nc = nccl_api();
  if (!nc) return FEM_ECUDA;
  FEM_NCCL_CHECK(nc->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, h->comm, st));
  return FEM_OK;
}

}  // namespace femb200

using namespace femb200;

extern "C" int fem_nccl_unique_id(void* id128_host) {
  FEM_REQUIRE(id128_host, "null pointer");
  const NcclApi* nc = nccl_api();
  if (!nc) return FEM_ECUDA;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  FEM_NCCL_CHECK(nc->GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128_host)));
  return FEM_OK;
}

extern "C" int fem_nccl_comm_create(int world, int rank, const void* id128_host, void** comm_out) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(id128_host && comm_out && world > 0 && rank >= 0 && rank < world, "bad communicator arguments");
  const NcclApi* nc = nccl_api();
  if (!nc) return FEM_ECUDA;
  ncclUniqueId id;
  memcpy(&id, id128_host, sizeof(id));
  ncclComm_t comm = nullptr;
  FEM_NCCL_CHECK(nc->CommInitRank(&comm, world, id, rank));
  *comm_out = comm;
  return FEM_OK;
}

extern "C" int fem_nccl_comm_destroy(void* comm) {
  const NcclApi* nc = nccl_api();
  if (!nc) return FEM_ECUDA;
  if (comm) FEM_NCCL_CHECK(nc->CommDestroy((ncclComm_t)comm));
  return FEM_OK;
}

extern "C" int fem_halo_create(void* nccl_comm, int vec, int n_neighbours, const int32_t* peer_host,
                               const int64_t* send_ptr_host, const int32_t* send_idx, const int64_t* recv_start_host,
                               const int64_t* recv_count_host, double* sendbuf, void** halo_out) {
  FEM_REQUIRE(nccl_comm && halo_out, "null pointer");
  FEM_REQUIRE(vec >= 1 && vec <= 3, "vec must be 1, 2 or 3");
  FEM_REQUIRE(n_neighbours >= 0 && n_neighbours <= 16, "at most 16 neighbour ranks");
  FEM_REQUIRE(n_neighbours == 0 || (peer_host && send_ptr_host && recv_start_host && recv_count_host), "null pointer");
  HaloPlan* h = new HaloPlan();
  h->comm = (ncclComm_t)nccl_comm;
  h->vec = vec;
  h->n_nb = n_neighbours;
  h->send_ptr[0] = 0;
  for (int k = 0; k < n_neighbours; ++k) {
    h->peer[k] = peer_host[k];
    h->send_ptr[k] = send_ptr_host[k];
    h->send_ptr[k + 1] = send_ptr_host[k + 1];
    h->recv_start[k] = recv_start_host[k];
    h->recv_count[k] = recv_count_host[k];
  }
  h->send_idx = send_idx;
  h->sendbuf = sendbuf;
  if (h->send_ptr[n_neighbours] > 0 && !(send_idx && sendbuf)) {
    delete h;
    set_error("invalid argument: send_idx / sendbuf missing");
    return FEM_EINVAL;
  }
  h->int_lo = h->int_hi = 0;
  h->mailbox = nullptr;
  h->n_ghost = h->n_owned_nodes = 0;
  h->p2p_ready = 0;
  h->p2p_seq = 0;
  for (int k = 0; k < 16; ++k) {
    h->peer_buf[k] = nullptr;
    h->peer_flag[k] = nullptr;
    h->peer_base[k] = nullptr;
    h->peer_stride[k] = 0;
  }
  h->comm_stream = nullptr;
  h->ev_packed = h->ev_arrived = nullptr;
  if (n_neighbours > 0) {
    cudaError_t e = cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_packed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_arrived, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      set_error("fem_halo_create: %s", cudaGetErrorString(e));
      delete h;
      return FEM_ECUDA;
    }
  }
  *halo_out = h;
  return FEM_OK;
}

extern "C" int fem_halo_destroy(void* halo) {
  HaloPlan* h = reinterpret_cast<HaloPlan*>(halo);
  if (h) {
    if (h->ev_packed) cudaEventDestroy(h->ev_packed);
    if (h->ev_arrived) cudaEventDestroy(h->ev_arrived);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    for (int k = 0; k < 16; ++k)
      if (h->peer_base[k]) cudaIpcCloseMemHandle(h->peer_base[k]);
    if (h->mailbox) cudaFree(h->mailbox);
  }
  delete h;
  return FEM_OK;
}

// ---- peer-memory exchange: setup ------------------------------------------------------------------------------------------
// 1. every rank: fem_halo_p2p_alloc -> its mailbox and the 64-byte IPC handle of it;
// 2. the ranks swap {handle, neighbour list, ghost blocks} by any host-side means (the Python side uses all_gather_object);
// 3. every rank: fem_halo_p2p_connect for each neighbour (maps the neighbour's mailbox: cudaIpcOpenMemHandle);
// 4. after a barrier: fem_halo_p2p_enable(halo, 1) on every rank or on none.
extern "C" int fem_halo_p2p_alloc(void* halo, int64_t n_owned_nodes, int64_t n_ghost_nodes, void* ipc_handle64_out) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(halo && ipc_handle64_out && n_ghost_nodes >= 0 && n_owned_nodes >= 0, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  HaloPlan* h = reinterpret_cast<HaloPlan*>(halo);
  FEM_REQUIRE(!h->mailbox, "mailbox already allocated");
  const size_t bytes = sizeof(unsigned long long) * kMailboxHeaderWords + sizeof(double) * 2 * (size_t)n_ghost_nodes * h->vec;
  FEM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->mailbox), bytes));
  FEM_CUDA_CHECK(cudaMemset(h->mailbox, 0, bytes));
  FEM_CUDA_CHECK(cudaDeviceSynchronize());
  h->n_ghost = n_ghost_nodes;
  h->n_owned_nodes = n_owned_nodes;
  FEM_CUDA_CHECK(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(ipc_handle64_out), h->mailbox));
  return FEM_OK;
}

extern "C" int fem_halo_p2p_connect(void* halo, int k, const void* peer_ipc_handle64, int64_t peer_n_ghost_nodes,
                                    int64_t peer_ghost_offset_nodes, int my_slot_in_peer) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(halo && peer_ipc_handle64, "null pointer");
  HaloPlan* h = reinterpret_cast<HaloPlan*>(halo);
  FEM_REQUIRE(k >= 0 && k < h->n_nb && my_slot_in_peer >= 0 && my_slot_in_peer < 16, "bad neighbour index");
  FEM_REQUIRE(peer_ghost_offset_nodes >= 0 && peer_ghost_offset_nodes + (h->send_ptr[k + 1] - h->send_ptr[k]) <= peer_n_ghost_nodes,
              "my block does not fit the neighbour's ghost range");
  cudaIpcMemHandle_t handle;
  memcpy(&handle, peer_ipc_handle64, sizeof(handle));
  void* base = nullptr;
  FEM_CUDA_CHECK(cudaIpcOpenMemHandle(&base, handle, cudaIpcMemLazyEnablePeerAccess));
  h->peer_base[k] = base;
  unsigned long long* words = static_cast<unsigned long long*>(base);
  h->peer_flag[k] = words + my_slot_in_peer;
  h->peer_buf[k] = reinterpret_cast<double*>(words + kMailboxHeaderWords) + peer_ghost_offset_nodes * h->vec;
  h->peer_stride[k] = peer_n_ghost_nodes * h->vec;
  return FEM_OK;
}

extern "C" int fem_halo_p2p_enable(void* halo, int on) {
  FEM_REQUIRE(halo, "null pointer");
  HaloPlan* h = reinterpret_cast<HaloPlan*>(halo);
  if (on) {
    FEM_REQUIRE(h->mailbox, "fem_halo_p2p_alloc first");
    for (int k = 0; k < h->n_nb; ++k) FEM_REQUIRE(h->peer_base[k], "fem_halo_p2p_connect every neighbour first");
  }
  h->p2p_ready = on ? 1 : 0;
  return FEM_OK;
}

extern "C" int fem_halo_set_interior(void* halo, int64_t node_lo, int64_t node_hi) {
  FEM_REQUIRE(halo && node_lo >= 0 && node_hi >= node_lo, "bad interior range");
  HaloPlan* h = reinterpret_cast<HaloPlan*>(halo);
  h->int_lo = node_lo;
  h->int_hi = node_hi;
  return FEM_OK;
}

extern "C" int fem_halo_exchange(void* halo, double* x, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(halo && x, "null pointer");
  return halo_exchange(reinterpret_cast<const HaloPlan*>(halo), x, (cudaStream_t)stream);
}

extern "C" int fem_allreduce_sum(void* halo, double* buf, int count, void* stream) {
  if (int e = check_device()) return e;
  FEM_REQUIRE(halo && buf && count > 0, "bad argument");
  return allreduce_sum(reinterpret_cast<const HaloPlan*>(halo), buf, count, (cudaStream_t)stream);
}
