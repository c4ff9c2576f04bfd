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
  }
  delete h;
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
