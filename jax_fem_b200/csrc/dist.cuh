// Multi-GPU plumbing of libfem_b200: NCCL (resolved at run time from the libnccl.so.2 the process already has, normally
// the copy PyTorch loaded) and the halo plan of one rank.
#pragma once
#include <nccl.h>
#include "common.cuh"

namespace femb200 {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
};
const NcclApi* nccl_api();   // nullptr (and fem_last_error set) when libnccl cannot be loaded

#define FEM_NCCL_CHECK(expr)                                                                              \
  do {                                                                                                    \
    ncclResult_t _r = (expr);                                                                             \
    if (_r != ncclSuccess) {                                                                              \
      femb200::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, nccl_api()->GetErrorString(_r)); \
      return FEM_ECUDA;                                                                                   \
    }                                                                                                     \
  } while (0)

// One rank's halo plan.  Local vectors hold the owned dofs first, then the ghosts grouped by owner rank, so that every
// neighbour's ghost block is one contiguous slice that ncclRecv writes into directly; what a neighbour needs from this
// rank is gathered into `sendbuf` by one pack kernel.
struct HaloPlan {
  ncclComm_t comm;
  int vec;
  int n_nb;
  int peer[16];
  int64_t send_ptr[17];     // neighbour k sends nodes send_idx[send_ptr[k] .. send_ptr[k+1])
  int64_t recv_start[16];   // first local node of the ghost block received from neighbour k
  int64_t recv_count[16];
  const int32_t* send_idx;  // device: local ids of the owned nodes to pack, all neighbours concatenated
  double* sendbuf;          // device: vec * send_ptr[n_nb] doubles
  // overlap of the exchange with the rows that need no ghost value (fem_halo_set_interior): owned nodes [int_lo, int_hi) have
  // no ghost neighbour; the sends / receives run on comm_stream between the two events
  int64_t int_lo, int_hi;
  cudaStream_t comm_stream;
  cudaEvent_t ev_packed, ev_arrived;
  // peer-memory exchange (fem_halo_p2p_*): the pack kernel stores the interface values straight into the neighbours' mailboxes
  // over NVLink and raises a flag there; the receiver's unpack kernel waits for its flags and copies the ghosts into x.
  // mailbox (cudaMalloc, exported with cudaIpcGetMemHandle): 16 flag words, an error word, a ticket, then two receive buffers
  // of n_ghost * vec doubles (alternating with the exchange counter).
  unsigned long long* mailbox;
  int64_t n_ghost, n_owned_nodes;        // ghost / owned nodes of this rank
  double* peer_buf[16];                  // mapped: start of my block in neighbour k's receive buffer 0
  int64_t peer_stride[16];               // doubles between neighbour k's two receive buffers
  unsigned long long* peer_flag[16];     // mapped: my flag word in neighbour k's mailbox
  void* peer_base[16];                   // what cudaIpcOpenMemHandle returned (closed in fem_halo_destroy)
  int p2p_ready;
  mutable unsigned long long p2p_seq;    // exchanges so far; identical on every rank
};
constexpr int kMailboxHeaderWords = 32;  // 16 flags, error, ticket, padding: 256 bytes

int halo_exchange(const HaloPlan* h, double* x, cudaStream_t st);
// split form: halo_begin packs on `st` and issues the sends / receives on the plan's own stream; work queued on `st` after
// it runs concurrently with the transfer and must not touch the ghost entries of x; halo_end makes `st` wait for the arrival.
int halo_begin(const HaloPlan* h, double* x, cudaStream_t st);
int halo_end(const HaloPlan* h, cudaStream_t st);
// peer-memory form of the same pair (valid when h->p2p_ready): no NCCL call, no second stream
int halo_p2p_send(const HaloPlan* h, const double* x, cudaStream_t st);
int halo_p2p_receive(const HaloPlan* h, double* x, cudaStream_t st);
int halo_p2p_status(const HaloPlan* h, cudaStream_t st);      // FEM_ECUDA if a receive ever gave up waiting
int allreduce_sum(const HaloPlan* h, double* buf, int count, cudaStream_t st);

}  // namespace femb200
