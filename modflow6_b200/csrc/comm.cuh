// comm.cuh -- one-process-per-GPU communication for the split-model path:
// NCCL send/recv halo exchange and small all-gathers for the Krylov / outer scalars.
//
// Replaces, for the path in scope, the reference's distributed backend:
//   MpiRouter (src/Distributed/MpiRouter.f90:238-343) halo of x / ibound for the interface cells
//   PETSc MPIAIJ VecScatter inside MatMult (PetscMatrix.F90:108-113, 455)
//   MPI_Allreduce of Krylov dots / convergence scalars (PetscConvergence.F90:126-158,
//   ParallelSolution.f90:54-235)
// NCCL is loaded with dlopen (libnccl.so.2: the copy torch already mapped, else the system
// one) so that the single-GPU library has no NCCL dependency.
#pragma once
#include "common.cuh"
#include "../../include/mf6gpu.h"

// Peer-memory mailboxes (NVLink / NVSwitch): every rank maps every other rank's mailbox with CUDA IPC
// and WRITES its contribution straight into the consumers' memory (st.global over NVLink), followed
// by a system-scope fence and a sequence flag; consumers spin on their own local flags.  Two parities
// alternate so that a fast rank can never overwrite data a slow rank has not consumed yet.
struct P2PLayout {
  int nranks = 0;
  size_t small_doubles = 16;   // capacity of one small all-gather record
  size_t halo_doubles = 0;     // capacity of one halo message
  size_t small_slot() const { return (small_doubles + 2) * sizeof(double); }              // data + flag + pad
  size_t halo_slot() const { return (halo_doubles + 2) * sizeof(double); }
  size_t small_off(int parity, int src) const { return ((size_t)parity * nranks + src) * small_slot(); }
  size_t halo_base() const { return 2 * (size_t)nranks * small_slot(); }
  size_t halo_off(int parity, int src) const { return halo_base() + ((size_t)parity * nranks + src) * halo_slot(); }
  size_t total() const { return halo_base() + 2 * (size_t)nranks * halo_slot(); }
};

struct mf6gpu_comm {
  int nranks = 1, rank = 0;
  void *nccl = nullptr;  // ncclComm_t
  cudaStream_t stream = 0;
  // peer-memory path
  bool p2p = false;
  P2PLayout lay;
  char *mailbox = nullptr;               // own mailbox (cudaMalloc)
  std::vector<char *> peer;              // [nranks] mapped mailboxes (peer[rank] == mailbox)
  mf6::DevBuf<char *> d_peer;            // the same table on the device
  mf6::DevBuf<int> d_err;                // set to 1 by a consumer that timed out
  unsigned long long small_seq = 0, halo_seq = 0;
};

namespace mf6 {

// halo pattern of one solution on one rank
struct HaloPlan {
  mf6gpu_comm *comm = nullptr;
  int n_own = 0, n_halo = 0;
  std::vector<int> nbr_rank, send_ptr, recv_ptr;  // per neighbour; *_ptr have size nnbr+1
  DevBuf<int> send_idx;                           // [send_ptr.back()] owned rows (final numbering) to pack
  DevBuf<int> d_nbr_rank, d_send_ptr, d_recv_ptr; // device copies for the peer-memory kernels
  DevBuf<unsigned int> ticket;
  DevBuf<double> sendbuf;
  bool active() const { return comm != nullptr && comm->nranks > 1; }
  // vec[n_own + recv range of neighbour k] <- neighbour k's owned values
  void exchange(double *vec, cudaStream_t s);
};

// Peer-memory small all-gather split in two: the push is issued here, the consumer kernel waits for
// the flags itself and reads the records in place (no pull kernel, no staging copy).
struct SmallGather {
  const char *base;          // first record of this parity in the own mailbox (nullptr: not the p2p path)
  size_t slot;               // byte stride between the records of consecutive ranks
  int cap;                   // doubles before the flag word
  unsigned long long seq;
  int *err;
};
SmallGather comm_small_push(mf6gpu_comm *c, const double *in, size_t count, cudaStream_t s);

#ifdef __CUDACC__
// device side of SmallGather: block until rank r's record of this round has landed, return it
__device__ __forceinline__ const double *small_gather_wait(const SmallGather &g, int r) {
  const double *src = reinterpret_cast<const double *>(g.base + (size_t)r * g.slot);
  const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(src + g.cap);
  for (unsigned int spin = 0; spin < (1u << 27); spin++) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= g.seq) return src;
    __nanosleep(20);
  }
  *g.err = 1;
  return src;
}
#endif

// throws if a peer-memory wait timed out (a rank died or fell out of step); call after a stream sync
void comm_check(mf6gpu_comm *c);

// out[rank*count .. ] <- every rank's in[0..count)  (count doubles; count <= 16 uses the peer mailboxes)
void comm_allgather(mf6gpu_comm *c, const double *in, double *out, size_t count, cudaStream_t s);

}  // namespace mf6
