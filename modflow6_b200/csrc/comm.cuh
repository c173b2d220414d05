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
  size_t small_slot() const { return 2 * small_doubles * sizeof(double); }   // two tagged words per double (LL)
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

// Peer-memory small all-gather split in two: the push is issued here, the consumer kernel waits for
// the flags itself and reads the records in place (no pull kernel, no staging copy).
struct SmallGather {
  const char *base;          // first record of this parity in the own mailbox (nullptr: not the p2p path)
  size_t slot;               // byte stride between the records of consecutive ranks
  int cap;                   // doubles before the flag word
  unsigned long long seq;
  int *err;
  int nranks;
};
SmallGather comm_small_push(mf6gpu_comm *c, const double *in, size_t count, cudaStream_t s);

// One reduction round of the fused peer-memory path: no launch of its own.  The LAST CTA of the kernel that
// produces this rank's partial result stores the 8-double record into every peer's mailbox (DistPush) and
// every CTA of the kernel that consumes the result waits for the nranks records in its own mailbox and
// combines them in rank order (SmallGather) -- bit-identical scalars on every rank, no finalize launch.
struct DistPush {
  char *const *peer;         // device table of the mapped mailboxes; nullptr = not the fused path
  int nranks;
  size_t off;                // byte offset of this rank's record of this round inside every mailbox
  int cap;
  unsigned long long seq;
};
struct DistRound {
  DistPush push;
  SmallGather pull;
};
DistRound comm_round(mf6gpu_comm *c);   // allocates the next sequence number (host side only)

// Halo exchange of the fused path.  Producer side (HaloPush): the last CTA of the kernel that finalises the
// vector gathers the send cells into the neighbours' mailboxes and publishes the flags.  Consumer side
// (HaloSrc): the SpMV reads halo columns straight out of the own mailbox; rows of slices that touch a halo
// column are deferred to the end of the kernel and wait for the flags there, so the exchange overlaps the
// interior rows (replaces MatMult's VecScatterBegin/End overlap, PetscMatrix.F90:108-113).
constexpr int kMaxHaloNbr = 8;
struct HaloPush {
  char *const *peer;         // nullptr = nothing to push
  int nnbr;
  const int *nbr_rank, *send_ptr, *send_idx;
  size_t off;                // byte offset of this rank's message slot of this round in the peers' mailboxes
  int cap;
  unsigned long long seq;
  unsigned int *ticket;
  // send entries grouped by the CTA that computes their source row in a `grid`-CTA grid-stride kernel of kBlock
  // threads (row r belongs to CTA (r / kBlock) % grid): every CTA pushes its own entries, the last one to
  // finish publishes the flags.  grid == 0: no grouping, the last CTA pushes everything.
  int grid;
  const int *cta_ptr, *cta_ent;
};
struct HaloSrc {
  int nnbr;                  // 0 = halo columns live behind the owned entries of the vector (legacy layout)
  int n_own;
  int recv_ptr[kMaxHaloNbr + 1];
  const double *msg[kMaxHaloNbr];              // neighbour k's message of this round in the own mailbox
  int cap;
  unsigned long long seq;
  int *err;
  const unsigned char *slice_halo;             // [nslices] 1 = the slice has a row with a halo column
  // rows with halo columns grouped by the CTA that owns them in a `grid`-CTA grid-stride kernel (see HaloPush)
  int grid;
  const int *def_ptr, *def_row;
};

// halo pattern of one solution on one rank
struct HaloPlan {
  mf6gpu_comm *comm = nullptr;
  int n_own = 0, n_halo = 0;
  std::vector<int> nbr_rank, send_ptr, recv_ptr;  // per neighbour; *_ptr have size nnbr+1
  DevBuf<int> send_idx;                           // [send_ptr.back()] owned rows (final numbering) to pack
  DevBuf<int> d_nbr_rank, d_send_ptr, d_recv_ptr; // device copies for the peer-memory kernels
  DevBuf<unsigned int> ticket;
  DevBuf<double> sendbuf;
  DevBuf<unsigned char> slice_halo;               // [nslices] slices with halo columns (fused SpMV)
  int own_grid = 0;                               // grid the ownership lists below were built for
  DevBuf<int> cta_ptr, cta_ent;                   // send entries grouped by owning CTA (see HaloPush)
  DevBuf<int> def_ptr, def_row;                   // rows with halo columns grouped by owning CTA (see HaloSrc)
  bool active() const { return comm != nullptr && comm->nranks > 1; }
  // fused peer-memory path available (mailboxes mapped, few enough neighbours)?
  bool fused() const {
    return active() && comm->p2p && !nbr_rank.empty() && (int)nbr_rank.size() <= kMaxHaloNbr && slice_halo.n > 0;
  }
  // vec[n_own + recv range of neighbour k] <- neighbour k's owned values
  void exchange(double *vec, cudaStream_t s);
  // fused path: one round = (producer argument, consumer argument); `push_now` launches the stand-alone
  // push kernel for producers that cannot carry the push themselves (ILU sweeps, residual)
  void round(HaloPush &push, HaloSrc &src);
  void push_now(const HaloPush &push, const double *vec, cudaStream_t s);
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin until *flag >= seq; gives up (and raises *err) after ~10 s so that a dead or out-of-step peer cannot
// hang the GPU (legitimate waits are microseconds, host-side launch skew between ranks milliseconds)
__device__ __forceinline__ bool wait_seq(const unsigned long long *flag, unsigned long long seq, int *err) {
  for (unsigned int spin = 0; spin < (1u << 23); spin++) {
    if (ld_acquire_sys_u64(flag) >= seq) return true;
    __nanosleep(spin < 4096 ? 20 : 1000);
  }
  *err = 1;
  return false;
}
// Small records travel in a low-latency format: every double is split into two 8-byte words
// {32 data bits | 32-bit round tag}.  An 8-byte store is atomic, so a word whose tag matches the round carries valid
// data by itself -- no fence and no separate flag between data and "ready", i.e. one NVLink trip instead of a
// store + system fence (round trip) + flag store.  (The protocol NCCL calls LL.)
__device__ __forceinline__ void ll_store(unsigned long long *dst, const double *src, int count, unsigned int tag) {
  for (int i = 0; i < count; i++) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(src[i]);
    const unsigned long long lo = (b & 0xffffffffull) | ((unsigned long long)tag << 32);
    const unsigned long long hi = (b >> 32) | ((unsigned long long)tag << 32);
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst + 2 * i), "l"(lo) : "memory");
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst + 2 * i + 1), "l"(hi) : "memory");
  }
}
// waits until the first `count` doubles of the record carry this round's tag, returns them in out[]
__device__ __forceinline__ bool ll_load_wait(const unsigned long long *src, double *out, int count, unsigned int tag,
                                             int *err) {
  for (int i = 0; i < count; i++) {
    unsigned long long lo = 0, hi = 0;
    bool ok = false;
    for (unsigned int spin = 0; spin < (1u << 23); spin++) {
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(lo) : "l"(src + 2 * i) : "memory");
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(hi) : "l"(src + 2 * i + 1) : "memory");
      if ((unsigned int)(lo >> 32) == tag && (unsigned int)(hi >> 32) == tag) {
        ok = true;
        break;
      }
      __nanosleep(spin < 4096 ? 20 : 1000);
    }
    if (!ok) {
      *err = 1;
      return false;
    }
    out[i] = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
  }
  return true;
}
// device side of SmallGather: block until rank r's record of this round has landed, copy `count` doubles of it
__device__ __forceinline__ void small_gather_read(const SmallGather &g, int r, double *out, int count) {
  ll_load_wait(reinterpret_cast<const unsigned long long *>(g.base + (size_t)r * g.slot), out, count,
               (unsigned int)g.seq, g.err);
}
// producer side of a fused round: lanes 0..nranks-1 of the first warp store the record (8 doubles, written by
// this CTA before the call and made visible with __syncthreads) into the peers' mailboxes
__device__ __forceinline__ void dist_push_record(const DistPush &P, const double *rec8) {
  const int q = threadIdx.x;
  if (q < P.nranks)
    ll_store(reinterpret_cast<unsigned long long *>(P.peer[q] + P.off), rec8, 8, (unsigned int)P.seq);
}
// value of halo column c (>= n_own) from the mailbox
__device__ __forceinline__ double halo_load(const HaloSrc &H, int c) {
  const int j = c - H.n_own;
  int k = 0;
#pragma unroll
  for (int u = 1; u < kMaxHaloNbr; u++)
    if (u < H.nnbr && j >= H.recv_ptr[u]) k = u;
  return __ldcg(H.msg[k] + (j - H.recv_ptr[k]));
}
// have all neighbours' messages of this round landed already?  (one thread, no waiting)
__device__ __forceinline__ bool halo_arrived(const HaloSrc &H) {
  bool ok = true;
  for (int k = 0; k < H.nnbr; k++)
    ok = ok && ld_acquire_sys_u64(reinterpret_cast<const unsigned long long *>(H.msg[k] + H.cap)) >= H.seq;
  return ok;
}
// all threads of the CTA: wait until every neighbour's message of this round has landed
__device__ __forceinline__ void halo_wait(const HaloSrc &H) {
  if (threadIdx.x < H.nnbr)
    wait_seq(reinterpret_cast<const unsigned long long *>(H.msg[threadIdx.x] + H.cap), H.seq, H.err);
  __syncthreads();
}
// "last CTA done" for kernels in which EVERY thread has made stores that the last CTA (or a peer GPU) must
// see: fence by every thread, barrier, then the ticket.  Returns true in all threads of the last CTA.
__device__ __forceinline__ bool last_block_all(unsigned int *counter, bool *sh_flag) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(counter, gridDim.x - 1);
    *sh_flag = (t == gridDim.x - 1);
  }
  __syncthreads();
  const bool f = *sh_flag;
  if (f) __threadfence_system();
  return f;
}
__device__ __forceinline__ void halo_push_entry(const HaloPush &P, const double *vec, int i) {
  int k = 0;
  while (i >= P.send_ptr[k + 1]) k++;
  double *dst = reinterpret_cast<double *>(P.peer[P.nbr_rank[k]] + P.off);
  dst[i - P.send_ptr[k]] = __ldcg(vec + P.send_idx[i]);
}
__device__ __forceinline__ void halo_publish(const HaloPush &P) {
  if (threadIdx.x < P.nnbr) {
    double *dst = reinterpret_cast<double *>(P.peer[P.nbr_rank[threadIdx.x]] + P.off);
    st_release_sys_u64(reinterpret_cast<unsigned long long *>(dst + P.cap), P.seq);
  }
}
// Tail of a kernel that has just written `vec` (every thread of every CTA calls it, after its last store to vec):
// each CTA pushes the send cells whose rows it computed itself, the last CTA to finish publishes the flags.
// Only threads that stored into a peer's memory pay for a system-scope fence.
__device__ __forceinline__ void halo_push_tail(const HaloPush &P, const double *vec, bool *sh_flag) {
  const bool owned = (P.grid == (int)gridDim.x);
  if (owned) {
    const int e0 = P.cta_ptr[blockIdx.x], e1 = P.cta_ptr[blockIdx.x + 1];
    if (e1 > e0) {
      __syncthreads();  // the CTA's own stores to vec are visible to all its threads
      bool pushed = false;
      for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        halo_push_entry(P, vec, P.cta_ent[e]);
        pushed = true;
      }
      if (pushed) __threadfence_system();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicInc(P.ticket, gridDim.x - 1);
    *sh_flag = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (*sh_flag) {
    __threadfence_system();
    if (!owned) {
      const int total = P.send_ptr[P.nnbr];
      for (int i = threadIdx.x; i < total; i += blockDim.x) halo_push_entry(P, vec, i);
      __threadfence_system();
      __syncthreads();
    }
    halo_publish(P);
  }
}
// stand-alone producer: one CTA pushes everything (vec complete before the launch)
__device__ __forceinline__ void halo_push_all(const HaloPush &P, const double *vec) {
  const int total = P.send_ptr[P.nnbr];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    halo_push_entry(P, vec, i);
}
#endif

// throws if a peer-memory wait timed out (a rank died or fell out of step); call after a stream sync
void comm_check(mf6gpu_comm *c);

// out[rank*count .. ] <- every rank's in[0..count)  (count doubles; count <= 16 uses the peer mailboxes)
void comm_allgather(mf6gpu_comm *c, const double *in, double *out, size_t count, cudaStream_t s);

}  // namespace mf6
