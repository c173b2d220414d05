// comm.cu -- NCCL plumbing (dlopen) for the split-model path; see comm.cuh
#include "comm.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <cstring>

namespace mf6 {

// minimal NCCL ABI (nccl.h, 2.x): opaque comm, 128-byte unique id, enums as ints
typedef struct { char internal[128]; } nccl_uid;
enum { NCCL_FLOAT64 = 8 };
typedef int (*fn_GetUniqueId)(nccl_uid *);
typedef int (*fn_CommInitRank)(void **, int, nccl_uid, int);
typedef int (*fn_CommDestroy)(void *);
typedef int (*fn_GroupStart)();
typedef int (*fn_GroupEnd)();
typedef int (*fn_Send)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_Recv)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_AllGather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef const char *(*fn_GetErrorString)(int);

struct Nccl {
  void *h = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_GroupStart GroupStart = nullptr;
  fn_GroupEnd GroupEnd = nullptr;
  fn_Send Send = nullptr;
  fn_Recv Recv = nullptr;
  fn_AllGather AllGather = nullptr;
  fn_GetErrorString GetErrorString = nullptr;
};

static Nccl &nccl() {
  static Nccl n;
  if (n.h) return n;
  // RTLD_NOLOAD first: reuse the library the process already mapped (torch's bundled NCCL)
  n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!n.h) n.h = dlopen("libnccl.so.2", RTLD_NOW);
  if (!n.h) n.h = dlopen("libnccl.so", RTLD_NOW);
  MF6_REQUIRE(n.h, "comm: libnccl.so.2 not found (needed only for multi-GPU runs)");
  auto sym = [&](const char *name) {
    void *p = dlsym(n.h, name);
    if (!p) throw Error(std::string("mf6gpu: comm: NCCL symbol missing: ") + name);
    return p;
  };
  n.GetUniqueId = (fn_GetUniqueId)sym("ncclGetUniqueId");
  n.CommInitRank = (fn_CommInitRank)sym("ncclCommInitRank");
  n.CommDestroy = (fn_CommDestroy)sym("ncclCommDestroy");
  n.GroupStart = (fn_GroupStart)sym("ncclGroupStart");
  n.GroupEnd = (fn_GroupEnd)sym("ncclGroupEnd");
  n.Send = (fn_Send)sym("ncclSend");
  n.Recv = (fn_Recv)sym("ncclRecv");
  n.AllGather = (fn_AllGather)sym("ncclAllGather");
  n.GetErrorString = (fn_GetErrorString)sym("ncclGetErrorString");
  return n;
}

#define MF6_NCCL(call)                                                                     \
  do {                                                                                     \
    int r__ = (call);                                                                      \
    if (r__ != 0)                                                                          \
      throw mf6::Error(std::string("mf6gpu: NCCL error: ") + mf6::nccl().GetErrorString(r__)); \
  } while (0)

__global__ void halo_pack_kernel(int n, const int *__restrict__ idx, const double *__restrict__ vec,
                                 double *__restrict__ buf) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    buf[i] = vec[idx[i]];
}

// ---- peer-memory kernels ---------------------------------------------------------------
__device__ __forceinline__ void st_flag(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// spin until *flag >= seq; gives up (and raises *err) after ~2^27 polls so a dead peer cannot hang the GPU
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long seq, int *err) {
  for (unsigned int spin = 0; spin < (1u << 23); spin++) {
    if (ld_flag(flag) >= seq) return true;
    __nanosleep(spin < 4096 ? 20 : 1000);
  }
  *err = 1;
  return false;
}

// small all-gather, push side: lane q writes this rank's record into rank q's mailbox (LL words, see comm.cuh)
__global__ void p2p_small_push_kernel(char *const *__restrict__ peer, int nranks, int rank, size_t off,
                                      const double *__restrict__ in, int count, int cap,
                                      unsigned long long seq) {
  const int q = threadIdx.x;
  if (q >= nranks) return;
  ll_store(reinterpret_cast<unsigned long long *>(peer[q] + off), in, count, (unsigned int)seq);
}

// small all-gather, pull side: wait for every rank's record, copy them out in rank order
__global__ void p2p_small_pull_kernel(const char *__restrict__ mailbox, int nranks, size_t off0, size_t slot,
                                      double *__restrict__ out, int count, int cap, unsigned long long seq,
                                      int *err) {
  const int r = threadIdx.x;
  if (r >= nranks) return;
  ll_load_wait(reinterpret_cast<const unsigned long long *>(mailbox + off0 + (size_t)r * slot),
               out + (size_t)r * count, count, (unsigned int)seq, err);
}

// halo, push side: gather the owned values every neighbour needs straight into THEIR mailbox; the last
// CTA to finish publishes the flags
__global__ void __launch_bounds__(kBlock)
p2p_halo_push_kernel(char *const *__restrict__ peer, int nnbr, const int *__restrict__ nbr_rank,
                     const int *__restrict__ send_ptr, const int *__restrict__ send_idx,
                     const double *__restrict__ vec, size_t off, int cap, unsigned long long seq,
                     unsigned int *ticket) {
  __shared__ bool last;
  const int total = send_ptr[nnbr];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= send_ptr[k + 1]) k++;
    double *dst = reinterpret_cast<double *>(peer[nbr_rank[k]] + off);
    dst[i - send_ptr[k]] = vec[send_idx[i]];
  }
  __threadfence_system();
  if (threadIdx.x == 0) {
    unsigned int t = atomicInc(ticket, gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < nnbr) {
    __threadfence_system();
    double *dst = reinterpret_cast<double *>(peer[nbr_rank[threadIdx.x]] + off);
    st_flag(reinterpret_cast<unsigned long long *>(dst + cap), seq);
  }
}

// halo, pull side: wait for the neighbours' flags, then move the messages into the halo region
__global__ void __launch_bounds__(kBlock)
p2p_halo_pull_kernel(const char *__restrict__ mailbox, int nnbr, const int *__restrict__ nbr_rank,
                     const int *__restrict__ recv_ptr, double *__restrict__ halo, size_t off0, size_t slot,
                     int cap, unsigned long long seq, int *err) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < nnbr) {
    const double *src = reinterpret_cast<const double *>(mailbox + off0 + (size_t)nbr_rank[threadIdx.x] * slot);
    if (!wait_flag(reinterpret_cast<const unsigned long long *>(src + cap), seq, err)) ok = 0;
  }
  __syncthreads();
  if (!ok) return;
  const int total = recv_ptr[nnbr];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= recv_ptr[k + 1]) k++;
    const double *src = reinterpret_cast<const double *>(mailbox + off0 + (size_t)nbr_rank[k] * slot);
    halo[i] = __ldcg(src + (i - recv_ptr[k]));
  }
}

void HaloPlan::exchange(double *vec, cudaStream_t s) {
  if (!active()) return;
  if (comm->p2p) {
    const int nnbr = (int)nbr_rank.size();
    if (nnbr == 0) return;
    const unsigned long long seq = ++comm->halo_seq;
    const int parity = (int)(seq & 1ull);
    const P2PLayout &L = comm->lay;
    const int nsend = send_ptr.back(), nrecv = recv_ptr.back();
    p2p_halo_push_kernel<<<grid_for(std::max(nsend, 1)), kBlock, 0, s>>>(
        comm->d_peer.p, nnbr, d_nbr_rank.p, d_send_ptr.p, send_idx.p, vec, L.halo_off(parity, comm->rank),
        (int)L.halo_doubles, seq, ticket.p);
    p2p_halo_pull_kernel<<<grid_for(std::max(nrecv, 1)), kBlock, 0, s>>>(
        comm->mailbox, nnbr, d_nbr_rank.p, d_recv_ptr.p, vec + n_own, L.halo_off(parity, 0), L.halo_slot(),
        (int)L.halo_doubles, seq, comm->d_err.p);
    return;
  }
  const int nsend = send_ptr.back();
  if (nsend > 0) halo_pack_kernel<<<grid_for(nsend), kBlock, 0, s>>>(nsend, send_idx.p, vec, sendbuf.p);
  Nccl &N = nccl();
  MF6_NCCL(N.GroupStart());
  for (size_t k = 0; k < nbr_rank.size(); k++) {
    const int sc = send_ptr[k + 1] - send_ptr[k], rc = recv_ptr[k + 1] - recv_ptr[k];
    if (sc > 0) MF6_NCCL(N.Send(sendbuf.p + send_ptr[k], (size_t)sc, NCCL_FLOAT64, nbr_rank[k], comm->nccl, s));
    if (rc > 0) MF6_NCCL(N.Recv(vec + n_own + recv_ptr[k], (size_t)rc, NCCL_FLOAT64, nbr_rank[k], comm->nccl, s));
  }
  MF6_NCCL(N.GroupEnd());
}

void comm_check(mf6gpu_comm *c) {
  if (!c || !c->p2p) return;
  int e = 0;
  MF6_CK(cudaMemcpy(&e, c->d_err.p, sizeof(int), cudaMemcpyDeviceToHost));
  MF6_REQUIRE(e == 0, "comm: a peer-memory wait timed out (another rank stopped or fell out of step)");
}

DistRound comm_round(mf6gpu_comm *c) {
  DistRound R{};
  if (!c || !c->p2p) return R;
  const unsigned long long seq = ++c->small_seq;
  const int parity = (int)(seq & 1ull);
  const P2PLayout &L = c->lay;
  R.push = DistPush{c->d_peer.p, c->nranks, L.small_off(parity, c->rank), (int)L.small_doubles, seq};
  R.pull = SmallGather{c->mailbox + L.small_off(parity, 0), L.small_slot(), (int)L.small_doubles, seq,
                       c->d_err.p, c->nranks};
  return R;
}

// stand-alone producer of a fused halo round (producers that cannot carry the push themselves): a few CTAs
// push, the last one to finish publishes the flags
__global__ void __launch_bounds__(kBlock) halo_push_kernel(HaloPush P, const double *__restrict__ vec) {
  __shared__ bool last;
  halo_push_all(P, vec);
  if (last_block_all(P.ticket, &last)) halo_publish(P);
}

void HaloPlan::round(HaloPush &push, HaloSrc &src) {
  const int nnbr = (int)nbr_rank.size();
  const unsigned long long seq = ++comm->halo_seq;
  const int parity = (int)(seq & 1ull);
  const P2PLayout &L = comm->lay;
  push = HaloPush{comm->d_peer.p, nnbr, d_nbr_rank.p, d_send_ptr.p, send_idx.p, L.halo_off(parity, comm->rank),
                  (int)L.halo_doubles, seq, ticket.p, own_grid, cta_ptr.p, cta_ent.p};
  src = HaloSrc{};
  src.nnbr = nnbr;
  src.n_own = n_own;
  for (int k = 0; k <= nnbr; k++) src.recv_ptr[k] = recv_ptr[k];
  for (int k = 0; k < nnbr; k++)
    src.msg[k] = reinterpret_cast<const double *>(comm->mailbox + L.halo_off(parity, nbr_rank[k]));
  src.cap = (int)L.halo_doubles;
  src.seq = seq;
  src.err = comm->d_err.p;
  src.slice_halo = slice_halo.p;
  src.grid = own_grid;
  src.def_ptr = def_ptr.p;
  src.def_row = def_row.p;
}

void HaloPlan::push_now(const HaloPush &push, const double *vec, cudaStream_t s) {
  const int g = std::max(1, std::min(64, (send_ptr.back() + kBlock - 1) / kBlock));
  halo_push_kernel<<<g, kBlock, 0, s>>>(push, vec);
}

SmallGather comm_small_push(mf6gpu_comm *c, const double *in, size_t count, cudaStream_t s) {
  SmallGather g{nullptr, 0, 0, 0, nullptr, 0};
  if (!c || !c->p2p || count > c->lay.small_doubles) return g;
  const unsigned long long seq = ++c->small_seq;
  const int parity = (int)(seq & 1ull);
  const P2PLayout &L = c->lay;
  p2p_small_push_kernel<<<1, 32, 0, s>>>(c->d_peer.p, c->nranks, c->rank, L.small_off(parity, c->rank), in,
                                         (int)count, (int)L.small_doubles, seq);
  g.base = c->mailbox + L.small_off(parity, 0);
  g.slot = L.small_slot();
  g.cap = (int)L.small_doubles;
  g.seq = seq;
  g.err = c->d_err.p;
  g.nranks = c->nranks;
  return g;
}

void comm_allgather(mf6gpu_comm *c, const double *in, double *out, size_t count, cudaStream_t s) {
  if (!c || c->nranks == 1) {
    MF6_CK(cudaMemcpyAsync(out, in, count * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return;
  }
  if (c->p2p && count <= c->lay.small_doubles) {
    const unsigned long long seq = ++c->small_seq;
    const int parity = (int)(seq & 1ull);
    const P2PLayout &L = c->lay;
    p2p_small_push_kernel<<<1, 32, 0, s>>>(c->d_peer.p, c->nranks, c->rank, L.small_off(parity, c->rank), in,
                                           (int)count, (int)L.small_doubles, seq);
    p2p_small_pull_kernel<<<1, 32, 0, s>>>(c->mailbox, c->nranks, L.small_off(parity, 0), L.small_slot(), out,
                                           (int)count, (int)L.small_doubles, seq, c->d_err.p);
    return;
  }
  MF6_NCCL(nccl().AllGather(in, out, count, NCCL_FLOAT64, c->nccl, s));
}

}  // namespace mf6

using namespace mf6;

extern "C" {

int mf6gpu_comm_unique_id(void *out128) {
  return guard([&] {
    MF6_REQUIRE(out128, "comm_unique_id: null argument");
    nccl_uid id;
    MF6_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(out128, &id, sizeof(id));
  });
}

int mf6gpu_comm_create(int32_t nranks, int32_t rank, const void *id128, mf6gpu_comm **out) {
  return guard([&] {
    MF6_REQUIRE(out && nranks >= 1 && rank >= 0 && rank < nranks, "comm_create: bad argument");
    auto *c = new mf6gpu_comm();
    c->nranks = nranks;
    c->rank = rank;
    if (nranks > 1) {
      MF6_REQUIRE(id128, "comm_create: unique id required");
      nccl_uid id;
      std::memcpy(&id, id128, sizeof(id));
      try {
        MF6_NCCL(nccl().CommInitRank(&c->nccl, nranks, id, rank));
      } catch (...) {
        delete c;
        throw;
      }
    }
    *out = c;
  });
}

// peer-memory mailboxes: export this rank's buffer ...
int mf6gpu_comm_p2p_export(mf6gpu_comm *c, int64_t halo_doubles, void *handle64) {
  return guard([&] {
    MF6_REQUIRE(c && handle64 && halo_doubles >= 0, "comm_p2p_export: bad argument");
    MF6_REQUIRE(c->nranks <= 32, "comm_p2p_export: at most 32 ranks");
    c->lay.nranks = c->nranks;
    c->lay.halo_doubles = (size_t)halo_doubles;
    MF6_CK(cudaMalloc((void **)&c->mailbox, c->lay.total()));
    MF6_CK(cudaMemset(c->mailbox, 0, c->lay.total()));
    MF6_CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    MF6_CK(cudaIpcGetMemHandle(&h, c->mailbox));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle64, &h, sizeof(h));
  });
}

// ... and map everybody else's (handles = nranks x 64 bytes in rank order)
int mf6gpu_comm_p2p_import(mf6gpu_comm *c, const void *handles) {
  return guard([&] {
    MF6_REQUIRE(c && handles && c->mailbox, "comm_p2p_import: export first");
    c->peer.assign((size_t)c->nranks, nullptr);
    for (int r = 0; r < c->nranks; r++) {
      if (r == c->rank) {
        c->peer[r] = c->mailbox;
        continue;
      }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, (const char *)handles + 64 * (size_t)r, sizeof(h));
      void *p = nullptr;
      MF6_CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      c->peer[r] = (char *)p;
    }
    c->d_peer.upload(c->peer);
    c->d_err.alloc_zero(1);
    c->p2p = true;
  });
}

int mf6gpu_comm_p2p_enabled(const mf6gpu_comm *c) { return (c && c->p2p) ? 1 : 0; }

int mf6gpu_comm_p2p_disable(mf6gpu_comm *c) {
  if (c) c->p2p = false;
  return 0;
}

int mf6gpu_comm_destroy(mf6gpu_comm *c) {
  return guard([&] {
    if (!c) return;
    for (int r = 0; r < (int)c->peer.size(); r++)
      if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->mailbox) cudaFree(c->mailbox);
    if (c->nccl) nccl().CommDestroy(c->nccl);
    delete c;
  });
}

int mf6gpu_comm_rank(const mf6gpu_comm *c) { return c ? c->rank : 0; }
int mf6gpu_comm_size(const mf6gpu_comm *c) { return c ? c->nranks : 1; }

}  // extern "C"
