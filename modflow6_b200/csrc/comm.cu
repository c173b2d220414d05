// comm.cu -- NCCL plumbing (dlopen) for the split-model path; see comm.cuh
#include "comm.cuh"
#include <dlfcn.h>
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

void HaloPlan::exchange(double *vec, cudaStream_t s) {
  if (!active()) return;
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

void comm_allgather(mf6gpu_comm *c, const double *in, double *out, size_t count, cudaStream_t s) {
  if (!c || c->nranks == 1) {
    MF6_CK(cudaMemcpyAsync(out, in, count * sizeof(double), cudaMemcpyDeviceToDevice, s));
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

int mf6gpu_comm_destroy(mf6gpu_comm *c) {
  return guard([&] {
    if (!c) return;
    if (c->nccl) nccl().CommDestroy(c->nccl);
    delete c;
  });
}

int mf6gpu_comm_rank(const mf6gpu_comm *c) { return c ? c->rank : 0; }
int mf6gpu_comm_size(const mf6gpu_comm *c) { return c ? c->nranks : 1; }

}  // extern "C"
