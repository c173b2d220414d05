// common.cuh -- error handling, device buffers and reduction helpers shared by
// every translation unit of libmf6gpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <stdexcept>
#include <string>
#include <vector>

namespace mf6 {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

std::string &last_error();

#define MF6_CK(call)                                                          \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      char b__[512];                                                          \
      snprintf(b__, sizeof(b__), "CUDA error %s at %s:%d: %s",                \
               cudaGetErrorName(e__), __FILE__, __LINE__,                     \
               cudaGetErrorString(e__));                                      \
      throw mf6::Error(b__);                                                  \
    }                                                                         \
  } while (0)

#define MF6_REQUIRE(cond, msg)                                                \
  do {                                                                        \
    if (!(cond)) throw mf6::Error(std::string("mf6gpu: ") + (msg));           \
  } while (0)

template <class F>
int guard(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    last_error() = e.what();
    return -1;
  } catch (...) {
    last_error() = "mf6gpu: unknown exception";
    return -1;
  }
}

// Owning device buffer.
template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) MF6_CK(cudaMalloc((void **)&p, count * sizeof(T)));
  }
  void alloc_zero(size_t count) {
    alloc(count);
    if (count) MF6_CK(cudaMemset(p, 0, count * sizeof(T)));
  }
  void upload(const T *h, size_t count, cudaStream_t s = 0) {
    if (count > n || !p) alloc(count);
    if (count) MF6_CK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T> &h, cudaStream_t s = 0) {
    upload(h.data(), h.size(), s);
    MF6_CK(cudaStreamSynchronize(s));
  }
  void download(T *h, size_t count, cudaStream_t s = 0) const {
    if (count) MF6_CK(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    MF6_CK(cudaStreamSynchronize(s));
  }
  void zero(cudaStream_t s = 0) {
    if (n) MF6_CK(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

// Pinned host scalar block for flag polling.
template <class T>
struct PinnedBuf {
  T *p = nullptr;
  size_t n = 0;
  ~PinnedBuf() {
    if (p) cudaFreeHost(p);
  }
  void alloc(size_t count) {
    if (p) cudaFreeHost(p);
    n = count;
    MF6_CK(cudaHostAlloc((void **)&p, count * sizeof(T), cudaHostAllocDefault));
  }
};

// layout of MaxLoc (below) for structs that the host must be able to size
struct MaxLocPOD {
  double a, v;
  int ord, idx;
};

constexpr int kBlock = 256;
// grid cap of the grid-stride kernels: 148 SMs x 8 resident 256-thread CTAs by default
// (MF6GPU_GRID_CAP overrides it for tuning experiments)
int max_blocks();
#define kMaxBlocks (mf6::max_blocks())

inline int grid_for(long long n) {
  long long b = (n + kBlock - 1) / kBlock;
  if (b < 1) b = 1;
  if (b > kMaxBlocks) b = kMaxBlocks;
  return (int)b;
}

#ifdef __CUDACC__
// signed value of largest magnitude + its location; ties resolved towards the
// smaller `ord` (the reference's sequential loop keeps the FIRST maximum,
// ImsLinearBase.f90:162-165) and zeros are never selected.
struct MaxLoc {
  double a;  // |v|
  double v;
  int ord;   // position in the reference's loop order
  int idx;   // row in device numbering
};

__device__ __forceinline__ MaxLoc maxloc_init() { return MaxLoc{0.0, 0.0, INT_MAX, -1}; }

__device__ __forceinline__ void maxloc_take(MaxLoc &b, double v, int ord, int idx) {
  double a = fabs(v);
  if (a > b.a || (a == b.a && a > 0.0 && ord < b.ord)) {
    b.a = a;
    b.v = v;
    b.ord = ord;
    b.idx = idx;
  }
}

__device__ __forceinline__ void maxloc_merge(MaxLoc &b, const MaxLoc &y) {
  if (y.a > b.a || (y.a == b.a && y.a > 0.0 && y.ord < b.ord)) b = y;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ MaxLoc warp_maxloc(MaxLoc m) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxLoc y;
    y.a = __shfl_down_sync(0xffffffffu, m.a, o);
    y.v = __shfl_down_sync(0xffffffffu, m.v, o);
    y.ord = __shfl_down_sync(0xffffffffu, m.ord, o);
    y.idx = __shfl_down_sync(0xffffffffu, m.idx, o);
    maxloc_merge(m, y);
  }
  return m;
}

// block-wide sum; result valid in thread 0.  blockDim.x == kBlock.
__device__ __forceinline__ double block_sum(double v, double *sh /*[8]*/) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

__device__ __forceinline__ double block_max(double v, double *sh) {
  v = warp_max(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
    r = warp_max(r);
  }
  __syncthreads();
  return r;
}

__device__ __forceinline__ MaxLoc block_maxloc(MaxLoc m, MaxLoc *sh /*[8]*/) {
  m = warp_maxloc(m);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = m;
  __syncthreads();
  MaxLoc r = maxloc_init();
  if (w == 0) {
    if (l < (blockDim.x >> 5)) r = sh[l];
    r = warp_maxloc(r);
  }
  __syncthreads();
  return r;
}

// "last block done" ticket: returns true in every thread of the block that
// finishes last.  `counter` wraps back to 0 so it can be reused by the next
// launch without a memset.
__device__ __forceinline__ bool last_block(unsigned int *counter, bool *sh_flag) {
  __threadfence();
  if (threadIdx.x == 0) {
    unsigned int t = atomicInc(counter, gridDim.x - 1);
    *sh_flag = (t == gridDim.x - 1);
  }
  __syncthreads();
  bool f = *sh_flag;
  if (f) __threadfence();
  return f;
}
#endif

}  // namespace mf6
