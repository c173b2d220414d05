// capi.cu -- library-level entry points and the VectorBaseType replacement
// (src/Utilities/Vector/SeqVector.f90).
#include "common.cuh"
#include "../../include/mf6gpu.h"

struct mf6gpu_vector {
  int n = 0;
  mf6::DevBuf<double> v;
  mf6::DevBuf<double> partial;
  mf6::DevBuf<unsigned int> ticket;
  mf6::DevBuf<double> out;
};

namespace mf6 {

__global__ void vec_axpy_kernel(int n, double alpha, const double *__restrict__ x,
                                double *__restrict__ y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    y[i] = y[i] + alpha * x[i];
}

// sum a[i]*b[i], deterministic two-stage reduction
__global__ void __launch_bounds__(kBlock)
vec_dot_kernel(int n, const double *__restrict__ a, const double *__restrict__ b,
               double *__restrict__ partial, unsigned int *ticket, double *out) {
  __shared__ double sh[8];
  __shared__ bool last;
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    s += a[i] * b[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (last_block(ticket, &last)) {
    double r = 0.0;
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) r += partial[i];
    r = block_sum(r, sh);
    if (threadIdx.x == 0) *out = r;
  }
}

}  // namespace mf6

using namespace mf6;

extern "C" {

int mf6gpu_abi_version(void) { return MF6GPU_ABI_VERSION; }

const char *mf6gpu_last_error(void) { return last_error().c_str(); }

size_t mf6gpu_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(mf6gpu_ims_settings);
    case 1: return sizeof(mf6gpu_sln_settings);
    case 2: return sizeof(mf6gpu_gwf_model);
    case 3: return sizeof(mf6gpu_bnd_package);
    case 4: return sizeof(mf6gpu_step_report);
  }
  return 0;
}

// Page-lock a host array the caller keeps for the whole simulation (the Fortran amat / rhs / x arrays of
// NumericalSolution: NumericalSolution.f90:415-430), so that the per-outer-iteration copies of
// mf6gpu_matrix_update / mf6gpu_solver_solve run at full PCIe rate instead of through a pageable staging copy.
int mf6gpu_host_register(void *ptr, size_t bytes) {
  return guard([&] {
    MF6_REQUIRE(ptr && bytes > 0, "host_register: null argument");
    MF6_CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  });
}

int mf6gpu_host_unregister(void *ptr) {
  return guard([&] {
    MF6_REQUIRE(ptr, "host_unregister: null argument");
    MF6_CK(cudaHostUnregister(ptr));
  });
}

int mf6gpu_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return c;
}

int mf6gpu_init(int device) {
  return guard([&] {
    int c = 0;
    MF6_CK(cudaGetDeviceCount(&c));
    MF6_REQUIRE(c > 0, "init: no CUDA device is visible; libmf6gpu has no CPU fallback");
    if (device >= 0) {
      MF6_REQUIRE(device < c, "init: device index out of range");
      MF6_CK(cudaSetDevice(device));
    }
    MF6_CK(cudaFree(0));
    cudaDeviceProp p;
    int dev;
    MF6_CK(cudaGetDevice(&dev));
    MF6_CK(cudaGetDeviceProperties(&p, dev));
    MF6_REQUIRE(p.major >= 10, "init: libmf6gpu is built for sm_100a (Blackwell) only");
  });
}

int mf6gpu_vector_create(int32_t n, mf6gpu_vector **out) {
  return guard([&] {
    MF6_REQUIRE(n > 0 && out, "vector_create: bad argument");
    auto *v = new mf6gpu_vector();
    try {
      v->n = n;
      v->v.alloc_zero((size_t)n);
      v->partial.alloc_zero((size_t)kMaxBlocks);
      v->ticket.alloc_zero(1);
      v->out.alloc_zero(1);
    } catch (...) {
      delete v;
      throw;
    }
    *out = v;
  });
}

int mf6gpu_vector_destroy(mf6gpu_vector *v) {
  return guard([&] { delete v; });
}

int mf6gpu_vector_set(mf6gpu_vector *v, const double *host) {
  return guard([&] {
    MF6_REQUIRE(v && host, "vector_set: null argument");
    MF6_CK(cudaMemcpy(v->v.p, host, sizeof(double) * (size_t)v->n, cudaMemcpyHostToDevice));
  });
}

int mf6gpu_vector_get(const mf6gpu_vector *v, double *host) {
  return guard([&] {
    MF6_REQUIRE(v && host, "vector_get: null argument");
    MF6_CK(cudaMemcpy(host, v->v.p, sizeof(double) * (size_t)v->n, cudaMemcpyDeviceToHost));
  });
}

int mf6gpu_vector_zero_entries(mf6gpu_vector *v) {
  return guard([&] {
    MF6_REQUIRE(v, "vector_zero_entries: null argument");
    MF6_CK(cudaMemset(v->v.p, 0, sizeof(double) * (size_t)v->n));
  });
}

int mf6gpu_vector_axpy(mf6gpu_vector *y, double alpha, const mf6gpu_vector *x) {
  return guard([&] {
    MF6_REQUIRE(y && x && y->n == x->n, "vector_axpy: size mismatch");
    vec_axpy_kernel<<<grid_for(y->n), kBlock>>>(y->n, alpha, x->v.p, y->v.p);
    MF6_CK(cudaGetLastError());
    MF6_CK(cudaDeviceSynchronize());
  });
}

int mf6gpu_vector_dot(const mf6gpu_vector *a, const mf6gpu_vector *b, double *result) {
  return guard([&] {
    MF6_REQUIRE(a && b && result && a->n == b->n, "vector_dot: size mismatch");
    vec_dot_kernel<<<grid_for(a->n), kBlock>>>(a->n, a->v.p, b->v.p, a->partial.p, a->ticket.p,
                                               a->out.p);
    MF6_CK(cudaGetLastError());
    MF6_CK(cudaMemcpy(result, a->out.p, sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int mf6gpu_vector_norm2(const mf6gpu_vector *v, double *result) {
  return guard([&] {
    double s = 0.0;
    MF6_REQUIRE(v && result, "vector_norm2: null argument");
    vec_dot_kernel<<<grid_for(v->n), kBlock>>>(v->n, v->v.p, v->v.p, v->partial.p, v->ticket.p,
                                               v->out.p);
    MF6_CK(cudaGetLastError());
    MF6_CK(cudaMemcpy(&s, v->out.p, sizeof(double), cudaMemcpyDeviceToHost));
    *result = sqrt(s);
  });
}

}  // extern "C"
