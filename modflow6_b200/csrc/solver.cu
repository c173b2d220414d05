// solver.cu -- device-resident IMS linear accelerators: PCG and BiCGSTAB with
// ILU0/MILU0, the IMS stopping rules and the ConvergenceSummary side channel.
//
// Restates on the device:
//   imslinear_ap       src/Solution/LinearMethods/ImsLinear.f90:617-750
//   ims_base_cg        src/Solution/LinearMethods/ImsLinearBase.f90:30-240
//   ims_base_bcgs      :249-549        ims_base_pcu      :761-864
//   ims_base_testcnvg  :1101-1146      ims_base_residual :1291-1312
//   ims_base_epfact    :1316-1333      ddot / dnrm2  blas1_d.f90:295-333, 387-480
//   is_close           src/Utilities/MathUtil.f90:45-86
//
// The inner loop never synchronises with the host: the recurrence scalars
// (rho, alpha, beta, omega), the reductions and the convergence decision live
// in a KState block in HBM; every kernel is a no-op once KState::done is set
// and the host polls that flag every few iterations.
// Reductions are deterministic: per-CTA partials in fixed slots, combined in a
// fixed order by the last CTA to finish (threadfence + ticket).
#include "solver.cuh"
#include "spmv.cuh"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace mf6 {
// The inner iteration replays a CUDA graph at every size: launch-bound small systems gain most (C1 0.090 -> 0.068 s),
// but the 10 launches of an iteration still cost ~2 us of gaps each at 1e7 rows (measured: 1594 -> 1537 ms per C2
// time step).  MF6GPU_GRAPH_MAX_ROWS lowers the limit for experiments.
static int graph_max_rows() {
  static int v = [] {
    const char *e = std::getenv("MF6GPU_GRAPH_MAX_ROWS");
    return e ? std::atoi(e) : INT_MAX;
  }();
  return v;
}

enum { TK_DOT = 0, TK_SPMV = 1, TK_UPD = 2, TK_NRM = 3 };
// finalisation modes of the reduction kernels
enum {
  FIN_CG_RHO = 0,     // rho = sum ; beta = rho / rho0
  FIN_CG_ALPHA = 1,   // den = sum + sign(eps) ; alpha = rho / den
  FIN_BCGS_RHO = 2,   // rho = sum ; beta = (rho/rho0)(alpha0/omega0)
  FIN_BCGS_ALPHA = 3, // alpha = rho / (sum + sign(eps))
  FIN_BCGS_OMEGA = 4, // omega = s0 / (s1 + sign(eps))
  FIN_NRM_MAX = 5,    // scale = max |d|
  FIN_NRM_SSQ = 6,    // l2norm0 = scale * sqrt(sum (d/scale)^2)
  FIN_PLAIN = 7       // out = sum
};

__device__ __forceinline__ double dsign(double a, double b) { return copysign(fabs(a), b); }

__device__ __forceinline__ bool is_close_dev(double a, double b) {
  if (a == b) return true;
  const double m = fmax(fabs(a), fabs(b));
  return fabs(a - b) <= fmax(100.0 * DBL_EPSILON * m, 0.0);
}

// ims_base_testcnvg, ImsLinearBase.f90:1101-1146
__device__ __forceinline__ void testcnvg_dev(int opt, int &icnvg, int iiter, double dvmax,
                                             double rmax, double rmax0, double epfact,
                                             double dvclose, double rclose) {
  if (opt == 0) {
    if (fabs(dvmax) <= dvclose && fabs(rmax) <= rclose) icnvg = 1;
  } else if (opt == 1) {
    if (fabs(dvmax) <= dvclose && fabs(rmax) <= rclose) icnvg = (iiter == 1) ? 1 : -1;
  } else if (opt == 2) {
    if (fabs(dvmax) <= dvclose || rmax <= rclose)
      icnvg = 1;
    else if (rmax <= rmax0 * epfact)
      icnvg = -1;
  } else if (opt == 3) {
    if (fabs(dvmax) <= dvclose)
      icnvg = 1;
    else if (rmax <= rmax0 * rclose)
      icnvg = -1;
  } else if (opt == 4) {
    if (fabs(dvmax) <= dvclose && rmax <= rclose)
      icnvg = 1;
    else if (rmax <= rmax0 * epfact)
      icnvg = -1;
  }
}

// scalar epilogue of a sum reduction, executed by one thread
__device__ void finalize_sum(int mode, double s0, double s1, KState *st, double *out) {
  switch (mode) {
    case FIN_CG_RHO:
      st->rho = s0;
      st->beta = s0 / st->rho0;  // unused on the first iteration
      break;
    case FIN_CG_ALPHA: {
      double den = s0;
      den = den + dsign(DBL_EPSILON, den);
      st->alpha = st->rho / den;
      break;
    }
    case FIN_BCGS_RHO:
      st->rho = s0;
      st->beta = (s0 / st->rho0) * (st->alpha0 / st->omega0);
      break;
    case FIN_BCGS_ALPHA: {
      double den = s0;
      den = den + dsign(DBL_EPSILON, den);
      st->alpha = st->rho / den;
      break;
    }
    case FIN_BCGS_OMEGA: {
      double den = s1;
      den = den + dsign(DBL_EPSILON, den);
      st->omega = s0 / den;
      break;
    }
    case FIN_NRM_MAX:
      *out = s0;
      break;
    case FIN_NRM_SSQ: {
      const double scale = out[0];
      st->l2norm0 = (scale == 0.0) ? 0.0 : scale * sqrt(s0);
      break;
    }
    default:
      *out = s0;
  }
}

// combine per-CTA partial sums (1 or 2 interleaved sums) in a fixed order.  Called by every thread of the
// last CTA.  Split-model path: the rank's partial goes to the peers (fused round: stored into their
// mailboxes right here; NCCL fallback: parked in st->red for the all-gather that follows).
template <int NS>
__device__ __forceinline__ void reduce_partials_and_finalize(int mode, const double *partial,
                                                             KState *st, double *out, double *sh,
                                                             const DistPush &push) {
  double a0 = 0.0, a1 = 0.0;
  for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    a0 += partial[NS * i];
    if (NS == 2) a1 += partial[NS * i + 1];
  }
  a0 = block_sum(a0, sh);
  if (NS == 2) a1 = block_sum(a1, sh);
  if (push.peer) {
    if (threadIdx.x == 0) {  // block_sum leaves its result in thread 0 only
      sh[0] = a0;
      sh[1] = a1;
      for (int i = 2; i < 8; i++) sh[i] = 0.0;
    }
    __syncthreads();
    dist_push_record(push, sh);
    return;
  }
  if (threadIdx.x == 0) {
    if (st->dist) {
      st->red.s[0] = a0;
      st->red.s[1] = a1;
    } else {
      finalize_sum(mode, a0, a1, st, out);
    }
  }
}

// Consumer side of a fused round (every CTA): lanes 0..nranks-1 of the first warp wait for one record each,
// the sums are then combined in rank order (identical on every CTA and every rank); sc[0..1] = the two sums.
__device__ __forceinline__ void dist_combine(const SmallGather &g, double *sc) {
  if (threadIdx.x < 32) {
    double v0 = 0.0, v1 = 0.0;
    if ((int)threadIdx.x < g.nranks) {
      double v[2];
      small_gather_read(g, threadIdx.x, v, 2);
      v0 = v[0];
      v1 = v[1];
    }
    double s0 = 0.0, s1 = 0.0;
    for (int r = 0; r < g.nranks; r++) {
      s0 += __shfl_sync(0xffffffffu, v0, r);
      s1 += __shfl_sync(0xffffffffu, v1, r);
    }
    if (threadIdx.x == 0) {
      sc[0] = s0;
      sc[1] = s1;
    }
  }
  __syncthreads();
}

// ---- dot product (ddot) ------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
dot_kernel(int n, const double *__restrict__ a, const double *__restrict__ b,
           double *__restrict__ partial, unsigned int *ticket, KState *st, int mode,
           double *out, int check_done, DistPush push) {
  __shared__ double sh[8];
  __shared__ bool last;
  if (check_done && st->done) return;
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    s += a[i] * b[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (last_block(ticket, &last)) reduce_partials_and_finalize<1>(mode, partial, st, out, sh, push);
}

// ---- dnrm2 in two passes (max |d|, then sum (d/scale)^2) ---------------------
__global__ void __launch_bounds__(kBlock)
nrm_max_kernel(int n, const double *__restrict__ a, double *__restrict__ partial,
               unsigned int *ticket, KState *st, double *out) {
  __shared__ double sh[8];
  __shared__ bool last;
  double m = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    m = fmax(m, fabs(a[i]));
  m = block_max(m, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = m;
  if (last_block(ticket, &last)) {
    double r = 0.0;
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) r = fmax(r, partial[i]);
    r = block_max(r, sh);
    if (threadIdx.x == 0) {
      if (st->dist)
        st->red.s[0] = r;
      else
        *out = r;
    }
  }
}

__global__ void __launch_bounds__(kBlock)
nrm_ssq_kernel(int n, const double *__restrict__ a, double *__restrict__ partial,
               unsigned int *ticket, KState *st, double *scale_io) {
  __shared__ double sh[8];
  __shared__ bool last;
  const double scale = *scale_io;
  double s = 0.0;
  if (scale > 0.0)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const double r = fabs(a[i]) / scale;
      s += r * r;
    }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (last_block(ticket, &last))
    reduce_partials_and_finalize<1>(FIN_NRM_SSQ, partial, st, scale_io, sh, DistPush{});
}

// ---- y = A x (+ fused dot products with the freshly computed y) --------------
// EPI 0: y = A x ; EPI 1: y = b - A x (ims_base_residual)
// NDOT 0: none ; 1: sum w[row]*y[row] ; 2: sum w[row]*y[row] and sum y[row]^2
// Fused split-model path (H.nnbr > 0): halo columns are read from the peer-memory mailbox; rows of slices
// that touch one are deferred behind the interior rows and wait for the neighbours' flags there.
// (A x)_row for such a row, same slot order as sell_row_dot / sell_row_dot_w
__device__ __forceinline__ double sell_row_dot_halo(int row, long long base, int len, const int *__restrict__ col,
                                                    const double *__restrict__ val, const double *__restrict__ x,
                                                    const HaloSrc &H) {
  double t = 0.0;
  for (int k = 0; k < len; k++) {
    const long long p = base + 32LL * k;
    const int c = __ldg(col + p);
    const double xv = (c < H.n_own) ? x[c] : halo_load(H, c);
    t = t + __ldg(val + p) * xv;
  }
  return t;
}

// fixed-width layout: all loads of the row issued up front like sell_row_dot_w
template <int W>
__device__ __forceinline__ double sell_row_dot_w_halo(int row, const int *__restrict__ col,
                                                      const double *__restrict__ val, const double *__restrict__ x,
                                                      const HaloSrc &H) {
  const long long base = (long long)(row >> 5) * (32 * W) + (row & 31);
  double v[W], xv[W];
  int c[W];
#pragma unroll
  for (int u = 0; u < W; u++) {
    v[u] = __ldg(val + base + 32 * u);
    c[u] = __ldg(col + base + 32 * u);
  }
#pragma unroll
  for (int u = 0; u < W; u++) xv[u] = (c[u] < H.n_own) ? x[c[u]] : halo_load(H, c[u]);
  double t = 0.0;
#pragma unroll
  for (int u = 0; u < W; u++) t = t + v[u] * xv[u];
  return t;
}

template <int EPI, int NDOT>
__global__ void __launch_bounds__(kBlock)
spmv_fused_kernel(int n, const int *__restrict__ slice_ptr,
                  const unsigned char *__restrict__ rowlen, const int *__restrict__ col,
                  const double *__restrict__ val, const double *__restrict__ x,
                  double *__restrict__ y, const double *__restrict__ b,
                  const double *__restrict__ w, double *__restrict__ partial,
                  unsigned int *ticket, KState *st, int mode, int check_done,
                  const __grid_constant__ HaloSrc H, DistPush push) {
  __shared__ double sh[8];
  __shared__ bool last;
  if (check_done && st->done) return;
  double s0 = 0.0, s1 = 0.0, h0 = 0.0, h1 = 0.0;  // interior rows / rows that read halo columns
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
    if (H.nnbr > 0 && H.slice_halo[row >> 5]) continue;  // deferred: waits for the neighbours below
    double t = sell_row_dot(row, slice_ptr, rowlen, col, val, x);
    if (EPI == 1) t = b[row] - t;
    y[row] = t;
    if (NDOT >= 1) s0 += w[row] * t;
    if (NDOT == 2) s1 += t * t;
  }
  if (H.nnbr > 0) {
    const bool listed = (H.grid == (int)gridDim.x);
    const int e0 = listed ? H.def_ptr[blockIdx.x] : 0, e1 = listed ? H.def_ptr[blockIdx.x + 1] : 1;
    if (e1 > e0) {
      halo_wait(H);
      if (listed) {
        for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
          const int row = H.def_row[e];
          if (row < 0) continue;
          double t = sell_row_dot_halo(row, (long long)slice_ptr[row >> 5] + (row & 31), rowlen[row], col, val, x, H);
          if (EPI == 1) t = b[row] - t;
          y[row] = t;
          if (NDOT >= 1) h0 += w[row] * t;
          if (NDOT == 2) h1 += t * t;
        }
      } else {
        for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
          if (!H.slice_halo[row >> 5]) continue;
          double t = sell_row_dot_halo(row, (long long)slice_ptr[row >> 5] + (row & 31), rowlen[row], col, val, x, H);
          if (EPI == 1) t = b[row] - t;
          y[row] = t;
          if (NDOT >= 1) h0 += w[row] * t;
          if (NDOT == 2) h1 += t * t;
        }
      }
    }
    s0 += h0;
    s1 += h1;
  }
  if (NDOT >= 1) {
    s0 = block_sum(s0, sh);
    if (NDOT == 2) s1 = block_sum(s1, sh);
    if (threadIdx.x == 0) {
      partial[NDOT * blockIdx.x] = s0;
      if (NDOT == 2) partial[NDOT * blockIdx.x + 1] = s1;
    }
    if (last_block(ticket, &last))
      reduce_partials_and_finalize<(NDOT == 2 ? 2 : 1)>(mode, partial, st, nullptr, sh, push);
  }
}

// fixed-width variant (see sell_row_dot_w): no slice_ptr / rowlen loads
template <int EPI, int NDOT, int W>
__global__ void __launch_bounds__(kBlock)
spmv_fused_w_kernel(int n, int ncols, const int *__restrict__ col, const int *__restrict__ soff,
                    const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
                    const double *__restrict__ b, const double *__restrict__ w,
                    double *__restrict__ partial, unsigned int *ticket, KState *st, int mode,
                    int check_done, const __grid_constant__ HaloSrc H, DistPush push) {
  __shared__ double sh[8];
  __shared__ bool last;
  if (check_done && st->done) return;
  double s0 = 0.0, s1 = 0.0, h0 = 0.0, h1 = 0.0;  // interior rows / rows that read halo columns
  // Fused split-model path: when the neighbours' messages are already there (the usual case: they are pushed by
  // the kernel that precedes this one on every rank) every row is handled in the one pass; otherwise the rows
  // that read halo columns are left for the end of the CTA that owns them and wait there.  Both modes give a
  // thread the same rows and keep the two partial sums apart, so the result does not depend on the timing.
  __shared__ int arrived;
  if (H.nnbr > 0) {
    if (threadIdx.x == 0) arrived = halo_arrived(H) ? 1 : 0;
    __syncthreads();
  }
  const bool inline_halo = H.nnbr > 0 && arrived != 0;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
    if (H.nnbr > 0 && H.slice_halo[row >> 5]) {
      if (!inline_halo) continue;
      double t = sell_row_dot_w_halo<W>(row, col, val, x, H);
      if (EPI == 1) t = b[row] - t;
      y[row] = t;
      if (NDOT >= 1) h0 += w[row] * t;
      if (NDOT == 2) h1 += t * t;
      continue;
    }
    double t = sell_row_dot_w<W>(row, col, val, x, soff, ncols);
    if (EPI == 1) t = b[row] - t;
    y[row] = t;
    if (NDOT >= 1) s0 += w[row] * t;
    if (NDOT == 2) s1 += t * t;
  }
  if (H.nnbr > 0 && !inline_halo) {
    const bool listed = (H.grid == (int)gridDim.x);
    const int e0 = listed ? H.def_ptr[blockIdx.x] : 0, e1 = listed ? H.def_ptr[blockIdx.x + 1] : 1;
    if (e1 > e0) {
      halo_wait(H);
      if (listed) {
        for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
          const int row = H.def_row[e];
          if (row < 0) continue;
          double t = sell_row_dot_w_halo<W>(row, col, val, x, H);
          if (EPI == 1) t = b[row] - t;
          y[row] = t;
          if (NDOT >= 1) h0 += w[row] * t;
          if (NDOT == 2) h1 += t * t;
        }
      } else {
        for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
          if (!H.slice_halo[row >> 5]) continue;
          double t = sell_row_dot_w_halo<W>(row, col, val, x, H);
          if (EPI == 1) t = b[row] - t;
          y[row] = t;
          if (NDOT >= 1) h0 += w[row] * t;
          if (NDOT == 2) h1 += t * t;
        }
      }
    }
  }
  s0 += h0;
  s1 += h1;
  if (NDOT >= 1) {
    s0 = block_sum(s0, sh);
    if (NDOT == 2) s1 = block_sum(s1, sh);
    if (threadIdx.x == 0) {
      partial[NDOT * blockIdx.x] = s0;
      if (NDOT == 2) partial[NDOT * blockIdx.x + 1] = s1;
    }
    if (last_block(ticket, &last))
      reduce_partials_and_finalize<(NDOT == 2 ? 2 : 1)>(mode, partial, st, nullptr, sh, push);
  }
}

template <int EPI, int NDOT>
static void launch_spmv_fused(const mf6gpu_matrix &A, int G, cudaStream_t S, const double *x, double *y,
                              const double *b, const double *w, double *partial, unsigned int *ticket,
                              KState *st, int mode, int check_done, const HaloSrc &H = HaloSrc{},
                              const DistPush &push = DistPush{}) {
  const int N = A.n;
#define MF6_SPMV_W(WW)                                                                              \
  case WW:                                                                                          \
    spmv_fused_w_kernel<EPI, NDOT, WW><<<G, kBlock, 0, S>>>(N, A.n_ext, A.col.p,                     \
                                                            A.slot_off.n ? A.slot_off.p : nullptr,  \
                                                            A.val.p, x, y, b, w, partial,           \
                                                            ticket, st, mode, check_done, H, push); \
    return;
  switch (A.uniform_w) {
    MF6_SPMV_W(4)
    MF6_SPMV_W(5)
    MF6_SPMV_W(6)
    MF6_SPMV_W(7)
    MF6_SPMV_W(8)
    MF6_SPMV_W(9)
    MF6_SPMV_W(10)
    default:
      break;
  }
#undef MF6_SPMV_W
  spmv_fused_kernel<EPI, NDOT><<<G, kBlock, 0, S>>>(N, A.slice_ptr.p, A.rowlen.p, A.col.p, A.val.p, x, y, b,
                                                    w, partial, ticket, st, mode, check_done, H, push);
}

// ---- vector updates -----------------------------------------------------------
// CG: P = Z (first) | P = Z + beta P                     ImsLinearBase.f90:118-127
// Fused split-model path: consumes the rho round (every CTA combines the ranks' records itself) and its last
// CTA pushes the halo cells of the new P to the neighbours.
__global__ void __launch_bounds__(kBlock)
cg_p_kernel(int n, const double *__restrict__ z, double *__restrict__ p, KState *st,
            int first, SmallGather pull, HaloPush hp) {
  __shared__ double sc[2];
  __shared__ bool last;
  if (st->done) return;
  double beta = st->beta;
  if (pull.base) {
    dist_combine(pull, sc);
    const double rho = sc[0];
    beta = rho / st->rho0;  // unused on the first iteration
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      st->rho = rho;
      st->beta = beta;
    }
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    p[i] = first ? z[i] : z[i] + beta * p[i];
  if (hp.peer) halo_push_tail(hp, p, &last);
}

// BCGS: P = D (first) | P = D + beta (P - omega0 V)      ImsLinearBase.f90:346-355
__global__ void __launch_bounds__(kBlock)
bcgs_p_kernel(int n, const double *__restrict__ d, const double *__restrict__ v,
              double *__restrict__ p, KState *st, int first, SmallGather pull) {
  __shared__ double sc[2];
  if (st->done) return;
  double beta = st->beta;
  const double omega0 = st->omega0;
  if (pull.base) {
    dist_combine(pull, sc);
    const double rho = sc[0];
    beta = (rho / st->rho0) * (st->alpha0 / omega0);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      st->rho = rho;
      st->beta = beta;
    }
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    p[i] = first ? d[i] : d[i] + beta * (p[i] - omega0 * v[i]);
}

// BCGS: Q = D - alpha V                                   ImsLinearBase.f90:380-382
__global__ void __launch_bounds__(kBlock)
bcgs_q_kernel(int n, const double *__restrict__ d, const double *__restrict__ v,
              double *__restrict__ q, KState *st, SmallGather pull) {
  __shared__ double sc[2];
  if (st->done) return;
  double alpha = st->alpha;
  if (pull.base) {
    dist_combine(pull, sc);
    double den = sc[0];
    den = den + dsign(DBL_EPSILON, den);
    alpha = st->rho / den;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->alpha = alpha;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    q[i] = d[i] - alpha * v[i];
}

struct SummaryPtrs {
  int *itinner, *locdv, *locr;
  double *dvmax, *rmax, *alpha, *omega;
};

// scalar tail of one inner iteration: everything between the update loop and
// "SAVE CURRENT INNER ITERATES" (ImsLinearBase.f90:183-235 / 480-544)
__device__ void finalize_iteration(KState *st, double ssq, MaxLoc mx, MaxLoc mr, int bcgs,
                                   SummaryPtrs sp) {
  const double l2norm = sqrt(ssq);
  st->deltax = mx.v;
  st->rmax = mr.v;
  st->l2norm = l2norm;
  st->xloc = mx.idx;
  st->rloc = mr.idx;
  st->iter += 1;
  st->sum_count += 1;
  const int k = st->sum_count - 1;
  if (k < st->sum_cap) {
    sp.itinner[k] = st->iter;
    sp.dvmax[k] = mx.v;
    sp.locdv[k] = mx.idx;
    sp.rmax[k] = mr.v;
    sp.locr[k] = mr.idx;
    sp.alpha[k] = st->alpha;
    sp.omega[k] = bcgs ? st->omega : 0.0;
  }
  const int opt = st->icnvgopt;
  const double rcnvg = (opt == 2 || opt == 3 || opt == 4) ? l2norm : mr.v;
  int icnvg = st->icnvg;
  testcnvg_dev(opt, icnvg, st->iter, mx.v, rcnvg, st->l2norm0, st->epfact, st->dvclose,
               st->rclose);
  if (rcnvg == 0.0) icnvg = 1;
  st->icnvg = icnvg;
  int done = 0;
  if (icnvg != 0) done = 1;
  if (!done && is_close_dev(st->rho, st->rho0)) done = 1;
  if (bcgs) {
    if (!done && is_close_dev(st->alpha, st->alpha0)) done = 1;
    if (!done && is_close_dev(st->omega, st->omega0)) done = 1;
    if (!done && st->rho * st->omega == 0.0) done = 1;
  } else {
    if (!done && st->rho == 0.0) done = 1;
  }
  if (!done) {
    st->rho0 = st->rho;
    st->alpha0 = st->alpha;
    st->omega0 = st->omega;
  }
  st->done = done;
}

// CG: X += alpha P ; D -= alpha Q ; max|alpha P|, max|D|, sum D^2   :137-183
// BCGS: X += alpha PHAT + omega QHAT ; D = Q - omega T ; same reductions :429-480
// VEC2: 128-bit loads / stores (two rows per access; needs 16-byte aligned vectors), two accesses in flight
template <int BCGS, int VEC2>
__global__ void __launch_bounds__(kBlock)
update_kernel(int n, double *__restrict__ x, double *__restrict__ d,
              const double *__restrict__ p, const double *__restrict__ q,
              const double *__restrict__ qhat, const double *__restrict__ t,
              const double *__restrict__ dscale, const int *__restrict__ ord,
              double *__restrict__ partial, MaxLoc *__restrict__ pmx, MaxLoc *__restrict__ pmr,
              unsigned int *ticket, KState *st, SummaryPtrs sp, SmallGather pull, DistPush push) {
  __shared__ double sh[8];
  __shared__ MaxLoc shm[8];
  __shared__ bool last;
  __shared__ RedRec shrec;
  if (st->done) return;
  double alpha = st->alpha, omega = st->omega;
  if (pull.base) {
    // fused split-model path: this kernel consumes the round of the SpMV that precedes it
    dist_combine(pull, sh);
    if (BCGS) {
      double den = sh[1];
      den = den + dsign(DBL_EPSILON, den);
      omega = sh[0] / den;
      if (blockIdx.x == 0 && threadIdx.x == 0) st->omega = omega;
    } else {
      double den = sh[0];
      den = den + dsign(DBL_EPSILON, den);
      alpha = st->rho / den;
      if (blockIdx.x == 0 && threadIdx.x == 0) st->alpha = alpha;
    }
    __syncthreads();
  }
  const int iscl = st->iscl;
  double ssq = 0.0;
  MaxLoc mx = maxloc_init(), mr = maxloc_init();
  // one row: the reference's update of X and D plus the three reductions, in the reference's operation order
  auto row = [&](int i, double xi, double di, double pi, double qi, double qh, double ti, double sc,
                 double &xo, double &dout) {
    double tv, rv;
    if (BCGS) {
      tv = alpha * pi + omega * qh;  // p = PHAT here
      xo = xi + tv;
      if (iscl != 0) tv = tv * sc;
      rv = qi - omega * ti;
      dout = rv;
      if (iscl != 0) rv = rv / sc;
    } else {
      tv = alpha * pi;
      xo = xi + tv;
      rv = di;
      rv = rv - alpha * qi;
      dout = rv;
    }
    // running maxima: only the row index is tracked; its position in the reference's loop order (`ord`, the
    // tie-breaker) is looked up on an exact tie and once at the end -- no index loads in the streaming loop
    const double atv = fabs(tv), arv = fabs(rv);
    if (atv > mx.a) {
      mx.a = atv;
      mx.v = tv;
      mx.idx = i;
    } else if (atv == mx.a && atv > 0.0 && ord && ord[i] < ord[mx.idx]) {
      mx.v = tv;
      mx.idx = i;
    }
    if (arv > mr.a) {
      mr.a = arv;
      mr.v = rv;
      mr.idx = i;
    } else if (arv == mr.a && arv > 0.0 && ord && ord[i] < ord[mr.idx]) {
      mr.v = rv;
      mr.idx = i;
    }
    ssq += rv * rv;
  };
  if (VEC2) {
    const int n2 = n >> 1;
    double2 *x2 = reinterpret_cast<double2 *>(x), *d2 = reinterpret_cast<double2 *>(d);
    const double2 *p2 = reinterpret_cast<const double2 *>(p), *q2 = reinterpret_cast<const double2 *>(q);
    const double2 *qh2 = reinterpret_cast<const double2 *>(qhat), *t2 = reinterpret_cast<const double2 *>(t);
    const double2 *s2 = reinterpret_cast<const double2 *>(dscale);
    const int stride = gridDim.x * blockDim.x;
    const double2 zero2 = make_double2(0.0, 0.0), one2 = make_double2(1.0, 1.0);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += 2 * stride) {
      const int j1 = j + stride;
      const bool two = j1 < n2;
      // all loads of both accesses first
      double2 X0 = x2[j], P0 = p2[j], Q0 = q2[j];
      double2 D0 = BCGS ? zero2 : d2[j];
      double2 H0 = BCGS ? qh2[j] : zero2, T0 = BCGS ? t2[j] : zero2;
      double2 S0 = (BCGS && iscl != 0) ? s2[j] : one2;
      double2 X1 = zero2, P1 = zero2, Q1 = zero2, D1 = zero2, H1 = zero2, T1 = zero2, S1 = one2;
      if (two) {
        X1 = x2[j1];
        P1 = p2[j1];
        Q1 = q2[j1];
        if (!BCGS) D1 = d2[j1];
        if (BCGS) {
          H1 = qh2[j1];
          T1 = t2[j1];
          if (iscl != 0) S1 = s2[j1];
        }
      }
      double2 XO, DO;
      row(2 * j, X0.x, D0.x, P0.x, Q0.x, H0.x, T0.x, S0.x, XO.x, DO.x);
      row(2 * j + 1, X0.y, D0.y, P0.y, Q0.y, H0.y, T0.y, S0.y, XO.y, DO.y);
      x2[j] = XO;
      d2[j] = DO;
      if (two) {
        row(2 * j1, X1.x, D1.x, P1.x, Q1.x, H1.x, T1.x, S1.x, XO.x, DO.x);
        row(2 * j1 + 1, X1.y, D1.y, P1.y, Q1.y, H1.y, T1.y, S1.y, XO.y, DO.y);
        x2[j1] = XO;
        d2[j1] = DO;
      }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      const int i = n - 1;
      double xo, dout;
      row(i, x[i], BCGS ? 0.0 : d[i], p[i], q[i], BCGS ? qhat[i] : 0.0, BCGS ? t[i] : 0.0,
          (BCGS && iscl != 0) ? dscale[i] : 1.0, xo, dout);
      x[i] = xo;
      d[i] = dout;
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      double xo, dout;
      row(i, x[i], BCGS ? 0.0 : d[i], p[i], q[i], BCGS ? qhat[i] : 0.0, BCGS ? t[i] : 0.0,
          (BCGS && iscl != 0) ? dscale[i] : 1.0, xo, dout);
      x[i] = xo;
      d[i] = dout;
    }
  }
  if (mx.idx >= 0) mx.ord = ord ? ord[mx.idx] : mx.idx;
  if (mr.idx >= 0) mr.ord = ord ? ord[mr.idx] : mr.idx;
  ssq = block_sum(ssq, sh);
  mx = block_maxloc(mx, shm);
  mr = block_maxloc(mr, shm);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = ssq;
    pmx[blockIdx.x] = mx;
    pmr[blockIdx.x] = mr;
  }
  if (last_block(ticket, &last)) {
    double a = 0.0;
    MaxLoc gx = maxloc_init(), gr = maxloc_init();
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
      a += partial[i];
      maxloc_merge(gx, pmx[i]);
      maxloc_merge(gr, pmr[i]);
    }
    a = block_sum(a, sh);
    gx = block_maxloc(gx, shm);
    gr = block_maxloc(gr, shm);
    if (threadIdx.x == 0) {
      if (push.peer) {
        shrec.s[0] = a;
        shrec.s[1] = 0.0;
        shrec.mx = MaxLocPOD{gx.a, gx.v, gx.ord, gx.idx};
        shrec.mr = MaxLocPOD{gr.a, gr.v, gr.ord, gr.idx};
      } else if (st->dist) {
        st->red.s[0] = a;
        st->red.mx = MaxLocPOD{gx.a, gx.v, gx.ord, gx.idx};
        st->red.mr = MaxLocPOD{gr.a, gr.v, gr.ord, gr.idx};
      } else {
        finalize_iteration(st, a, gx, gr, BCGS, sp);
      }
    }
    if (push.peer) {
      __syncthreads();
      dist_push_record(push, reinterpret_cast<const double *>(&shrec));
    }
  }
}

enum { FIN_UPDATE = 100 };

// Per-model maxima of the iteration the update kernel has just finished (ImsLinearBase.f90:143-176, 186-197):
// grid.y = model.  dx and the residual are recomputed from the same operands (alpha, omega are final, P / PHAT /
// QHAT untouched until the next iteration, D holds the new residual), so the values are the ones the update
// kernel saw.  Record index = summary%iter_cnt - 1, filled once per iteration; a no-op rerun after the loop has
// ended rewrites the same record with the same values.
__global__ void __launch_bounds__(kBlock)
model_summary_kernel(int n, int nmod, const int *__restrict__ modid, const double *__restrict__ d,
                     const double *__restrict__ p, const double *__restrict__ qhat,
                     const double *__restrict__ dscale, const int *__restrict__ ord, const KState *st, int bcgs,
                     MaxLoc *__restrict__ pmx, MaxLoc *__restrict__ pmr, unsigned int *tickets,
                     double *__restrict__ odv, int *__restrict__ olocdv, double *__restrict__ orm,
                     int *__restrict__ olocr) {
  __shared__ MaxLoc shm[8];
  __shared__ bool last;
  const int im = blockIdx.y;
  const double alpha = st->alpha, omega = st->omega;
  const int iscl = st->iscl;
  MaxLoc mx = maxloc_init(), mr = maxloc_init();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (modid[i] != im) continue;
    double tv = bcgs ? alpha * p[i] + omega * qhat[i] : alpha * p[i];
    double rv = d[i];
    if (bcgs && iscl != 0) {
      tv = tv * dscale[i];
      rv = rv / dscale[i];
    }
    const double atv = fabs(tv), arv = fabs(rv);
    if (atv >= mx.a && atv > 0.0) maxloc_take(mx, tv, ord ? ord[i] : i, i);
    if (arv >= mr.a && arv > 0.0) maxloc_take(mr, rv, ord ? ord[i] : i, i);
  }
  mx = block_maxloc(mx, shm);
  mr = block_maxloc(mr, shm);
  if (threadIdx.x == 0) {
    pmx[(size_t)im * gridDim.x + blockIdx.x] = mx;
    pmr[(size_t)im * gridDim.x + blockIdx.x] = mr;
  }
  if (last_block(tickets + im, &last)) {
    MaxLoc gx = maxloc_init(), gr = maxloc_init();
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
      maxloc_merge(gx, pmx[(size_t)im * gridDim.x + i]);
      maxloc_merge(gr, pmr[(size_t)im * gridDim.x + i]);
    }
    gx = block_maxloc(gx, shm);
    gr = block_maxloc(gr, shm);
    if (threadIdx.x == 0) {
      const int k = st->sum_count - 1;
      if (k >= 0 && k < st->sum_cap) {
        const size_t q = (size_t)k * nmod + im;
        odv[q] = gx.v;
        olocdv[q] = gx.idx;
        orm[q] = gr.v;
        olocr[q] = gr.idx;
      }
    }
  }
}

// split-model path: combine the gathered per-rank records in rank order (deterministic and
// identical on every rank) and run the same scalar epilogue the single-GPU kernels run inline
__global__ void global_finalize_kernel(int mode, const RedRec *__restrict__ all_in, int nranks,
                                       KState *st, double *out, int bcgs, SummaryPtrs sp, SmallGather sg) {
  // launched with one warp.  Peer-memory path: lane r waits for rank r's flag and copies the record out of
  // the mailbox (waited for even when the loop is already done, so that every rank consumes every round)
  __shared__ RedRec rec[32];
  const RedRec *all = all_in;
  const bool krylov = !(mode == FIN_NRM_MAX || mode == FIN_NRM_SSQ);
  if (krylov && st->done) return;  // every rank takes the same decision: nobody pushed, nobody waits
  if (sg.base) {
    if ((int)threadIdx.x < nranks)
      small_gather_read(sg, threadIdx.x, reinterpret_cast<double *>(&rec[threadIdx.x]),
                        (int)(sizeof(RedRec) / sizeof(double)));
    __syncwarp();
    all = rec;
  }
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s0 = 0.0, s1 = 0.0;
  MaxLoc mx = maxloc_init(), mr = maxloc_init();
  for (int r = 0; r < nranks; r++) {
    if (mode == FIN_NRM_MAX) {
      s0 = fmax(s0, all[r].s[0]);
    } else {
      s0 += all[r].s[0];
      s1 += all[r].s[1];
    }
    if (mode == FIN_UPDATE) {
      maxloc_merge(mx, MaxLoc{all[r].mx.a, all[r].mx.v, all[r].mx.ord, all[r].mx.ord});
      maxloc_merge(mr, MaxLoc{all[r].mr.a, all[r].mr.v, all[r].mr.ord, all[r].mr.ord});
    }
  }
  if (mode == FIN_UPDATE)
    finalize_iteration(st, s0, mx, mr, bcgs, sp);  // locations are GLOBAL cell ids here
  else
    finalize_sum(mode, s0, s1, st, out);
}

__global__ void copy_kernel(int n, const double *__restrict__ a, double *__restrict__ b) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    b[i] = a[i];
}

__global__ void init_state_kernel(KState *st, double epfact, double dvclose, double rclose,
                                  int icnvgopt, int sum_cap, int iscl, int reset_count, int dist) {
  st->dist = dist;
  st->rho = st->rho0 = st->alpha = st->alpha0 = st->omega = st->omega0 = st->beta = 0.0;
  st->rho_acc = 0.0;
  st->epfact = epfact;
  st->dvclose = dvclose;
  st->rclose = rclose;
  st->deltax = st->rmax = st->l2norm = 0.0;
  st->xloc = st->rloc = -1;
  st->icnvgopt = icnvgopt;
  st->icnvg = 0;
  st->done = 0;
  st->iter = 0;
  if (reset_count) st->sum_count = 0;
  st->sum_cap = sum_cap;
  st->iscl = iscl;
}

// ims_base_scale, symmetric diagonal scaling (ISCL = 1) :644-663, 727-752
__global__ void scale1_vec_kernel(int n, const int *__restrict__ slice_ptr,
                                  const double *__restrict__ val, double *__restrict__ dscale) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const long long base = (long long)slice_ptr[r >> 5] + (r & 31);
    dscale[r] = 1.0 / sqrt(fabs(val[base]));
  }
}

// ---- SCALING_METHOD L2NORM (ims_base_scale, ImsLinearBase.f90:676-721) -------------------------
// pass 1: dscale(n) = 1 / ||row n||_2 (1 for an empty row), row scaled in place
__global__ void scale2_rows_kernel(int n, const int *__restrict__ slice_ptr,
                                   const unsigned char *__restrict__ rowlen, double *__restrict__ val,
                                   double *__restrict__ dscale) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const long long base = (long long)slice_ptr[r >> 5] + (r & 31);
    const int len = rowlen[r];
    double c1 = 0.0;
    for (int k = 0; k < len; k++) {
      const double a = val[base + 32LL * k];
      c1 = c1 + a * a;
    }
    c1 = sqrt(c1);
    c1 = (c1 == 0.0) ? 1.0 : 1.0 / c1;
    dscale[r] = c1;
    for (int k = 0; k < len; k++) val[base + 32LL * k] = c1 * val[base + 32LL * k];
  }
}

// pass 2: dscale2(j) = 1 / ||column j||_2 of the row-scaled matrix.  The entries of column j are the
// transposed positions of row j's entries (structurally symmetric matrix); they are gathered in
// ascending row order -- lower neighbours, diagonal, upper neighbours -- like the reference's row loop
// accumulates them, so no atomics and a fixed summation order.
__global__ void scale2_cols_kernel(int n, const int *__restrict__ slice_ptr,
                                   const unsigned char *__restrict__ rowlen,
                                   const unsigned char *__restrict__ nlow, const int *__restrict__ col,
                                   const double *__restrict__ val, double *__restrict__ dscale2) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const long long base = (long long)slice_ptr[j >> 5] + (j & 31);
    const int len = rowlen[j], lo = nlow[j];
    double c2 = 0.0;
    for (int k = 1; k <= len; k++) {
      // visiting order: slots 1..lo, then the diagonal (slot 0), then slots lo+1..len-1
      const int slot = (k <= lo) ? k : (k == lo + 1 ? 0 : k - 1);
      double a = 0.0;
      if (slot == 0) {
        a = val[base];
      } else {
        const int i = col[base + 32LL * slot];
        if (i < n) {  // (halo columns have no row here)
          const long long bi = (long long)slice_ptr[i >> 5] + (i & 31);
          const int leni = rowlen[i];
          for (int k2 = 1; k2 < leni; k2++)
            if (col[bi + 32LL * k2] == j) {
              a = val[bi + 32LL * k2];
              break;
            }
        }
      }
      c2 = c2 + a * a;
    }
    dscale2[j] = (c2 == 0.0) ? 1.0 : 1.0 / sqrt(c2);
  }
}

// pass 3: column scaling of the matrix, then x / dscale2 and b * dscale (ImsLinearBase.f90:712-727)
__global__ void scale2_apply_kernel(int n, const int *__restrict__ slice_ptr,
                                    const unsigned char *__restrict__ rowlen, const int *__restrict__ col,
                                    double *__restrict__ val, const double *__restrict__ ds,
                                    const double *__restrict__ ds2, double *__restrict__ x,
                                    double *__restrict__ b) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const long long base = (long long)slice_ptr[r >> 5] + (r & 31);
    const int len = rowlen[r];
    for (int k = 0; k < len; k++) {
      const long long p = base + 32LL * k;
      val[p] = ds2[col[p]] * val[p];
    }
    x[r] = x[r] / ds2[r];
    b[r] = b[r] * ds[r];
  }
}

__global__ void scale_apply_kernel(int n, const int *__restrict__ slice_ptr,
                                   const unsigned char *__restrict__ rowlen,
                                   const int *__restrict__ col, double *__restrict__ val,
                                   const double *__restrict__ ds, const double *__restrict__ ds2,
                                   double *__restrict__ x, double *__restrict__ b, int unscale) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const long long base = (long long)slice_ptr[r >> 5] + (r & 31);
    const int len = rowlen[r];
    const double c1 = ds[r];
    for (int k = 0; k < len; k++) {
      const long long p = base + 32LL * k;
      const double c2 = ds2[col[p]];
      if (!unscale)
        val[p] = c1 * val[p] * c2;
      else
        val[p] = (1.0 / c1) * val[p] * (1.0 / c2);
    }
    const double c2 = ds2[r];
    if (!unscale) {
      x[r] = x[r] / c2;
      b[r] = b[r] * c1;
    } else {
      x[r] = x[r] * c2;
      b[r] = b[r] / c1;
    }
  }
}

}  // namespace mf6

using namespace mf6;

// ImsLinearBase.f90:1316-1333 (0.01 / 0.10 are single-precision literals there)
static double epfact_of(int icnvgopt, int kstp) {
  if (icnvgopt == 2) return kstp == 1 ? (double)0.01f : (double)0.10f;
  if (icnvgopt == 4) return 1.0e-4;
  return 1.0;
}

void mf6gpu_solver::prof_begin(int cls) {
  if (!profiling || !prof_on) return;
  if (ev_used + 2 > ev_pool.size()) {
    const size_t old = ev_pool.size();
    ev_pool.resize(old + 1024);
    for (size_t i = old; i < ev_pool.size(); i++) MF6_CK(cudaEventCreate(&ev_pool[i]));
  }
  ev_cls.push_back(cls);
  MF6_CK(cudaEventRecord(ev_pool[ev_used++], stream));
}

void mf6gpu_solver::prof_end() {
  if (!profiling || !prof_on) return;
  MF6_CK(cudaEventRecord(ev_pool[ev_used++], stream));
}

void mf6gpu_solver::prof_collect() {
  if (!profiling) return;
  MF6_CK(cudaStreamSynchronize(stream));
  for (size_t k = 0; k < ev_cls.size(); k++) {
    float ms = 0.f;
    MF6_CK(cudaEventElapsedTime(&ms, ev_pool[2 * k], ev_pool[2 * k + 1]));
    prof_ms[ev_cls[k]] += ms;
    prof_cnt[ev_cls[k]] += 1;
  }
  ev_cls.clear();
  ev_used = 0;
}

void mf6gpu_solver::reduce_finalize(int mode, double *out, int bcgs) {
  if (!halo || !halo->active()) return;
  SummaryPtrs sp{sum_itinner.p, sum_locdv.p, sum_locr.p, sum_dvmax.p, sum_rmax.p, sum_alpha.p, sum_omega.p};
  const size_t cnt = sizeof(RedRec) / sizeof(double);
  SmallGather sg = comm_small_push(halo->comm, reinterpret_cast<const double *>(&st.p->red), cnt, stream);
  if (!sg.base) comm_allgather(halo->comm, reinterpret_cast<const double *>(&st.p->red), red_all.p, cnt, stream);
  global_finalize_kernel<<<1, 32, 0, stream>>>(mode, reinterpret_cast<const RedRec *>(red_all.p),
                                              halo->comm->nranks, st.p, out, bcgs, sp, sg);
  launches += 2;
}

int mf6gpu_solver::precond(const double *rin, double *dd, cudaStream_t S, const mf6::IluDotArgs *dot) {
  if (ilut) return ilut->apply(rin, dd, &st.p->done, S);
  return ilu0_apply(*A, lu.p, rin, dd, &st.p->done, S, dot);
}

// ims_base_pcu, ImsLinearBase.f90:808-858
int mf6gpu_solver::factor() {
  if (ilut) {  // ILUT / MILUT: sequential factorisation on the host, see ilut.cuh
    prof_begin(PC_FACTOR);
    const int c = ilut->factor(*A, A->val.p, s.relax, stream);
    prof_end();
    launches += 1;
    return c;
  }
  int ipcflag = 0, icount = 0;
  double delta = 0.0;
  for (;;) {
    failflag.zero(stream);
    prof_begin(PC_FACTOR);
    launches += ilu0_factor(*A, A->val.p, lu.p, s.relax, delta, ipcflag, failflag.p, stream);
    prof_end();
    MF6_CK(cudaGetLastError());
    MF6_CK(cudaMemcpyAsync(h_flag.p, failflag.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MF6_CK(cudaStreamSynchronize(stream));
    ipcflag = h_flag.p[0] ? 1 : 0;
    if (ipcflag < 1) break;
    delta = 1.5 * delta + 1.0e-3;
    ipcflag = 0;
    if (delta > 0.5) {
      delta = 0.5;
      ipcflag = 2;
    }
    icount++;
    if (icount > 10) break;
  }
  return icount;
}

void mf6gpu_solver::solve_device(int kiter, int kstp, double *x_dev, double *b_dev, int *iters,
                                 int *icnvg_out) {
  const int N = n;
  const int G = grid_for(N);
  cudaStream_t S = stream;
  const bool bcgs = (s.ilinmeth == 2);
  const bool dist = halo && halo->active();
  MF6_REQUIRE(!(dist && s.iscl != 0), "solver: SCALING_METHOD is not available on the split-model path");
  SummaryPtrs sp{sum_itinner.p, sum_locdv.p, sum_locr.p, sum_dvmax.p, sum_rmax.p, sum_alpha.p, sum_omega.p};
  launches = 0;
  MF6_CK(cudaEventRecord(ev[0], S));
  // -- scale (ImsLinear.f90:645-650)
  if (s.iscl == 1) {
    scale1_vec_kernel<<<G, kBlock, 0, S>>>(N, A->slice_ptr.p, A->val.p, dscale.p);
    copy_kernel<<<G, kBlock, 0, S>>>(N, dscale.p, dscale2.p);
    scale_apply_kernel<<<G, kBlock, 0, S>>>(N, A->slice_ptr.p, A->rowlen.p, A->col.p, A->val.p,
                                            dscale.p, dscale2.p, x_dev, b_dev, 0);
    launches += 3;
  } else if (s.iscl == 2) {
    scale2_rows_kernel<<<G, kBlock, 0, S>>>(N, A->slice_ptr.p, A->rowlen.p, A->val.p, dscale.p);
    scale2_cols_kernel<<<G, kBlock, 0, S>>>(N, A->slice_ptr.p, A->rowlen.p, A->nlow.p, A->col.p, A->val.p,
                                            dscale2.p);
    scale2_apply_kernel<<<G, kBlock, 0, S>>>(N, A->slice_ptr.p, A->rowlen.p, A->col.p, A->val.p, dscale.p,
                                             dscale2.p, x_dev, b_dev);
    launches += 3;
  }
  // -- preconditioner (ImsLinear.f90:669-673)
  npivfix = factor();
  prof_collect();
  MF6_CK(cudaEventRecord(ev[1], S));
  // -- initial residual and its norm (ImsLinear.f90:676-699)
  init_state_kernel<<<1, 1, 0, S>>>(st.p, epfact_of(s.icnvgopt, kstp), s.dvclose, s.rclose,
                                    s.icnvgopt, sum_cap, s.iscl, kiter == 1 ? 1 : 0, dist ? 1 : 0);
  p.zero(S);
  q.zero(S);
  z.zero(S);
  if (dist) halo->exchange(x_dev, S);
  launch_spmv_fused<1, 0>(*A, G, S, x_dev, d.p, b_dev, nullptr, nullptr, nullptr, st.p, 0, 0);
  double *scale_slot = partial.p + 3 * kMaxBlocks;  // scratch scalar
  nrm_max_kernel<<<G, kBlock, 0, S>>>(N, d.p, partial.p, tickets.p + TK_NRM, st.p, scale_slot);
  reduce_finalize(FIN_NRM_MAX, scale_slot, 0);
  nrm_ssq_kernel<<<G, kBlock, 0, S>>>(N, d.p, partial.p, tickets.p + TK_NRM, st.p, scale_slot);
  reduce_finalize(FIN_NRM_SSQ, scale_slot, 0);
  launches += 4;
  MF6_CK(cudaGetLastError());
  MF6_CK(cudaMemcpyAsync(h_st.p, st.p, sizeof(KState), cudaMemcpyDeviceToHost, S));
  MF6_CK(cudaStreamSynchronize(S));
  l2norm0 = h_st.p->l2norm0;
  int itmax = s.iter1;
  int icnvg = 0;
  if (l2norm0 == 0.0) {
    itmax = 0;
    icnvg = 1;
  }
  int innerit = 0;
  const int *ord = A->ord_ptr();
  const bool fuse_dot = (A->nlevels <= 16) && !ilut;
  IluDotArgs idot;
  idot.partial = ilu_partial.p;
  idot.cta_sums = partial.p + 2 * kMaxBlocks;
  idot.ticket = tickets.p + 4;
  idot.rho_out = dist ? &st.p->red.s[0] : &st.p->rho;
  idot.beta_out = dist ? &st.p->red.s[1] : &st.p->beta;
  idot.rho0 = &st.p->rho0;
  // polling cadence: cheap iterations (small n) are batched deeper
  const int batch = (N > 2000000) ? 4 : 16;
  if (itmax > 0) {
    if (bcgs) copy_kernel<<<G, kBlock, 0, S>>>(N, d.p, dhat.p);
    int launched = 0;
    bool finished = false;
    // one inner iteration = a fixed sequence of launches whose arguments never change (the scalars live in
    // KState): small systems are launch-bound on the host, so iterations 2.. replay a CUDA graph of the
    // sequence captured once per solve (not on the split-model path, while profiling, or with NORTH > 0)
    // Split-model path, two transports.  Fused (peer mailboxes mapped): the partial result of every reduction
    // is stored into the peers' mailboxes by the last CTA of the kernel that produces it and combined by every
    // CTA of the kernel that consumes it; the halo of the search direction is pushed by the last CTA of the
    // kernel that writes it and read in place by the SpMV, whose boundary rows wait while the interior runs.
    // One extra launch per iteration (the scalar tail of the update).  Fallback (NCCL): pack + send/recv,
    // all-gather + finalize launch after every reduction.
    const bool fused = dist && halo->fused() && !std::getenv("MF6GPU_NO_FUSED_EXCHANGE");
    fused_exchange = fused;
    auto round = [&]() { return fused ? comm_round(halo->comm) : DistRound{}; };
    auto finalize_update = [&](const DistRound &r, int bc) {
      if (fused) {
        global_finalize_kernel<<<1, 32, 0, S>>>(FIN_UPDATE, nullptr, halo->comm->nranks, st.p, nullptr, bc, sp,
                                                r.pull);
        launches += 1;
      } else {
        reduce_finalize(FIN_UPDATE, nullptr, bc);
      }
    };
    // halo of `vec` for the SpMV that follows; `carried` = the producer kernel pushes it itself
    auto halo_round = [&](HaloPush &hp, HaloSrc &hs, const double *vec, bool carried) {
      hp = HaloPush{};
      hs = HaloSrc{};
      if (!dist) return;
      if (fused) {
        halo->round(hp, hs);
        if (!carried) {
          halo->push_now(hp, vec, S);
          launches += 1;
        }
      } else {
        halo->exchange(const_cast<double *>(vec), S);
      }
    };
    // 128-bit vector accesses in the update kernel need 16-byte aligned vectors (cudaMalloc gives 256)
    const bool vec2 = ((reinterpret_cast<uintptr_t>(x_dev) & 15) == 0) && !std::getenv("MF6GPU_UPDATE_SCALAR");
    auto model_summary = [&](int bc) {
      if (nmod <= 1 || sum_cap <= 0 || dist) return;
      model_summary_kernel<<<dim3(G, nmod), kBlock, 0, S>>>(
          N, nmod, modid.p, d.p, bc ? phat.p : p.p, bc ? qhat.p : nullptr, dscale.p, ord, st.p, bc, mpmx.p, mpmr.p,
          mtickets.p, msum_dvmax.p, msum_locdv.p, msum_rmax.p, msum_locr.p);
      launches += 1;
    };
    auto enqueue_iteration = [&](int first) {
      HaloPush hp;
      HaloSrc hs;
      if (!bcgs) {
        const DistRound r1 = round();
        prof_begin(PC_ILU);
        if (fuse_dot) {
          // z = M^-1 d with rho = d.z (and beta = rho/rho0) accumulated by the same launches
          idot.push = r1.push;
          launches += precond(d.p, z.p, S, &idot);
          if (!fused) reduce_finalize(FIN_CG_RHO, nullptr, 0);
          prof_end();
        } else {
          launches += precond(d.p, z.p, S);
          prof_end();
          prof_begin(PC_DOT);
          dot_kernel<<<G, kBlock, 0, S>>>(N, d.p, z.p, partial.p, tickets.p + TK_DOT, st.p,
                                    FIN_CG_RHO, nullptr, 1, r1.push);
          if (!fused) reduce_finalize(FIN_CG_RHO, nullptr, 0);
          prof_end();
          launches += 1;
        }
        prof_begin(PC_PUPD);
        if (fused) halo->round(hp, hs); else { hp = HaloPush{}; hs = HaloSrc{}; }
        cg_p_kernel<<<G, kBlock, 0, S>>>(N, z.p, p.p, st.p, first, r1.pull, hp);
        if (dist && !fused) halo->exchange(p.p, S);
        prof_end();
        const DistRound r2 = round();
        prof_begin(PC_SPMV);
        launch_spmv_fused<0, 1>(*A, G, S, p.p, q.p, nullptr, p.p, partial.p, tickets.p + TK_SPMV, st.p,
                            FIN_CG_ALPHA, 1, hs, r2.push);
        if (!fused) reduce_finalize(FIN_CG_ALPHA, nullptr, 0);
        prof_end();
        const DistRound r3 = round();
        prof_begin(PC_UPD);
        if (vec2)
          update_kernel<0, 1><<<G, kBlock, 0, S>>>(N, x_dev, d.p, p.p, q.p, nullptr, nullptr, nullptr,
                                             ord, partial.p, pmx.p, pmr.p, tickets.p + TK_UPD,
                                             st.p, sp, r2.pull, r3.push);
        else
          update_kernel<0, 0><<<G, kBlock, 0, S>>>(N, x_dev, d.p, p.p, q.p, nullptr, nullptr, nullptr,
                                             ord, partial.p, pmx.p, pmr.p, tickets.p + TK_UPD,
                                             st.p, sp, r2.pull, r3.push);
        finalize_update(r3, 0);
        model_summary(0);
        prof_end();
        launches += 3;
      } else {
        const DistRound r1 = round();
        dot_kernel<<<G, kBlock, 0, S>>>(N, dhat.p, d.p, partial.p, tickets.p + TK_DOT, st.p,
                                  FIN_BCGS_RHO, nullptr, 1, r1.push);
        if (!fused) reduce_finalize(FIN_BCGS_RHO, nullptr, 1);
        bcgs_p_kernel<<<G, kBlock, 0, S>>>(N, d.p, v.p, p.p, st.p, first, r1.pull);
        prof_begin(PC_ILU);
        launches += precond(p.p, phat.p, S);
        halo_round(hp, hs, phat.p, false);
        prof_end();
        const DistRound r2 = round();
        prof_begin(PC_SPMV);
        launch_spmv_fused<0, 1>(*A, G, S, phat.p, v.p, nullptr, dhat.p, partial.p, tickets.p + TK_SPMV,
                            st.p, FIN_BCGS_ALPHA, 1, hs, r2.push);
        if (!fused) reduce_finalize(FIN_BCGS_ALPHA, nullptr, 1);
        prof_end();
        bcgs_q_kernel<<<G, kBlock, 0, S>>>(N, d.p, v.p, q.p, st.p, r2.pull);
        prof_begin(PC_ILU);
        launches += precond(q.p, qhat.p, S);
        halo_round(hp, hs, qhat.p, false);
        prof_end();
        const DistRound r3 = round();
        prof_begin(PC_SPMV);
        launch_spmv_fused<0, 2>(*A, G, S, qhat.p, t.p, nullptr, q.p, partial.p, tickets.p + TK_SPMV, st.p,
                            FIN_BCGS_OMEGA, 1, hs, r3.push);
        if (!fused) reduce_finalize(FIN_BCGS_OMEGA, nullptr, 1);
        prof_end();
        const DistRound r4 = round();
        if (vec2)
          update_kernel<1, 1><<<G, kBlock, 0, S>>>(N, x_dev, d.p, phat.p, q.p, qhat.p, t.p, dscale.p,
                                             ord, partial.p, pmx.p, pmr.p, tickets.p + TK_UPD,
                                             st.p, sp, r3.pull, r4.push);
        else
          update_kernel<1, 0><<<G, kBlock, 0, S>>>(N, x_dev, d.p, phat.p, q.p, qhat.p, t.p, dscale.p,
                                             ord, partial.p, pmx.p, pmr.p, tickets.p + TK_UPD,
                                             st.p, sp, r3.pull, r4.push);
        finalize_update(r4, 1);
        model_summary(1);
        launches += 6;
      }
    };
    const bool use_graph = !dist && !profiling && s.north == 0 && N <= graph_max_rows() && itmax > 1 &&
                           !std::getenv("MF6GPU_NO_GRAPH");
    cudaGraphExec_t gexec = nullptr;
    int graph_launches = 0;
    while (!finished && launched < itmax) {
      const int nb = std::min(batch, itmax - launched);
      for (int b = 0; b < nb; b++) {
        const int iiter = launched + b + 1;
        prof_on = (b == 0);  // time one iteration per polling batch
        if (use_graph && iiter > 1) {
          if (!gexec) {
            cudaGraph_t graph = nullptr;
            const int before = launches;
            if (!cap_stream) MF6_CK(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
            const cudaStream_t work = S;
            S = cap_stream;  // the launches of enqueue_iteration go to S
            MF6_CK(cudaStreamBeginCapture(S, cudaStreamCaptureModeThreadLocal));
            enqueue_iteration(0);
            const cudaError_t cap_rc = cudaStreamEndCapture(S, &graph);
            S = work;
            MF6_CK(cap_rc);
            graph_launches = launches - before;
            launches = before;
            MF6_CK(cudaGraphInstantiate(&gexec, graph, 0));
            MF6_CK(cudaGraphDestroy(graph));
          }
          MF6_CK(cudaGraphLaunch(gexec, S));
          launches += graph_launches;
        } else {
          enqueue_iteration(iiter == 1 ? 1 : 0);
        }
        if (s.north > 0 && ((iiter + 1) % s.north == 0)) {
          if (dist) halo->exchange(x_dev, S);
          launch_spmv_fused<1, 0>(*A, G, S, x_dev, d.p, b_dev, nullptr, nullptr, nullptr, st.p, 0, 1);
          launches++;
        }
      }
      launched += nb;
      MF6_CK(cudaGetLastError());
      MF6_CK(cudaMemcpyAsync(h_st.p, st.p, sizeof(KState), cudaMemcpyDeviceToHost, S));
      MF6_CK(cudaStreamSynchronize(S));
      if (h_st.p->done) finished = true;
      if (dist) comm_check(halo->comm);
      prof_on = true;
      if (profiling) {
        // the timed iteration is the first of the batch: it executed unless the loop had already ended
        if (h_st.p->iter - (launched - nb) < 1) ev_cls.clear();
        prof_collect();
      }
    }
    if (gexec) MF6_CK(cudaGraphExecDestroy(gexec));
    innerit = h_st.p->iter;
    icnvg = h_st.p->icnvg;
  }
  if (icnvg < 0) icnvg = 0;
  // -- unscale (ImsLinear.f90:740-745)
  if (s.iscl != 0) {
    scale_apply_kernel<<<G, kBlock, 0, S>>>(N, A->slice_ptr.p, A->rowlen.p, A->col.p, A->val.p,
                                            dscale.p, dscale2.p, x_dev, b_dev, 1);
    launches++;
  }
  MF6_CK(cudaEventRecord(ev[2], S));
  MF6_CK(cudaEventSynchronize(ev[2]));
  float ms01 = 0.f, ms12 = 0.f;
  MF6_CK(cudaEventElapsedTime(&ms01, ev[0], ev[1]));
  MF6_CK(cudaEventElapsedTime(&ms12, ev[1], ev[2]));
  t_factor = 1e-3 * ms01;
  t_krylov = 1e-3 * ms12;
  *iters = innerit;
  *icnvg_out = icnvg;
}

static void check_settings(mf6gpu_ims_settings &s) {
  // the same spirit as petsc_check_settings (PetscSolver.F90:123-154): options the
  // backend cannot honour are rejected (ILUT) or downgraded (reordering)
  MF6_REQUIRE(s.ilinmeth == 1 || s.ilinmeth == 2, "solver: LINEAR_ACCELERATION must be CG (1) or BICGSTAB (2)");
  MF6_REQUIRE(s.level >= 0 && s.droptol >= 0.0, "solver: PRECONDITIONER_LEVELS / DROP_TOLERANCE must be >= 0");
  MF6_REQUIRE(s.iscl >= 0 && s.iscl <= 2, "solver: SCALING_METHOD must be NONE (0), DIAGONAL (1) or L2NORM (2)");
  MF6_REQUIRE(s.relax >= 0.0 && s.relax <= 1.0, "solver: RELAXATION_FACTOR must be in [0,1]");
  MF6_REQUIRE(s.north >= 0, "solver: NUMBER_ORTHOGONALIZATIONS must be >= 0");
  MF6_REQUIRE(s.iter1 > 0, "solver: INNER_MAXIMUM must be > 0");
  if (s.iord != 0) s.iord = 0;  // REORDERING_METHOD ignored (own level-sorted ordering)
}

extern "C" {

int mf6gpu_solver_create(mf6gpu_matrix *m, const mf6gpu_ims_settings *settings,
                         int32_t summary_capacity, mf6gpu_solver **out) {
  return guard([&] {
    MF6_REQUIRE(m && settings && out, "solver_create: null argument");
    auto *s = new mf6gpu_solver();
    try {
      s->A = m;
      s->s = *settings;
      check_settings(s->s);
      // ImsLinear.f90:178-185: LEVEL > 0 or DROPTOL > 0 selects ILUT, RELAX > 0 the modified variants
      s->ipc = ((s->s.level > 0 || s->s.droptol > 0.0) ? 3 : 1) + ((s->s.relax > 0.0) ? 1 : 0);
      s->n = m->n;
      s->stream = m->stream;
      const size_t n = (size_t)m->n_ext;  // vectors carry the halo region behind the owned rows
      if (s->ipc >= 3) {
        s->ilut.reset(new mf6::IlutPlan());
        s->ilut->build(*m, s->s.level, s->s.droptol);
      } else {
        s->lu.alloc_zero((size_t)m->nslots);
      }
      s->x.alloc_zero(n);
      s->b.alloc_zero(n);
      s->d.alloc_zero(n);
      s->p.alloc_zero(n);
      s->q.alloc_zero(n);
      s->z.alloc_zero(n);
      if (s->s.ilinmeth == 2) {
        s->t.alloc_zero(n);
        s->v.alloc_zero(n);
        s->dhat.alloc_zero(n);
        s->phat.alloc_zero(n);
        s->qhat.alloc_zero(n);
      }
      if (s->s.iscl != 0) {
        s->dscale.alloc_zero(n);
        s->dscale2.alloc_zero(n);
      }
      s->hx.alloc(n);
      s->hb.alloc(n);
      s->st.alloc_zero(1);
      s->partial.alloc_zero(4 * (size_t)kMaxBlocks);
      s->ilu_partial.alloc_zero((n / kBlock + (size_t)m->nlevels + 2) * (kBlock / 32) + mf6::ilu0_block_dot_slots(*m));
      s->pmx.alloc_zero((size_t)kMaxBlocks);
      s->pmr.alloc_zero((size_t)kMaxBlocks);
      s->tickets.alloc_zero(8);
      s->failflag.alloc_zero(1);
      s->red_all.alloc_zero(8 * 64);
      s->sum_cap = summary_capacity > 0 ? summary_capacity : 0;
      const size_t c = (size_t)(s->sum_cap > 0 ? s->sum_cap : 1);
      s->sum_itinner.alloc_zero(c);
      s->sum_locdv.alloc_zero(c);
      s->sum_locr.alloc_zero(c);
      s->sum_dvmax.alloc_zero(c);
      s->sum_rmax.alloc_zero(c);
      s->sum_alpha.alloc_zero(c);
      s->sum_omega.alloc_zero(c);
      s->h_st.alloc(1);
      s->h_flag.alloc(4);
      for (auto &e : s->ev) MF6_CK(cudaEventCreate(&e));
    } catch (...) {
      delete s;
      throw;
    }
    *out = s;
  });
}

int mf6gpu_solver_destroy(mf6gpu_solver *s) {
  return guard([&] {
    if (!s) return;
    for (auto &e : s->ev)
      if (e) cudaEventDestroy(e);
    for (auto &e : s->ev_pool) cudaEventDestroy(e);
    if (s->cap_stream) cudaStreamDestroy(s->cap_stream);
    delete s;
  });
}

int mf6gpu_solver_solve(mf6gpu_solver *s, int32_t kiter, int32_t kstp, const double *rhs,
                        double *x, int32_t *iteration_number, int32_t *is_converged) {
  return guard([&] {
    MF6_REQUIRE(s && rhs && x && iteration_number && is_converged, "solver_solve: null argument");
    const size_t nb = sizeof(double) * (size_t)s->n;
    cudaStream_t S = s->stream;
    MF6_CK(cudaMemcpyAsync(s->hx.p, x, nb, cudaMemcpyHostToDevice, S));
    MF6_CK(cudaMemcpyAsync(s->hb.p, rhs, nb, cudaMemcpyHostToDevice, S));
    launch_gather(s->n, s->A->d_perm.p, s->hx.p, s->x.p, S);
    launch_gather(s->n, s->A->d_perm.p, s->hb.p, s->b.p, S);
    int it = 0, cv = 0;
    s->solve_device(kiter, kstp, s->x.p, s->b.p, &it, &cv);
    launch_scatter(s->n, s->A->d_perm.p, s->x.p, s->hx.p, S);
    MF6_CK(cudaGetLastError());
    MF6_CK(cudaMemcpyAsync(x, s->hx.p, nb, cudaMemcpyDeviceToHost, S));
    MF6_CK(cudaStreamSynchronize(S));
    *iteration_number = it;
    *is_converged = cv;
  });
}

int mf6gpu_solver_get_summary(mf6gpu_solver *s, int32_t cap, int32_t *itinner, double *dvmax,
                              int32_t *locdv, double *rmax, int32_t *locr, double *alpha,
                              double *omega) {
  int count = 0;
  int rc = guard([&] {
    MF6_REQUIRE(s, "solver_get_summary: null argument");
    MF6_CK(cudaMemcpy(s->h_st.p, s->st.p, sizeof(KState), cudaMemcpyDeviceToHost));
    count = std::min(std::min(s->h_st.p->sum_count, s->sum_cap), (int)cap);
    if (count <= 0) {
      count = 0;
      return;
    }
    const size_t c = (size_t)count;
    std::vector<int> li(c);
    if (itinner) s->sum_itinner.download(itinner, c);
    if (dvmax) s->sum_dvmax.download(dvmax, c);
    if (rmax) s->sum_rmax.download(rmax, c);
    if (alpha) s->sum_alpha.download(alpha, c);
    if (omega) s->sum_omega.download(omega, c);
    if (locdv) {
      s->sum_locdv.download(li.data(), c);
      const bool dist = s->halo && s->halo->active();
      for (size_t i = 0; i < c; i++) locdv[i] = li[i] >= 0 ? (dist ? li[i] : s->A->perm[li[i]]) + 1 : 0;
    }
    if (locr) {
      s->sum_locr.download(li.data(), c);
      const bool dist = s->halo && s->halo->active();
      for (size_t i = 0; i < c; i++) locr[i] = li[i] >= 0 ? (dist ? li[i] : s->A->perm[li[i]]) + 1 : 0;
    }
  });
  return rc < 0 ? rc : count;
}

// convmodstart [nmod + 1]: first row of every model in the solution's row numbering, ascending, last = n
int mf6gpu_solver_set_models(mf6gpu_solver *s, int32_t nmod, const int32_t *convmodstart, int32_t index_base) {
  return guard([&] {
    MF6_REQUIRE(s && convmodstart && nmod >= 1 && nmod <= 1024, "solver_set_models: bad argument");
    MF6_REQUIRE(convmodstart[0] - index_base == 0 && convmodstart[nmod] - index_base == s->n,
                "solver_set_models: convmodstart must span all rows");
    std::vector<int> mid((size_t)s->n);
    for (int im = 0; im < nmod; im++) {
      MF6_REQUIRE(convmodstart[im + 1] >= convmodstart[im], "solver_set_models: convmodstart must ascend");
      for (int r = convmodstart[im] - index_base; r < convmodstart[im + 1] - index_base; r++)
        mid[(size_t)s->A->iperm[r]] = im;
    }
    s->nmod = nmod;
    s->modid.upload(mid);
    const size_t c = (size_t)std::max(s->sum_cap, 1) * (size_t)nmod;
    s->msum_dvmax.alloc_zero(c);
    s->msum_rmax.alloc_zero(c);
    s->msum_locdv.alloc_zero(c);
    s->msum_locr.alloc_zero(c);
    s->mpmx.alloc_zero((size_t)nmod * (size_t)kMaxBlocks);
    s->mpmr.alloc_zero((size_t)nmod * (size_t)kMaxBlocks);
    s->mtickets.alloc_zero((size_t)nmod);
  });
}

// per-model records of the iterations recorded so far: arrays [cap * nmod] laid out like the Fortran
// convdvmax(nmod, niter) (model index fastest); locations 1-based original rows, 0 = none.  Returns the count.
int mf6gpu_solver_get_model_summary(mf6gpu_solver *s, int32_t cap, double *convdvmax, int32_t *convlocdv,
                                    double *convrmax, int32_t *convlocr) {
  int count = 0;
  int rc = guard([&] {
    MF6_REQUIRE(s && s->nmod >= 1, "solver_get_model_summary: call mf6gpu_solver_set_models first");
    MF6_CK(cudaMemcpy(s->h_st.p, s->st.p, sizeof(KState), cudaMemcpyDeviceToHost));
    count = std::min(std::min(s->h_st.p->sum_count, s->sum_cap), (int)cap);
    if (count <= 0) {
      count = 0;
      return;
    }
    const size_t c = (size_t)count * (size_t)s->nmod;
    if (s->nmod == 1) {  // one model: its records are the solution-wide ones
      std::vector<int> li((size_t)count);
      if (convdvmax) s->sum_dvmax.download(convdvmax, (size_t)count);
      if (convrmax) s->sum_rmax.download(convrmax, (size_t)count);
      if (convlocdv) {
        s->sum_locdv.download(li.data(), (size_t)count);
        for (int i = 0; i < count; i++) convlocdv[i] = li[i] >= 0 ? s->A->perm[li[i]] + 1 : 0;
      }
      if (convlocr) {
        s->sum_locr.download(li.data(), (size_t)count);
        for (int i = 0; i < count; i++) convlocr[i] = li[i] >= 0 ? s->A->perm[li[i]] + 1 : 0;
      }
      return;
    }
    std::vector<int> li(c);
    if (convdvmax) s->msum_dvmax.download(convdvmax, c);
    if (convrmax) s->msum_rmax.download(convrmax, c);
    if (convlocdv) {
      s->msum_locdv.download(li.data(), c);
      for (size_t i = 0; i < c; i++) convlocdv[i] = li[i] >= 0 ? s->A->perm[li[i]] + 1 : 0;
    }
    if (convlocr) {
      s->msum_locr.download(li.data(), c);
      for (size_t i = 0; i < c; i++) convlocr[i] = li[i] >= 0 ? s->A->perm[li[i]] + 1 : 0;
    }
  });
  return rc < 0 ? rc : count;
}

double mf6gpu_solver_stat(const mf6gpu_solver *s, int what) {
  if (!s) return -1.0;
  switch (what) {
    case 0: return s->l2norm0;
    case 1: return (double)s->npivfix;
    case 2: return s->t_factor;
    case 3: return s->t_krylov;
    case 4: return (double)s->launches;
    case 5: return s->fused_exchange ? 1.0 : 0.0;
  }
  return -1.0;
}

int mf6gpu_solver_profile(mf6gpu_solver *s, int32_t enable) {
  return guard([&] {
    MF6_REQUIRE(s, "solver_profile: null argument");
    s->profiling = enable != 0;
    for (int c = 0; c < mf6gpu_solver::PC_N; c++) {
      s->prof_ms[c] = 0.0;
      s->prof_cnt[c] = 0;
    }
    s->ev_cls.clear();
    s->ev_used = 0;
  });
}

int mf6gpu_solver_profile_get(mf6gpu_solver *s, int32_t cls, double *total_ms, int64_t *count) {
  return guard([&] {
    MF6_REQUIRE(s && cls >= 0 && cls < mf6gpu_solver::PC_N && total_ms && count, "solver_profile_get: bad argument");
    *total_ms = s->prof_ms[cls];
    *count = s->prof_cnt[cls];
  });
}

int mf6gpu_solver_factor(mf6gpu_solver *s, int32_t *npivot_fixes) {
  return guard([&] {
    MF6_REQUIRE(s, "solver_factor: null argument");
    int c = s->factor();
    if (npivot_fixes) *npivot_fixes = c;
  });
}

int mf6gpu_solver_apply_preconditioner(mf6gpu_solver *s, const double *r, double *z) {
  return guard([&] {
    MF6_REQUIRE(s && r && z, "solver_apply_preconditioner: null argument");
    const size_t nb = sizeof(double) * (size_t)s->n;
    cudaStream_t S = s->stream;
    MF6_CK(cudaMemcpyAsync(s->hb.p, r, nb, cudaMemcpyHostToDevice, S));
    launch_gather(s->n, s->A->d_perm.p, s->hb.p, s->d.p, S);
    if (s->ilut)
      s->ilut->apply(s->d.p, s->z.p, nullptr, S);
    else
      ilu0_apply(*s->A, s->lu.p, s->d.p, s->z.p, nullptr, S);
    launch_scatter(s->n, s->A->d_perm.p, s->z.p, s->hx.p, S);
    MF6_CK(cudaGetLastError());
    MF6_CK(cudaMemcpyAsync(z, s->hx.p, nb, cudaMemcpyDeviceToHost, S));
    MF6_CK(cudaStreamSynchronize(S));
  });
}

}  // extern "C"
