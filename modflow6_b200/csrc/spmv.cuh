// spmv.cuh -- the SELL-32 row product shared by every kernel that needs (A x)_row.
#pragma once
#include "common.cuh"

namespace mf6 {
#ifdef __CUDACC__
// (A x)_row with the accumulation order of amux (sparsekit.f90:44-57): slot 0
// (diagonal) first, then the row's remaining entries in storage order.
// Loads of 8 slots are issued before use to keep many requests in flight.
__device__ __forceinline__ double sell_row_dot(int row, const int *__restrict__ slice_ptr,
                                               const unsigned char *__restrict__ rowlen,
                                               const int *__restrict__ col,
                                               const double *__restrict__ val,
                                               const double *__restrict__ x) {
  const int len = rowlen[row];
  const long long base = (long long)slice_ptr[row >> 5] + (row & 31);
  double t = 0.0;
  for (int k0 = 0; k0 < len; k0 += 8) {
    double v[8], xv[8];
    int c[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const bool ok = (k0 + u) < len;
      const long long p = base + (long long)(k0 + u) * 32;
      v[u] = ok ? __ldg(val + p) : 0.0;
      c[u] = ok ? __ldg(col + p) : row;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) xv[u] = x[c[u]];
#pragma unroll
    for (int u = 0; u < 8; u++)
      if ((k0 + u) < len) t = t + v[u] * xv[u];
  }
  return t;
}

// Uniform-width variant: every SELL slice has the same width W (true for structured DIS grids and
// regular DISV tilings), so the slot address needs neither slice_ptr nor rowlen and all 2W loads of
// a row are issued at once.  Padding slots hold val = 0, col = own row: they add +0 at the END of the
// row sum, so the result is still bit-identical to amux.
// soff (may be null): stencil table, see matrix.cuh -- slots whose column is row + constant for the
// whole slice take the column from the table (one broadcast load per warp) instead of the col array.
template <int W>
__device__ __forceinline__ double sell_row_dot_w(int row, const int *__restrict__ col,
                                                 const double *__restrict__ val,
                                                 const double *__restrict__ x,
                                                 const int *__restrict__ soff, int ncols) {
  const long long base = (long long)(row >> 5) * (32 * W) + (row & 31);
  double v[W], xv[W];
  int c[W];
#pragma unroll
  for (int u = 0; u < W; u++) v[u] = __ldg(val + base + 32 * u);
  if (soff) {
    const int *so = soff + (row >> 5) * W;
#pragma unroll
    for (int u = 0; u < W; u++) {
      const int o = __ldg(so + u);
      if (o != INT_MIN)
        c[u] = min(max(row + o, 0), ncols - 1);
      else
        c[u] = __ldg(col + base + 32 * u);
    }
  } else {
#pragma unroll
    for (int u = 0; u < W; u++) c[u] = __ldg(col + base + 32 * u);
  }
#pragma unroll
  for (int u = 0; u < W; u++) xv[u] = x[c[u]];
  double t = 0.0;
#pragma unroll
  for (int u = 0; u < W; u++) t = t + v[u] * xv[u];
  return t;
}
#endif
}  // namespace mf6
