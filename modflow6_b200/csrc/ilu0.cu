// ilu0.cu -- level-scheduled ILU0 / MILU0 factorisation and triangular solves.
//
// Restates on the device (one thread per row, rows of one dependency level per
// launch; arithmetic inside a row in the reference's order):
//   ims_base_pcilu0  src/Solution/LinearMethods/ImsLinearBase.f90:928-1042
//   ims_base_ilu0a   src/Solution/LinearMethods/ImsLinearBase.f90:1049-1092
// The factor shares the matrix's SELL-32 structure (matrix.cuh): slot 0 holds
// the inverse pivot APC(n), lower slots the L multipliers, upper slots U.
#include "ilu0.cuh"

namespace mf6 {

template <int MAXLEN>
__global__ void __launch_bounds__(128)
ilu0_factor_level_kernel(int r0, int r1, const int *__restrict__ slice_ptr,
                         const unsigned char *__restrict__ rowlen,
                         const unsigned char *__restrict__ nlow,
                         const int *__restrict__ col, const double *__restrict__ aval,
                         double *__restrict__ lu, double relax, double delta, int ipcflag,
                         int *__restrict__ failflag) {
  const int r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= r1) return;
  const int len = rowlen[r], lo = nlow[r];
  const long long base = (long long)slice_ptr[r >> 5] + (r & 31);
  double w[MAXLEN];
  int c[MAXLEN];
  for (int k = 0; k < len; k++) {
    w[k] = aval[base + 32LL * k];
    c[k] = col[base + 32LL * k];
  }
  double rs = 0.0;
  for (int j = 1; j <= lo; j++) {
    const int jcol = c[j];
    const long long jb = (long long)slice_ptr[jcol >> 5] + (jcol & 31);
    const int jlen = rowlen[jcol], jlo = nlow[jcol];
    const double tl = w[j] * lu[jb];
    w[j] = tl;
    for (int jj = jlo + 1; jj < jlen; jj++) {
      const int jjcol = col[jb + 32LL * jj];
      const double u = lu[jb + 32LL * jj];
      int pos = -1;
      for (int k = 0; k < len; k++)
        if (c[k] == jjcol) pos = k;
      if (pos >= 0)
        w[pos] = w[pos] - tl * u;
      else
        rs = rs + tl * u;
    }
  }
  const double d = w[0];
  double tl = (1.0 + delta) * d - (relax * rs);
  const double sd1 = copysign(fabs(d), tl);
  bool bad = false;
  if (sd1 != d) {
    if (ipcflag > 1)
      tl = copysign(1.0e-6, d);
    else
      bad = true;
  }
  if (fabs(tl) == 0.0) {
    if (ipcflag > 1)
      tl = copysign(1.0e-6, d);
    else
      bad = true;
  }
  if (bad) {
    *failflag = 1;               // reference: IPCFLAG = 1 ; EXIT MAIN (result discarded)
    tl = copysign(1.0e-6, d);    // keep later rows finite
    if (tl == 0.0) tl = 1.0e-6;
  }
  lu[base] = 1.0 / tl;
  for (int k = 1; k < len; k++) lu[base + 32LL * k] = w[k];
}

// forward sweep of one level: d(n) = r(n) - sum_lower APC(j) d(col)
__global__ void __launch_bounds__(kBlock)
ilu0_fwd_level_kernel(int r0, int r1, const int *__restrict__ slice_ptr,
                      const unsigned char *__restrict__ nlow, const int *__restrict__ col,
                      const double *__restrict__ lu, const double *__restrict__ rin,
                      double *__restrict__ d, const int *__restrict__ done) {
  const int r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= r1) return;
  if (done && *done) return;
  const int lo = nlow[r];
  const long long base = (long long)slice_ptr[r >> 5] + (r & 31);
  double tv = rin[r];
  for (int k0 = 1; k0 <= lo; k0 += 4) {
    double v[4], dv[4];
    int c[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const bool ok = (k0 + u) <= lo;
      const long long p = base + 32LL * (k0 + u);
      v[u] = ok ? __ldg(lu + p) : 0.0;
      c[u] = ok ? __ldg(col + p) : r;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) dv[u] = ((k0 + u) <= lo) ? d[c[u]] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; u++)
      if ((k0 + u) <= lo) tv = tv - v[u] * dv[u];
  }
  d[r] = tv;
}

// backward sweep of one level: d(n) = (d(n) - sum_upper APC(j) d(col)) * APC(n)
__global__ void __launch_bounds__(kBlock)
ilu0_bwd_level_kernel(int r0, int r1, const int *__restrict__ slice_ptr,
                      const unsigned char *__restrict__ rowlen,
                      const unsigned char *__restrict__ nlow, const int *__restrict__ col,
                      const double *__restrict__ lu, double *__restrict__ d,
                      const int *__restrict__ done) {
  const int r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= r1) return;
  if (done && *done) return;
  const int len = rowlen[r], lo = nlow[r];
  const long long base = (long long)slice_ptr[r >> 5] + (r & 31);
  double tv = d[r];
  const double piv = __ldg(lu + base);
  for (int k0 = lo + 1; k0 < len; k0 += 4) {
    double v[4], dv[4];
    int c[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const bool ok = (k0 + u) < len;
      const long long p = base + 32LL * (k0 + u);
      v[u] = ok ? __ldg(lu + p) : 0.0;
      c[u] = ok ? __ldg(col + p) : r;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) dv[u] = ((k0 + u) < len) ? d[c[u]] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; u++)
      if ((k0 + u) < len) tv = tv - v[u] * dv[u];
  }
  d[r] = tv * piv;
}

int ilu0_factor(const mf6gpu_matrix &A, const double *aval, double *lu, double relax,
                double delta, int ipcflag, int *d_failflag, cudaStream_t s) {
  MF6_REQUIRE(A.maxlen <= 64, "ILU0: more than 64 entries in a row is not supported");
  int launches = 0;
  for (int l = 0; l < A.nlevels; l++) {
    const int r0 = A.level_ptr[l], r1 = A.level_ptr[l + 1];
    if (r1 <= r0) continue;
    const int blocks = (r1 - r0 + 127) / 128;
    if (A.maxlen <= 8)
      ilu0_factor_level_kernel<8><<<blocks, 128, 0, s>>>(r0, r1, A.slice_ptr.p, A.rowlen.p, A.nlow.p,
                                                         A.col.p, aval, lu, relax, delta, ipcflag, d_failflag);
    else if (A.maxlen <= 16)
      ilu0_factor_level_kernel<16><<<blocks, 128, 0, s>>>(r0, r1, A.slice_ptr.p, A.rowlen.p, A.nlow.p,
                                                          A.col.p, aval, lu, relax, delta, ipcflag, d_failflag);
    else if (A.maxlen <= 32)
      ilu0_factor_level_kernel<32><<<blocks, 128, 0, s>>>(r0, r1, A.slice_ptr.p, A.rowlen.p, A.nlow.p,
                                                          A.col.p, aval, lu, relax, delta, ipcflag, d_failflag);
    else
      ilu0_factor_level_kernel<64><<<blocks, 128, 0, s>>>(r0, r1, A.slice_ptr.p, A.rowlen.p, A.nlow.p,
                                                          A.col.p, aval, lu, relax, delta, ipcflag, d_failflag);
    launches++;
  }
  return launches;
}

int ilu0_apply(const mf6gpu_matrix &A, const double *lu, const double *rin, double *d,
               const int *done, cudaStream_t s) {
  int launches = 0;
  for (int l = 0; l < A.nlevels; l++) {
    const int r0 = A.level_ptr[l], r1 = A.level_ptr[l + 1];
    if (r1 <= r0) continue;
    ilu0_fwd_level_kernel<<<(r1 - r0 + kBlock - 1) / kBlock, kBlock, 0, s>>>(
        r0, r1, A.slice_ptr.p, A.nlow.p, A.col.p, lu, rin, d, done);
    launches++;
  }
  for (int l = A.nlevels - 1; l >= 0; l--) {
    const int r0 = A.level_ptr[l], r1 = A.level_ptr[l + 1];
    if (r1 <= r0) continue;
    ilu0_bwd_level_kernel<<<(r1 - r0 + kBlock - 1) / kBlock, kBlock, 0, s>>>(
        r0, r1, A.slice_ptr.p, A.rowlen.p, A.nlow.p, A.col.p, lu, d, done);
    launches++;
  }
  return launches;
}

}  // namespace mf6
