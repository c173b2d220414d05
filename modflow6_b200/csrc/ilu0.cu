// ilu0.cu -- level-scheduled ILU0 / MILU0 factorisation and triangular solves.
//
// Restates on the device (one thread per row, rows of one dependency level per
// launch; arithmetic inside a row in the reference's order):
//   ims_base_pcilu0  src/Solution/LinearMethods/ImsLinearBase.f90:928-1042
//   ims_base_ilu0a   src/Solution/LinearMethods/ImsLinearBase.f90:1049-1092
// The factor shares the matrix's SELL-32 structure (matrix.cuh): slot 0 holds
// the inverse pivot APC(n), lower slots the L multipliers, upper slots U.
#include "ilu0.cuh"
#include <algorithm>

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

// ---- triangular solves ---------------------------------------------------------
// Level-scheduled ims_base_ilu0a with three fusions that remove whole vector passes:
//  * rows of level 0 have no lower entries, so their forward value is r itself: the
//    level-0 forward launch is elided and readers take rin[col] for such neighbours;
//  * rows of the LAST level have no upper entries, so their backward step is only the
//    multiplication by the inverse pivot: it is fused into their forward launch;
//  * (CG) rho = r.z is accumulated by the launches that finalise z (IluDot).
struct IluDot {
  double *partial;  // this launch's partial slots (one per WARP), nullptr = no dot
};

// per-warp partial of rho: no block-level barrier, so CTAs retire as soon as their rows are done
__device__ __forceinline__ void ilu_dot_finish(double s, const IluDot &D) {
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) D.partial[blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5)] = s;
}

// rho = sum of the per-warp partials of every finalising launch: fixed chunking and order
// (deterministic); the last CTA to finish combines the per-CTA sums and writes rho, beta
__global__ void __launch_bounds__(kBlock)
ilu_dot_reduce_kernel(int nslots, const double *__restrict__ partial, double *__restrict__ cta_sums,
                      unsigned int *ticket, double *rho_out, double *beta_out, const double *rho0,
                      const int *__restrict__ done) {
  __shared__ double sh[8];
  __shared__ bool last;
  if (done && *done) return;
  const int chunk = (nslots + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * chunk, i1 = min(nslots, i0 + chunk);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int i = i0 + threadIdx.x;
  for (; i + 3 * kBlock < i1; i += 4 * kBlock) {
    a0 += partial[i];
    a1 += partial[i + kBlock];
    a2 += partial[i + 2 * kBlock];
    a3 += partial[i + 3 * kBlock];
  }
  for (; i < i1; i += kBlock) a0 += partial[i];
  double a = block_sum((a0 + a1) + (a2 + a3), sh);
  if (threadIdx.x == 0) cta_sums[blockIdx.x] = a;
  if (last_block(ticket, &last)) {
    double t = (threadIdx.x < gridDim.x) ? cta_sums[threadIdx.x] : 0.0;
    t = block_sum(t, sh);
    if (threadIdx.x == 0) {
      *rho_out = t;
      *beta_out = t / *rho0;
    }
  }
}

// forward sweep of one level (l >= 1): d(n) = r(n) - sum_lower APC(j) d(col)
// FINAL: the level is the last one -> d(n) = (...) * APC(n) is already the result
template <bool FINAL>
__global__ void __launch_bounds__(kBlock, 8)
ilu0_fwd_level_kernel(int r0, int r1, int lvl1_start, const int *__restrict__ slice_ptr,
                      const unsigned char *__restrict__ nlow, const int *__restrict__ col,
                      const double *__restrict__ lu, const double *__restrict__ rin,
                      double *__restrict__ d, const int *__restrict__ done, IluDot D) {
  if (done && *done) return;
  double dot = 0.0;
  const int r = r0 + blockIdx.x * blockDim.x + threadIdx.x;  // one row per thread: every row's loads in flight at once
  if (r < r1) {
    const int lo = nlow[r];
    const long long base = (long long)slice_ptr[r >> 5] + (r & 31);
    const double rr = rin[r];
    double tv = rr;
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
      for (int u = 0; u < 4; u++) {
        dv[u] = 0.0;
        if ((k0 + u) <= lo) dv[u] = (c[u] < lvl1_start) ? rin[c[u]] : d[c[u]];
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if ((k0 + u) <= lo) tv = tv - v[u] * dv[u];
    }
    if (FINAL) {
      tv = tv * __ldg(lu + base);
      dot += rr * tv;
    }
    d[r] = tv;
  }
  if (FINAL && D.partial) ilu_dot_finish(dot, D);
}

// backward sweep of one level: d(n) = (d(n) - sum_upper APC(j) d(col)) * APC(n)
// LEVEL0: the forward value of the row is rin(n) itself
template <bool LEVEL0>
__global__ void __launch_bounds__(kBlock, 8)
ilu0_bwd_level_kernel(int r0, int r1, const int *__restrict__ slice_ptr,
                      const unsigned char *__restrict__ rowlen,
                      const unsigned char *__restrict__ nlow, const int *__restrict__ col,
                      const double *__restrict__ lu, const double *__restrict__ rin,
                      double *__restrict__ d, const int *__restrict__ done, IluDot D) {
  if (done && *done) return;
  double dot = 0.0;
  const int r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (r < r1) {
    const int len = rowlen[r], lo = nlow[r];
    const long long base = (long long)slice_ptr[r >> 5] + (r & 31);
    const double rr = rin[r];
    double tv = LEVEL0 ? rr : d[r];
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
    tv = tv * piv;
    d[r] = tv;
    dot += rr * tv;
  }
  if (D.partial) ilu_dot_finish(dot, D);
}

// ---- two-level (bipartite multicolour) fast path on fixed-width SELL ---------------------------
// With exactly two levels every off-diagonal of a level-1 row is a lower entry and every
// off-diagonal of a level-0 row an upper entry, so each sweep is a full-row product without
// nlow / rowlen / slice_ptr loads.  Padding and halo slots hold lu = 0 and only add -0*finite.
//   PHASE 0 (rows of level 1):  z = (r - sum_k lu_k r[col_k]) * piv     (forward + its trivial backward)
//   PHASE 1 (rows of level 0):  z = (r - sum_k lu_k z[col_k]) * piv     (backward; forward value = r)
template <int W, int PHASE>
__global__ void __launch_bounds__(kBlock, 8)
ilu0_two_level_kernel(int r0, int r1, int ncols, const int *__restrict__ col,
                      const int *__restrict__ soff, const double *__restrict__ lu,
                      const double *__restrict__ rin, double *__restrict__ d,
                      const int *__restrict__ done, IluDot D) {
  if (done && *done) return;
  double dot = 0.0;
  const int r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (r < r1) {
    const long long base = (long long)(r >> 5) * (32 * W) + (r & 31);
    double v[W], dv[W];
    int c[W];
#pragma unroll
    for (int u = 0; u < W; u++) v[u] = __ldg(lu + base + 32 * u);
    if (soff) {  // stencil-compressed columns (matrix.cuh)
      const int *so = soff + (r >> 5) * W;
#pragma unroll
      for (int u = 1; u < W; u++) {
        const int o = __ldg(so + u);
        if (o != INT_MIN)
          c[u] = min(max(r + o, 0), ncols - 1);
        else
          c[u] = __ldg(col + base + 32 * u);
      }
    } else {
#pragma unroll
      for (int u = 1; u < W; u++) c[u] = __ldg(col + base + 32 * u);
    }
    const double rr = rin[r];
#pragma unroll
    for (int u = 1; u < W; u++) dv[u] = (PHASE == 0) ? rin[c[u]] : d[c[u]];
    double tv = rr;
#pragma unroll
    for (int u = 1; u < W; u++) tv = tv - v[u] * dv[u];
    tv = tv * v[0];
    d[r] = tv;
    dot = rr * tv;
  }
  if (D.partial) ilu_dot_finish(dot, D);
}

template <int W>
static int launch_two_level(const mf6gpu_matrix &A, const double *lu, const double *rin, double *d,
                            const int *done, cudaStream_t s, const IluDotArgs *dot) {
  const int lvl1 = A.level_ptr[1];
  const int gA = (A.n - lvl1 + kBlock - 1) / kBlock, gB = (lvl1 + kBlock - 1) / kBlock;
  IluDot DA{dot ? dot->partial : nullptr};
  IluDot DB{dot ? dot->partial + gA * (kBlock / 32) : nullptr};
  const int *so = A.slot_off.n ? A.slot_off.p : nullptr;
  ilu0_two_level_kernel<W, 0><<<gA, kBlock, 0, s>>>(lvl1, A.n, A.n_ext, A.col.p, so, lu, rin, d, done, DA);
  ilu0_two_level_kernel<W, 1><<<gB, kBlock, 0, s>>>(0, lvl1, A.n_ext, A.col.p, so, lu, rin, d, done, DB);
  int launches = 2;
  if (dot) {
    const int slot = (gA + gB) * (kBlock / 32);
    const int rb = std::max(1, std::min(148, slot / (4 * kBlock)));
    ilu_dot_reduce_kernel<<<rb, kBlock, 0, s>>>(slot, dot->partial, dot->cta_sums, dot->ticket, dot->rho_out,
                                                dot->beta_out, dot->rho0, done);
    launches++;
  }
  return launches;
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
      ilu0_factor_level_kernel<8><<<blocks, 128, 0, s>>>(r0, r1, A.slice_ptr.p, A.rowlen_loc(), A.nlow.p,
                                                         A.col.p, aval, lu, relax, delta, ipcflag, d_failflag);
    else if (A.maxlen <= 16)
      ilu0_factor_level_kernel<16><<<blocks, 128, 0, s>>>(r0, r1, A.slice_ptr.p, A.rowlen_loc(), A.nlow.p,
                                                          A.col.p, aval, lu, relax, delta, ipcflag, d_failflag);
    else if (A.maxlen <= 32)
      ilu0_factor_level_kernel<32><<<blocks, 128, 0, s>>>(r0, r1, A.slice_ptr.p, A.rowlen_loc(), A.nlow.p,
                                                          A.col.p, aval, lu, relax, delta, ipcflag, d_failflag);
    else
      ilu0_factor_level_kernel<64><<<blocks, 128, 0, s>>>(r0, r1, A.slice_ptr.p, A.rowlen_loc(), A.nlow.p,
                                                          A.col.p, aval, lu, relax, delta, ipcflag, d_failflag);
    launches++;
  }
  return launches;
}

int ilu0_apply(const mf6gpu_matrix &A, const double *lu, const double *rin, double *d,
               const int *done, cudaStream_t s, const IluDotArgs *dot) {
  int launches = 0;
  const int L = A.nlevels;
  if (L == 2 && A.level_ptr[1] > 0 && A.level_ptr[1] < A.n) {
    switch (A.uniform_w) {
      case 4: return launch_two_level<4>(A, lu, rin, d, done, s, dot);
      case 5: return launch_two_level<5>(A, lu, rin, d, done, s, dot);
      case 6: return launch_two_level<6>(A, lu, rin, d, done, s, dot);
      case 7: return launch_two_level<7>(A, lu, rin, d, done, s, dot);
      case 8: return launch_two_level<8>(A, lu, rin, d, done, s, dot);
      case 9: return launch_two_level<9>(A, lu, rin, d, done, s, dot);
      case 10: return launch_two_level<10>(A, lu, rin, d, done, s, dot);
      default: break;
    }
  }
  const int lvl1 = (L > 1) ? A.level_ptr[1] : A.n;
  int slot = 0;  // next free partial slot
  // forward: levels 1 .. L-1 (level 0 is elided); the last one also finalises its rows
  for (int l = 1; l < L; l++) {
    const int r0 = A.level_ptr[l], r1 = A.level_ptr[l + 1];
    if (r1 <= r0) continue;
    const int g = (r1 - r0 + kBlock - 1) / kBlock;
    if (l == L - 1) {
      IluDot D{dot ? dot->partial + slot : nullptr};
      slot += g * (kBlock / 32);
      ilu0_fwd_level_kernel<true><<<g, kBlock, 0, s>>>(r0, r1, lvl1, A.slice_ptr.p, A.nlow.p, A.col.p, lu,
                                                        rin, d, done, D);
    } else {
      ilu0_fwd_level_kernel<false><<<g, kBlock, 0, s>>>(r0, r1, lvl1, A.slice_ptr.p, A.nlow.p, A.col.p, lu,
                                                         rin, d, done, IluDot{nullptr});
    }
    launches++;
  }
  // backward: levels L-2 .. 0 (L == 1: the single level is level 0)
  for (int l = (L > 1 ? L - 2 : 0); l >= 0; l--) {
    const int r0 = A.level_ptr[l], r1 = A.level_ptr[l + 1];
    if (r1 <= r0) continue;
    const int g = (r1 - r0 + kBlock - 1) / kBlock;
    IluDot D{dot ? dot->partial + slot : nullptr};
    slot += g * (kBlock / 32);
    if (l == 0)
      ilu0_bwd_level_kernel<true><<<g, kBlock, 0, s>>>(r0, r1, A.slice_ptr.p, A.rowlen_loc(), A.nlow.p, A.col.p,
                                                        lu, rin, d, done, D);
    else
      ilu0_bwd_level_kernel<false><<<g, kBlock, 0, s>>>(r0, r1, A.slice_ptr.p, A.rowlen_loc(), A.nlow.p, A.col.p,
                                                         lu, rin, d, done, D);
    launches++;
  }
  if (dot) {
    const int rb = std::max(1, std::min(148, slot / (4 * kBlock)));
    ilu_dot_reduce_kernel<<<rb, kBlock, 0, s>>>(slot, dot->partial, dot->cta_sums, dot->ticket, dot->rho_out,
                                                dot->beta_out, dot->rho0, done);
    launches++;
  }
  return launches;
}

}  // namespace mf6
