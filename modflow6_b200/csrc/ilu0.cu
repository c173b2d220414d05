// ilu0.cu -- level-scheduled ILU0 / MILU0 factorisation and triangular solves.
//
// Restates on the device (one thread per row, rows of one dependency level per
// launch; arithmetic inside a row in the reference's order):
//   ims_base_pcilu0  src/Solution/LinearMethods/ImsLinearBase.f90:928-1042
//   ims_base_ilu0a   src/Solution/LinearMethods/ImsLinearBase.f90:1049-1092
// The factor shares the matrix's SELL-32 structure (matrix.cuh): slot 0 holds
// the inverse pivot APC(n), lower slots the L multipliers, upper slots U.
#include "ilu0.cuh"
#include <type_traits>
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
  if ((threadIdx.x & 31) == 0) D.partial[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = s;
}

// rho = sum of the per-warp partials of every finalising launch: fixed chunking and order
// (deterministic); the last CTA to finish combines the per-CTA sums and writes rho, beta
__global__ void __launch_bounds__(kBlock)
ilu_dot_reduce_kernel(int nslots, const double *__restrict__ partial, double *__restrict__ cta_sums,
                      unsigned int *ticket, double *rho_out, double *beta_out, const double *rho0,
                      const int *__restrict__ done, DistPush push) {
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
    if (push.peer) {
      if (threadIdx.x < 8) sh[threadIdx.x] = (threadIdx.x == 0) ? t : 0.0;
      __syncthreads();
      dist_push_record(push, sh);
    } else if (threadIdx.x == 0) {
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

// ---- fixed-width SELL variants of the level kernels (any number of levels) ---------------------
// Same arithmetic as ilu0_fwd/bwd_level_kernel; the slot address needs neither slice_ptr nor rowlen
// (padding slots hold lu = 0) and the columns come from the stencil table where the slice allows it.
template <int W>
__device__ __forceinline__ int ilu_col(int r, int u, long long base, const int *__restrict__ col,
                                       const int *__restrict__ so, int ncols) {
  if (so) {
    const int o = __ldg(so + u);
    if (o != INT_MIN) return min(max(r + o, 0), ncols - 1);
  }
  return __ldg(col + base + 32 * u);
}

template <int W, bool FINAL>
__global__ void __launch_bounds__(kBlock, 8)
ilu0_fwd_w_kernel(int r0, int r1, int lvl1_start, int ncols, const unsigned char *__restrict__ nlow,
                  const int *__restrict__ col, const int *__restrict__ soff,
                  const double *__restrict__ lu, const double *__restrict__ rin,
                  double *__restrict__ d, const int *__restrict__ done, IluDot D) {
  if (done && *done) return;
  double dot = 0.0;
  const int r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (r < r1) {
    const int lo = nlow[r];
    const long long base = (long long)(r >> 5) * (32 * W) + (r & 31);
    const int *so = soff ? soff + (r >> 5) * W : nullptr;
    double v[W], dv[W];
#pragma unroll
    for (int u = 1; u < W; u++) {
      v[u] = 0.0;
      dv[u] = 0.0;
      if (u <= lo) {
        v[u] = __ldg(lu + base + 32 * u);
        const int c = ilu_col<W>(r, u, base, col, so, ncols);
        dv[u] = (c < lvl1_start) ? rin[c] : d[c];
      }
    }
    const double rr = rin[r];
    double tv = rr;
#pragma unroll
    for (int u = 1; u < W; u++)
      if (u <= lo) tv = tv - v[u] * dv[u];
    if (FINAL) {
      tv = tv * __ldg(lu + base);
      dot = rr * tv;
    }
    d[r] = tv;
  }
  if (FINAL && D.partial) ilu_dot_finish(dot, D);
}

template <int W, bool LEVEL0>
__global__ void __launch_bounds__(kBlock, 8)
ilu0_bwd_w_kernel(int r0, int r1, int ncols, const unsigned char *__restrict__ nlow,
                  const int *__restrict__ col, const int *__restrict__ soff,
                  const double *__restrict__ lu, const double *__restrict__ rin,
                  double *__restrict__ d, const int *__restrict__ done, IluDot D) {
  if (done && *done) return;
  double dot = 0.0;
  const int r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (r < r1) {
    const int lo = nlow[r];
    const long long base = (long long)(r >> 5) * (32 * W) + (r & 31);
    const int *so = soff ? soff + (r >> 5) * W : nullptr;
    double v[W], dv[W];
    const double piv = __ldg(lu + base);
#pragma unroll
    for (int u = 1; u < W; u++) {
      v[u] = 0.0;
      dv[u] = 0.0;
      if (u > lo) {
        v[u] = __ldg(lu + base + 32 * u);
        const int c = ilu_col<W>(r, u, base, col, so, ncols);
        dv[u] = d[c];
      }
    }
    const double rr = rin[r];
    double tv = LEVEL0 ? rr : d[r];
#pragma unroll
    for (int u = 1; u < W; u++)
      if (u > lo) tv = tv - v[u] * dv[u];
    tv = tv * piv;
    d[r] = tv;
    dot = rr * tv;
  }
  if (D.partial) ilu_dot_finish(dot, D);
}

template <int W>
static int launch_levels_w(const mf6gpu_matrix &A, const double *lu, const double *rin, double *d,
                           const int *done, cudaStream_t s, const IluDotArgs *dot) {
  int launches = 0;
  const int L = A.nlevels;
  const int lvl1 = (L > 1) ? A.level_ptr[1] : A.n;
  const int *so = A.slot_off.n ? A.slot_off.p : nullptr;
  int slot = 0;
  for (int l = 1; l < L; l++) {
    const int r0 = A.level_ptr[l], r1 = A.level_ptr[l + 1];
    if (r1 <= r0) continue;
    const int g = (r1 - r0 + kBlock - 1) / kBlock;
    if (l == L - 1) {
      IluDot D{dot ? dot->partial + slot : nullptr};
      slot += g * (kBlock / 32);
      ilu0_fwd_w_kernel<W, true><<<g, kBlock, 0, s>>>(r0, r1, lvl1, A.n_ext, A.nlow.p, A.col.p, so, lu, rin, d,
                                                      done, D);
    } else {
      ilu0_fwd_w_kernel<W, false><<<g, kBlock, 0, s>>>(r0, r1, lvl1, A.n_ext, A.nlow.p, A.col.p, so, lu, rin, d,
                                                       done, IluDot{nullptr});
    }
    launches++;
  }
  for (int l = (L > 1 ? L - 2 : 0); l >= 0; l--) {
    const int r0 = A.level_ptr[l], r1 = A.level_ptr[l + 1];
    if (r1 <= r0) continue;
    const int g = (r1 - r0 + kBlock - 1) / kBlock;
    IluDot D{dot ? dot->partial + slot : nullptr};
    slot += g * (kBlock / 32);
    if (l == 0)
      ilu0_bwd_w_kernel<W, true><<<g, kBlock, 0, s>>>(r0, r1, A.n_ext, A.nlow.p, A.col.p, so, lu, rin, d, done, D);
    else
      ilu0_bwd_w_kernel<W, false><<<g, kBlock, 0, s>>>(r0, r1, A.n_ext, A.nlow.p, A.col.p, so, lu, rin, d, done, D);
    launches++;
  }
  if (dot) {
    const int rb = std::max(1, std::min(148, slot / (4 * kBlock)));
    ilu_dot_reduce_kernel<<<rb, kBlock, 0, s>>>(slot, dot->partial, dot->cta_sums, dot->ticket, dot->rho_out,
                                                dot->beta_out, dot->rho0, done, dot->push);
    launches++;
  }
  return launches;
}

// ---- block sweeps (BLOCK_MULTICOLOR on fixed-width SELL) ---------------------------------------
// A block is a vertical cell column whose cells form a chain in elimination order; blocks of one
// colour are not connected.  The sweep of a colour therefore splits into
//   gather: for every row of the colour at once (one thread per row, like a level kernel) the products
//           with the neighbours in OTHER blocks -- earlier colours in the forward sweep, later ones
//           in the backward sweep -- accumulated into d;
//   chain:  one thread per block runs the first-order recurrence along its chain, all operands
//           (d, chain multiplier, pivot, r) loaded up front, consecutive threads touching
//           consecutive rows of one layer (coalesced);
// i.e. at most two launches per colour and sweep whatever the number of layers.  The first colour
// needs no forward gather, the last colour no backward gather and its two chain passes fuse.
// The chain term is applied after the others, so the summation order differs from the level kernels
// (rounding only) when the chain neighbour is not the last slot of its half.
// AFFINE (regular colour, matrix.cuh): grid.y = cell index k along the chain, row = kbase[k] + q
template <int W, bool LOWER, bool AFFINE>
__global__ void __launch_bounds__(kBlock, 8)
ilu0_blk_gather_kernel(int nb, int maxk, int ncols, const int *__restrict__ brow,
                       const unsigned char *__restrict__ bnlow, const unsigned char *__restrict__ bchain,
                       const int *__restrict__ kbase, const unsigned char *__restrict__ nlow,
                       const int *__restrict__ col, const int *__restrict__ soff,
                       const double *__restrict__ lu, const double *__restrict__ rin, double *d,
                       const int *__restrict__ done) {
  if (done && *done) return;
  int r, lo, skip;
  if (AFFINE) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (q >= nb) return;
    r = __ldg(kbase + k) + q;
    lo = __ldg(nlow + r);
    skip = LOWER ? (k > 0 ? lo : 0) : (k + 1 < maxk ? lo + 1 : 0);
  } else {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nb * maxk) return;
    r = __ldg(brow + t);
    if (r < 0) return;
    lo = __ldg(bnlow + t);
    const int ch = __ldg(bchain + t);
    skip = LOWER ? (ch & 15) : (ch >> 4);
  }
  const long long base = (long long)(r >> 5) * (32 * W) + (r & 31);
  const int *so = soff ? soff + (r >> 5) * W : nullptr;
  double acc = LOWER ? rin[r] : d[r];
  double v[W], dv[W];
  int o[W];
#pragma unroll
  for (int u = 1; u < W; u++) {
    const bool want = (LOWER == (u <= lo)) && u != skip;
    v[u] = want ? __ldg(lu + base + 32 * u) : 0.0;
    o[u] = (soff && want) ? __ldg(so + u) : INT_MIN;
  }
#pragma unroll
  for (int u = 1; u < W; u++) {
    const bool want = (LOWER == (u <= lo)) && u != skip;
    int c = min(max(r + o[u], 0), ncols - 1);
    if (want && o[u] == INT_MIN) c = __ldg(col + base + 32 * u);
    o[u] = c;
  }
#pragma unroll
  for (int u = 1; u < W; u++) {
    const bool want = (LOWER == (u <= lo)) && u != skip;
    dv[u] = want ? d[o[u]] : 0.0;
  }
#pragma unroll
  for (int u = 1; u < W; u++) {
    const bool want = (LOWER == (u <= lo)) && u != skip;
    acc = want ? acc - v[u] * dv[u] : acc;
  }
  d[r] = acc;
}

// MODE 0: forward   d(k) = src(k) - L(k,k-1) d(k-1)            src = rin (first colour) or d (gathered)
// MODE 1: backward  d(k) = (d(k) - U(k,k+1) d(k+1)) * piv(k)   + rho partial
// MODE 2: both in one pass (last colour, block of at most KC cells)
__device__ __forceinline__ double ld_f64(const double *p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_nc_f64(const double *p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_nc_i32(const int *p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_nc_u8(const unsigned char *p) {
  unsigned int v;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return (int)v;
}

// ADDR 0: rows and chain slots from the block tables; 1: regular colour, bases from kbase[]; 2: regular
// colour whose chain fits one chunk of <= 16 cells: bases as kernel parameters (no dependent load at all
// in front of the operand loads)
struct ChainBase {
  int r0[16];
};
// FROM_RIN: the forward pass starts from rin (first colour: nothing was gathered into d)
template <int MODE, int KC, int ADDR, bool FROM_RIN>
__global__ void __launch_bounds__(kBlock, 2)
ilu0_blk_chain_kernel(int nb, int maxk, int W, int opaque_zero, const int *__restrict__ brow,
                      const unsigned char *__restrict__ bchain, const int *__restrict__ kbase,
                      const __grid_constant__ ChainBase kb, const unsigned char *__restrict__ nlow,
                      const double *__restrict__ lu, const double *__restrict__ rin, double *d,
                      const int *__restrict__ done, IluDot D) {
  if (done && *done) return;
  double dot = 0.0;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nb) {
    const int nchunks = (maxk + KC - 1) / KC;
    double carry = 0.0;  // value of the chain neighbour in the chunk handled before
    for (int it = 0; it < nchunks; it++) {
      const int k0 = ((MODE == 1) ? nchunks - 1 - it : it) * KC;
      int rk[KC], ck[KC];
      double acc[KC], ml[KC], mu[KC], pk[KC], rr[KC];
      // Operand loads are volatile asm in program order: the compiler otherwise sinks each load next to
      // its consumer to save registers, and a consumer between two loads stalls the warp (in-order issue)
      // -- KC dependent round trips instead of two.  Invalid cells read row 0 and are masked afterwards.
#pragma unroll
      for (int k = 0; k < KC; k++) {
        const bool in = (k0 + k < maxk);
        if (ADDR != 0) {
          rk[k] = !in ? -1 : (ADDR == 2 ? kb.r0[(k0 + k) & 15] : __ldg(kbase + k0 + k)) + q;
        } else {
          rk[k] = in ? ld_nc_i32(brow + (size_t)(k0 + k) * nb + q) : -1;
        }
      }
#pragma unroll
      for (int k = 0; k < KC; k++) {
        const bool in = (k0 + k < maxk);
        if (ADDR != 0) {
          const int lo = ld_nc_u8(nlow + max(rk[k], 0));
          ck[k] = !in ? 0 : ((k0 + k > 0 ? lo : 0) | ((k0 + k + 1 < maxk ? lo + 1 : 0) << 4));
        } else {
          ck[k] = in ? ld_nc_u8(bchain + (size_t)(k0 + k) * nb + q) : 0;
        }
      }
#pragma unroll
      for (int k = 0; k < KC; k++) {
        const int r = max(rk[k], 0);
        const long long base = (long long)(r >> 5) * (32 * W) + (r & 31);
        acc[k] = (MODE != 1 && FROM_RIN) ? ld_nc_f64(rin + r) : ld_f64(d + r);
        if (MODE != 0) {
          pk[k] = ld_nc_f64(lu + base);
          if (!(MODE == 2 && FROM_RIN)) rr[k] = ld_nc_f64(rin + r);   // else acc[k] is r itself
        }
      }
#pragma unroll
      for (int k = 0; k < KC; k++) {
        const int r = max(rk[k], 0);
        const long long base = (long long)(r >> 5) * (32 * W) + (r & 31);
        const int slo = ck[k] & 15, sup = ck[k] >> 4;
        if (MODE != 1) ml[k] = ld_nc_f64(lu + base + 32 * slo);   // slot 0 (the pivot) when there is no chain entry
        if (MODE != 0) mu[k] = ld_nc_f64(lu + base + 32 * sup);
      }
      // the recurrence must not start before the last operand load has been ISSUED: its start value is
      // made to depend (through `opaque_zero`, always 0) on the values loaded last
      long long after = 0;
#pragma unroll
      for (int k = 0; k < KC; k++) {
        const bool valid = rk[k] >= 0;
        const int slo = ck[k] & 15, sup = ck[k] >> 4;
        if (MODE != 1) after ^= __double_as_longlong(ml[k]);
        if (MODE != 0) after ^= __double_as_longlong(mu[k]) ^ __double_as_longlong(pk[k]);
        if (MODE != 0 && !(MODE == 2 && FROM_RIN)) after ^= __double_as_longlong(rr[k]);
        if (!valid) acc[k] = 0.0;
        if (MODE != 1 && !(valid && slo)) ml[k] = 0.0;
        if (MODE != 0 && !(valid && sup)) mu[k] = 0.0;
      }
      carry = __longlong_as_double(__double_as_longlong(carry) | (after & (long long)opaque_zero));
      double fw[KC];  // forward values
      if (MODE != 1) {
#pragma unroll
        for (int k = 0; k < KC; k++) {
          fw[k] = 0.0;
          if (rk[k] >= 0) {
            carry = acc[k] - ml[k] * carry;
            fw[k] = carry;
            if (MODE == 0) d[rk[k]] = carry;
          }
        }
      }
      if (MODE != 0) {
        if (MODE == 2) carry = 0.0;
#pragma unroll
        for (int k = KC - 1; k >= 0; k--)
          if (rk[k] >= 0) {
            carry = ((MODE == 2 ? fw[k] : acc[k]) - mu[k] * carry) * pk[k];
            d[rk[k]] = carry;
            dot += ((MODE == 2 && FROM_RIN) ? acc[k] : rr[k]) * carry;
          }
      }
    }
  }
  if (MODE != 0 && D.partial) ilu_dot_finish(dot, D);
}

// rho partials the block sweeps write (one per warp of every finalising launch)
size_t ilu0_block_dot_slots(const mf6gpu_matrix &A) {
  size_t n = 0;
  for (int c = 0; c < A.blk_ncolors; c++) n += (size_t)((A.blk_nb[c] + kBlock - 1) / kBlock) * (kBlock / 32);
  return n;
}

// cells per chunk: the operands of a whole chunk live in registers (2 doubles per cell in MODE 0, 4 in MODE 1,
// 5 in MODE 2; 128 registers per thread), longer chains run chunk after chunk
static inline int chain_chunk(int mode, int maxk) {
  const int cap = (mode == 0) ? 16 : (mode == 1 ? 12 : 10);
  if (maxk > cap) return 8;
  return maxk <= 4 ? 4 : maxk <= 6 ? 6 : maxk <= 8 ? 8 : maxk <= 10 ? 10 : maxk <= 12 ? 12 : 16;
}
static inline bool chain_fits_one_chunk(int mode, int maxk) { return maxk <= chain_chunk(mode, maxk); }

// CTA size of the chain kernels.  They hold a whole column's operands in registers (25 % occupancy), and all warps of
// a CTA move through load -> recurrence -> store in step; small CTAs stagger those phases across the SM and shrink
// the last partial wave (MF6GPU_CHAIN_BLOCK overrides for tuning)
static int chain_block() {
  static int v = [] {
    const char *e = std::getenv("MF6GPU_CHAIN_BLOCK");
    const int c = e ? std::atoi(e) : 0;
    return (c == 32 || c == 64 || c == 128 || c == 256) ? c : 64;
  }();
  return v;
}

struct ChainArgs {
  int g, nb, maxk, W;
  const int *brow;
  const unsigned char *bch;
  const int *kb;
  ChainBase kbv;
  const unsigned char *nlow;
  const double *lu, *rin;
  double *d;
  const int *done;
  IluDot D;
  cudaStream_t s;
};

template <int MODE, int KC>
static void launch_chain_kc(const ChainArgs &a, int addr, bool from_rin) {
  constexpr int cap = (MODE == 0) ? 16 : (MODE == 1 ? 12 : 10);
  if constexpr (KC <= cap) {
    auto go = [&](auto addr_c, auto rin_c) {
      ilu0_blk_chain_kernel<MODE, KC, decltype(addr_c)::value, decltype(rin_c)::value><<<a.g, chain_block(), 0, a.s>>>(
          a.nb, a.maxk, a.W, 0, a.brow, a.bch, a.kb, a.kbv, a.nlow, a.lu, a.rin, a.d, a.done, a.D);
    };
    auto with_addr = [&](auto rin_c) {
      if (addr == 2)
        go(std::integral_constant<int, 2>{}, rin_c);
      else if (addr == 1)
        go(std::integral_constant<int, 1>{}, rin_c);
      else
        go(std::integral_constant<int, 0>{}, rin_c);
    };
    if (from_rin && MODE != 1)
      with_addr(std::bool_constant<(MODE != 1)>{});
    else
      with_addr(std::false_type{});
  }
}

template <int MODE>
static void launch_chain(const mf6gpu_matrix &A, int c, bool from_rin, const double *lu, const double *rin,
                         double *d, const int *done, cudaStream_t s, IluDot D) {
  ChainArgs a{};
  a.nb = A.blk_nb[c];
  a.maxk = A.blk_maxk[c];
  a.g = (a.nb + chain_block() - 1) / chain_block();
  a.W = A.uniform_w;
  a.brow = A.blk_rows.p + A.blk_off[c];
  a.bch = A.blk_chain.p + A.blk_off[c];
  a.kb = A.blk_base.p + A.blk_base_off[c];
  a.nlow = A.nlow.p;
  a.lu = lu;
  a.rin = rin;
  a.d = d;
  a.done = done;
  a.D = D;
  a.s = s;
  const int addr = !A.blk_affine[c] ? 0 : (a.maxk <= 16 ? 2 : 1);
  if (addr == 2)
    for (int k = 0; k < a.maxk; k++) a.kbv.r0[k] = A.blk_base_h[A.blk_base_off[c] + k];
  switch (chain_chunk(MODE, a.maxk)) {
    case 4: launch_chain_kc<MODE, 4>(a, addr, from_rin); break;
    case 6: launch_chain_kc<MODE, 6>(a, addr, from_rin); break;
    case 8: launch_chain_kc<MODE, 8>(a, addr, from_rin); break;
    case 10: launch_chain_kc<MODE, 10>(a, addr, from_rin); break;
    case 12: launch_chain_kc<MODE, 12>(a, addr, from_rin); break;
    case 16: launch_chain_kc<MODE, 16>(a, addr, from_rin); break;
  }
}

template <int W>
static int launch_blocks_w(const mf6gpu_matrix &A, const double *lu, const double *rin, double *d,
                           const int *done, cudaStream_t s, const IluDotArgs *dot) {
  const int C = A.blk_ncolors;
  const int *so = A.slot_off.n ? A.slot_off.p : nullptr;
  int launches = 0, slot = 0;
  auto gather = [&](auto lower, int c) {
    constexpr bool LOWER = decltype(lower)::value;
    const int nb = A.blk_nb[c], maxk = A.blk_maxk[c];
    const int *brow = A.blk_rows.p + A.blk_off[c];
    const unsigned char *bnl = A.blk_nlow.p + A.blk_off[c], *bch = A.blk_chain.p + A.blk_off[c];
    const int *kb = A.blk_base.p + A.blk_base_off[c];
    if (A.blk_affine[c] && maxk <= 65535) {
      const dim3 g((nb + kBlock - 1) / kBlock, maxk);
      ilu0_blk_gather_kernel<W, LOWER, true><<<g, kBlock, 0, s>>>(nb, maxk, A.n_ext, brow, bnl, bch, kb, A.nlow.p,
                                                                  A.col.p, so, lu, rin, d, done);
    } else {
      const long long ncell = (long long)nb * maxk;
      ilu0_blk_gather_kernel<W, LOWER, false><<<(unsigned)((ncell + kBlock - 1) / kBlock), kBlock, 0, s>>>(
          nb, maxk, A.n_ext, brow, bnl, bch, kb, A.nlow.p, A.col.p, so, lu, rin, d, done);
    }
    launches++;
  };
  // the last colour's two chain passes fuse when nothing of its U half lies outside the chains and
  // a block fits one chunk
  const bool fuse_last = !A.blk_has_upper[C - 1] && chain_fits_one_chunk(2, A.blk_maxk[C - 1]);
  for (int c = 0; c < C; c++) {
    if (A.blk_nb[c] == 0) continue;
    if (A.blk_has_lower[c]) gather(std::true_type{}, c);
    if (c == C - 1 && fuse_last) break;
    launch_chain<0>(A, c, !A.blk_has_lower[c], lu, rin, d, done, s, IluDot{nullptr});
    launches++;
  }
  for (int c = C - 1; c >= 0; c--) {
    if (A.blk_nb[c] == 0) continue;
    IluDot D{dot ? dot->partial + slot : nullptr};
    slot += ((A.blk_nb[c] + kBlock - 1) / kBlock) * (kBlock / 32);
    if (c == C - 1 && fuse_last) {
      launch_chain<2>(A, c, !A.blk_has_lower[c], lu, rin, d, done, s, D);
    } else {
      if (A.blk_has_upper[c]) gather(std::false_type{}, c);
      launch_chain<1>(A, c, false, lu, rin, d, done, s, D);
    }
    launches++;
  }
  if (dot) {
    const int rb = std::max(1, std::min(148, slot / (4 * kBlock)));
    ilu_dot_reduce_kernel<<<rb, kBlock, 0, s>>>(slot, dot->partial, dot->cta_sums, dot->ticket, dot->rho_out,
                                                dot->beta_out, dot->rho0, done, dot->push);
    launches++;
  }
  return launches;
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
                                                dot->beta_out, dot->rho0, done, dot->push);
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
  if (A.blk_ncolors > 0 && A.blk_rows.n > 0 && A.blk_chain_ok && !std::getenv("MF6GPU_NO_BLOCK_SWEEP")) {
    switch (A.uniform_w) {
      case 4: return launch_blocks_w<4>(A, lu, rin, d, done, s, dot);
      case 5: return launch_blocks_w<5>(A, lu, rin, d, done, s, dot);
      case 6: return launch_blocks_w<6>(A, lu, rin, d, done, s, dot);
      case 7: return launch_blocks_w<7>(A, lu, rin, d, done, s, dot);
      case 8: return launch_blocks_w<8>(A, lu, rin, d, done, s, dot);
      case 9: return launch_blocks_w<9>(A, lu, rin, d, done, s, dot);
      case 10: return launch_blocks_w<10>(A, lu, rin, d, done, s, dot);
      default: break;
    }
  }
  if (L <= 64) {  // fixed-width level kernels (few, wide levels); thin natural-order levels keep the generic path
    switch (A.uniform_w) {
      case 4: return launch_levels_w<4>(A, lu, rin, d, done, s, dot);
      case 5: return launch_levels_w<5>(A, lu, rin, d, done, s, dot);
      case 6: return launch_levels_w<6>(A, lu, rin, d, done, s, dot);
      case 7: return launch_levels_w<7>(A, lu, rin, d, done, s, dot);
      case 8: return launch_levels_w<8>(A, lu, rin, d, done, s, dot);
      case 9: return launch_levels_w<9>(A, lu, rin, d, done, s, dot);
      case 10: return launch_levels_w<10>(A, lu, rin, d, done, s, dot);
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
                                                dot->beta_out, dot->rho0, done, dot->push);
    launches++;
  }
  return launches;
}

}  // namespace mf6
