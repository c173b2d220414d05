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
                                                dot->beta_out, dot->rho0, done);
    launches++;
  }
  return launches;
}

// ---- block sweeps (BLOCK_MULTICOLOR on fixed-width SELL) ---------------------------------------
// One thread owns one block (a vertical cell column) and walks its cells in elimination order, so the
// whole forward sweep of a colour is ONE launch whatever the number of layers; consecutive threads
// touch consecutive rows of the same layer (coalesced).  Lower neighbours outside the block belong to
// earlier colours (finished by earlier launches), upper ones to later colours.  Values of the thread's
// own block are read back through global memory in program order (same-thread RAW).
// MODE 0: forward sweep            d(n) = r(n) - sum_lower L d
// MODE 1: backward sweep           d(n) = (d(n) - sum_upper U d) * piv          (+ rho partial)
// MODE 2: forward then backward in one pass -- valid for the LAST colour, whose upper entries all lie
//         inside the block
template <int W, int MODE, int MAXK>
__global__ void __launch_bounds__(kBlock)
ilu0_block_kernel(int nb, int maxk, int ncols, const int *__restrict__ brow,
                  const unsigned char *__restrict__ nlow, const int *__restrict__ col,
                  const int *__restrict__ soff, const double *__restrict__ lu,
                  const double *__restrict__ rin, double *d, const int *__restrict__ done, IluDot D) {
  if (done && *done) return;
  double dot = 0.0;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nb) {
    // Phase 1 -- everything that does not depend on the block's own recurrence, for all cells of the
    // block at once (independent loads, issued back to back): row ids, the sums over the neighbours in
    // OTHER blocks, the multipliers that couple consecutive cells of the block, the inverse pivots.
    int rk[MAXK];
    double sk[MAXK];   // forward:  r - sum_{lower, other blocks} L d     backward: fwd - sum_{upper, other blocks} U d
    double lk[MAXK];   // L multiplier towards the previous cell of the block (0 if none)
    double uk[MAXK];   // U entry towards the next cell of the block (0 if none)      [MODE 1, 2]
    double pk[MAXK];   // inverse pivot                                              [MODE 1, 2]
#pragma unroll
    for (int k = 0; k < MAXK; k++) rk[k] = (k < maxk) ? __ldg(brow + (size_t)k * nb + q) : -1;
#pragma unroll
    for (int k = 0; k < MAXK; k++) {
      sk[k] = 0.0;
      lk[k] = 0.0;
      uk[k] = 0.0;
      pk[k] = 0.0;
      const int r = rk[k];
      if (r < 0) continue;
      const int rprev = (k > 0) ? rk[k - 1] : -1;
      const int rnext = (k + 1 < MAXK) ? rk[k + 1] : -1;
      const int lo = nlow[r];
      const long long base = (long long)(r >> 5) * (32 * W) + (r & 31);
      const int *so = soff ? soff + (r >> 5) * W : nullptr;
      double acc = (MODE == 1) ? d[r] : rin[r];
      if (MODE != 0) pk[k] = __ldg(lu + base);
#pragma unroll
      for (int u = 1; u < W; u++) {
        const bool lower = (u <= lo);
        if ((MODE == 0 && !lower) || (MODE == 1 && lower)) continue;
        const double v = __ldg(lu + base + 32 * u);
        const int c = ilu_col<W>(r, u, base, col, so, ncols);
        // a padding slot (v == 0) may carry a stencil column that happens to equal rprev / rnext: it
        // must not overwrite the real multiplier
        if (lower) {
          if (c == rprev) {
            if (v != 0.0) lk[k] = v;
          } else
            acc = acc - v * d[c];
        } else {
          if (c == rnext) {
            if (v != 0.0) uk[k] = v;
          } else if (MODE == 1)
            acc = acc - v * d[c];
          // MODE 2 (last colour): every upper entry lies inside the block, nothing else to subtract
        }
      }
      sk[k] = acc;
    }
    // Phase 2 -- the recurrences along the block
    if (MODE == 0 || MODE == 2) {
      double prev = 0.0;
#pragma unroll
      for (int k = 0; k < MAXK; k++) {
        if (rk[k] < 0) continue;
        const double tv = sk[k] - lk[k] * prev;
        sk[k] = tv;
        prev = tv;
        if (MODE == 0) d[rk[k]] = tv;
      }
    }
    if (MODE == 1 || MODE == 2) {
      double next = 0.0;
#pragma unroll
      for (int k = MAXK - 1; k >= 0; k--) {
        if (rk[k] < 0) continue;
        const double tv = (sk[k] - uk[k] * next) * pk[k];
        next = tv;
        d[rk[k]] = tv;
        dot += rin[rk[k]] * tv;
      }
    }
  }
  if (MODE != 0 && D.partial) ilu_dot_finish(dot, D);
}

template <int W, int MODE>
static void launch_block_kernel(int g, cudaStream_t s, int nb, int maxk, int ncols, const int *brow,
                                const unsigned char *nlow, const int *col, const int *so, const double *lu,
                                const double *rin, double *d, const int *done, IluDot D) {
  if (maxk <= 4)
    ilu0_block_kernel<W, MODE, 4><<<g, kBlock, 0, s>>>(nb, maxk, ncols, brow, nlow, col, so, lu, rin, d, done, D);
  else if (maxk <= 8)
    ilu0_block_kernel<W, MODE, 8><<<g, kBlock, 0, s>>>(nb, maxk, ncols, brow, nlow, col, so, lu, rin, d, done, D);
  else if (maxk <= 12)
    ilu0_block_kernel<W, MODE, 12><<<g, kBlock, 0, s>>>(nb, maxk, ncols, brow, nlow, col, so, lu, rin, d, done, D);
  else if (maxk <= 16)
    ilu0_block_kernel<W, MODE, 16><<<g, kBlock, 0, s>>>(nb, maxk, ncols, brow, nlow, col, so, lu, rin, d, done, D);
  else
    ilu0_block_kernel<W, MODE, 32><<<g, kBlock, 0, s>>>(nb, maxk, ncols, brow, nlow, col, so, lu, rin, d, done, D);
}

template <int W>
static int launch_blocks_w(const mf6gpu_matrix &A, const double *lu, const double *rin, double *d,
                           const int *done, cudaStream_t s, const IluDotArgs *dot) {
  const int C = A.blk_ncolors;
  const int *so = A.slot_off.n ? A.slot_off.p : nullptr;
  int launches = 0, slot = 0;
  auto grid = [&](int c) { return (A.blk_nb[c] + kBlock - 1) / kBlock; };
  for (int c = 0; c < C - 1; c++) {
    launch_block_kernel<W, 0>(grid(c), s, A.blk_nb[c], A.blk_maxk[c], A.n_ext, A.blk_rows.p + A.blk_off[c], A.nlow.p,
                              A.col.p, so, lu, rin, d, done, IluDot{nullptr});
    launches++;
  }
  for (int c = C - 1; c >= 0; c--) {
    IluDot D{dot ? dot->partial + slot : nullptr};
    slot += grid(c) * (kBlock / 32);
    if (c == C - 1)
      launch_block_kernel<W, 2>(grid(c), s, A.blk_nb[c], A.blk_maxk[c], A.n_ext, A.blk_rows.p + A.blk_off[c],
                                A.nlow.p, A.col.p, so, lu, rin, d, done, D);
    else
      launch_block_kernel<W, 1>(grid(c), s, A.blk_nb[c], A.blk_maxk[c], A.n_ext, A.blk_rows.p + A.blk_off[c],
                                A.nlow.p, A.col.p, so, lu, rin, d, done, D);
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
                                                dot->beta_out, dot->rho0, done);
    launches++;
  }
  return launches;
}

}  // namespace mf6
