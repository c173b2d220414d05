// ilut.cu -- see ilut.cuh
#include "ilut.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace mf6 {

// ---------------------------------------------------------------- host factorisation -------------
// 1-based arrays (element 0 unused) so that the index arithmetic is the reference's own.
// qsplit (ilut.f90:484-548): |a(i)| >= |a(ncut)| for i < ncut, <= for i > ncut
static void quick_split(int n, double *a, int *ind, int ncut) {
  int first = 1, last = n;
  if (ncut < first || ncut > last) return;
  for (;;) {
    int mid = first;
    const double abskey = std::fabs(a[mid]);
    for (int j = first + 1; j <= last; j++)
      if (std::fabs(a[j]) > abskey) {
        mid++;
        std::swap(a[mid], a[j]);
        std::swap(ind[mid], ind[j]);
      }
    std::swap(a[mid], a[first]);
    std::swap(ind[mid], ind[first]);
    if (mid == ncut) return;
    if (mid > ncut)
      last = mid - 1;
    else
      first = mid + 1;
  }
}

// one pass of ilut (ilut.f90:48-428) with fixed delta; izero is the reference's in/out flag
static int ilut_pass(IlutPlan &P, const double *a, double relax, int &izero, double delta) {
  const int n = P.n, lfil = P.lfil;
  const long long iwk = P.iwk;
  const double droptol = P.droptol;
  double *alu = P.alu.data(), *w = P.w.data();
  int *jlu = P.jlu.data(), *ju = P.ju.data(), *jw = P.jw.data();
  const int *ia = P.e_ia.data(), *ja = P.e_ja.data();
  if (lfil < 0) return -4;
  long long ju0 = (long long)n + 2;
  jlu[1] = (int)ju0;
  for (int j = 1; j <= n; j++) jw[n + j] = 0;
  for (int ii = 1; ii <= n; ii++) {
    const int j1 = ia[ii - 1] + 1, j2 = ia[ii];
    double dropsum = 0.0, tnorm = 0.0;
    for (int k = j1; k <= j2; k++) tnorm = tnorm + std::fabs(a[k - 1]);
    if (tnorm == 0.0) return -5;
    tnorm = tnorm / (double)(j2 - j1 + 1);
    int lenu = 1, lenl = 0;
    jw[ii] = ii;
    w[ii] = 0.0;
    jw[n + ii] = ii;
    for (int j = j1; j <= j2; j++) {
      const int k = ja[j - 1] + 1;
      const double t = a[j - 1];
      if (k < ii) {
        lenl++;
        jw[lenl] = k;
        w[lenl] = t;
        jw[n + k] = lenl;
      } else if (k == ii) {
        w[ii] = t;
      } else {
        lenu++;
        const int jpos = ii + lenu - 1;
        jw[jpos] = k;
        w[jpos] = t;
        jw[n + k] = jpos;
      }
    }
    int ilen = 0;
    for (int jj = 1; jj <= lenl; jj++) {  // lenl may grow inside the loop (fill-in)
      // eliminate in ascending column order: select the smallest column among jw(jj..lenl)
      int jrow = jw[jj], k = jj;
      for (int j = jj + 1; j <= lenl; j++)
        if (jw[j] < jrow) {
          jrow = jw[j];
          k = j;
        }
      if (k != jj) {
        const int j = jw[jj];
        jw[jj] = jw[k];
        jw[k] = j;
        jw[n + jrow] = jj;
        jw[n + j] = k;
        std::swap(w[jj], w[k]);
      }
      jw[n + jrow] = 0;
      const double fact = w[jj] * alu[jrow];
      if (std::fabs(fact) <= droptol) {
        dropsum = dropsum + w[jj];
        continue;
      }
      for (int kk = ju[jrow]; kk <= jlu[jrow + 1] - 1; kk++) {
        const double s = fact * alu[kk];
        const int j = jlu[kk];
        const int jpos = jw[n + j];
        if (j >= ii) {
          if (jpos == 0) {
            lenu++;
            if (lenu > n) return -1;
            const int i = ii + lenu - 1;
            jw[i] = j;
            jw[n + j] = i;
            w[i] = -s;
          } else {
            w[jpos] = w[jpos] - s;
          }
        } else {
          if (jpos == 0) {
            lenl++;
            if (lenl > n) return -1;
            jw[lenl] = j;
            jw[n + j] = lenl;
            w[lenl] = -s;
          } else {
            w[jpos] = w[jpos] - s;
          }
        }
      }
      ilen++;
      w[ilen] = fact;
      jw[ilen] = jrow;
    }
    for (int k = 1; k <= lenu; k++) jw[n + jw[ii + k - 1]] = 0;
    lenl = ilen;
    ilen = std::min(lenl, lfil);
    quick_split(lenl, w, jw, ilen);
    for (int k = 1; k <= ilen; k++) {
      if (ju0 > iwk) return -2;
      alu[ju0] = w[k];
      jlu[ju0] = jw[k];
      ju0++;
    }
    ju[ii] = (int)ju0;
    ilen = 0;
    for (int k = 1; k <= lenu - 1; k++) {
      if (std::fabs(w[ii + k]) > droptol * tnorm) {
        ilen++;
        w[ii + ilen] = w[ii + k];
        jw[ii + ilen] = jw[ii + k];
      } else {
        dropsum = dropsum + w[ii + k];
      }
    }
    lenu = ilen + 1;
    ilen = std::min(lenu, lfil);
    quick_split(lenu - 1, w + ii, jw + ii, ilen);
    if (ilen + ju0 > iwk) return -3;
    for (int k = ii + 1; k <= ii + ilen - 1; k++) {
      jlu[ju0] = jw[k];
      alu[ju0] = w[k];
      ju0++;
    }
    const double diag = w[ii];
    double diag_working = (1.0 + delta) * diag + (relax * dropsum);
    const double sign_check = std::copysign(std::fabs(diag), diag_working);
    if (sign_check != diag) {
      if (izero > 1) {
        diag_working = std::copysign(1.0, diag) * (1.0e-4 + droptol) * tnorm;
      } else {
        izero = 1;
        return 0;
      }
    }
    if (std::fabs(diag_working) == 0.0) {
      if (izero > 1) {
        diag_working = std::copysign(1.0, diag) * (1.0e-4 + droptol) * tnorm;
      } else {
        izero = 1;
        return 0;
      }
    }
    w[ii] = diag_working;
    alu[ii] = 1.0 / w[ii];
    jlu[ii + 1] = (int)ju0;
  }
  return 0;
}

// ---------------------------------------------------------------- plan ---------------------------
void IlutPlan::build(const mf6gpu_matrix &A, int level, double droptol_) {
  n = A.n;
  lfil = level;
  droptol = droptol_;
  // host copies of the SELL structure (one-time)
  std::vector<int> slice_ptr((size_t)A.nslices + 1), col((size_t)A.nslots);
  std::vector<unsigned char> rl((size_t)n);
  A.slice_ptr.download(slice_ptr.data(), slice_ptr.size());
  A.col.download(col.data(), col.size());
  MF6_CK(cudaMemcpy(rl.data(), A.rowlen_loc(), (size_t)n, cudaMemcpyDeviceToHost));
  row_of_e.resize((size_t)n);
  std::vector<int> e_of_row((size_t)n);
  for (int e = 0; e < n; e++) {
    const int r = A.iperm[A.elim[e]];
    row_of_e[e] = r;
    e_of_row[r] = e;
  }
  e_ia.assign((size_t)n + 1, 0);
  for (int e = 0; e < n; e++) e_ia[e + 1] = e_ia[e] + rl[row_of_e[e]];
  const int nnz = e_ia[n];
  e_ja.resize((size_t)nnz);
  std::vector<int> src((size_t)nnz);
  int maxrow = 0;
  for (int e = 0; e < n; e++) {
    const int r = row_of_e[e], len = rl[r];
    maxrow = std::max(maxrow, len);
    const long long base = (long long)slice_ptr[r >> 5] + (r & 31);
    for (int k = 0; k < len; k++) {  // slot 0 = diagonal, then ascending elimination order
      const long long slot = base + 32LL * k;
      e_ja[(size_t)e_ia[e] + k] = e_of_row[col[slot]];
      src[(size_t)e_ia[e] + k] = (int)slot;
    }
  }
  d_src.upload(src);
  d_val.alloc((size_t)std::max(nnz, 1));
  h_val.alloc((size_t)std::max(nnz, 1));
  // ims_calc_pcdims (ImsLinearBase.f90:1171-1190)
  iwk = (lfil > 0) ? (long long)n * (lfil * 2 + 1) : (long long)n * maxrow;
  iwk = std::max(iwk, (long long)n + 2);
  MF6_REQUIRE(iwk < (long long)INT_MAX - 2, "ILUT: factor storage exceeds 2^31 entries");
  alu.assign((size_t)iwk + 2, 0.0);
  jlu.assign((size_t)iwk + 2, 0);
  ju.assign((size_t)n + 2, 0);
  w.assign((size_t)n + 3, 0.0);
  jw.assign(2 * (size_t)n + 2, 0);
  erow.upload(row_of_e);
}

__global__ void ilut_gather_kernel(int nnz, const int *__restrict__ src, const double *__restrict__ val,
                                   double *__restrict__ out) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += gridDim.x * blockDim.x) out[p] = val[src[p]];
}

// group the levels of one sweep: wide levels get a grid launch each, runs of narrow levels one CTA
static void make_groups(const std::vector<int> &lev_ptr, std::vector<IlutPlan::Group> &groups,
                        std::vector<int> &sizes) {
  groups.clear();
  sizes.clear();
  const int nlev = (int)lev_ptr.size() - 1;
  int l = 0;
  while (l < nlev) {
    const int sz = lev_ptr[l + 1] - lev_ptr[l];
    if (sz > kBlock) {
      groups.push_back({lev_ptr[l], sz, 1, 0});
      l++;
      continue;
    }
    IlutPlan::Group g{lev_ptr[l], 0, 0, (int)sizes.size()};
    while (l < nlev && lev_ptr[l + 1] - lev_ptr[l] <= kBlock && g.nlev < 4096) {
      sizes.push_back(lev_ptr[l + 1] - lev_ptr[l]);
      g.count += lev_ptr[l + 1] - lev_ptr[l];
      g.nlev++;
      l++;
    }
    groups.push_back(g);
  }
}

int IlutPlan::factor(const mf6gpu_matrix &A, const double *val, double relax, cudaStream_t s) {
  const int nnz = e_ia[n];
  ilut_gather_kernel<<<grid_for(nnz), kBlock, 0, s>>>(nnz, d_src.p, val, d_val.p);
  MF6_CK(cudaGetLastError());
  MF6_CK(cudaMemcpyAsync(h_val.p, d_val.p, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, s));
  MF6_CK(cudaStreamSynchronize(s));
  // ims_base_pcu (ImsLinearBase.f90:808-858)
  int ipcflag = 0, icount = 0;
  double delta = 0.0;
  for (;;) {
    ierr = ilut_pass(*this, h_val.p, relax, ipcflag, delta);
    if (ierr != 0) break;
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
  static const char *cerr[] = {"", "elimination process has generated a row in L or U whose length is > n",
                               "the matrix L overflows the array al", "the matrix U overflows the array alu",
                               "illegal value for lfil", "zero row encountered"};
  if (ierr != 0) throw Error(std::string("mf6gpu: ILUT: ") + (ierr < 0 && ierr >= -5 ? cerr[-ierr] : "zero pivot"));
  // device factor: L / U rows of elimination row e with DEVICE rows as columns, dependency levels of both sweeps
  std::vector<int> lp((size_t)n + 1, 0), up((size_t)n + 1, 0);
  for (int i = 1; i <= n; i++) {
    lp[i] = lp[i - 1] + (ju[i] - jlu[i]);
    up[i] = up[i - 1] + (jlu[i + 1] - ju[i]);
  }
  nnz_l = lp[n];
  nnz_u = up[n];
  std::vector<int> lc((size_t)std::max<long long>(nnz_l, 1)), uc((size_t)std::max<long long>(nnz_u, 1));
  std::vector<double> lv(lc.size()), uv(uc.size()), pv((size_t)n);
  std::vector<int> flev((size_t)n, 0), blev((size_t)n, 0);
  for (int i = 1; i <= n; i++) {
    int lvl = 0, q = lp[i - 1];
    for (int k = jlu[i]; k <= ju[i] - 1; k++, q++) {
      const int ec = jlu[k] - 1;
      lc[q] = row_of_e[ec];
      lv[q] = alu[k];
      lvl = std::max(lvl, flev[ec] + 1);
    }
    flev[i - 1] = lvl;
    pv[i - 1] = alu[i];
  }
  for (int i = n; i >= 1; i--) {
    int lvl = 0, q = up[i - 1];
    for (int k = ju[i]; k <= jlu[i + 1] - 1; k++, q++) {
      const int ec = jlu[k] - 1;
      uc[q] = row_of_e[ec];
      uv[q] = alu[k];
      lvl = std::max(lvl, blev[ec] + 1);
    }
    blev[i - 1] = lvl;
  }
  auto sort_levels = [&](const std::vector<int> &lev, std::vector<int> &list, std::vector<int> &ptr) {
    int nl = 0;
    for (int e = 0; e < n; e++) nl = std::max(nl, lev[e] + 1);
    ptr.assign((size_t)nl + 1, 0);
    for (int e = 0; e < n; e++) ptr[lev[e] + 1]++;
    for (int l = 0; l < nl; l++) ptr[l + 1] += ptr[l];
    list.resize((size_t)n);
    std::vector<int> cur(ptr.begin(), ptr.end() - 1);
    for (int e = 0; e < n; e++) list[cur[lev[e]]++] = e;
  };
  std::vector<int> fl, fp, bl, bp, fsz, bsz;
  sort_levels(flev, fl, fp);
  sort_levels(blev, bl, bp);
  nflev = (int)fp.size() - 1;
  nblev = (int)bp.size() - 1;
  make_groups(fp, fgroups, fsz);
  make_groups(bp, bgroups, bsz);
  if (fsz.empty()) fsz.push_back(0);
  if (bsz.empty()) bsz.push_back(0);
  lptr.upload(lp);
  uptr.upload(up);
  lcol.upload(lc);
  ucol.upload(uc);
  lval.upload(lv);
  uval.upload(uv);
  piv.upload(pv);
  flist.upload(fl);
  blist.upload(bl);
  flev_sz.upload(fsz);
  blev_sz.upload(bsz);
  (void)A;
  return icount;
}

// ---------------------------------------------------------------- device apply (lusol) -----------
// forward:  x(i) = y(i) - sum_k L(i,k) x(k)              ilut.f90:462-467
// backward: x(i) = (x(i) - sum_k U(i,k) x(k)) * alu(i)   :471-476
// one thread per row, the row's entries in storage order (the reference's summation order)
template <bool FWD>
__device__ __forceinline__ void ilut_row(int e, const int *__restrict__ ptr, const int *__restrict__ col,
                                         const double *__restrict__ val, const double *__restrict__ piv,
                                         const int *__restrict__ erow, const double *__restrict__ rin,
                                         double *x) {
  const int r = erow[e];
  double acc = FWD ? rin[r] : x[r];
  for (int k = ptr[e]; k < ptr[e + 1]; k++) acc = acc - val[k] * x[col[k]];
  x[r] = FWD ? acc : piv[e] * acc;
}

template <bool FWD>
__global__ void __launch_bounds__(kBlock)
ilut_level_kernel(int first, int count, const int *__restrict__ list, const int *__restrict__ ptr,
                  const int *__restrict__ col, const double *__restrict__ val, const double *__restrict__ piv,
                  const int *__restrict__ erow, const double *__restrict__ rin, double *x,
                  const int *__restrict__ done) {
  if (done && *done) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) ilut_row<FWD>(list[first + t], ptr, col, val, piv, erow, rin, x);
}

// a run of narrow levels (each <= kBlock rows) in ONE CTA: barrier between levels instead of a launch
template <bool FWD>
__global__ void __launch_bounds__(kBlock)
ilut_run_kernel(int first, int nlev, const int *__restrict__ lev_sz, const int *__restrict__ list,
                const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val,
                const double *__restrict__ piv, const int *__restrict__ erow, const double *__restrict__ rin,
                double *x, const int *__restrict__ done) {
  if (done && *done) return;
  int pos = first;
  for (int l = 0; l < nlev; l++) {
    const int sz = lev_sz[l];
    if ((int)threadIdx.x < sz) ilut_row<FWD>(list[pos + threadIdx.x], ptr, col, val, piv, erow, rin, x);
    pos += sz;
    __threadfence_block();
    __syncthreads();
  }
}

int IlutPlan::apply(const double *rin, double *d, const int *done, cudaStream_t s) const {
  int launches = 0;
  for (const Group &g : fgroups) {
    if (g.count == 0) continue;
    if (g.nlev == 1 && g.count > kBlock)
      ilut_level_kernel<true><<<(g.count + kBlock - 1) / kBlock, kBlock, 0, s>>>(
          g.first, g.count, flist.p, lptr.p, lcol.p, lval.p, piv.p, erow.p, rin, d, done);
    else
      ilut_run_kernel<true><<<1, kBlock, 0, s>>>(g.first, g.nlev, flev_sz.p + g.lev_off, flist.p, lptr.p, lcol.p,
                                                 lval.p, piv.p, erow.p, rin, d, done);
    launches++;
  }
  for (const Group &g : bgroups) {
    if (g.count == 0) continue;
    if (g.nlev == 1 && g.count > kBlock)
      ilut_level_kernel<false><<<(g.count + kBlock - 1) / kBlock, kBlock, 0, s>>>(
          g.first, g.count, blist.p, uptr.p, ucol.p, uval.p, piv.p, erow.p, rin, d, done);
    else
      ilut_run_kernel<false><<<1, kBlock, 0, s>>>(g.first, g.nlev, blev_sz.p + g.lev_off, blist.p, uptr.p, ucol.p,
                                                  uval.p, piv.p, erow.p, rin, d, done);
    launches++;
  }
  return launches;
}

}  // namespace mf6
