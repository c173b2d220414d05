// solution.cu -- device-resident GWF formulate (NPF / STO / boundary packages),
// the pre-solve fix-ups and the outer (Picard / Newton) iteration.
//
// Restates on the device (one thread per matrix row, "row-gather" form: every
// row evaluates its own connections, so there are no atomics and the result is
// deterministic; arithmetic per entry in the reference's order):
//   NumericalSolution.f90  solve :1482-1837, sln_buildsystem :1941-1991, sln_reset :2389-2396,
//       sln_ls fix-ups :2434-2573, sln_calc_ptc :2936-2962, sln_calc_residual :2966-2982,
//       sln_calcdx :2912-2932, sln_underrelax :2989-3114, sln_get_dxmax :3122-3153
//   gwf.f90  gwf_ad :396-442, gwf_cf :446-462, gwf_fc :466-555, gwf_ptc :625-687, gwf_cq :741-778,
//       gwf_bd :785-824
//   gwf-npf.f90  npf_cf :444-470, npf_fc :474-574, npf_fn :578-698, npf_nur :705-741,
//       npf_cq/qcalc :745-865, calc_condsat :1950-2037
//   gwf-sto.f90  sto_fc :226-345, sto_fn :353-439, sto_cq :447-564
//   BoundaryPackage.f90  bnd_fc :453-472, bnd_cq_simrate :583-619 ; gwf-{wel,riv,rch,ghb,drn,chd}.f90
//   Budget.f90 :259-267, :631-648 ; Sparse.f90 csr_diagsum :262-281
#include "solver.cuh"
#include "gwf_device.cuh"
#include "spmv.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <nvtx3/nvToolsExt.h>
#include <cstring>
#include <numeric>

namespace mf6 {

struct ModelOpts {
  int icellavg, inewton, inewtonur, iperched, ivarcv, idewatcv, insto;
  int istor_coef, iconf_ss, iorig_ss;
  int all_confined;  // no convertible cell, no perched, no Newton: cond == condsat
  double satomega;
};

// per-cell / per-slot / per-connection device views passed by value to kernels
struct ModelView {
  int n;
  const int *slice_ptr;
  const unsigned char *rowlen;
  const int *col;
  const int *slot_conn;       // (jas << 1) | (neighbour has the higher original index) ; -1 diag/pad
  const double *slot_condsat;
  const double *condsat, *cl1, *cl2, *hwva;
  const int *ihc;
  const double *top, *bot, *area, *k11, *k33, *ss, *sy;
  const int *icelltype, *ibound, *iconvert;
  const double *hyc;  // [2*njas] hy_eff of the lower- / higher-numbered cell along the connection; null = k11 / k33
  ModelOpts o;
};

// conductance of the connection between row r and its neighbour c, evaluated
// with the argument order of the reference (n = lower original index).
__device__ __forceinline__ double conn_cond(const ModelView &M, int r, int c, int up, int jas,
                                            double slot_csat, const double *__restrict__ h,
                                            const double *__restrict__ sat) {
  const int n = up ? r : c, m = up ? c : r;
  const int ihc = M.ihc[jas];
  if (ihc == 0)
    return vcond(M.ibound[n], M.ibound[m], M.icelltype[n], M.icelltype[m], M.o.ivarcv,
                 M.o.idewatcv, slot_csat, h[n], h[m], M.hyc ? M.hyc[2 * jas] : M.k33[n],
                 M.hyc ? M.hyc[2 * jas + 1] : M.k33[m], sat[n], sat[m], M.top[n], M.top[m], M.bot[n], M.bot[m],
                 M.hwva[jas]);
  return hcond(M.ibound[n], M.ibound[m], M.icelltype[n], M.icelltype[m], M.o.inewton, ihc,
               M.o.icellavg, slot_csat, h[n], h[m], sat[n], sat[m], M.hyc ? M.hyc[2 * jas] : M.k11[n],
               M.hyc ? M.hyc[2 * jas + 1] : M.k11[m], M.top[n], M.top[m], M.bot[n], M.bot[m], M.cl1[jas],
               M.cl2[jas], M.hwva[jas]);
}

// calc_condsat (gwf-npf.f90:1950-2037), upper triangle, no THICKSTRT (sat = 1)
// sat0 (may be null = 1): initial saturation of a THICKSTRT cell (calc_initial_sat, gwf-npf.f90:2046-2057)
__global__ void condsat_kernel(int njas, const int *__restrict__ conn_n,
                               const int *__restrict__ conn_m, ModelView M, const double *__restrict__ sat0,
                               double *__restrict__ condsat) {
  for (int jj = blockIdx.x * blockDim.x + threadIdx.x; jj < njas; jj += gridDim.x * blockDim.x) {
    const int n = conn_n[jj], m = conn_m[jj];
    const int ihc = M.ihc[jj];
    const double topn = M.top[n], botn = M.bot[n], topm = M.top[m], botm = M.bot[m];
    const double satn = sat0 ? sat0[n] : 1.0, satm = sat0 ? sat0[m] : 1.0;
    double csat;
    if (ihc == 0)
      csat = vcond(1, 1, 1, 1, 1, 1, 1.0, botn, botm, M.hyc ? M.hyc[2 * jj] : M.k33[n],
                   M.hyc ? M.hyc[2 * jj + 1] : M.k33[m], satn, satm, topn, topm, botn, botm, M.hwva[jj]);
    else
      csat = hcond(1, 1, 1, 1, 0, ihc, M.o.icellavg, 1.0, topn, topm, satn, satm,
                   M.hyc ? M.hyc[2 * jj] : M.k11[n], M.hyc ? M.hyc[2 * jj + 1] : M.k11[m], topn, topm, botn, botm,
                   M.cl1[jj], M.cl2[jj], M.hwva[jj]);
    condsat[jj] = csat;
  }
}

__global__ void slot_condsat_kernel(long long nslots, const int *__restrict__ slot_conn,
                                    const double *__restrict__ condsat,
                                    double *__restrict__ slot_condsat) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < nslots;
       s += (long long)gridDim.x * blockDim.x) {
    const int c = slot_conn[s];
    slot_condsat[s] = (c >= 0) ? condsat[c >> 1] : 0.0;
  }
}

// ---- HFB, horizontal flow barriers (gwf-hfb.f90) ------------------------------------------------------------------
struct HfbView {
  int nhfb;
  const int *rn, *rm;        // final rows of noden / nodem
  const int *jas;            // symmetric connection
  const int *slot_nm, *slot_mn;  // SELL slots of the entries (n, m) and (m, n)
  const double *hydchr;
  double *csatsav, *condsav;
};

__device__ __forceinline__ double hfb_faheight(const ModelView &M, int n, int m, int jj, const double *h) {
  double topn = M.top[n], topm = M.top[m];
  const double botn = M.bot[n], botm = M.bot[m];
  if (h) {
    if (M.icelltype[n] != 0 && h[n] < topn) topn = h[n];
    if (M.icelltype[m] != 0 && h[m] < topm) topm = h[m];
  }
  if (M.ihc[jj] == 2) return fmin(topn, topm) - fmax(botn, botm);
  return 0.5 * ((topn - botn) + (topm - botm));
}

// condsat_modify (gwf-hfb.f90:789-832) / condsat_reset (:770-781): one thread per barrier
__global__ void hfb_condsat_kernel(HfbView H, ModelView M, double *__restrict__ condsat,
                                   double *__restrict__ slot_condsat, int reset) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H.nhfb; i += gridDim.x * blockDim.x) {
    const int jj = H.jas[i], n = H.rn[i], m = H.rm[i];
    double cond;
    if (reset) {
      cond = H.csatsav[i];
    } else {
      cond = condsat[jj];
      H.csatsav[i] = cond;
      if (M.o.inewton == 1 || (M.icelltype[n] == 0 && M.icelltype[m] == 0)) {
        if (H.hydchr[i] > 0.0) {
          const double condhfb = H.hydchr[i] * M.hwva[jj] * hfb_faheight(M, n, m, jj, nullptr);
          cond = cond * condhfb / (cond + condhfb);
        } else {
          cond = -cond * H.hydchr[i];
        }
      }
    }
    condsat[jj] = cond;
    slot_condsat[H.slot_nm[i]] = cond;
    slot_condsat[H.slot_mn[i]] = cond;
  }
}

// hfb_fc without XT3D (gwf-hfb.f90:296-345), Picard with a convertible cell on either side: one thread per DISTINCT
// row walks the barriers of that row in barrier order (the reference's loop order): the row's own off-diagonal entry
// becomes cond, its diagonal gains aterm - cond.  Both rows of a barrier compute the same cond from the same inputs.
__global__ void hfb_fc_kernel(int nrows, const int *__restrict__ ev_row, const int *__restrict__ ev_ptr,
                              const int *__restrict__ ev_hfb, HfbView H, ModelView M, const double *__restrict__ h,
                              double *__restrict__ val) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nrows; e += gridDim.x * blockDim.x) {
    const int row = ev_row[e];
    const long long dslot = (long long)M.slice_ptr[row >> 5] + (row & 31);
    double diag = val[dslot];
    for (int q = ev_ptr[e]; q < ev_ptr[e + 1]; q++) {
      const int i = ev_hfb[q];
      const int n = H.rn[i], m = H.rm[i], jj = H.jas[i];
      if (M.ibound[n] == 0 || M.ibound[m] == 0) continue;
      if (M.icelltype[n] == 0 && M.icelltype[m] == 0) continue;
      const int slot = (row == n) ? H.slot_nm[i] : H.slot_mn[i];
      const double aterm = val[slot];
      double cond;
      if (H.hydchr[i] > 0.0) {
        const double condhfb = H.hydchr[i] * 1.0 * M.hwva[jj] * hfb_faheight(M, n, m, jj, h);
        cond = aterm * condhfb / (aterm + condhfb);
      } else {
        cond = -aterm * H.hydchr[i];
      }
      if (row == n) H.condsav[i] = cond;
      diag = diag + (aterm - cond);
      val[slot] = cond;
    }
    val[dslot] = diag;
  }
}

// hfb_cq without XT3D (gwf-hfb.f90:432-449)
__global__ void hfb_cq_kernel(HfbView H, ModelView M, const double *__restrict__ h, double *__restrict__ flowja) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H.nhfb; i += gridDim.x * blockDim.x) {
    const int n = H.rn[i], m = H.rm[i];
    if (M.ibound[n] == 0 || M.ibound[m] == 0) continue;
    if (M.icelltype[n] == 0 && M.icelltype[m] == 0) continue;
    const double qnm = H.condsav[i] * (h[m] - h[n]);
    flowja[H.slot_nm[i]] = qnm;
    flowja[H.slot_mn[i]] = -qnm;
  }
}

// ---- GNC, ghost node correction, EXPLICIT variant (src/Exchange/GhostNode.f90) -------------------------------------
struct GncView {
  int ngnc, numj;
  const int *rn, *rm;            // final rows of noden / nodem
  const int *jas;                // symmetric connection (n, m)
  const int *slot_nm, *slot_mn;  // SELL slots of the entries (n, m) and (m, n)
  const int *rj;                 // [ngnc * numj] final rows of the contributing cells, < 0 = none
  const double *alpha;           // [ngnc * numj]
  double *cond;                  // conductance of (n, m) saved by the last formulate (gnc_fmsav :251-273)
};

// explicit branch of gnc_fc (:280-324): one thread per DISTINCT row walks the corrections of that row in list order;
// row n loses rterm = alpha cond (h_n - h_j), row m gains it.  Both read the (n, m) entry npf_fc / hfb_fc left.
__global__ void gnc_fc_kernel(int nrows, const int *__restrict__ ev_row, const int *__restrict__ ev_ptr,
                              const int *__restrict__ ev_gnc, GncView Gv, ModelView M, const double *__restrict__ h,
                              const double *__restrict__ val, double *__restrict__ rhs) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nrows; e += gridDim.x * blockDim.x) {
    const int row = ev_row[e];
    double r = rhs[row];
    for (int q = ev_ptr[e]; q < ev_ptr[e + 1]; q++) {
      const int i = ev_gnc[q];
      const int n = Gv.rn[i], m = Gv.rm[i];
      const double cond = val[Gv.slot_nm[i]];
      if (row == n) Gv.cond[i] = cond;
      if (M.ibound[n] == 0 || M.ibound[m] == 0) continue;
      for (int k = 0; k < Gv.numj; k++) {
        const int j = Gv.rj[i * Gv.numj + k];
        if (j < 0) continue;
        const double alpha = Gv.alpha[i * Gv.numj + k];
        if (alpha == 0.0) continue;
        const double aterm = alpha * cond;
        const double rterm = aterm * (h[n] - h[j]);
        r = (row == n) ? (r - rterm) : (r + rterm);
      }
    }
    rhs[row] = r;
  }
}

// gnc_fn (:340-443), single-model arguments (gwf.f90:521-528): row n owns its diagonal and the (n, m) entry, row m
// its diagonal and (m, n)
__global__ void gnc_fn_kernel(int nrows, const int *__restrict__ ev_row, const int *__restrict__ ev_ptr,
                              const int *__restrict__ ev_gnc, GncView Gv, ModelView M, const double *__restrict__ h,
                              double *__restrict__ val, double *__restrict__ rhs) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nrows; e += gridDim.x * blockDim.x) {
    const int row = ev_row[e];
    const long long dslot = (long long)M.slice_ptr[row >> 5] + (row & 31);
    double r = rhs[row], diag = val[dslot];
    for (int q = ev_ptr[e]; q < ev_ptr[e + 1]; q++) {
      const int i = ev_gnc[q];
      const int n = Gv.rn[i], m = Gv.rm[i], jj = Gv.jas[i];
      if (M.ibound[n] == 0 || M.ibound[m] == 0) continue;
      const int ihc = M.ihc[jj];
      if (ihc == 0 && M.o.ivarcv == 0) continue;
      const int iups = (h[m] > h[n]) ? 1 : 0;
      const int up = iups ? m : n;
      if (M.icelltype[up] == 0) continue;
      double topup = M.top[up], botup = M.bot[up];
      const double xup = h[up];
      if (ihc == 2) {
        topup = fmin(M.top[n], M.top[m]);
        botup = fmax(M.bot[n], M.bot[m]);
      }
      const double csat = M.condsat[jj];
      for (int k = 0; k < Gv.numj; k++) {
        const int j = Gv.rj[i * Gv.numj + k];
        if (j < 0) continue;
        if (M.ibound[j] == 0) continue;
        const double alpha = Gv.alpha[i * Gv.numj + k];
        if (alpha == 0.0) continue;
        const double consterm = csat * alpha * (h[n] - h[j]);
        const double derv = sQuadraticSaturationDerivative(topup, botup, xup, 1.0e-6);
        const double term = consterm * derv;
        if (row == n) {
          if (iups == 0) {
            diag = diag + term;
            r = r + term * h[n];
          } else {
            if (M.ibound[n] > 0) val[Gv.slot_nm[i]] = val[Gv.slot_nm[i]] + term;
            r = r + term * h[m];
          }
        } else {
          if (iups == 0) {
            if (M.ibound[m] > 0) val[Gv.slot_mn[i]] = val[Gv.slot_mn[i]] + (-term);
            r = r - term * h[n];
          } else {
            diag = diag + (-term);
            r = r - term * h[m];
          }
        }
      }
    }
    rhs[row] = r;
    val[dslot] = diag;
  }
}

// gnc_cq (:478-503) with deltaQgnc (:509-542)
__global__ void gnc_cq_kernel(GncView Gv, ModelView M, const double *__restrict__ h, double *__restrict__ flowja) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Gv.ngnc; i += gridDim.x * blockDim.x) {
    const int n = Gv.rn[i], m = Gv.rm[i];
    double dq = 0.0;
    if (M.ibound[n] != 0 && M.ibound[m] != 0) {
      double sigalj = 0.0, hd = 0.0;
      for (int k = 0; k < Gv.numj; k++) {
        const int j = Gv.rj[i * Gv.numj + k];
        if (j < 0) continue;
        if (M.ibound[j] == 0) continue;
        const double alpha = Gv.alpha[i * Gv.numj + k];
        sigalj = sigalj + alpha;
        hd = hd + alpha * h[j];
      }
      const double aterm = sigalj * h[n] - hd;
      dq = aterm * Gv.cond[i];
    }
    flowja[Gv.slot_nm[i]] = flowja[Gv.slot_nm[i]] + dq;
    flowja[Gv.slot_mn[i]] = flowja[Gv.slot_mn[i]] - dq;
  }
}

// sgwf_npf_wetdry without rewetting (gwf-npf.f90:2061-2158), the first thing npf_cf does for a model without
// NEWTON: a convertible cell whose saturated thickness is gone becomes inactive for good (ibound0 keeps it so in
// later stress periods), its head the dry value; a constant-head cell going dry is fatal (flag)
__global__ void npf_wd_kernel(ModelView M, double *__restrict__ x, int *__restrict__ ibound,
                              int *__restrict__ ibound0, int *__restrict__ flag) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M.n; r += gridDim.x * blockDim.x) {
    const int ib = ibound[r];
    if (ib == 0 || M.icelltype[r] == 0) continue;
    double ttop = M.top[r];
    if (x[r] < ttop) ttop = x[r];
    if (ttop - M.bot[r] <= 0.0) {
      if (ib < 0) {
        *flag = 1;
        continue;
      }
      x[r] = -1.0e30;  // DHDRY
      ibound[r] = 0;
      ibound0[r] = 0;
    }
  }
}

// ---- rewetting (rewet_check, gwf-npf.f90:2167-2223, driven by the first loop of sgwf_npf_wetdry :2096-2107) ----------
// The reference sweeps the cells in their natural order and a cell wetted earlier in the sweep (ibound 30000) already
// counts as a wet neighbour, i.e. the outcome for cell n depends on the outcome for its lower-numbered dry neighbours.
// Reproduced exactly in passes: a candidate (dry, WETDRY != 0) is evaluated in the first pass in which all its
// lower-numbered neighbouring candidates have been decided in EARLIER passes; its neighbours are visited in ascending
// original number like the reference's ja order and the first one that qualifies wets it.  Higher-numbered
// neighbours are seen in their pre-sweep state (a 30000 there means "was dry").  state: 0 undecided, else the pass
// in which the cell was decided (1 = not a candidate).
__global__ void rewet_begin_kernel(int n, const double *__restrict__ wetdry, const int *__restrict__ ibound,
                                   int *__restrict__ state) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    state[r] = (ibound[r] == 0 && wetdry[r] != 0.0) ? 0 : 1;
}

__global__ void rewet_pass_kernel(ModelView M, const int *__restrict__ orig, const double *__restrict__ wetdry,
                                  double *x, int *ibound, int *state, int pass, double wetfct, int ihdwet,
                                  int *__restrict__ pending) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M.n; r += gridDim.x * blockDim.x) {
    if (state[r] != 0) continue;
    const long long base = (long long)M.slice_ptr[r >> 5] + (r & 31);
    const int len = M.rowlen[r], o_r = orig[r];
    // neighbours in ascending original number (insertion sort of at most 15 entries; longer rows keep the tail order)
    int nb[16], no[16], nh[16], cnt = 0;
    bool wait = false;
    for (int k = 1; k < len && cnt < 16; k++) {
      const int c = M.col[base + 32LL * k];
      if (c >= M.n) continue;  // halo column (not on this path)
      const int sc = M.slot_conn[base + 32LL * k];
      if (sc < 0) continue;
      const int o_c = orig[c];
      if (o_c < o_r) {
        const int stc = state[c];
        if (stc == 0 || stc >= pass) wait = true;  // a lower-numbered candidate is not decided yet
      }
      int q = cnt++;
      while (q > 0 && no[q - 1] > o_c) {
        nb[q] = nb[q - 1];
        no[q] = no[q - 1];
        nh[q] = nh[q - 1];
        q--;
      }
      nb[q] = c;
      no[q] = o_c;
      nh[q] = M.ihc[sc >> 1];
    }
    if (wait) {
      *pending = 1;
      continue;
    }
    const double bbot = M.bot[r], wd = wetdry[r];
    const double awd = (wd < 0.0) ? -wd : wd;
    const double turnon = bbot + awd;
    for (int q = 0; q < cnt; q++) {
      const int c = nb[q];
      int ibd = ibound[c];
      if (no[q] > o_r && ibd == 30000) ibd = 0;  // not reached yet by the reference's sweep: still dry
      const double hm = x[c];
      bool wet = false;
      if (nh[q] == 0)
        wet = (ibd > 0 && hm >= turnon);
      else if (wd > 0.0)
        wet = (ibd > 0 && hm >= turnon);
      if (wet) {
        x[r] = (ihdwet == 0) ? bbot + wetfct * (hm - bbot) : bbot + wetfct * awd;
        ibound[r] = 30000;
        break;
      }
    }
    state[r] = pass;
  }
}

// last loop of sgwf_npf_wetdry (:2150-2153) + what the next chd_rp restores (ibound0 = the cell's own state)
__global__ void rewet_end_kernel(int n, int *__restrict__ ibound, int *__restrict__ ibound0) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    int ib = ibound[r];
    if (ib == 30000) ibound[r] = ib = 1;
    if (ib >= 0) ibound0[r] = ib;
  }
}

// npf_ad (gwf-npf.f90:393-408): a dry wettable cell starts the time step with hold = bottom and hnew = HDRY
__global__ void npf_ad_rewet_kernel(int n, const double *__restrict__ wetdry, const int *__restrict__ ibound,
                                    const double *__restrict__ bot, double *__restrict__ x,
                                    double *__restrict__ xold) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    if (wetdry[r] == 0.0 || ibound[r] != 0) continue;
    xold[r] = bot[r];
    x[r] = -1.0e30;
  }
}

// npf_cf (gwf-npf.f90:444-470) + thksat (:775-794)
__global__ void npf_cf_kernel(ModelView M, const double *__restrict__ h, double *__restrict__ sat) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M.n; r += gridDim.x * blockDim.x) {
    if (M.icelltype[r] != 0) {
      double satn;
      if (M.ibound[r] == 0) {
        satn = 0.0;
      } else {
        const double hn = h[r];
        if (hn >= M.top[r])
          satn = 1.0;
        else
          satn = (hn - M.bot[r]) / (M.top[r] - M.bot[r]);
        if (M.o.inewton != 0) satn = sQuadraticSaturation(M.top[r], M.bot[r], hn, M.o.satomega);
      }
      sat[r] = satn;
    }
  }
}

// ---- boundary packages: all bounds of all packages concatenated -------------
struct BndView {
  int nb;
  const unsigned char *type;
  const unsigned char *flag;  // WEL: iflowred
  const int *node;            // final numbering
  const double *b1, *b2, *b3, *fred;
  double *hcof, *rhs, *simvals, *ratein, *rateout;
  int *eff;                   // cell the bound acts on (RCH without FIXED_CELL: the highest active cell)
  const int *below;           // cell under every cell (final numbering, -1 = bottom), null when no cell can be inactive
};

// *_cf of WEL/RIV/RCH/GHB/DRN (gwf-wel.f90:296-332, gwf-riv.f90:270-299, gwf-rch.f90:303-353,
// gwf-ghb.f90:245-265, gwf-drn.f90:340-373 incl. the drainage-depth scaling); CHD has no cf terms
__global__ void bnd_cf_kernel(BndView B, ModelView M, const double *__restrict__ x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B.nb; i += gridDim.x * blockDim.x) {
    const int node = B.node[i];
    const int ib = M.ibound[node];
    double hcof = 0.0, rhs = 0.0;
    int eff = node;
    switch (B.type[i]) {
      case MF6GPU_PKG_WEL:
        if (ib > 0) {
          double q = B.b1[i];
          if (B.flag[i] != 0 && q < 0.0 && M.icelltype[node] != 0) {
            double tp = M.top[node];
            const double bt = M.bot[node], thick = tp - bt;
            tp = bt + B.fred[i] * thick;
            q = q * sQSaturation(tp, bt, x[node]);
          }
          rhs = -q;
        }
        break;
      case MF6GPU_PKG_RIV:
        if (ib > 0) {
          const double hriv = B.b1[i], criv = B.b2[i], rbot = B.b3[i];
          if (x[node] <= rbot) {
            rhs = -criv * (hriv - rbot);
            hcof = 0.0;
          } else {
            rhs = -criv * hriv;
            hcof = -criv;
          }
        }
        break;
      case MF6GPU_PKG_RCH: {
        // rch_cf (gwf-rch.f90:303-353); flag = FIXED_CELL.  Otherwise an inactive cell hands its recharge down
        // the column to the first cell that is not inactive (highest_active, DiscretizationBase.f90:1077-1112)
        int nd = node;
        if (B.flag[i] == 0 && B.below && ib == 0) {
          for (;;) {
            const int b = B.below[nd];
            if (b < 0) break;
            nd = b;
            if (M.ibound[nd] != 0) break;
          }
        }
        eff = nd;
        rhs = -B.b1[i] * M.area[nd];
        if (M.ibound[nd] <= 0) rhs = 0.0;
        break;
      }
      case MF6GPU_PKG_GHB:
        if (ib > 0) {
          hcof = -B.b2[i];
          rhs = -B.b2[i] * B.b1[i];
        }
        break;
      case MF6GPU_PKG_DRN:
        // drn_cf + get_drain_factor (gwf-drn.f90:340-373, 534-574): b3 = drainage depth (DDRN auxiliary), flag =
        // DEV_CUBIC_SCALING
        if (ib > 0) {
          const double cdrn = B.b2[i], drndepth = B.b3[i];
          double drntop, drnbot, fact;
          drain_elevations(B.b1[i], drndepth, drntop, drnbot);
          if (drndepth != 0.0) {
            if (B.flag[i] != 0)
              fact = sQSaturationC(drntop, drnbot, x[node], -1.0, 2.0);
            else
              fact = sQuadraticSaturation(drntop, drnbot, x[node], 0.0);
          } else {
            fact = (x[node] <= drnbot) ? 0.0 : 1.0;
          }
          rhs = -fact * cdrn * drnbot;
          hcof = -fact * cdrn;
        }
        break;
      default:
        break;
    }
    B.hcof[i] = hcof;
    B.rhs[i] = rhs;
    B.eff[i] = eff;
  }
}

// bnd_fc / bnd_fn / bnd_cq scatter: one thread per DISTINCT node walks that
// node's bounds in (package, bound) order -> same accumulation order as the
// reference's package loop, no atomics.
// mode 0: bnd_fc (BoundaryPackage.f90:453-472) ; 1: wel_fn (gwf-wel.f90:378-424) + drn_fn (gwf-drn.f90:420-470) ;
// mode 2: bnd_cq_simrate (:583-619)
__global__ void bnd_scatter_kernel(int nseg, const int *__restrict__ seg_node,
                                   const int *__restrict__ seg_ptr, const int *__restrict__ seg_idx,
                                   BndView B, ModelView M, const double *__restrict__ x,
                                   double *__restrict__ val, double *__restrict__ rhsv,
                                   double *__restrict__ flowja, int mode) {
  for (int sidx = blockIdx.x * blockDim.x + threadIdx.x; sidx < nseg; sidx += gridDim.x * blockDim.x) {
    const int node = seg_node[sidx];
    const long long dslot = (long long)M.slice_ptr[node >> 5] + (node & 31);
    const int ib = M.ibound[node];
    if (mode == 0) {
      double diag = val[dslot], r = rhsv[node];
      for (int e = seg_ptr[sidx]; e < seg_ptr[sidx + 1]; e++) {
        const int i = seg_idx[e];
        if (B.type[i] == MF6GPU_PKG_CHD) continue;  // chd_fc is a no-op (gwf-chd.f90:238-246)
        if (B.eff[i] != node) continue;              // recharge handed down the column: rch_moved_kernel
        r = r + B.rhs[i];
        diag = diag + B.hcof[i];
      }
      val[dslot] = diag;
      rhsv[node] = r;
    } else if (mode == 1) {
      if (ib <= 0) continue;
      double diag = val[dslot], r = rhsv[node];
      for (int e = seg_ptr[sidx]; e < seg_ptr[sidx + 1]; e++) {
        const int i = seg_idx[e];
        if (B.type[i] == MF6GPU_PKG_DRN) {  // drn_fn (gwf-drn.f90:420-470)
          const double drndepth = B.b3[i];
          if (drndepth != 0.0) {
            double drntop, drnbot;
            drain_elevations(B.b1[i], drndepth, drntop, drnbot);
            double drterm = sQSaturationDerivativeC(drntop, drnbot, x[node], -1.0, 2.0);
            drterm = drterm * B.b2[i] * (drnbot - x[node]);
            diag = diag + drterm;
            r = r + drterm * x[node];
          }
          continue;
        }
        if (B.type[i] != MF6GPU_PKG_WEL) continue;
        if (B.flag[i] != 0 && M.icelltype[node] != 0) {
          const double q = -B.rhs[i];
          if (q < 0.0) {
            double tp = M.top[node];
            const double bt = M.bot[node], thick = tp - bt;
            tp = bt + B.fred[i] * thick;
            double drterm = sQSaturationDerivative(tp, bt, x[node]);
            drterm = drterm * B.b1[i];
            diag = diag + drterm;
            r = r + drterm * x[node];
          }
        }
      }
      val[dslot] = diag;
      rhsv[node] = r;
    } else {
      double fd = flowja[dslot];
      for (int e = seg_ptr[sidx]; e < seg_ptr[sidx + 1]; e++) {
        const int i = seg_idx[e];
        if (B.type[i] == MF6GPU_PKG_CHD) continue;
        if (B.eff[i] != node) continue;
        double rrate = 0.0;
        if (ib > 0) rrate = B.hcof[i] * x[node] - B.rhs[i];
        fd = fd + rrate;
        B.simvals[i] = rrate;
      }
      flowja[dslot] = fd;
    }
  }
}

// recharge that rch_cf handed down to another cell than the listed one (rare: dry or inactive top cells): its
// rhs / rate goes to the row of the cell it acts on.  hcof of RCH is 0.  Distinct columns give distinct targets;
// bounds of several RCH packages meeting in one cell are summed with atomicAdd.
// mode 0: bnd_fc ; mode 2: bnd_cq_simrate
__global__ void rch_moved_kernel(BndView B, ModelView M, double *__restrict__ rhsv, double *__restrict__ flowja,
                                 int mode) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B.nb; i += gridDim.x * blockDim.x) {
    const int nd = B.eff[i];
    if (B.type[i] != MF6GPU_PKG_RCH || nd == B.node[i]) continue;
    if (mode == 0) {
      atomicAdd(&rhsv[nd], B.rhs[i]);
    } else {
      const double rrate = (M.ibound[nd] > 0) ? -B.rhs[i] : 0.0;
      atomicAdd(&flowja[(long long)M.slice_ptr[nd >> 5] + (nd & 31)], rrate);
      B.simvals[i] = rrate;
    }
  }
}

// chd_ad (gwf-chd.f90:175-197): x(node) = head ; xold(node) = x(node)
__global__ void chd_ad_kernel(BndView B, double *__restrict__ x, double *__restrict__ xold) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B.nb; i += gridDim.x * blockDim.x) {
    if (B.type[i] != MF6GPU_PKG_CHD) continue;
    const int node = B.node[i];
    x[node] = B.b1[i];
    xold[node] = B.b1[i];
  }
}

// chd_rp (gwf-chd.f90:143-155): ibound(node) = -ibcnum
__global__ void chd_ibound_kernel(BndView B, const int *__restrict__ pkgid, int *__restrict__ ibound) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B.nb; i += gridDim.x * blockDim.x)
    if (B.type[i] == MF6GPU_PKG_CHD) ibound[B.node[i]] = -(pkgid[i] + 1);
}

// chd_rp, first loop (gwf-chd.f90:134-138): "Reset previous CHDs to active cell"
__global__ void chd_release_kernel(BndView B, int *__restrict__ ibound0) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B.nb; i += gridDim.x * blockDim.x)
    if (B.type[i] == MF6GPU_PKG_CHD) ibound0[B.node[i]] = 1;
}

// calc_chd_rate (gwf-chd.f90:264-320)
__global__ void chd_rate_kernel(BndView B, ModelView M, double *__restrict__ flowja) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B.nb; i += gridDim.x * blockDim.x) {
    if (B.type[i] != MF6GPU_PKG_CHD) continue;
    const int node = B.node[i];
    const long long base = (long long)M.slice_ptr[node >> 5] + (node & 31);
    const int len = M.rowlen[node];
    double rate = 0.0, ratein = 0.0, rateout = 0.0;
    for (int k = 1; k < len; k++) {
      const double q = flowja[base + 32LL * k];
      rate = rate - q;
      const int n2 = M.col[base + 32LL * k];
      if (M.ibound[n2] > 0) {
        if (q < 0.0)
          ratein = ratein - q;
        else
          rateout = rateout + q;
      }
    }
    B.rhs[i] = -rate;
    B.hcof[i] = 0.0;
    B.simvals[i] = rate;
    B.ratein[i] = ratein;
    B.rateout[i] = rateout;
    flowja[base] = flowja[base] + rate;
  }
}

// ---- sln_reset + npf_fc + sto_fc, one thread per row --------------------------
template <bool CONFINED>
__global__ void __launch_bounds__(kBlock)
assemble_rows_kernel(ModelView M, const double *__restrict__ h, const double *__restrict__ hold,
                     const double *__restrict__ sat, double *__restrict__ val,
                     double *__restrict__ rhsv, int transient, double tled) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M.n; r += gridDim.x * blockDim.x) {
    const int len = M.rowlen[r];
    const long long base = (long long)M.slice_ptr[r >> 5] + (r & 31);
    const int ibr = M.ibound[r];
    double diag = 0.0, rhs = 0.0;
    for (int k = 1; k < len; k++) {
      const long long slot = base + 32LL * k;
      const int c = __ldg(M.col + slot);
      const double csat = __ldg(M.slot_condsat + slot);
      double cond;
      if (CONFINED) {
        cond = (ibr == 0 || M.ibound[c] == 0) ? 0.0 : csat;
        val[slot] = cond;
        diag = diag + (-cond);
      } else {
        const int conn = __ldg(M.slot_conn + slot);
        const int up = conn & 1, jas = conn >> 1;
        cond = conn_cond(M, r, c, up, jas, csat, h, sat);
        bool perched = false;
        if (M.o.iperched != 0 && M.ihc[jas] == 0) {
          const int m = up ? c : r;  // lower cell of the vertical pair has the higher index
          if (M.icelltype[m] != 0 && h[m] < M.top[m]) perched = true;
        }
        if (perched) {  // gwf-npf.f90:523-540
          const int n = up ? r : c;
          if (up) {
            rhs = rhs - cond * M.bot[n];
            diag = diag + (-cond);
            val[slot] = 0.0;
          } else {
            val[slot] = cond;
            rhs = rhs + cond * M.bot[n];
          }
        } else {
          val[slot] = cond;
          diag = diag + (-cond);
        }
      }
    }
    // sto_fc (gwf-sto.f90:226-345)
    if (transient && ibr >= 1) {
      const double tp = M.top[r], bt = M.bot[r];
      const int icv = M.iconvert[r];
      double snold = 1.0, snnew = 1.0;
      if (icv != 0) {
        snold = sQuadraticSaturation(tp, bt, hold[r], M.o.satomega);
        snnew = sQuadraticSaturation(tp, bt, h[r], M.o.satomega);
      }
      const double sc1 = SsCapacity(M.o.istor_coef, tp, bt, M.area[r], M.ss[r]);
      const double rho1 = sc1 * tled;
      double aterm, rhsterm, rate;
      SsTerms(icv, M.o.iorig_ss, M.o.iconf_ss, tp, bt, rho1, rho1, snnew, snold, h[r], hold[r],
              aterm, rhsterm, rate);
      diag = diag + aterm;
      rhs = rhs + rhsterm;
      if (icv != 0) {
        rhsterm = 0.0;
        const double sc2 = M.sy[r] * M.area[r];
        const double rho2 = sc2 * tled;
        SyTerms(tp, bt, rho2, rho2, snnew, snold, aterm, rhsterm, rate);
        diag = diag + aterm;
        rhs = rhs + rhsterm;
      }
    }
    val[base] = diag;
    rhsv[r] = rhs;
  }
}

// ---- npf_fn + sto_fn, one thread per row ----------------------------------------
__global__ void __launch_bounds__(kBlock)
newton_rows_kernel(ModelView M, const double *__restrict__ h, double *__restrict__ val,
                   double *__restrict__ rhsv, int transient, double tled) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M.n; r += gridDim.x * blockDim.x) {
    const int len = M.rowlen[r];
    const long long base = (long long)M.slice_ptr[r >> 5] + (r & 31);
    const int ibr = M.ibound[r];
    double diag = val[base], rhs = rhsv[r];
    for (int k = 1; k < len; k++) {
      const long long slot = base + 32LL * k;
      const int c = M.col[slot];
      const int conn = M.slot_conn[slot];
      const int up = conn & 1, jas = conn >> 1;
      const int ihc = M.ihc[jas];
      if (ihc == 0 && M.o.ivarcv == 0) continue;
      const int n = up ? r : c, m = up ? c : r;
      int iups = m;
      if (h[m] < h[n]) iups = n;
      const int idn = (iups == n) ? m : n;
      if (M.icelltype[iups] == 0) continue;
      double topup = M.top[iups], botup = M.bot[iups];
      if (ihc == 2) {
        topup = fmin(M.top[n], M.top[m]);
        botup = fmax(M.bot[n], M.bot[m]);
      }
      const double cond = M.condsat[jas];
      const double consterm = -cond * (h[iups] - h[idn]);
      const double derv = sQuadraticSaturationDerivative(topup, botup, h[iups], M.o.satomega);
      if (iups == n) {
        const double term = consterm * derv;
        if (up) {  // this row is n
          rhs = rhs + term * h[n];
          diag = diag + term;
        } else {   // this row is m
          rhs = rhs - term * h[n];
          if (ibr > 0) val[slot] = val[slot] + (-term);
        }
      } else {
        const double term = -consterm * derv;
        if (up) {
          rhs = rhs + term * h[m];
          if (ibr > 0) val[slot] = val[slot] + term;
        } else {
          rhs = rhs - term * h[m];
          diag = diag + (-term);
        }
      }
    }
    // sto_fn (gwf-sto.f90:353-439); the smoothing calls there use the default eps = 1e-6
    if (transient && M.o.insto && ibr > 0) {
      const double tp = M.top[r], bt = M.bot[r], tthk = tp - bt, hh = h[r];
      const double snnew = sQuadraticSaturation(tp, bt, hh, 1.0e-6);
      const double sc1 = SsCapacity(M.o.istor_coef, tp, bt, M.area[r], M.ss[r]);
      const double sc2 = M.sy[r] * M.area[r];
      const double rho1 = sc1 * tled, rho2 = sc2 * tled;
      if (M.iconvert[r] != 0) {
        const double derv = sQuadraticSaturationDerivative(tp, bt, hh, 1.0e-6);
        double drterm;
        if (M.o.iconf_ss == 0) {
          if (M.o.iorig_ss == 0)
            drterm = -rho1 * derv * (hh - bt) + rho1 * tthk * snnew * derv;
          else
            drterm = -(rho1 * derv * hh);
          diag = diag + drterm;
          rhs = rhs + drterm * hh;
        }
        if (snnew < 1.0) {
          if (snnew > 0.0) {
            const double rterm = -rho2 * tthk * snnew;
            drterm = -rho2 * tthk * derv;
            diag = diag + (drterm + rho2);
            rhs = rhs - rterm + drterm * hh + rho2 * bt;
          }
        }
      }
    }
    val[base] = diag;
    rhsv[r] = rhs;
  }
}

// ---- pre-solve fix-ups of sln_ls (NumericalSolution.f90:2434-2478) -----------
__global__ void __launch_bounds__(kBlock)
ls_fixup_kernel(ModelView M, const double *__restrict__ x, double *__restrict__ xtemp,
                double *__restrict__ val, double *__restrict__ rhsv, int isymmetric) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M.n; r += gridDim.x * blockDim.x) {
    const int len = M.rowlen[r];
    const long long base = (long long)M.slice_ptr[r >> 5] + (r & 31);
    const double xr = x[r];
    xtemp[r] = xr;
    if (M.ibound[r] > 0) {
      double rhs = rhsv[r];
      bool touched = false;
      if (fabs(val[base]) < 1.0e-15) {
        val[base] = -1.0;
        rhs = rhs + (-1.0) * xr;
        touched = true;
      }
      if (isymmetric) {
        for (int k = 1; k < len; k++) {
          const long long slot = base + 32LL * k;
          const int c = M.col[slot];
          if (M.ibound[c] < 0) {
            rhs = rhs - (val[slot] * x[c]);
            val[slot] = 0.0;
            touched = true;
          }
        }
      }
      if (touched) rhsv[r] = rhs;
    } else {
      val[base] = 1.0;
      for (int k = 1; k < len; k++) val[base + 32LL * k] = 0.0;
      rhsv[r] = xr;
    }
  }
}

// NB: the reference walks a row in CSR order (ascending column) when it moves
// constant-head columns to the right-hand side; the slots are in that order too.

// ---- outer-iteration vector kernels ---------------------------------------------
// every field is a double so that the record can be all-gathered across ranks as is
struct OuterState {
  double hncg;      // signed largest |x - xtemp|
  double loc;       // device row of it (single GPU)
  double loc_ord;   // tie-break / global cell id of it
  double nur_flag;
  double dxold_max;
  double ptc_max;   // max |r| / volume
  double l2;        // sum r^2
  double rin, rout;
  double pad;
};

// sln_get_dxmax (:3122-3153)
__global__ void __launch_bounds__(kBlock)
dxmax_kernel(int n, const double *__restrict__ x, const double *__restrict__ xtemp,
             const int *__restrict__ ibound, const int *__restrict__ ord, MaxLoc *__restrict__ pm,
             unsigned int *ticket, OuterState *os) {
  __shared__ MaxLoc shm[8];
  __shared__ bool last;
  MaxLoc m = maxloc_init();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (ibound[i] < 1) continue;
    const double hdif = x[i] - xtemp[i];
    const double a = fabs(hdif);
    if (a >= m.a && a > 0.0) maxloc_take(m, hdif, ord ? ord[i] : i, i);
  }
  m = block_maxloc(m, shm);
  if (threadIdx.x == 0) pm[blockIdx.x] = m;
  if (last_block(ticket, &last)) {
    MaxLoc g = maxloc_init();
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) maxloc_merge(g, pm[i]);
    g = block_maxloc(g, shm);
    if (threadIdx.x == 0) {
      os->hncg = g.v;
      os->loc = (double)g.idx;
      os->loc_ord = (double)g.ord;
    }
  }
}

// sln_calcdx (:2912-2932)
__global__ void calcdx_kernel(int n, const int *__restrict__ ibound, const double *__restrict__ x,
                              const double *__restrict__ xtemp, double *__restrict__ dx) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    dx[i] = (ibound[i] < 1) ? 0.0 : x[i] - xtemp[i];
}

// sln_underrelax simple / cooley (:3010-3064): x = xtemp + f * (x - xtemp)
__global__ void relax_kernel(int n, const int *__restrict__ ibound, double *__restrict__ x,
                             const double *__restrict__ xtemp, double *__restrict__ dxold, double f) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (ibound[i] < 1) continue;
    const double delx = x[i] - xtemp[i];
    dxold[i] = delx;
    x[i] = xtemp[i] + f * delx;
  }
}

// sln_underrelax delta-bar-delta (:3066-3112)
__global__ void dbd_kernel(int n, const int *__restrict__ ibound, double *__restrict__ x,
                           const double *__restrict__ xtemp, double *__restrict__ dxold,
                           double *__restrict__ wsave, double *__restrict__ hchold,
                           double *__restrict__ deold, int kiter, double theta, double akappa,
                           double gamma, double amomentum) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (ibound[i] < 1) continue;
    double delx = x[i] - xtemp[i];
    double ws = wsave[i], hc = hchold[i], de = deold[i];
    if (kiter == 1) {
      ws = 1.0;
      hc = 1.0e-20;
      de = 0.0;
    }
    double ww;
    if (de * delx < 0.0)
      ww = theta * ws;
    else
      ww = ws + akappa;
    if (ww > 1.0) ww = 1.0;
    wsave[i] = ww;
    if (kiter == 1)
      hc = delx;
    else
      hc = (1.0 - gamma) * delx + gamma * hc;
    hchold[i] = hc;
    deold[i] = delx;
    dxold[i] = delx;
    double amom = 0.0;
    if (kiter > 4) amom = amomentum;
    delx = delx * ww + amom * hc;
    x[i] = xtemp[i] + delx;
  }
}

// apply_backtracking (NumericalSolution.f90:2828-2842): x = xtemp + breduc (x - xtemp) on active cells
__global__ void backtrack_kernel(int n, const int *__restrict__ ibound, double *__restrict__ x,
                                 const double *__restrict__ xtemp, double breduc) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (ibound[i] < 1) continue;
    const double delx = breduc * (x[i] - xtemp[i]);
    x[i] = xtemp[i] + delx;
  }
}

// npf_nur (gwf-npf.f90:705-741)
__global__ void nur_kernel(ModelView M, const int *__restrict__ ibotnode, double *__restrict__ x,
                           const double *__restrict__ xtemp, double *__restrict__ dx,
                           OuterState *os) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M.n; r += gridDim.x * blockDim.x) {
    if (M.ibound[r] < 1) continue;
    if (M.icelltype[r] > 0) {
      const double botm = M.bot[ibotnode[r]];
      if (x[r] < botm) {
        os->nur_flag = 1.0;
        const double xx = xtemp[r] * (1.0 - 0.9) + botm * 0.9;
        x[r] = xx;
        dx[r] = 0.0;
      }
    }
  }
}

// max |a| (sln_maxval) -> os->dxold_max
__global__ void __launch_bounds__(kBlock)
absmax_kernel(int n, const double *__restrict__ a, double *__restrict__ partial,
              unsigned int *ticket, OuterState *os) {
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
    if (threadIdx.x == 0) os->dxold_max = r;
  }
}

// sln_calc_residual (:2966-2982) + gwf_ptc (gwf.f90:625-687) + l2 norm:
// r = A x - b (0 for inactive) ; max |r| / V ; sum r^2
__global__ void __launch_bounds__(kBlock)
ptc_resid_kernel(ModelView M, const double *__restrict__ val, const double *__restrict__ x,
                 const double *__restrict__ rhsv, double *__restrict__ partial,
                 unsigned int *ticket, OuterState *os) {
  __shared__ double sh[8];
  __shared__ bool last;
  double mx = 0.0, ssq = 0.0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M.n; r += gridDim.x * blockDim.x) {
    double t = sell_row_dot(r, M.slice_ptr, M.rowlen, M.col, val, x);
    t = t + (-1.0) * rhsv[r];
    if (M.ibound[r] < 1) t = 0.0;
    ssq += t * t;
    if (M.ibound[r] >= 1) {
      const double v = M.area[r] * (M.top[r] - M.bot[r]);
      mx = fmax(mx, fabs(t) / v);
    }
  }
  mx = block_max(mx, sh);
  ssq = block_sum(ssq, sh);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = mx;
    partial[2 * blockIdx.x + 1] = ssq;
  }
  if (last_block(ticket, &last)) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
      a = fmax(a, partial[2 * i]);
      b += partial[2 * i + 1];
    }
    a = block_max(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x == 0) {
      os->ptc_max = a;
      os->l2 = b;
    }
  }
}

// PTC terms (:2556-2565): diag -= ptcval ; rhs -= ptcval * x   for active rows
__global__ void ptc_apply_kernel(ModelView M, double *__restrict__ val, double *__restrict__ rhsv,
                                 const double *__restrict__ x, double ptcval) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M.n; r += gridDim.x * blockDim.x) {
    if (M.ibound[r] > 0) {
      const long long base = (long long)M.slice_ptr[r >> 5] + (r & 31);
      val[base] = val[base] + (-ptcval);
      rhsv[r] = rhsv[r] - ptcval * x[r];
    }
  }
}

// ---- flows: npf_cq/qcalc (gwf-npf.f90:745-865) + sto_cq (gwf-sto.f90:447-564) ----
__global__ void __launch_bounds__(kBlock)
flow_rows_kernel(ModelView M, const double *__restrict__ h, const double *__restrict__ hold,
                 const double *__restrict__ sat, double *__restrict__ flowja,
                 double *__restrict__ strgss, double *__restrict__ strgsy, int transient,
                 double tled) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M.n; r += gridDim.x * blockDim.x) {
    const int len = M.rowlen[r];
    const long long base = (long long)M.slice_ptr[r >> 5] + (r & 31);
    const int ibr = M.ibound[r];
    for (int k = 1; k < len; k++) {
      const long long slot = base + 32LL * k;
      const int c = M.col[slot];
      const int conn = M.slot_conn[slot];
      const int up = conn & 1, jas = conn >> 1;
      const double csat = M.slot_condsat[slot];
      double cond;
      if (M.o.all_confined)
        cond = (ibr == 0 || M.ibound[c] == 0) ? 0.0 : csat;
      else
        cond = conn_cond(M, r, c, up, jas, csat, h, sat);
      const int n = up ? r : c, m = up ? c : r;
      double hn = h[n], hm = h[m];
      if (M.o.iperched != 0 && M.ihc[jas] == 0) {
        // qcalc is called with n < m, so only its "else" branch applies (:855-858)
        if (M.icelltype[m] != 0)
          if (hm < M.top[m]) hm = M.bot[n];
      }
      const double qnm = cond * (hm - hn);  // flow into n
      flowja[slot] = up ? qnm : -qnm;
    }
    double fd = 0.0, rss = 0.0, rsy = 0.0;
    if (transient && ibr > 0) {
      const double tp = M.top[r], bt = M.bot[r];
      const int icv = M.iconvert[r];
      double snold = 1.0, snnew = 1.0;
      if (icv != 0) {
        snold = sQuadraticSaturation(tp, bt, hold[r], M.o.satomega);
        snnew = sQuadraticSaturation(tp, bt, h[r], M.o.satomega);
      }
      const double sc1 = SsCapacity(M.o.istor_coef, tp, bt, M.area[r], M.ss[r]);
      const double rho1 = sc1 * tled;
      double aterm, rhsterm, rate;
      SsTerms(icv, M.o.iorig_ss, M.o.iconf_ss, tp, bt, rho1, rho1, snnew, snold, h[r], hold[r],
              aterm, rhsterm, rate);
      rss = rate;
      fd = fd + rate;
      rate = 0.0;
      if (icv != 0) {
        const double sc2 = M.sy[r] * M.area[r];
        const double rho2 = sc2 * tled;
        SyTerms(tp, bt, rho2, rho2, snnew, snold, aterm, rhsterm, rate);
      }
      rsy = rate;
      fd = fd + rate;
    }
    flowja[base] = fd;
    if (strgss) {
      strgss[r] = rss;
      strgsy[r] = rsy;
    }
  }
}

// csr_diagsum (Sparse.f90:262-281)
__global__ void diagsum_kernel(ModelView M, double *__restrict__ flowja) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M.n; r += gridDim.x * blockDim.x) {
    const int len = M.rowlen[r];
    const long long base = (long long)M.slice_ptr[r >> 5] + (r & 31);
    double d = flowja[base];
    for (int k = 1; k < len; k++) d = d + flowja[base + 32LL * k];
    flowja[base] = d;
  }
}

// rate_accumulator (Budget.f90:631-648) over a[i0:i1)
__global__ void __launch_bounds__(kBlock)
posneg_kernel(int i0, int i1, const double *__restrict__ a, double *__restrict__ partial,
              unsigned int *ticket, OuterState *os) {
  __shared__ double sh[8];
  __shared__ bool last;
  double rin = 0.0, rout = 0.0;
  for (int i = i0 + blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += gridDim.x * blockDim.x) {
    const double f = a[i];
    if (f < 0.0)
      rout = rout - f;
    else
      rin = rin + f;
  }
  rin = block_sum(rin, sh);
  rout = block_sum(rout, sh);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = rin;
    partial[2 * blockIdx.x + 1] = rout;
  }
  if (last_block(ticket, &last)) {
    double a0 = 0.0, b0 = 0.0;
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
      a0 += partial[2 * i];
      b0 += partial[2 * i + 1];
    }
    a0 = block_sum(a0, sh);
    b0 = block_sum(b0, sh);
    if (threadIdx.x == 0) {
      os->rin = a0;
      os->rout = b0;
    }
  }
}

__global__ void copy_d_kernel(int n, const double *__restrict__ a, double *__restrict__ b) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    b[i] = a[i];
}

__global__ void i2d_kernel(int n, const int *__restrict__ a, double *__restrict__ b) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) b[i] = (double)a[i];
}

__global__ void d2i_kernel(int n, const double *__restrict__ a, int *__restrict__ b) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) b[i] = (int)a[i];
}

__global__ void copy_i_kernel(int n, const int *__restrict__ a, int *__restrict__ b) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    b[i] = a[i];
}

}  // namespace mf6

using namespace mf6;

struct PkgRange {
  int type, i0, i1;
};

struct mf6gpu_solution {
  mf6gpu_matrix *A = nullptr;
  mf6gpu_solver *S = nullptr;
  int n = 0, nja = 0, njas = 0;   // n = owned cells (matrix rows)
  int n_ext = 0;                  // owned + halo cells
  mf6::HaloPlan halo;             // split-model path only
  DevBuf<double> os_all;          // [nranks] gathered OuterState
  DevBuf<double> ibd_tmp;         // [n_ext] ibound as doubles for the halo exchange
  ModelOpts o{};
  mf6gpu_sln_settings ss{};
  int isymmetric = 0;
  cudaStream_t stream = 0;
  DevBuf<double> top, bot, area, k11, k33, ssv, syv;
  DevBuf<double> hyc;            // NPF anisotropy: see ModelView::hyc (empty = isotropic in the plane)
  DevBuf<double> strt;
  DevBuf<double> x, xold, sat, rhs, xtemp, dxold, wsave, hchold, deold, strgss, strgsy;
  DevBuf<int> icelltype, ibound, ibound0, iconvert, ibotnode;
  DevBuf<int> below;     // cell under every cell (final numbering); only when a cell can be inactive
  DevBuf<int> b_eff;     // [nb] cell every bound acts on
  DevBuf<int> wd_flag;   // [1] a constant-head cell went dry
  bool do_wd = false;    // npf wet/dry conversion applies (no NEWTON, convertible cells)
  // THICKSTRT: initial saturation per cell (empty = 1 everywhere)
  DevBuf<double> sat0;
  // HFB: barriers of the current period
  int ngnc = 0, gnc_numj = 0, gnc_nrows = 0;
  DevBuf<int> gnc_rn, gnc_rm, gnc_jas, gnc_slot_nm, gnc_slot_mn, gnc_rj, gnc_ev_row, gnc_ev_ptr, gnc_ev_gnc;
  DevBuf<double> gnc_alpha, gnc_cond;
  GncView gview() const {
    return GncView{ngnc, gnc_numj, gnc_rn.p, gnc_rm.p, gnc_jas.p, gnc_slot_nm.p, gnc_slot_mn.p, gnc_rj.p, gnc_alpha.p,
                   gnc_cond.p};
  }
  int nhfb = 0, hfb_nrows = 0;
  DevBuf<int> hfb_rn, hfb_rm, hfb_jas, hfb_slot_nm, hfb_slot_mn, hfb_ev_row, hfb_ev_ptr, hfb_ev_hfb;
  DevBuf<double> hfb_hydchr, hfb_csatsav, hfb_condsav;
  std::vector<int> h_ia, h_ja;   // host copy of the CSR pattern (barrier lookup)
  int index_base = 0;
  HfbView hview() const {
    return HfbView{nhfb, hfb_rn.p, hfb_rm.p, hfb_jas.p, hfb_slot_nm.p, hfb_slot_mn.p, hfb_hydchr.p, hfb_csatsav.p,
                   hfb_condsav.p};
  }
  // REWET (rewet_check): wetdry per cell, the REWET record, pass bookkeeping
  DevBuf<double> wetdry;
  DevBuf<int> rw_state, rw_pending;
  int irewet = 0, iwetit = 1, ihdwet = 0, kiter_cur = 1;
  double wetfct = 1.0;
  void wetdry_sweep(int kiter);  // sgwf_npf_wetdry: rewetting passes, drying, 30000 -> 1
  bool moving_rch = false;  // some RCH package without FIXED_CELL and cells that can be inactive
  DevBuf<int> slot_conn;
  DevBuf<double> slot_condsat, flowja;
  DevBuf<double> condsat, cl1, cl2, hwva;
  DevBuf<int> ihc;
  DevBuf<double> hstage;  // [max(n, nja)] staging in original order
  // packages (concatenated)
  int nb = 0, nseg = 0;
  std::vector<PkgRange> ranges;
  DevBuf<unsigned char> b_type, b_flag;
  DevBuf<int> b_node, b_pkg;
  DevBuf<double> b_b1, b_b2, b_b3, b_fred, b_hcof, b_rhs, b_sim, b_rin, b_rout;
  DevBuf<int> seg_node, seg_ptr, seg_idx;
  // scratch
  DevBuf<double> partial;
  DevBuf<MaxLoc> pm;
  DevBuf<unsigned int> tickets;
  DevBuf<OuterState> os;
  PinnedBuf<OuterState> h_os;
  // outer state
  double relaxold = 1.0, bigchold = 0.0, bigch = 0.0;
  double ptcdel = 0.0, l2norm0 = 0.0;
  double res_prev = 0.0, res_new = 0.0;  // backtracking
  int nbacktracks = 0;
  double delt = 1.0;
  int iss = 1;
  int icnvg = 0;
  long long nl = 0;  // kernel launches of the current time step
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  std::vector<int> h_conn_jas;  // csr position -> jas (original), for get_condsat ordering

  ModelView view() const {
    ModelView M;
    M.n = n;  // owned rows
    M.slice_ptr = A->slice_ptr.p;
    M.rowlen = A->rowlen.p;
    M.col = A->col.p;
    M.slot_conn = slot_conn.p;
    M.slot_condsat = slot_condsat.p;
    M.condsat = condsat.p;
    M.cl1 = cl1.p;
    M.cl2 = cl2.p;
    M.hwva = hwva.p;
    M.ihc = ihc.p;
    M.top = top.p;
    M.bot = bot.p;
    M.area = area.p;
    M.k11 = k11.p;
    M.k33 = k33.p;
    M.hyc = hyc.n ? hyc.p : nullptr;
    M.ss = ssv.p;
    M.sy = syv.p;
    M.icelltype = icelltype.p;
    M.ibound = ibound.p;
    M.iconvert = iconvert.p;
    M.o = o;
    return M;
  }
  BndView bview() const {
    BndView B;
    B.nb = nb;
    B.type = b_type.p;
    B.flag = b_flag.p;
    B.node = b_node.p;
    B.b1 = b_b1.p;
    B.b2 = b_b2.p;
    B.b3 = b_b3.p;
    B.fred = b_fred.p;
    B.hcof = b_hcof.p;
    B.rhs = b_rhs.p;
    B.simvals = b_sim.p;
    B.ratein = b_rin.p;
    B.rateout = b_rout.p;
    B.eff = b_eff.p;
    B.below = below.n ? below.p : nullptr;
    return B;
  }
  // local record, or (split-model path) the records of all ranks combined on the host in rank
  // order: identical on every rank, like the MPI_Allreduce calls of ParallelSolution.f90:54-235
  OuterState fetch_os() {
    if (!halo.active()) {
      MF6_CK(cudaMemcpyAsync(h_os.p, os.p, sizeof(OuterState), cudaMemcpyDeviceToHost, stream));
      MF6_CK(cudaStreamSynchronize(stream));
      return *h_os.p;
    }
    const int nr = halo.comm->nranks;
    const size_t cnt = sizeof(OuterState) / sizeof(double);
    comm_allgather(halo.comm, reinterpret_cast<const double *>(os.p), os_all.p, cnt, stream);
    MF6_CK(cudaMemcpyAsync(h_os.p, os_all.p, sizeof(OuterState) * (size_t)nr, cudaMemcpyDeviceToHost, stream));
    MF6_CK(cudaStreamSynchronize(stream));
    OuterState g = h_os.p[0];
    for (int r = 1; r < nr; r++) {
      const OuterState &y = h_os.p[r];
      const double ay = std::fabs(y.hncg), ag = std::fabs(g.hncg);
      if (ay > ag || (ay == ag && ay > 0.0 && y.loc_ord < g.loc_ord)) {
        g.hncg = y.hncg;
        g.loc_ord = y.loc_ord;
      }
      g.nur_flag = std::fmax(g.nur_flag, y.nur_flag);
      g.dxold_max = std::fmax(g.dxold_max, y.dxold_max);
      g.ptc_max = std::fmax(g.ptc_max, y.ptc_max);
      g.l2 += y.l2;
      g.rin += y.rin;
      g.rout += y.rout;
    }
    g.loc = g.loc_ord;  // reported as a global cell id
    return g;
  }
  void exchange_x() {
    if (halo.active()) halo.exchange(x.p, stream);
  }
  void buildsystem(int inewton);
  void calc_ptc(int &iptc, double &ptcf);
  void ls_fixups(int kiter, int kstp, int kper, int iptc, double ptcf);
  int solve_outer(int kiter, int kstp, int kper, double &hncg, int &lrch, double &tf, double &tl);
  void posneg(const double *a, int i0, int i1, double &rin, double &rout);
  double residual_l2();
  void backtracking(int kiter);
};

template <class T>
static std::vector<T> permuted(const T *src, const std::vector<int> &perm, T dflt) {
  std::vector<T> out(perm.size(), dflt);
  if (src)
    for (size_t r = 0; r < perm.size(); r++) out[r] = src[perm[r]];
  return out;
}

// sgwf_npf_wetdry (gwf-npf.f90:2061-2158): [rewetting sweep] + drying + [30000 -> 1]
void mf6gpu_solution::wetdry_sweep(int kiter) {
  const ModelView M = view();
  const int G = grid_for(n);
  const bool rewet = irewet > 0 && wetdry.n > 0 && (kiter % iwetit) == 0;
  if (rewet) {
    rewet_begin_kernel<<<G, kBlock, 0, stream>>>(n, wetdry.p, ibound.p, rw_state.p);
    nl++;
    for (int pass = 2; pass < n + 3; pass++) {
      rw_pending.zero(stream);
      rewet_pass_kernel<<<G, kBlock, 0, stream>>>(M, A->d_perm.p, wetdry.p, x.p, ibound.p, rw_state.p, pass, wetfct,
                                                  ihdwet, rw_pending.p);
      nl++;
      int pend = 0;
      MF6_CK(cudaMemcpyAsync(&pend, rw_pending.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
      MF6_CK(cudaStreamSynchronize(stream));
      if (!pend) break;
    }
  }
  npf_wd_kernel<<<G, kBlock, 0, stream>>>(M, x.p, ibound.p, ibound0.p, wd_flag.p);
  nl++;
  if (irewet > 0 && wetdry.n > 0) {
    rewet_end_kernel<<<G, kBlock, 0, stream>>>(n, ibound.p, ibound0.p);
    nl++;
  }
  MF6_CK(cudaGetLastError());
}

// sln_buildsystem (:1941-1991): sln_reset + gwf_cf + gwf_fc [+ Newton terms]
void mf6gpu_solution::buildsystem(int inewton) {
  const ModelView M = view();
  const BndView B = bview();
  const int G = grid_for(n);
  const int transient = (iss == 0 && o.insto) ? 1 : 0;
  const double tled = 1.0 / delt;
  exchange_x();  // halo heads (STG_BFR_EXG_CF synchronisation of the reference)
  nl += (o.all_confined ? 0 : 1) + (nb > 0 ? 1 : 0) + 1 + (nseg > 0 ? 1 : 0);
  if (inewton && o.inewton) nl += 1 + (nseg > 0 ? 1 : 0);
  if (!o.all_confined) {
    if (do_wd) {
      wetdry_sweep(kiter_cur);
      if (halo.active()) {
        // split-model path: the neighbours' copies of the cells that have just gone dry (ibound 0, head = HDRY)
        // are refreshed before the conductances are formed -- what the reference re-synchronises per outer
        // iteration for the interface model (GwfGwfConnection.f90:205-228, VirtualGwfModel ibound / x)
        const int nh = n_ext - n;
        i2d_kernel<<<grid_for(n), kBlock, 0, stream>>>(n, ibound.p, ibd_tmp.p);
        halo.exchange(ibd_tmp.p, stream);
        if (nh > 0) d2i_kernel<<<grid_for(nh), kBlock, 0, stream>>>(nh, ibd_tmp.p + n, ibound.p + n);
        exchange_x();
        nl += 4;
      }
    }
    ModelView Me = M;
    Me.n = n_ext;  // saturation of the halo cells too
    npf_cf_kernel<<<grid_for(n_ext), kBlock, 0, stream>>>(Me, x.p, sat.p);
  }
  if (nb > 0) bnd_cf_kernel<<<grid_for(nb), kBlock, 0, stream>>>(B, M, x.p);
  if (o.all_confined)
    assemble_rows_kernel<true><<<G, kBlock, 0, stream>>>(M, x.p, xold.p, sat.p, A->val.p, rhs.p, transient, tled);
  else
    assemble_rows_kernel<false><<<G, kBlock, 0, stream>>>(M, x.p, xold.p, sat.p, A->val.p, rhs.p, transient, tled);
  if (nhfb > 0 && o.inewton == 0 && !o.all_confined) {
    // gwf_fc: hfb_fc follows npf_fc (gwf.f90:500-506).  Here STO is already in the diagonal (fused into the row
    // kernel), so with storage the diagonal is summed in another order than the reference's: last-bit only
    hfb_fc_kernel<<<grid_for(hfb_nrows), kBlock, 0, stream>>>(hfb_nrows, hfb_ev_row.p, hfb_ev_ptr.p, hfb_ev_hfb.p,
                                                              hview(), M, x.p, A->val.p);
    nl++;
  }
  if (ngnc > 0) {  // gwf_fc: gnc_fc follows hfb_fc (gwf.f90:496)
    gnc_fc_kernel<<<grid_for(gnc_nrows), kBlock, 0, stream>>>(gnc_nrows, gnc_ev_row.p, gnc_ev_ptr.p, gnc_ev_gnc.p,
                                                              gview(), M, x.p, A->val.p, rhs.p);
    nl++;
  }
  if (nseg > 0)
    bnd_scatter_kernel<<<grid_for(nseg), kBlock, 0, stream>>>(nseg, seg_node.p, seg_ptr.p, seg_idx.p, B, M,
                                                              x.p, A->val.p, rhs.p, flowja.p, 0);
  if (moving_rch) {
    rch_moved_kernel<<<grid_for(nb), kBlock, 0, stream>>>(B, M, rhs.p, flowja.p, 0);
    nl++;
  }
  if (inewton && o.inewton) {
    newton_rows_kernel<<<G, kBlock, 0, stream>>>(M, x.p, A->val.p, rhs.p, transient, tled);
    if (ngnc > 0) {
      gnc_fn_kernel<<<grid_for(gnc_nrows), kBlock, 0, stream>>>(gnc_nrows, gnc_ev_row.p, gnc_ev_ptr.p, gnc_ev_gnc.p,
                                                                gview(), M, x.p, A->val.p, rhs.p);
      nl++;
    }
    if (nseg > 0)
      bnd_scatter_kernel<<<grid_for(nseg), kBlock, 0, stream>>>(nseg, seg_node.p, seg_ptr.p, seg_idx.p, B,
                                                                M, x.p, A->val.p, rhs.p, flowja.p, 1);
  }
  MF6_CK(cudaGetLastError());
}

// sln_calc_ptc (:2936-2962) + gwf_ptc (gwf.f90:625-687)
void mf6gpu_solution::calc_ptc(int &iptc, double &ptcf) {
  iptc = 0;
  ptcf = 0.0;
  int iptct = 0;
  if (iss > 0) iptct = o.inewton;
  if (iptct > 0) {
    nl++;
    ptc_resid_kernel<<<grid_for(n), kBlock, 0, stream>>>(view(), A->val.p, x.p, rhs.p, partial.p,
                                                         tickets.p + 1, os.p);
    MF6_CK(cudaGetLastError());
    OuterState h = fetch_os();
    ptcf = h.ptc_max;
    if (ptcf == 0.0) ptcf = 1.0 / (delt * 10.0);
    iptc = 1;
  }
}

void mf6gpu_solution::ls_fixups(int kiter, int kstp, int kper, int iptc, double ptcf) {
  const ModelView M = view();
  const int G = grid_for(n);
  nl++;
  ls_fixup_kernel<<<G, kBlock, 0, stream>>>(M, x.p, xtemp.p, A->val.p, rhs.p, isymmetric);
  MF6_CK(cudaGetLastError());
  int iallowptc;
  if (ss.iallowptc < 0)
    iallowptc = (kper > 1) ? 1 : 0;
  else
    iallowptc = ss.iallowptc;
  int iptct = iptc * iallowptc;
  double l2norm = 0.0;
  if (iptct != 0) {
    nl++;
    ptc_resid_kernel<<<G, kBlock, 0, stream>>>(M, A->val.p, x.p, rhs.p, partial.p, tickets.p + 1, os.p);
    MF6_CK(cudaGetLastError());
    l2norm = std::sqrt(fetch_os().l2);
    if (kiter == 1) {
      if (kper > 1 || kstp > 1)
        if (l2norm <= l2norm0) iptc = 0;
    } else {
      const double a = l2norm, b = l2norm0;
      bool same = (a == b) || std::fabs(a - b) <= 100.0 * 2.220446049250313e-16 * std::fmax(std::fabs(a), std::fabs(b));
      if (same) iptc = 0;
    }
  }
  iptct = iptc * iallowptc;
  if (iptct != 0) {
    if (kiter == 1) {
      ptcdel = 1.0 / ptcf;
    } else {
      if (l2norm > 0.0)
        ptcdel = ptcdel * std::pow(l2norm0 / l2norm, 1.0);
      else
        ptcdel = 0.0;
    }
    const double ptcval = (ptcdel > 0.0) ? 1.0 / ptcdel : 1.0;
    nl++;
    ptc_apply_kernel<<<G, kBlock, 0, stream>>>(M, A->val.p, rhs.p, x.p, ptcval);
    MF6_CK(cudaGetLastError());
    l2norm0 = l2norm;
  }
}

// sln_l2norm (:2855-2871): || A x - b || with inactive rows zeroed
double mf6gpu_solution::residual_l2() {
  exchange_x();
  nl++;
  ptc_resid_kernel<<<grid_for(n), kBlock, 0, stream>>>(view(), A->val.p, x.p, rhs.p, partial.p, tickets.p + 1, os.p);
  MF6_CK(cudaGetLastError());
  return std::sqrt(fetch_os().l2);
}

// sln_backtracking (:2680-2776)
void mf6gpu_solution::backtracking(int kiter) {
  const int G = grid_for(n);
  buildsystem(0);
  if (kiter == 1) {
    res_prev = residual_l2();
  } else {
    res_new = residual_l2();
  }
  if (kiter > 1) {
    if (res_new > res_prev * ss.btol) {
      for (int nbt = 1; nbt <= ss.numtrack; nbt++) {
        // get_backtracking_flag (:2790-2822)
        nl++;
        dxmax_kernel<<<G, kBlock, 0, stream>>>(n, x.p, xtemp.p, ibound.p, A->ord_ptr(), pm.p, tickets.p, os.p);
        MF6_CK(cudaGetLastError());
        const double dx_abs_max = std::fabs(fetch_os().hncg);
        if (!(ss.breduc * dx_abs_max >= ss.dvclose)) break;
        nl++;
        backtrack_kernel<<<G, kBlock, 0, stream>>>(n, ibound.p, x.p, xtemp.p, ss.breduc);
        nbacktracks++;
        buildsystem(0);
        res_new = residual_l2();
        if (nbt == ss.numtrack) break;
        if (res_new < res_prev * ss.btol) break;
        if (res_new < ss.res_lim) break;
      }
    }
    res_prev = res_new;
  }
}

// solve(kiter) (:1482-1837).  The two NVTX ranges carry the names of the reference's own profiler sections
// ("Formulate", "Linear solve": NumericalSolution.f90:1581-1601), so a timeline lines up with its listing.
int mf6gpu_solution::solve_outer(int kiter, int kstp, int kper, double &hncg, int &lrch, double &tf,
                                 double &tl) {
  const int G = grid_for(n);
  MF6_CK(cudaEventRecord(ev[0], stream));
  nvtxRangePushA("Formulate");
  kiter_cur = kiter;
  if (ss.numtrack > 0) backtracking(kiter);
  buildsystem(1);
  int iptc;
  double ptcf;
  calc_ptc(iptc, ptcf);
  nvtxRangePop();
  MF6_CK(cudaEventRecord(ev[1], stream));
  nvtxRangePushA("Linear solve");
  ls_fixups(kiter, kstp, kper, iptc, ptcf);
  int iter = 0, icnvg_lin = 0;
  try {
    S->solve_device(kiter, kstp, x.p, rhs.p, &iter, &icnvg_lin);
  } catch (...) {
    nvtxRangePop();
    throw;
  }
  nvtxRangePop();
  nl += S->launches + 1;  // + dxmax
  MF6_CK(cudaEventRecord(ev[2], stream));
  dxmax_kernel<<<G, kBlock, 0, stream>>>(n, x.p, xtemp.p, ibound.p, A->ord_ptr(), pm.p, tickets.p, os.p);
  MF6_CK(cudaGetLastError());
  OuterState h = fetch_os();
  float ms0 = 0.f, ms1 = 0.f;
  MF6_CK(cudaEventElapsedTime(&ms0, ev[0], ev[1]));
  MF6_CK(cudaEventElapsedTime(&ms1, ev[1], ev[2]));
  tf += 1e-3 * ms0;
  tl += 1e-3 * ms1;
  hncg = h.hncg;
  lrch = (int)h.loc;
  icnvg = 0;
  if (std::fabs(hncg) <= ss.dvclose) icnvg = 1;
  if (icnvg != 1) {
    nl += 1 + ((o.inewton != 0 && o.inewtonur != 0) ? 2 : 0);
    if (ss.nonmeth == 1) {
      relax_kernel<<<G, kBlock, 0, stream>>>(n, ibound.p, x.p, xtemp.p, dxold.p, ss.gamma);
    } else if (ss.nonmeth == 2) {
      double relax;
      bigch = hncg;
      if (kiter == 1) {
        relax = 1.0;
        relaxold = 1.0;
        bigchold = hncg;
      } else {
        const double es = bigch / (bigchold * relaxold);
        const double aes = std::fabs(es);
        if (es < -1.0)
          relax = 0.5 / aes;
        else
          relax = (3.0 + es) / (3.0 + aes);
      }
      relaxold = relax;
      bigchold = (1.0 - ss.gamma) * bigch + ss.gamma * bigchold;
      if (relax < 1.0) relax_kernel<<<G, kBlock, 0, stream>>>(n, ibound.p, x.p, xtemp.p, dxold.p, relax);
    } else if (ss.nonmeth == 3) {
      dbd_kernel<<<G, kBlock, 0, stream>>>(n, ibound.p, x.p, xtemp.p, dxold.p, wsave.p, hchold.p, deold.p,
                                           kiter, ss.theta, ss.akappa, ss.gamma, ss.amomentum);
    } else {
      calcdx_kernel<<<G, kBlock, 0, stream>>>(n, ibound.p, x.p, xtemp.p, dxold.p);
    }
    if (o.inewton != 0 && o.inewtonur != 0) {
      MF6_CK(cudaMemsetAsync(&os.p->nur_flag, 0, sizeof(double), stream));
      nur_kernel<<<G, kBlock, 0, stream>>>(view(), ibotnode.p, x.p, xtemp.p, dxold.p, os.p);
      absmax_kernel<<<G, kBlock, 0, stream>>>(n, dxold.p, partial.p, tickets.p + 2, os.p);
      MF6_CK(cudaGetLastError());
      OuterState h2 = fetch_os();
      if (h2.nur_flag != 0.0) {
        if (std::fabs(h2.dxold_max) <= ss.dvclose && std::fabs(hncg) <= ss.dvclose) {
          icnvg = 1;
          dxmax_kernel<<<G, kBlock, 0, stream>>>(n, x.p, xtemp.p, ibound.p, A->ord_ptr(), pm.p, tickets.p, os.p);
          OuterState h3 = fetch_os();
          hncg = h3.hncg;
          lrch = (int)h3.loc;
        }
      }
    }
    MF6_CK(cudaGetLastError());
  }
  return iter;
}

void mf6gpu_solution::posneg(const double *a, int i0, int i1, double &rin, double &rout) {
  rin = rout = 0.0;
  if (i1 <= i0 && !halo.active()) return;
  posneg_kernel<<<grid_for(std::max(i1 - i0, 1)), kBlock, 0, stream>>>(i0, i1, a, partial.p, tickets.p + 3, os.p);
  MF6_CK(cudaGetLastError());
  OuterState h = fetch_os();
  rin = h.rin;
  rout = h.rout;
}

extern "C" {

// split-model description of one rank (all arrays in the model's LOCAL numbering / index base)
struct DistArgs {
  mf6gpu_comm *comm;
  int n_own;
  int nnbr;
  const int32_t *nbr_rank;   // [nnbr]
  const int32_t *send_ptr;   // [nnbr+1]
  const int32_t *send_idx;   // [send_ptr[nnbr]] owned cells whose values neighbour k needs
  const int32_t *recv_ptr;   // [nnbr+1] ranges inside the halo region (cells n_own + ...)
  const int32_t *global_id;  // [nodes] global cell id of every local cell (owned + halo)
};

// hyeff (src/Utilities/HGeoUtil.f90:29-108, iavgmeth = 0 -- the only value the reference sets) and hy_eff
// (gwf-npf.f90:2280-2355), evaluated once on the host per connection side: the effective conductivity of a cell
// along a connection depends only on static data (K11, K22, K33, the rotation angles, the connection normal)
static double hyeff_host(double k11, double k22, double k33, double ang1, double ang2, double ang3, double vg1,
                         double vg2, double vg3) {
  const double s1 = std::sin(ang1), c1 = std::cos(ang1), s2 = std::sin(ang2), c2 = std::cos(ang2);
  const double s3 = std::sin(ang3), c3 = std::cos(ang3);
  const double r11 = c1 * c2, r12 = c1 * s2 * s3 - s1 * c3, r13 = -c1 * s2 * c3 - s1 * s3;
  const double r21 = s1 * c2, r22 = s1 * s2 * s3 + c1 * c3, r23 = -s1 * s2 * c3 + c1 * s3;
  const double r31 = s2, r32 = -c2 * s3, r33 = c2 * c3;
  const double ve1 = r11 * vg1 + r21 * vg2 + r31 * vg3;
  const double ve2 = r12 * vg1 + r22 * vg2 + r32 * vg3;
  const double ve3 = r13 * vg1 + r23 * vg2 + r33 * vg3;
  double dnum = 1.0, d1 = ve1 * ve1, d2 = ve2 * ve2, d3 = ve3 * ve3;
  if (ve1 != 0.0) {
    dnum = dnum * k11;
    d2 = d2 * k11;
    d3 = d3 * k11;
  }
  if (ve2 != 0.0) {
    dnum = dnum * k22;
    d1 = d1 * k22;
    d3 = d3 * k22;
  }
  if (ve3 != 0.0) {
    dnum = dnum * k33;
    d1 = d1 * k33;
    d2 = d2 * k33;
  }
  const double denom = d1 + d2 + d3;
  return denom > 0.0 ? dnum / denom : 0.0;
}

static double hy_eff_host(const mf6gpu_gwf_model *md, int n, int ihc, double vg1, double vg2, double vg3) {
  const double hy11 = md->k11[n], hy22 = md->k22 ? md->k22[n] : md->k11[n];
  const double hy33 = md->k33 ? md->k33[n] : md->k11[n];
  if (ihc == 0) {
    if (!md->angle2) return hy33;
    return hyeff_host(hy11, hy22, hy33, md->angle1 ? md->angle1[n] : 0.0, md->angle2[n],
                      md->angle3 ? md->angle3[n] : 0.0, vg1, vg2, vg3);
  }
  if (!md->k22) return hy11;
  double a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (md->angle1) {
    a1 = md->angle1[n];
    if (md->angle2) {
      a2 = md->angle2[n];
      if (md->angle3) a3 = md->angle3[n];
    }
  }
  return hyeff_host(hy11, hy22, hy33, a1, a2, a3, vg1, vg2, vg3);
}

// blocks of the BLOCK_MULTICOLOR ordering: the vertical cell columns (chains of ihc == 0 connections)
static std::vector<int32_t> model_column_blocks(const mf6gpu_gwf_model *m, int n_own) {
  const int base0 = m->index_base;
  std::vector<int32_t> block((size_t)n_own);
  for (int v = 0; v < n_own; v++) {
    block[v] = v;
    for (int p = m->ia[v] - base0 + 1; p < m->ia[v + 1] - base0; p++) {
      const int u = m->ja[p] - base0;
      if (u < v && m->ihc[m->jas[p] - base0] == 0) {
        block[v] = block[u];
        break;
      }
    }
  }
  return block;
}

static void create_solution(const mf6gpu_gwf_model *m, const mf6gpu_sln_settings *sln,
                            const mf6gpu_ims_settings *ims, const DistArgs *da, mf6gpu_solution **out) {
    MF6_REQUIRE(m && sln && ims && out, "solution_create: null argument");
    if (!da) MF6_REQUIRE(m->njas * 2 == m->nja - m->nodes, "solution_create: nja/njas/nodes are inconsistent");
    MF6_REQUIRE((long long)m->njas < (1LL << 30), "solution_create: too many connections for one GPU");
    MF6_REQUIRE(!(m->inewton != 0 && ims->ilinmeth == 1),
                "solution_create: NEWTON needs an asymmetric accelerator (BICGSTAB), cf. NumericalSolution.f90:942-958");
    const int n_own = da ? da->n_own : m->nodes;
    MF6_REQUIRE(n_own > 0 && n_own <= m->nodes, "solution_create: bad number of owned cells");
    const int nja_own = m->ia[n_own] - m->index_base;  // halo rows (diagonal only) are not matrix rows
    auto *s = new mf6gpu_solution();
    s->n = n_own;
    s->n_ext = m->nodes;
    s->nja = nja_own;
    s->njas = m->njas;
    s->ss = *sln;
    s->isymmetric = (ims->ilinmeth == 1) ? 1 : 0;  // NumericalSolution.f90:914-916
    std::vector<int32_t> block;
    if (ims->gpu_ordering == MF6GPU_ORDER_BLOCK_MULTICOLOR) block = model_column_blocks(m, n_own);
    if (mf6gpu_matrix_create_blocked(n_own, m->nodes, nja_own, m->ia, m->ja, m->index_base, ims->gpu_ordering,
                                     da ? da->global_id : nullptr, block.empty() ? nullptr : block.data(),
                                     &s->A) != 0) {
      const std::string keep = last_error();
      delete s;
      throw Error(keep);
    }
    try {
      const int base = m->index_base;
      const int n = m->nodes, nja = nja_own, njas = m->njas;  // n = extended cell count here
      mf6gpu_matrix *A = s->A;
      s->stream = A->stream;
      if (mf6gpu_solver_create(A, ims, 0, &s->S) != 0) throw Error(last_error());
      if (da) {
        mf6::HaloPlan &H = s->halo;
        H.comm = da->comm;
        H.n_own = n_own;
        H.n_halo = n - n_own;
        H.nbr_rank.assign(da->nbr_rank, da->nbr_rank + da->nnbr);
        H.send_ptr.assign(da->send_ptr, da->send_ptr + da->nnbr + 1);
        H.recv_ptr.assign(da->recv_ptr, da->recv_ptr + da->nnbr + 1);
        MF6_REQUIRE(H.recv_ptr.back() == H.n_halo, "solution_create: recv ranges must cover the halo region");
        std::vector<int> sidx((size_t)H.send_ptr.back());
        for (size_t i = 0; i < sidx.size(); i++) {
          const int c = da->send_idx[i] - base;
          MF6_REQUIRE(c >= 0 && c < n_own, "solution_create: send_idx must name owned cells");
          sidx[i] = A->iperm[c];
        }
        if (sidx.empty()) sidx.push_back(0);
        H.send_idx.upload(sidx);
        H.sendbuf.alloc_zero(sidx.size());
        {
          std::vector<int> nr = H.nbr_rank;
          if (nr.empty()) nr.push_back(0);
          H.d_nbr_rank.upload(nr);
          H.d_send_ptr.upload(H.send_ptr);
          H.d_recv_ptr.upload(H.recv_ptr);
          H.ticket.alloc_zero(1);
          if (da->comm->p2p) {
            for (int k = 0; k < da->nnbr; k++)
              MF6_REQUIRE((size_t)(H.send_ptr[k + 1] - H.send_ptr[k]) <= da->comm->lay.halo_doubles &&
                              (size_t)(H.recv_ptr[k + 1] - H.recv_ptr[k]) <= da->comm->lay.halo_doubles,
                          "solution_create_dist: halo message larger than the peer mailbox");
          }
        }
        {
          // slices with a halo column: their rows wait for the neighbours' messages inside the fused SpMV
          std::vector<unsigned char> sh((size_t)A->nslices, 0);
          for (int v = 0; v < n_own; v++)
            for (int p = m->ia[v] - base + 1; p < m->ia[v + 1] - base; p++)
              if (m->ja[p] - base >= n_own) {
                sh[(size_t)(A->iperm[v] >> 5)] = 1;
                break;
              }
          H.slice_halo.upload(sh);
          // send entries grouped by the CTA that computes their row in the vector-update kernels
          const int G = grid_for(n_own);
          std::vector<int> ent(sidx.size()), cp((size_t)G + 1, 0);
          const int nsend = H.send_ptr.back();
          for (int i = 0; i < nsend; i++) cp[(size_t)((sidx[i] / kBlock) % G) + 1]++;
          for (int b = 0; b < G; b++) cp[b + 1] += cp[b];
          {
            std::vector<int> cur(cp.begin(), cp.end() - 1);
            for (int i = 0; i < nsend; i++) ent[(size_t)cur[(sidx[i] / kBlock) % G]++] = i;
          }
          H.cta_ptr.upload(cp);
          H.cta_ent.upload(ent);
          H.own_grid = G;
          // chunks (kBlock rows) with halo-touching slices, grouped the same way, for the deferred pass of the
          // fused SpMV: entry [chunk][t] = the row thread t handles (the same thread as in the one-pass mode, so
          // that both modes accumulate identical per-thread sums) or -1
          std::vector<int> dp((size_t)G + 1, 0), dr;
          const int nchunk = (n_own + kBlock - 1) / kBlock;
          std::vector<char> cflag((size_t)nchunk, 0);
          for (int r = 0; r < n_own; r++)
            if (sh[(size_t)(r >> 5)]) cflag[(size_t)(r / kBlock)] = 1;
          for (int c = 0; c < nchunk; c++)
            if (cflag[c]) dp[(size_t)(c % G) + 1] += kBlock;
          for (int b = 0; b < G; b++) dp[b + 1] += dp[b];
          dr.assign((size_t)std::max(dp[G], 1), -1);
          {
            std::vector<int> cur(dp.begin(), dp.end() - 1);
            for (int c = 0; c < nchunk; c++) {
              if (!cflag[c]) continue;
              const int b = c % G;
              for (int t = 0; t < kBlock; t++) {
                const int r = c * kBlock + t;
                if (r < n_own && sh[(size_t)(r >> 5)]) dr[(size_t)cur[b] + t] = r;
              }
              cur[b] += kBlock;
            }
          }
          H.def_ptr.upload(dp);
          H.def_row.upload(dr);
        }
        s->S->halo = &s->halo;
        s->os_all.alloc_zero(sizeof(OuterState) / sizeof(double) * (size_t)da->comm->nranks);
        s->h_os.alloc((size_t)da->comm->nranks);
        s->ibd_tmp.alloc_zero((size_t)n);
      }
      const std::vector<int> &perm = A->perm;
      // options
      ModelOpts &o = s->o;
      o.icellavg = m->icellavg;
      o.inewton = m->inewton;
      o.inewtonur = m->inewtonur;
      o.iperched = m->iperched;
      o.ivarcv = m->ivarcv;
      o.idewatcv = m->idewatcv;
      o.insto = m->insto;
      o.istor_coef = m->istor_coef;
      o.iconf_ss = m->iconf_ss;
      o.iorig_ss = m->iorig_ss;
      o.satomega = (m->inewton > 0) ? 1.0e-6 : 0.0;
      // prepcheck (gwf-npf.f90:1838-1882): a negative ICELLTYPE means "convertible" without THICKSTRT; with THICKSTRT
      // the cell is confined with the saturated thickness of its STARTING head (calc_initial_sat :2046-2057)
      std::vector<int> ict(m->icelltype, m->icelltype + n);
      std::vector<double> sat0;
      for (int i = 0; i < n; i++) {
        if (ict[i] >= 0) continue;
        if (m->ithickstrt != 0) {
          if (sat0.empty()) sat0.assign((size_t)n, 1.0);
          if (!m->ibound || m->ibound[i] != 0) {
            const double tp = m->top[i], bt = m->bot[i], hn = m->strt[i];
            double sn = (hn >= tp) ? 1.0 : (hn - bt) / (tp - bt);
            if (m->inewton != 0) {  // sQuadraticSaturation with satomega = 1e-6 (thksat, :775-794)
              const double eps = 1.0e-6, b = tp - bt;
              if (b > 0.0) {
                const double br = (hn < bt) ? 0.0 : (hn > tp ? 1.0 : (hn - bt) / b);
                const double av = 1.0 / (1.0 - eps), bri = 1.0 - br;
                if (br < eps)
                  sn = av * 0.5 * (br * br) / eps;
                else if (br < (1.0 - eps))
                  sn = av * br + 0.5 * (1.0 - av);
                else if (br < 1.0)
                  sn = 1.0 - ((av * 0.5 * (bri * bri)) / eps);
                else
                  sn = 1.0;
              } else {
                sn = (hn < bt) ? 0.0 : 1.0;
              }
            }
            sat0[i] = sn;
          }
          ict[i] = 0;
        } else {
          ict[i] = 1;
        }
      }
      if (!sat0.empty()) s->sat0.upload(permuted(sat0.data(), perm, 1.0));
      bool anyconv = false;
      for (int i = 0; i < n; i++)
        if (ict[i] != 0) anyconv = true;
      o.all_confined = (!anyconv && m->iperched == 0 && m->inewton == 0) ? 1 : 0;
      // cells can become / be inactive: wet-dry conversion (no NEWTON, convertible cells) or IDOMAIN holes.  Recharge then needs
      // the cell under every cell (highest_active walks the m > n vertical connections)
      s->do_wd = (m->inewton == 0) && anyconv;
      s->wd_flag.alloc_zero(1);
      if (m->wetdry && m->irewet) {
        MF6_REQUIRE(!da, "solution_create: REWET is not available on the split-model path");
        s->wetdry.upload(permuted(m->wetdry, perm, 0.0));
        s->rw_state.alloc_zero((size_t)n_own);
        s->rw_pending.alloc_zero(1);
        s->irewet = 1;
        s->iwetit = m->iwetit > 0 ? m->iwetit : 1;
        s->ihdwet = m->ihdwet;
        s->wetfct = m->wetfct;
      }
      bool anyinactive = false;
      for (int i = 0; i < n_own; i++)
        if (m->ibound && m->ibound[i] == 0) anyinactive = true;
      if (s->do_wd || anyinactive) {  // (split-model path: a cell column never straddles two ranks)
        std::vector<int> bel((size_t)n_own, -1);
        for (int v = 0; v < n_own; v++)
          for (int p = m->ia[v] - base + 1; p < m->ia[v + 1] - base; p++) {
            const int u = m->ja[p] - base;
            if (u > v && u < n_own && m->ihc[m->jas[p] - base] == 0) {
              bel[(size_t)A->iperm[v]] = A->iperm[u];
              break;
            }
          }
        s->below.upload(bel);
      }
      // per-cell arrays in final numbering
      s->top.upload(permuted(m->top, perm, 0.0));
      s->bot.upload(permuted(m->bot, perm, 0.0));
      s->area.upload(permuted(m->area, perm, 0.0));
      s->k11.upload(permuted(m->k11, perm, 0.0));
      s->k33.upload(permuted(m->k33 ? m->k33 : m->k11, perm, 0.0));
      s->ssv.upload(permuted(m->ss, perm, 0.0));
      s->syv.upload(permuted(m->sy, perm, 0.0));
      s->icelltype.upload(permuted(ict.data(), perm, 0));
      s->iconvert.upload(permuted(m->iconvert, perm, 0));
      s->ibound0.upload(permuted(m->ibound, perm, 1));
      s->ibound.upload(permuted(m->ibound, perm, 1));
      {
        std::vector<int> ib(n);
        for (int r = 0; r < n; r++) {
          int o_ = perm[r];
          int b = m->ibotnode ? m->ibotnode[o_] - base : o_;
          ib[r] = A->iperm[b];
        }
        s->ibotnode.upload(ib);
      }
      s->strt.upload(permuted(m->strt, perm, 0.0));
      s->x.upload(permuted(m->strt, perm, 0.0));
      s->xold.upload(permuted(m->strt, perm, 0.0));
      if (!sat0.empty()) {
        s->sat.upload(permuted(sat0.data(), perm, 1.0));  // a THICKSTRT cell keeps its initial saturation (prepcheck)
      } else {
        std::vector<double> ones((size_t)n, 1.0);
        s->sat.upload(ones);
      }
      s->rhs.alloc_zero(n);
      s->xtemp.alloc_zero(n);
      s->dxold.alloc_zero(n);
      s->wsave.alloc_zero(n);
      s->hchold.alloc_zero(n);
      s->deold.alloc_zero(n);
      s->strgss.alloc_zero(n);
      s->strgsy.alloc_zero(n);
      s->hstage.alloc((size_t)std::max(nja, n));
      // per-connection arrays (original jas numbering)
      {
        std::vector<int> ihc(m->ihc, m->ihc + njas);
        s->ihc.upload(ihc);
        s->cl1.upload(std::vector<double>(m->cl1, m->cl1 + njas));
        s->cl2.upload(std::vector<double>(m->cl2, m->cl2 + njas));
        s->hwva.upload(std::vector<double>(m->hwva, m->hwva + njas));
      }
      // connection endpoints + slot -> connection map
      std::vector<int> conn_n((size_t)njas, 0), conn_m((size_t)njas, 0);
      std::vector<int> slot_conn((size_t)A->nslots, -1);
      std::vector<int> csr2sell((size_t)nja);
      A->csr2sell.download(csr2sell.data(), (size_t)nja);
      s->h_conn_jas.assign((size_t)nja, -1);
      s->index_base = base;
      s->h_ia.assign(m->ia, m->ia + n_own + 1);
      s->h_ja.assign(m->ja, m->ja + nja);
      for (int v = 0; v < n_own; v++) {
        const int i0 = m->ia[v] - base, i1 = m->ia[v + 1] - base;
        for (int p = i0 + 1; p < i1; p++) {
          const int u = m->ja[p] - base;
          const int jj = m->jas[p] - base;
          MF6_REQUIRE(jj >= 0 && jj < njas, "solution_create: jas out of range");
          s->h_conn_jas[p] = jj;
          // argument order of the reference: n = the cell with the lower (GLOBAL) number
          const int up = da ? ((da->global_id[u] > da->global_id[v]) ? 1 : 0) : ((u > v) ? 1 : 0);
          if (up) {
            conn_n[jj] = A->iperm[v];
            conn_m[jj] = A->iperm[u];
          } else if (u >= n_own) {  // exchange connection seen only from this side
            conn_n[jj] = A->iperm[u];
            conn_m[jj] = A->iperm[v];
          }
          slot_conn[csr2sell[p]] = (jj << 1) | up;
        }
      }
      s->slot_conn.upload(slot_conn);
      if (m->k22 || m->angle1 || m->angle2) {
        MF6_REQUIRE(!da, "solution_create: K22 / rotation angles are not available on the split-model path");
        MF6_REQUIRE(!(m->k22 || m->angle1) || (m->conn_nx && m->conn_ny),
                    "solution_create: K22 / ANGLE1 need the connection normals conn_nx / conn_ny");
        std::vector<double> hy(2 * (size_t)std::max(njas, 1), 0.0);
        for (int v = 0; v < n_own; v++)
          for (int p = m->ia[v] - base + 1; p < m->ia[v + 1] - base; p++) {
            const int u = m->ja[p] - base;
            if (u < v) continue;
            const int jj = m->jas[p] - base, hc = m->ihc[jj];
            const double nx = (hc != 0 && m->conn_nx) ? m->conn_nx[jj] : 0.0;
            const double ny = (hc != 0 && m->conn_ny) ? m->conn_ny[jj] : 0.0;
            // connection_normal: seen from the cell itself; u lies below v for a vertical connection
            hy[2 * (size_t)jj] = hy_eff_host(m, v, hc, nx, ny, hc == 0 ? -1.0 : 0.0);
            hy[2 * (size_t)jj + 1] = hy_eff_host(m, u, hc, -nx, -ny, hc == 0 ? 1.0 : 0.0);
          }
        s->hyc.upload(hy);
      }
      s->condsat.alloc_zero((size_t)std::max(njas, 1));
      s->slot_condsat.alloc_zero((size_t)A->nslots);
      s->flowja.alloc_zero((size_t)A->nslots);
      s->partial.alloc_zero(4 * (size_t)kMaxBlocks);
      s->pm.alloc_zero((size_t)kMaxBlocks);
      s->tickets.alloc_zero(8);
      s->os.alloc_zero(1);
      if (!da) s->h_os.alloc(1);
      for (auto &e : s->ev) MF6_CK(cudaEventCreate(&e));
      if (njas > 0) {
        DevBuf<int> dn, dm;
        dn.upload(conn_n);
        dm.upload(conn_m);
        condsat_kernel<<<grid_for(njas), kBlock, 0, s->stream>>>(njas, dn.p, dm.p, s->view(),
                                                                 s->sat0.n ? s->sat0.p : nullptr, s->condsat.p);
        slot_condsat_kernel<<<grid_for(A->nslots), kBlock, 0, s->stream>>>(A->nslots, s->slot_conn.p,
                                                                           s->condsat.p, s->slot_condsat.p);
        MF6_CK(cudaGetLastError());
        MF6_CK(cudaStreamSynchronize(s->stream));
      }
      // prepcheck (gwf-npf.f90:1817-1821): without NEWTON the wet/dry routine runs once on the initial heads
      // (kiter = 0); the split-model path does it in its first formulate, after the first halo exchange
      if (s->do_wd && !da) {
        s->wetdry_sweep(0);
        MF6_CK(cudaStreamSynchronize(s->stream));
      }
    } catch (const std::exception &e) {
      const std::string keep = e.what();
      mf6gpu_solution_destroy(s);
      throw Error(keep);
    }
    *out = s;
}

int mf6gpu_solution_create(const mf6gpu_gwf_model *m, const mf6gpu_sln_settings *sln,
                           const mf6gpu_ims_settings *ims, mf6gpu_solution **out) {
  return guard([&] { create_solution(m, sln, ims, nullptr, out); });
}

int mf6gpu_solution_create_dist(const mf6gpu_gwf_model *m, const mf6gpu_sln_settings *sln,
                                const mf6gpu_ims_settings *ims, mf6gpu_comm *comm, int32_t n_own,
                                int32_t nnbr, const int32_t *nbr_rank, const int32_t *send_ptr,
                                const int32_t *send_idx, const int32_t *recv_ptr,
                                const int32_t *global_id, mf6gpu_solution **out) {
  return guard([&] {
    MF6_REQUIRE(comm && global_id && (nnbr == 0 || (nbr_rank && send_ptr && send_idx && recv_ptr)),
                "solution_create_dist: null argument");
    DistArgs da{comm, n_own, nnbr, nbr_rank, send_ptr, send_idx, recv_ptr, global_id};
    create_solution(m, sln, ims, &da, out);
  });
}

int mf6gpu_solution_destroy(mf6gpu_solution *s) {
  return guard([&] {
    if (!s) return;
    if (s->S) mf6gpu_solver_destroy(s->S);
    if (s->A) mf6gpu_matrix_destroy(s->A);
    for (auto &e : s->ev)
      if (e) cudaEventDestroy(e);
    delete s;
  });
}

// hfb_rp (gwf-hfb.f90:149-201): condsat_reset, the new barrier list, condsat_modify
int mf6gpu_solution_set_hfb(mf6gpu_solution *s, int32_t nhfb, const int32_t *noden, const int32_t *nodem,
                            const double *hydchr, int32_t index_base) {
  return guard([&] {
    MF6_REQUIRE(s && nhfb >= 0 && (nhfb == 0 || (noden && nodem && hydchr)), "solution_set_hfb: bad argument");
    MF6_REQUIRE(!s->halo.active(), "solution_set_hfb: HFB is not available on the split-model path");
    const ModelView M = s->view();
    if (s->nhfb > 0) {
      hfb_condsat_kernel<<<grid_for(s->nhfb), kBlock, 0, s->stream>>>(s->hview(), M, s->condsat.p,
                                                                      s->slot_condsat.p, 1);
      MF6_CK(cudaGetLastError());
      MF6_CK(cudaStreamSynchronize(s->stream));
    }
    s->nhfb = nhfb;
    if (nhfb == 0) return;
    std::vector<int> csr2sell((size_t)s->nja);
    s->A->csr2sell.download(csr2sell.data(), (size_t)s->nja);
    const int b0 = s->index_base;
    auto find = [&](int v, int u) {  // CSR position of (v, u)
      for (int p = s->h_ia[v] - b0 + 1; p < s->h_ia[v + 1] - b0; p++)
        if (s->h_ja[p] - b0 == u) return p;
      return -1;
    };
    std::vector<int> rn(nhfb), rm(nhfb), jas(nhfb), snm(nhfb), smn(nhfb);
    std::vector<double> hc(hydchr, hydchr + nhfb);
    std::vector<std::pair<int, int>> ev;  // (row, barrier)
    for (int i = 0; i < nhfb; i++) {
      const int v = noden[i] - index_base, u = nodem[i] - index_base;
      MF6_REQUIRE(v >= 0 && v < s->n && u >= 0 && u < s->n, "solution_set_hfb: cell out of range");
      const int pnm = find(v, u), pmn = find(u, v);
      MF6_REQUIRE(pnm >= 0 && pmn >= 0, "solution_set_hfb: the two cells of a barrier are not connected");
      rn[i] = s->A->iperm[v];
      rm[i] = s->A->iperm[u];
      jas[i] = s->h_conn_jas[pnm];
      snm[i] = csr2sell[pnm];
      smn[i] = csr2sell[pmn];
      ev.emplace_back(rn[i], i);
      ev.emplace_back(rm[i], i);
    }
    std::sort(ev.begin(), ev.end());
    std::vector<int> ev_row, ev_ptr, ev_hfb;
    for (size_t e = 0; e < ev.size(); e++) {
      if (e == 0 || ev[e].first != ev[e - 1].first) {
        ev_row.push_back(ev[e].first);
        ev_ptr.push_back((int)e);
      }
      ev_hfb.push_back(ev[e].second);
    }
    ev_ptr.push_back((int)ev.size());
    s->hfb_nrows = (int)ev_row.size();
    s->hfb_rn.upload(rn);
    s->hfb_rm.upload(rm);
    s->hfb_jas.upload(jas);
    s->hfb_slot_nm.upload(snm);
    s->hfb_slot_mn.upload(smn);
    s->hfb_hydchr.upload(hc);
    s->hfb_csatsav.alloc_zero((size_t)nhfb);
    s->hfb_condsav.alloc_zero((size_t)nhfb);
    s->hfb_ev_row.upload(ev_row);
    s->hfb_ev_ptr.upload(ev_ptr);
    s->hfb_ev_hfb.upload(ev_hfb);
    hfb_condsat_kernel<<<grid_for(nhfb), kBlock, 0, s->stream>>>(s->hview(), M, s->condsat.p, s->slot_condsat.p, 0);
    MF6_CK(cudaGetLastError());
    MF6_CK(cudaStreamSynchronize(s->stream));
  });
}

// GNC6 package data (GhostNode.f90 read_data :739-864), EXPLICIT correction only
int mf6gpu_solution_set_gnc(mf6gpu_solution *s, int32_t ngnc, int32_t numj, const int32_t *noden,
                            const int32_t *nodem, const int32_t *nodesj, const double *alphasj, int32_t index_base) {
  return guard([&] {
    MF6_REQUIRE(s && ngnc >= 0 && numj >= 0 && (ngnc == 0 || (noden && nodem)), "solution_set_gnc: bad argument");
    MF6_REQUIRE(ngnc == 0 || numj == 0 || (nodesj && alphasj), "solution_set_gnc: bad argument");
    MF6_REQUIRE(!s->halo.active(), "solution_set_gnc: GNC is not available on the split-model path");
    s->ngnc = ngnc;
    s->gnc_numj = numj;
    if (ngnc == 0) return;
    std::vector<int> csr2sell((size_t)s->nja);
    s->A->csr2sell.download(csr2sell.data(), (size_t)s->nja);
    const int b0 = s->index_base;
    auto find = [&](int v, int u) {
      for (int p = s->h_ia[v] - b0 + 1; p < s->h_ia[v + 1] - b0; p++)
        if (s->h_ja[p] - b0 == u) return p;
      return -1;
    };
    std::vector<int> rn(ngnc), rm(ngnc), jas(ngnc), snm(ngnc), smn(ngnc), rj((size_t)ngnc * std::max(numj, 1), -1);
    std::vector<double> al((size_t)ngnc * std::max(numj, 1), 0.0);
    std::vector<std::pair<int, int>> ev;
    for (int i = 0; i < ngnc; i++) {
      const int v = noden[i] - index_base, u = nodem[i] - index_base;
      MF6_REQUIRE(v >= 0 && v < s->n && u >= 0 && u < s->n, "solution_set_gnc: cell out of range");
      const int pnm = find(v, u), pmn = find(u, v);
      MF6_REQUIRE(pnm >= 0 && pmn >= 0, "solution_set_gnc: GHOST NODE ERROR, the two cells are not connected");
      rn[i] = s->A->iperm[v];
      rm[i] = s->A->iperm[u];
      jas[i] = s->h_conn_jas[pnm];
      snm[i] = csr2sell[pnm];
      smn[i] = csr2sell[pmn];
      for (int k = 0; k < numj; k++) {
        const int j = nodesj[(size_t)i * numj + k] - index_base;
        MF6_REQUIRE(j < s->n, "solution_set_gnc: contributing cell out of range");
        rj[(size_t)i * numj + k] = (j >= 0) ? s->A->iperm[j] : -1;
        al[(size_t)i * numj + k] = alphasj[(size_t)i * numj + k];
      }
      ev.emplace_back(rn[i], i);
      ev.emplace_back(rm[i], i);
    }
    std::sort(ev.begin(), ev.end());
    std::vector<int> ev_row, ev_ptr, ev_gnc;
    for (size_t e = 0; e < ev.size(); e++) {
      if (e == 0 || ev[e].first != ev[e - 1].first) {
        ev_row.push_back(ev[e].first);
        ev_ptr.push_back((int)e);
      }
      ev_gnc.push_back(ev[e].second);
    }
    ev_ptr.push_back((int)ev.size());
    s->gnc_nrows = (int)ev_row.size();
    s->gnc_rn.upload(rn);
    s->gnc_rm.upload(rm);
    s->gnc_jas.upload(jas);
    s->gnc_slot_nm.upload(snm);
    s->gnc_slot_mn.upload(smn);
    s->gnc_rj.upload(rj);
    s->gnc_alpha.upload(al);
    s->gnc_cond.alloc_zero((size_t)ngnc);
    s->gnc_ev_row.upload(ev_row);
    s->gnc_ev_ptr.upload(ev_ptr);
    s->gnc_ev_gnc.upload(ev_gnc);
  });
}

int mf6gpu_solution_set_packages(mf6gpu_solution *s, int32_t npkg, const mf6gpu_bnd_package *pk) {
  return guard([&] {
    MF6_REQUIRE(s && (npkg == 0 || pk), "solution_set_packages: null argument");
    MF6_REQUIRE(npkg <= MF6GPU_MAX_BUDGET_TERMS - 2, "solution_set_packages: too many packages");
    // chd_rp (gwf-chd.f90:134-138): the cells of the PREVIOUS constant-head list become ordinary active cells
    if (s->nb > 0) {
      chd_release_kernel<<<grid_for(s->nb), kBlock, 0, s->stream>>>(s->bview(), s->ibound0.p);
      MF6_CK(cudaGetLastError());
      MF6_CK(cudaStreamSynchronize(s->stream));
    }
    int nb = 0;
    for (int k = 0; k < npkg; k++) nb += pk[k].nbound;
    std::vector<unsigned char> type(nb), flag(nb);
    std::vector<int> node(nb), pkg(nb);
    std::vector<double> b1(nb, 0.0), b2(nb, 0.0), b3(nb, 0.0), fred(nb, 0.0);
    s->ranges.clear();
    int g = 0;
    for (int k = 0; k < npkg; k++) {
      const mf6gpu_bnd_package &p = pk[k];
      MF6_REQUIRE(p.type >= MF6GPU_PKG_CHD && p.type <= MF6GPU_PKG_DRN, "solution_set_packages: unknown package type");
      s->ranges.push_back(PkgRange{p.type, g, g + p.nbound});
      for (int i = 0; i < p.nbound; i++, g++) {
        const int nd = p.nodelist[i] - p.index_base;
        MF6_REQUIRE(nd >= 0 && nd < s->n, "solution_set_packages: node out of range");
        type[g] = (unsigned char)p.type;
        flag[g] = (unsigned char)(p.iflowred != 0);
        fred[g] = p.flowred;
        node[g] = s->A->iperm[nd];
        pkg[g] = k;
        b1[g] = p.b1 ? p.b1[i] : 0.0;
        b2[g] = p.b2 ? p.b2[i] : 0.0;
        b3[g] = p.b3 ? p.b3[i] : 0.0;
      }
    }
    s->nb = nb;
    // segments of equal node, bounds in (package, bound) order inside a segment
    std::vector<int> order(nb);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return node[a] < node[b]; });
    std::vector<int> seg_node, seg_ptr;
    for (int e = 0; e < nb; e++) {
      if (e == 0 || node[order[e]] != node[order[e - 1]]) {
        seg_node.push_back(node[order[e]]);
        seg_ptr.push_back(e);
      }
    }
    seg_ptr.push_back(nb);
    s->nseg = (int)seg_node.size();
    const size_t c = (size_t)std::max(nb, 1);
    auto up = [&](auto &buf, auto &vec) {
      vec.resize(c);
      buf.upload(vec);
    };
    up(s->b_type, type);
    up(s->b_flag, flag);
    up(s->b_node, node);
    up(s->b_pkg, pkg);
    up(s->b_b1, b1);
    up(s->b_b2, b2);
    up(s->b_b3, b3);
    up(s->b_fred, fred);
    s->b_eff.upload(node);   // until the first bnd_cf: every bound acts on its listed cell
    s->moving_rch = false;
    for (int k = 0; k < npkg; k++)
      if (pk[k].type == MF6GPU_PKG_RCH && pk[k].iflowred == 0 && pk[k].nbound > 0 && s->below.n > 0)
        s->moving_rch = true;
    s->b_hcof.alloc_zero(c);
    s->b_rhs.alloc_zero(c);
    s->b_sim.alloc_zero(c);
    s->b_rin.alloc_zero(c);
    s->b_rout.alloc_zero(c);
    if (seg_node.empty()) seg_node.push_back(0);
    order.resize(c);
    s->seg_node.upload(seg_node);
    s->seg_ptr.upload(seg_ptr);
    s->seg_idx.upload(order);
    // ibound: reset to the initial state, then mark constant heads (chd_rp)
    copy_i_kernel<<<grid_for(s->n), kBlock, 0, s->stream>>>(s->n, s->ibound0.p, s->ibound.p);
    if (nb > 0)
      chd_ibound_kernel<<<grid_for(nb), kBlock, 0, s->stream>>>(s->bview(), s->b_pkg.p, s->ibound.p);
    if (s->halo.active()) {
      // ibound of the halo cells (constant heads of the neighbour), cf. VirtualGwfModel.f90:117-122
      const int nh = s->n_ext - s->n;
      i2d_kernel<<<grid_for(s->n), kBlock, 0, s->stream>>>(s->n, s->ibound.p, s->ibd_tmp.p);
      s->halo.exchange(s->ibd_tmp.p, s->stream);
      if (nh > 0) d2i_kernel<<<grid_for(nh), kBlock, 0, s->stream>>>(nh, s->ibd_tmp.p + s->n, s->ibound.p + s->n);
    }
    MF6_CK(cudaGetLastError());
    MF6_CK(cudaStreamSynchronize(s->stream));
  });
}

int mf6gpu_solution_formulate(mf6gpu_solution *s, int32_t kiter, double delt, int32_t iss) {
  return guard([&] {
    MF6_REQUIRE(s, "solution_formulate: null argument");
    s->delt = delt;
    s->iss = iss;
    s->kiter_cur = kiter;
    if (s->nb > 0) chd_ad_kernel<<<grid_for(s->nb), kBlock, 0, s->stream>>>(s->bview(), s->x.p, s->xold.p);
    s->buildsystem(1);
    int iptc;
    double ptcf;
    s->calc_ptc(iptc, ptcf);
    s->ls_fixups(kiter, 1, 1, iptc, ptcf);
    MF6_CK(cudaStreamSynchronize(s->stream));
  });
}

int mf6gpu_solution_timestep(mf6gpu_solution *s, int32_t kper, int32_t kstp, double delt,
                             int32_t iss, mf6gpu_step_report *rep) {
  return guard([&] {
    MF6_REQUIRE(s, "solution_timestep: null argument");
    MF6_REQUIRE(iss != 0 || delt > 0.0, "solution_timestep: DELT must be > 0 for a transient step (gwf-sto.f90:263-268)");
    s->delt = delt;
    s->iss = iss;
    const int n = s->n;
    const int G = grid_for(n);
    cudaStream_t st = s->stream;
    s->nl = 2 + 4 + (s->nb > 0 ? 3 : 0);  // prepareSolve + finalizeSolve kernels (budget reductions not counted)
    // prepareSolve: gwf_ad (xold = x) ; chd_ad
    copy_d_kernel<<<G, kBlock, 0, st>>>(n, s->x.p, s->xold.p);
    if (s->irewet > 0)
      npf_ad_rewet_kernel<<<G, kBlock, 0, st>>>(n, s->wetdry.p, s->ibound.p, s->bot.p, s->x.p, s->xold.p);
    if (s->nb > 0) chd_ad_kernel<<<grid_for(s->nb), kBlock, 0, st>>>(s->bview(), s->x.p, s->xold.p);
    MF6_CK(cudaGetLastError());
    int kiter, inner_total = 0, lrch = -1;
    double hncg = 0.0, tf = 0.0, tl = 0.0;
    s->icnvg = 0;
    s->nbacktracks = 0;
    for (kiter = 1; kiter <= s->ss.mxiter; kiter++) {
      inner_total += s->solve_outer(kiter, kstp, kper, hncg, lrch, tf, tl);
      if (s->icnvg == 1) break;
    }
    if (kiter > s->ss.mxiter) kiter = s->ss.mxiter;
    // finalizeSolve: gwf_cq
    const ModelView M = s->view();
    const BndView B = s->bview();
    const int transient = (iss == 0 && s->o.insto) ? 1 : 0;
    s->exchange_x();
    // npf_cq uses the saturation of the LAST formulate (this%sat is only updated in npf_cf, gwf-npf.f90:444-470)
    flow_rows_kernel<<<G, kBlock, 0, st>>>(M, s->x.p, s->xold.p, s->sat.p, s->flowja.p, s->strgss.p,
                                           s->strgsy.p, transient, 1.0 / delt);
    if (s->nhfb > 0 && s->o.inewton == 0 && !s->o.all_confined)
      hfb_cq_kernel<<<grid_for(s->nhfb), kBlock, 0, st>>>(s->hview(), M, s->x.p, s->flowja.p);
    if (s->ngnc > 0) gnc_cq_kernel<<<grid_for(s->ngnc), kBlock, 0, st>>>(s->gview(), M, s->x.p, s->flowja.p);
    if (s->nb > 0) {
      bnd_cf_kernel<<<grid_for(s->nb), kBlock, 0, st>>>(B, M, s->x.p);
      bnd_scatter_kernel<<<grid_for(s->nseg), kBlock, 0, st>>>(s->nseg, s->seg_node.p, s->seg_ptr.p,
                                                               s->seg_idx.p, B, M, s->x.p, s->A->val.p,
                                                               s->rhs.p, s->flowja.p, 2);
      if (s->moving_rch) rch_moved_kernel<<<grid_for(s->nb), kBlock, 0, st>>>(B, M, s->rhs.p, s->flowja.p, 2);
    }
    // gwf_bd: csr_diagsum, then the budget entries (chd_bd computes the CHD rates)
    diagsum_kernel<<<G, kBlock, 0, st>>>(M, s->flowja.p);
    if (s->nb > 0) chd_rate_kernel<<<grid_for(s->nb), kBlock, 0, st>>>(B, M, s->flowja.p);
    MF6_CK(cudaGetLastError());
    if (rep) {
      std::memset(rep, 0, sizeof(*rep));
      int nt = 0;
      double totrin = 0.0, totrot = 0.0, rin, rout, dum;
      if (s->o.insto) {
        s->posneg(s->strgss.p, 0, n, rin, rout);
        rep->term_id[nt] = 100; rep->term_in[nt] = rin; rep->term_out[nt] = rout; nt++;
        totrin += rin; totrot += rout;
        s->posneg(s->strgsy.p, 0, n, rin, rout);
        rep->term_id[nt] = 101; rep->term_in[nt] = rin; rep->term_out[nt] = rout; nt++;
        totrin += rin; totrot += rout;
      }
      for (const PkgRange &r : s->ranges) {
        if (nt >= MF6GPU_MAX_BUDGET_TERMS) break;
        if (r.type == MF6GPU_PKG_CHD) {
          s->posneg(s->b_rin.p, r.i0, r.i1, rin, dum);
          s->posneg(s->b_rout.p, r.i0, r.i1, rout, dum);
        } else {
          s->posneg(s->b_sim.p, r.i0, r.i1, rin, rout);
        }
        rep->term_id[nt] = r.type; rep->term_in[nt] = rin; rep->term_out[nt] = rout; nt++;
        totrin += rin; totrot += rout;
      }
      rep->nterms = nt;
      rep->totrin = totrin;
      rep->totrot = totrot;
      const double avgrat = (totrin + totrot) / 2.0;
      rep->pdiffr = (avgrat != 0.0) ? 100.0 * (totrin - totrot) / avgrat : 0.0;
      rep->converged = s->icnvg;
      rep->outer_iterations = kiter;
      rep->inner_iterations = inner_total;
      rep->max_dv = hncg;
      rep->max_dv_loc = lrch >= 0 ? (s->halo.active() ? lrch : s->A->perm[lrch]) + 1 : 0;
      rep->npivot_fixes = s->S->npivfix;
      rep->nbacktracks = s->nbacktracks;
      rep->t_formulate = tf;
      rep->t_linsolve = tl;
    }
    MF6_CK(cudaStreamSynchronize(st));
    if (s->do_wd) {
      int dry = 0;
      MF6_CK(cudaMemcpy(&dry, s->wd_flag.p, sizeof(int), cudaMemcpyDeviceToHost));
      MF6_REQUIRE(dry == 0, "CONSTANT-HEAD CELL WENT DRY -- SIMULATION ABORTED (gwf-npf.f90:2137-2146)");
    }
  });
}

static void get_cell_vector(mf6gpu_solution *s, const double *dev, double *host) {
  launch_scatter(s->n, s->A->d_perm.p, dev, s->hstage.p, s->stream);
  MF6_CK(cudaGetLastError());
  s->hstage.download(host, (size_t)s->n, s->stream);
}

int mf6gpu_solution_get_x(mf6gpu_solution *s, double *x) {
  return guard([&] {
    MF6_REQUIRE(s && x, "solution_get_x: null argument");
    get_cell_vector(s, s->x.p, x);
  });
}

int mf6gpu_solution_set_x(mf6gpu_solution *s, const double *x) {
  return guard([&] {
    MF6_REQUIRE(s && x, "solution_set_x: null argument");
    MF6_CK(cudaMemcpyAsync(s->hstage.p, x, sizeof(double) * (size_t)s->n, cudaMemcpyHostToDevice, s->stream));
    launch_gather(s->n, s->A->d_perm.p, s->hstage.p, s->x.p, s->stream);
    MF6_CK(cudaGetLastError());
    MF6_CK(cudaStreamSynchronize(s->stream));
  });
}

// heads back to the initial condition (IC strt) without touching the host
int mf6gpu_solution_reset_x(mf6gpu_solution *s) {
  return guard([&] {
    MF6_REQUIRE(s, "solution_reset_x: null argument");
    copy_d_kernel<<<grid_for(s->n), kBlock, 0, s->stream>>>(s->n, s->strt.p, s->x.p);
    MF6_CK(cudaGetLastError());
    MF6_CK(cudaStreamSynchronize(s->stream));
  });
}

int mf6gpu_solution_get_rhs(mf6gpu_solution *s, double *rhs) {
  return guard([&] {
    MF6_REQUIRE(s && rhs, "solution_get_rhs: null argument");
    get_cell_vector(s, s->rhs.p, rhs);
  });
}

int mf6gpu_solution_get_amat(mf6gpu_solution *s, double *amat) {
  return guard([&] {
    MF6_REQUIRE(s && amat, "solution_get_amat: null argument");
    if (mf6gpu_matrix_get_values(s->A, amat) != 0) throw Error(last_error());
  });
}

namespace mf6 {
__global__ void sell_gather_kernel(int nja, const int *__restrict__ map,
                                   const double *__restrict__ sell, double *__restrict__ csr) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nja; p += gridDim.x * blockDim.x)
    csr[p] = sell[map[p]];
}
}  // namespace mf6

int mf6gpu_solution_get_flowja(mf6gpu_solution *s, double *flowja) {
  return guard([&] {
    MF6_REQUIRE(s && flowja, "solution_get_flowja: null argument");
    sell_gather_kernel<<<grid_for(s->nja), kBlock, 0, s->stream>>>(s->nja, s->A->csr2sell.p, s->flowja.p,
                                                                   s->hstage.p);
    MF6_CK(cudaGetLastError());
    s->hstage.download(flowja, (size_t)s->nja, s->stream);
  });
}

int mf6gpu_solution_get_simvals(mf6gpu_solution *s, int32_t cap, double *simvals, int32_t *count) {
  return guard([&] {
    MF6_REQUIRE(s && count, "solution_get_simvals: null argument");
    *count = s->nb;
    if (simvals && s->nb > 0) {
      MF6_REQUIRE(cap >= s->nb, "solution_get_simvals: buffer too small");
      MF6_CK(cudaMemcpyAsync(simvals, s->b_sim.p, sizeof(double) * (size_t)s->nb, cudaMemcpyDeviceToHost, s->stream));
      MF6_CK(cudaStreamSynchronize(s->stream));
    }
  });
}

int mf6gpu_solution_get_nodes(mf6gpu_solution *s, int32_t cap, int32_t *nodes, int32_t *count) {
  return guard([&] {
    MF6_REQUIRE(s && count, "solution_get_nodes: null argument");
    *count = s->nb;
    if (nodes && s->nb > 0) {
      MF6_REQUIRE(cap >= s->nb, "solution_get_nodes: buffer too small");
      std::vector<int> eff((size_t)s->nb);
      MF6_CK(cudaMemcpyAsync(eff.data(), s->b_eff.p, sizeof(int) * (size_t)s->nb, cudaMemcpyDeviceToHost, s->stream));
      MF6_CK(cudaStreamSynchronize(s->stream));
      for (int i = 0; i < s->nb; i++) nodes[i] = s->A->perm[eff[(size_t)i]];
    }
  });
}

int mf6gpu_solution_get_storage(mf6gpu_solution *s, double *strgss, double *strgsy) {
  return guard([&] {
    MF6_REQUIRE(s && strgss && strgsy, "solution_get_storage: null argument");
    MF6_REQUIRE(s->strgss.n > 0, "solution_get_storage: the model has no STO package");
    get_cell_vector(s, s->strgss.p, strgss);
    get_cell_vector(s, s->strgsy.p, strgsy);
  });
}

int mf6gpu_solution_get_condsat(mf6gpu_solution *s, double *condsat) {
  return guard([&] {
    MF6_REQUIRE(s && condsat, "solution_get_condsat: null argument");
    s->condsat.download(condsat, (size_t)s->njas, s->stream);
  });
}

int mf6gpu_solution_get_permutation(mf6gpu_solution *s, int32_t *perm) {
  return guard([&] {
    MF6_REQUIRE(s && perm, "solution_get_permutation: null argument");
    std::memcpy(perm, s->A->elim.data(), sizeof(int) * (size_t)s->n);
  });
}

// Host-only: the elimination order mf6gpu_solution_create would use for this model (perm[k] = cell
// eliminated k-th); needs no device, so a CPU checker can be run on the same permuted system anywhere.
int mf6gpu_model_elimination_order(const mf6gpu_gwf_model *m, int32_t gpu_ordering, int32_t *perm) {
  return guard([&] {
    MF6_REQUIRE(m && perm, "model_elimination_order: null argument");
    std::vector<int32_t> block;
    if (gpu_ordering == MF6GPU_ORDER_BLOCK_MULTICOLOR) block = model_column_blocks(m, m->nodes);
    if (mf6gpu_ordering_compute(m->nodes, m->nodes, m->nja, m->ia, m->ja, m->index_base, gpu_ordering,
                                block.empty() ? nullptr : block.data(), perm) != 0)
      throw Error(last_error());
  });
}

mf6gpu_solver *mf6gpu_solution_solver(mf6gpu_solution *s) { return s ? s->S : nullptr; }

double mf6gpu_solution_stat(const mf6gpu_solution *s, int what) {
  if (!s) return -1.0;
  switch (what) {
    case 0: return (double)s->nl;
    case 1: return (double)s->A->nlevels;
    case 2: return (double)s->A->nslots;
    case 3: return s->A->slot_off_hit;
    case 4: return (double)s->A->uniform_w;
    case 5: {
      int na = 0;
      for (char a : s->A->blk_affine) na += a ? 1 : 0;
      return s->A->blk_chain_ok ? (double)na : -1.0;
    }
  }
  return -1.0;
}

}  // extern "C"
