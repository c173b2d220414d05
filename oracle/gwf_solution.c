/*
 * gwf_solution.c -- oracle restatement of the GWF formulate path and of the
 * NumericalSolution outer (Picard/Newton) iteration for ONE GWF model.
 * TEST INFRASTRUCTURE ONLY (see mf6_oracle.h).
 *
 * Follows (reference file:line):
 *   src/Solution/NumericalSolution.f90  solve :1482-1837, sln_buildsystem :1941-1991,
 *       sln_reset :2389-2396, sln_ls :2404-2615, sln_calc_ptc :2936-2962,
 *       sln_calc_residual :2966-2982, sln_l2norm :2855-2871, sln_calcdx :2912-2932,
 *       sln_underrelax :2989-3114, sln_get_dxmax :3122-3153
 *   src/Model/GroundWaterFlow/gwf.f90   gwf_ad :396-442, gwf_cf :446-462, gwf_fc :466-555,
 *       gwf_ptc :625-687, gwf_nur :696-734, gwf_cq :741-778, gwf_bd :785-824
 *   src/Model/GroundWaterFlow/gwf-npf.f90  npf_cf :444-470, npf_fc :474-574, npf_fn :578-698,
 *       npf_nur :705-741, npf_cq :745-771, thksat :775-794, qcalc :798-865,
 *       calc_condsat :1950-2037
 *   src/Model/ModelUtilities/GwfConductanceUtils.f90 hcond :43-86, vcond :149-222, condmean :226-284
 *   src/Utilities/SmoothingFunctions.f90 :275-324, :364-406, :412-516
 *   src/Model/GroundWaterFlow/gwf-sto.f90 sto_fc :226-345, sto_fn :353-439, sto_cq :447-564
 *   src/Model/ModelUtilities/GwfStorageUtils.f90 :32-177
 *   src/Model/ModelUtilities/BoundaryPackage.f90 bnd_fc :453-472, bnd_cq_simrate :583-619
 *   src/Model/GroundWaterFlow/gwf-{wel,riv,rch,ghb,drn,chd}.f90 (*_cf, wel_fn, chd_ad, calc_chd_rate)
 *   src/Utilities/Budget.f90 :259-267, :631-648 ; src/Utilities/Sparse.f90 :262-281
 */
#include "mf6_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define DEM15 1.0e-15
#define DEM20 1.0e-20
#define DEM6 1.0e-6
#define DP9 0.9

typedef struct {
  int type, nbound, iflowred;
  double flowred;
  int *nodelist;  /* the cell each bound acts on (RCH: reset to the highest active cell by rch_cf) */
  int *nodetop;   /* RCH without FIXED_CELL: the cell of the input list (nodesontop, gwf-rch.f90:283-297) */
  double *b1, *b2, *b3;
  double *hcof, *rhs, *simvals, *ratein, *rateout;
} pkg_t;

struct orc_solution {
  int nodes, nja, njas;
  int *ia, *ja, *jas, *isym, *ihc;
  double *cl1, *cl2, *hwva, *top, *bot, *area, *k11, *k33, *ss, *sy;
  double *hyc; /* [2*njas] effective K of the lower / higher cell along every connection (hy_eff), NULL = k11 / k33 */
  int *icelltype, *ibound0, *ibound, *ibotnode, *iconvert;
  int icellavg, inewton, inewtonur, iperched, ivarcv, idewatcv, insto;
  int istor_coef, iconf_ss, iorig_ss;
  double satomega;
  double *x, *xold, *sat, *condsat, *amat, *rhs, *xtemp, *dxold, *wsave, *hchold,
      *deold, *flowja, *strgss, *strgsy, *resid;
  int npkg;
  pkg_t *pkg;
  mf6gpu_sln_settings ss_;
  orc_imslinear *ims;
  int isymmetric;
  int dry_chd; /* a constant-head cell went dry (fatal in the reference) */
  /* THICKSTRT (gwf-npf.f90:1838-1882): initial saturation of every cell (1 unless flagged) */
  double *sat0;
  /* HFB (gwf-hfb.f90): barriers between cells noden / nodem, hydraulic characteristic */
  /* GNC, ghost node correction (src/Exchange/GhostNode.f90), EXPLICIT variant */
  int ngnc, gnc_numj, *gnc_n, *gnc_m, *gnc_j, *gnc_pos;
  double *gnc_alpha, *gnc_cond;
  int nhfb, *hfb_n, *hfb_m, *hfb_pos;
  double *hfb_hydchr, *hfb_condsav, *hfb_csatsav;
  /* REWET (gwf-npf.f90:2061-2223) */
  double *wetdry, wetfct;
  int irewet, iwetit, ihdwet, kiter_cur;
  /* cooley */
  double relaxold, bigchold, bigch;
  /* ptc */
  double ptcdel, l2norm0;
  /* backtracking */
  double res_prev, res_new;
  int nbacktracks;
  int icnvg;
  double t_form, t_ls;
  double delt;
  int iss;
};

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---------------- SmoothingFunctions.f90 -------------------------------- */
static double sQuadraticSaturation(double top, double bot, double x, double eps) {
  double y;
  double b = top - bot;
  if (b > 0.0) {
    double br;
    if (x < bot)
      br = 0.0;
    else if (x > top)
      br = 1.0;
    else
      br = (x - bot) / b;
    double av = 1.0 / (1.0 - eps);
    double bri = 1.0 - br;
    if (br < eps)
      y = av * 0.5 * (br * br) / eps;
    else if (br < (1.0 - eps))
      y = av * br + 0.5 * (1.0 - av);
    else if (br < 1.0)
      y = 1.0 - ((av * 0.5 * (bri * bri)) / eps);
    else
      y = 1.0;
  } else {
    y = (x < bot) ? 0.0 : 1.0;
  }
  return y;
}

static double sQuadraticSaturationDerivative(double top, double bot, double x,
                                             double eps) {
  double b = top - bot, br, y;
  if (x < bot)
    br = 0.0;
  else if (x > top)
    br = 1.0;
  else
    br = (x - bot) / b;
  double av = 1.0 / (1.0 - eps);
  double bri = 1.0 - br;
  if (br < eps)
    y = av * br / eps;
  else if (br < (1.0 - eps))
    y = av;
  else if (br < 1.0)
    y = av * bri / eps;
  else
    y = 0.0;
  return y / b;
}

static double sQSaturation(double top, double bot, double x) {
  double w = x - bot, b = top - bot, s = w / b;
  double cof1 = -2.0 / (b * b * b), cof2 = 3.0 / (b * b);
  if (s < 0.0) return 0.0;
  if (s < 1.0) return cof1 * (w * w * w) + cof2 * (w * w);
  return 1.0;
}

static double sQSaturationDerivative(double top, double bot, double x) {
  double w = x - bot, b = top - bot, s = w / b;
  double cof1 = -2.0 * 3.0 / (b * b * b), cof2 = 3.0 * 2.0 / (b * b);
  if (s < 0.0) return 0.0;
  if (s < 1.0) return cof1 * (w * w) + cof2 * w;
  return 0.0;
}

/* the same two functions with explicit c1 / c2 (SmoothingFunctions.f90:412-516; DRN passes -1, 2) */
static double sQSaturationC(double top, double bot, double x, double c1, double c2) {
  double w = x - bot, b = top - bot, s = w / b;
  double cof1 = c1 / (b * b * b), cof2 = c2 / (b * b);
  if (s < 0.0) return 0.0;
  if (s < 1.0) return cof1 * (w * w * w) + cof2 * (w * w);
  return 1.0;
}

static double sQSaturationDerivativeC(double top, double bot, double x, double c1, double c2) {
  double w = x - bot, b = top - bot, s = w / b;
  double cof1 = c1 * 3.0 / (b * b * b), cof2 = c2 * 2.0 / (b * b);
  if (s < 0.0) return 0.0;
  if (s < 1.0) return cof1 * (w * w) + cof2 * w;
  return 0.0;
}

/* get_drain_elevations, gwf-drn.f90:501-530 (b1 = elev, b3 = the DDRN auxiliary value, 0 = none) */
static void drain_elevations(double drnelev, double drndepth, double *drntop, double *drnbot) {
  if (drndepth != 0.0) {
    double elev = drnelev + drndepth;
    *drntop = elev > drnelev ? elev : drnelev;
    *drnbot = elev < drnelev ? elev : drnelev;
  } else {
    *drntop = drnelev;
    *drnbot = drnelev;
  }
}

/* ---------------- GwfConductanceUtils.f90 -------------------------------- */
static double logmean(double d1, double d2) {
  double drat = d2 / d1;
  if (drat <= 0.995 || drat >= 1.005) return (d2 - d1) / log(drat);
  return 0.5 * (d1 + d2);
}

static double condmean(double k1, double k2, double thick1, double thick2,
                       double cl1, double cl2, double width, int iavgmeth) {
  double t1 = k1 * thick1, t2 = k2 * thick2, tmean, kmean, denom;
  switch (iavgmeth) {
  case 0:
    if (t1 * t2 > 0.0) return width * t1 * t2 / (t1 * cl2 + t2 * cl1);
    return 0.0;
  case 1:
    tmean = (t1 * t2 > 0.0) ? logmean(t1, t2) : 0.0;
    return tmean * width / (cl1 + cl2);
  case 2:
    kmean = (k1 * k2 > 0.0) ? logmean(k1, k2) : 0.0;
    return kmean * 0.5 * (thick1 + thick2) * width / (cl1 + cl2);
  case 3:
    denom = (k1 * cl2 + k2 * cl1);
    kmean = (denom > 0.0) ? k1 * k2 / denom : 0.0;
    return kmean * 0.5 * (thick1 + thick2) * width;
  }
  return 0.0;
}

static double staggered_thkfrac(double top, double bot, double sat, double topc,
                                double botc) {
  /* GwfConductanceUtils.f90:staggered_thkfrac */
  double sill_top = fmin(top, topc), sill_bot = fmax(bot, botc);
  double tp = bot + sat * (top - bot);
  tp = fmin(tp, sill_top);
  double res = tp - sill_bot;
  return res > 0.0 ? res : 0.0;
}

static double hcond(int ibdn, int ibdm, int ictn, int ictm, int iupstream,
                    int ihc, int icellavg, double condsat, double hn, double hm,
                    double satn, double satm, double hkn, double hkm, double topn,
                    double topm, double botn, double botm, double cln, double clm,
                    double fawidth) {
  if (ibdn == 0 || ibdm == 0) return 0.0;
  if (ictn == 0 && ictm == 0) return condsat;
  if (iupstream == 1) {
    double sat_up = (hn > hm) ? satn : satm;
    return sat_up * condsat;
  }
  double thksatn, thksatm;
  if (ihc == 2) {
    thksatn = staggered_thkfrac(topn, botn, satn, topm, botm);
    thksatm = staggered_thkfrac(topm, botm, satm, topn, botn);
  } else {
    thksatn = satn * (topn - botn);
    thksatm = satm * (topm - botm);
  }
  return condmean(hkn, hkm, thksatn, thksatm, cln, clm, fawidth, icellavg);
}

static double vcond(int ibdn, int ibdm, int ictn, int ictm, int inewton,
                    int ivarcv, int idewatcv, double condsat, double hn,
                    double hm, double vkn, double vkm, double satn, double satm,
                    double topn, double topm, double botn, double botm,
                    double flowarea) {
  (void)inewton;
  if (ibdn == 0 || ibdm == 0) return 0.0;
  if (ivarcv == 0) return condsat;
  if (ictn == 0 && ictm == 0) return condsat;
  if (hn >= topn && hm >= topm) return condsat;
  double satntmp = satn, satmtmp = satm;
  if (idewatcv == 0) {
    if (botn > botm)
      satmtmp = 1.0;
    else
      satntmp = 1.0;
  }
  double bovk1 = satntmp * (topn - botn) * 0.5 / vkn;
  double bovk2 = satmtmp * (topm - botm) * 0.5 / vkm;
  double denom = bovk1 + bovk2;
  return (denom != 0.0) ? flowarea / denom : 0.0;
}

/* ---------------- NPF ------------------------------------------------------ */
static double thksat(const orc_solution *S, int n, double hn) {
  double t;
  if (hn >= S->top[n])
    t = 1.0;
  else
    t = (hn - S->bot[n]) / (S->top[n] - S->bot[n]);
  if (S->inewton != 0) t = sQuadraticSaturation(S->top[n], S->bot[n], hn, S->satomega);
  return t;
}

/* hyeff, src/Utilities/HGeoUtil.f90:29-108 (iavgmeth = 0, the only value the reference ever sets) */
static double hyeff(double k11, double k22, double k33, double ang1, double ang2, double ang3, double vg1,
                    double vg2, double vg3) {
  double s1 = sin(ang1), c1 = cos(ang1), s2 = sin(ang2), c2 = cos(ang2), s3 = sin(ang3), c3 = cos(ang3);
  double r11 = c1 * c2, r12 = c1 * s2 * s3 - s1 * c3, r13 = -c1 * s2 * c3 - s1 * s3;
  double r21 = s1 * c2, r22 = s1 * s2 * s3 + c1 * c3, r23 = -s1 * s2 * c3 + c1 * s3;
  double r31 = s2, r32 = -c2 * s3, r33 = c2 * c3;
  double ve1 = r11 * vg1 + r21 * vg2 + r31 * vg3;
  double ve2 = r12 * vg1 + r22 * vg2 + r32 * vg3;
  double ve3 = r13 * vg1 + r23 * vg2 + r33 * vg3;
  double K = 0.0, dnum = 1.0, d1 = ve1 * ve1, d2 = ve2 * ve2, d3 = ve3 * ve3;
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
  double denom = d1 + d2 + d3;
  if (denom > 0.0) K = dnum / denom;
  return K;
}

/* hy_eff, gwf-npf.f90:2280-2355: effective K of cell n along its connection to m (vg = normal seen from n) */
static double hy_eff(const mf6gpu_gwf_model *md, const double *k33, int n, int ihc, double vg1, double vg2,
                     double vg3) {
  double hy11 = md->k11[n], hy22 = md->k22 ? md->k22[n] : md->k11[n], hy33 = k33[n];
  if (ihc == 0) {
    if (!md->angle2) return hy33;
    return hyeff(hy11, hy22, hy33, md->angle1 ? md->angle1[n] : 0.0, md->angle2[n], md->angle3 ? md->angle3[n] : 0.0,
                 vg1, vg2, vg3);
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
  return hyeff(hy11, hy22, hy33, a1, a2, a3, vg1, vg2, vg3);
}

/* K of the two cells of every connection along it (only with K22 / rotation angles) */
static void calc_hyc(orc_solution *S, const mf6gpu_gwf_model *md) {
  S->hyc = NULL;
  if (!md->k22 && !md->angle1 && !md->angle2) return;
  S->hyc = (double *)calloc(2 * (size_t)(S->njas ? S->njas : 1), sizeof(double));
  for (int n = 0; n < S->nodes; n++)
    for (int ii = S->ia[n] + 1; ii < S->ia[n + 1]; ii++) {
      int m = S->ja[ii];
      if (m < n) continue;
      int jj = S->jas[ii], ihc = S->ihc[jj];
      double nx = (ihc != 0 && md->conn_nx) ? md->conn_nx[jj] : 0.0, ny = (ihc != 0 && md->conn_ny) ? md->conn_ny[jj] : 0.0;
      /* connection_normal: from n towards m; vertical: m below n => -1 seen from n, +1 seen from m */
      S->hyc[2 * jj] = hy_eff(md, S->k33, n, ihc, nx, ny, ihc == 0 ? -1.0 : 0.0);
      S->hyc[2 * jj + 1] = hy_eff(md, S->k33, m, ihc, -nx, -ny, ihc == 0 ? 1.0 : 0.0);
    }
}

/* gwf-npf.f90:1950-2037 with upperOnly = .true., no THICKSTRT (sat = 1) */
static void calc_condsat(orc_solution *S) {
  for (int n = 0; n < S->nodes; n++) {
    for (int ii = S->ia[n] + 1; ii < S->ia[n + 1]; ii++) {
      int m = S->ja[ii];
      if (m < n) continue;
      int jj = S->jas[ii];
      int ihc = S->ihc[jj];
      double topn = S->top[n], botn = S->bot[n], topm = S->top[m], botm = S->bot[m];
      double csat;
      if (ihc == 0) {
        csat = vcond(1, 1, 1, 1, 0, 1, 1, 1.0, botn, botm, S->hyc ? S->hyc[2 * jj] : S->k33[n],
                     S->hyc ? S->hyc[2 * jj + 1] : S->k33[m], S->sat0[n], S->sat0[m], topn, topm, botn, botm,
                     S->hwva[jj]);
      } else {
        csat = hcond(1, 1, 1, 1, 0, ihc, S->icellavg, 1.0, topn, topm, S->sat0[n], S->sat0[m],
                     S->hyc ? S->hyc[2 * jj] : S->k11[n], S->hyc ? S->hyc[2 * jj + 1] : S->k11[m], topn, topm,
                     botn, botm, S->cl1[jj], S->cl2[jj], S->hwva[jj]);
      }
      S->condsat[jj] = csat;
    }
  }
}

/* sgwf_npf_wetdry (gwf-npf.f90:2061-2158) with rewet_check (:2167-2223): first the sequential rewetting sweep (a
 * dry wettable cell turns wet when a neighbour's head has reached its wetting elevation -- a cell wetted earlier in
 * the same sweep, ibound 30000, already counts as a wet neighbour), then the drying of cells whose saturated
 * thickness is gone (a constant head going dry is fatal there), then 30000 -> 1 */
#define DHDRY (-1.0e30)
static void npf_wd(orc_solution *S, int kiter) {
  if (S->irewet > 0 && S->wetdry && (kiter % S->iwetit) == 0) {
    for (int n = 0; n < S->nodes; n++) {
      for (int ii = S->ia[n] + 1; ii < S->ia[n + 1]; ii++) {
        if (S->ibound[n] != 0 || S->wetdry[n] == 0.0) break; /* rewet_check returns at once for such a cell */
        const int m = S->ja[ii];
        const int ihc = S->ihc[S->jas[ii]];
        const double hm = S->x[m];
        const int ibdm = S->ibound[m];
        const double bbot = S->bot[n], wd = S->wetdry[n];
        double awd = wd;
        if (wd < 0) awd = -wd;
        const double turnon = bbot + awd;
        int irewet = 0;
        if (ihc == 0) {
          if (ibdm > 0 && hm >= turnon) irewet = 1;
        } else if (wd > 0.0) {
          if (ibdm > 0 && hm >= turnon) irewet = 1;
        }
        if (irewet == 1) {
          if (S->ihdwet == 0)
            S->x[n] = bbot + S->wetfct * (hm - bbot);
          else
            S->x[n] = bbot + S->wetfct * awd;
          S->ibound[n] = 30000;
        }
      }
    }
  }
  for (int n = 0; n < S->nodes; n++) {
    if (S->ibound[n] == 0 || S->icelltype[n] == 0) continue;
    double ttop = S->top[n];
    if (S->x[n] < ttop) ttop = S->x[n];
    if (ttop - S->bot[n] <= 0.0) {
      if (S->ibound[n] < 0) {
        S->dry_chd = 1;
        continue;
      }
      S->x[n] = DHDRY;
      S->ibound[n] = 0;
    }
  }
  for (int n = 0; n < S->nodes; n++) {
    if (S->ibound[n] == 30000) S->ibound[n] = 1;
    /* what the next chd_rp restores: the cell's own state without the constant-head marks */
    if (S->ibound[n] >= 0) S->ibound0[n] = S->ibound[n];
  }
}

static void npf_cf(orc_solution *S) {
  if (!S->inewton) npf_wd(S, S->kiter_cur); /* npf_cf :454-457 */
  for (int n = 0; n < S->nodes; n++) {
    if (S->icelltype[n] != 0) {
      double satn = (S->ibound[n] == 0) ? 0.0 : thksat(S, n, S->x[n]);
      S->sat[n] = satn;
    }
  }
}

static double conn_cond(const orc_solution *S, int n, int m, int ii, double hn,
                        double hm) {
  int jj = S->jas[ii];
  int ihc = S->ihc[jj];
  if (ihc == 0)
    return vcond(S->ibound[n], S->ibound[m], S->icelltype[n], S->icelltype[m],
                 S->inewton, S->ivarcv, S->idewatcv, S->condsat[jj], hn, hm,
                 S->hyc ? S->hyc[2 * jj + (n > m)] : S->k33[n], S->hyc ? S->hyc[2 * jj + (m > n)] : S->k33[m],
                 S->sat[n], S->sat[m], S->top[n], S->top[m], S->bot[n], S->bot[m], S->hwva[jj]);
  return hcond(S->ibound[n], S->ibound[m], S->icelltype[n], S->icelltype[m],
               S->inewton, ihc, S->icellavg, S->condsat[jj], hn, hm, S->sat[n],
               S->sat[m], S->hyc ? S->hyc[2 * jj + (n > m)] : S->k11[n],
               S->hyc ? S->hyc[2 * jj + (m > n)] : S->k11[m], S->top[n], S->top[m], S->bot[n],
               S->bot[m], S->cl1[jj], S->cl2[jj], S->hwva[jj]);
}

static void npf_fc(orc_solution *S) {
  double *amat = S->amat, *rhs = S->rhs, *hnew = S->x;
  for (int n = 0; n < S->nodes; n++) {
    for (int ii = S->ia[n] + 1; ii < S->ia[n + 1]; ii++) {
      int m = S->ja[ii];
      if (m < n) continue;
      int ihc = S->ihc[S->jas[ii]];
      double cond = conn_cond(S, n, m, ii, hnew[n], hnew[m]);
      if (ihc == 0 && S->iperched != 0) {
        if (S->icelltype[m] != 0 && hnew[m] < S->top[m]) {
          int idiag = S->ia[n];
          rhs[n] = rhs[n] - cond * S->bot[n];
          amat[idiag] += -cond;
          int isymcon = S->isym[ii];
          amat[isymcon] += cond;
          rhs[m] = rhs[m] + cond * S->bot[n];
          continue;
        }
      }
      int idiag = S->ia[n];
      amat[ii] += cond;
      amat[idiag] += -cond;
      int isymcon = S->isym[ii], idiagm = S->ia[m];
      amat[isymcon] += cond;
      amat[idiagm] += -cond;
    }
  }
}

static void npf_fn(orc_solution *S) {
  double *amat = S->amat, *rhs = S->rhs, *hnew = S->x;
  for (int n = 0; n < S->nodes; n++) {
    int idiag = S->ia[n];
    for (int ii = S->ia[n] + 1; ii < S->ia[n + 1]; ii++) {
      int m = S->ja[ii];
      int isymcon = S->isym[ii];
      if (m < n) continue;
      int jj = S->jas[ii];
      if (S->ihc[jj] == 0 && S->ivarcv == 0) continue;
      int iups = m;
      if (hnew[m] < hnew[n]) iups = n;
      int idn = n;
      if (iups == n) idn = m;
      if (S->icelltype[iups] == 0) continue;
      double topup = S->top[iups], botup = S->bot[iups];
      if (S->ihc[jj] == 2) {
        topup = fmin(S->top[n], S->top[m]);
        botup = fmax(S->bot[n], S->bot[m]);
      }
      double cond = S->condsat[jj];
      double consterm = -cond * (hnew[iups] - hnew[idn]);
      double derv = sQuadraticSaturationDerivative(topup, botup, hnew[iups], S->satomega);
      int idiagm = S->ia[m];
      if (iups == n) {
        double term = consterm * derv;
        rhs[n] = rhs[n] + term * hnew[n];
        rhs[m] = rhs[m] - term * hnew[n];
        amat[idiag] += term;
        if (S->ibound[m] > 0) amat[isymcon] += -term;
      } else {
        double term = -consterm * derv;
        rhs[n] = rhs[n] + term * hnew[m];
        rhs[m] = rhs[m] - term * hnew[m];
        if (S->ibound[n] > 0) amat[ii] += term;
        amat[idiagm] += -term;
      }
    }
  }
}

/* gwf-npf.f90:705-741 */
static void npf_nur(orc_solution *S, int *inewtonur, double *dxmax, int *locmax) {
  for (int n = 0; n < S->nodes; n++) {
    if (S->ibound[n] < 1) continue;
    if (S->icelltype[n] > 0) {
      double botm = S->bot[S->ibotnode[n]];
      if (S->x[n] < botm) {
        *inewtonur = 1;
        double xx = S->xtemp[n] * (1.0 - DP9) + botm * DP9;
        double dxx = S->x[n] - xx;
        if (fabs(dxx) > fabs(*dxmax)) {
          *locmax = n;
          *dxmax = dxx;
        }
        S->x[n] = xx;
        S->dxold[n] = 0.0;
      }
    }
  }
}

/* gwf-npf.f90:745-771, 798-865 */
static void npf_cq(orc_solution *S) {
  double *hnew = S->x;
  for (int n = 0; n < S->nodes; n++) {
    for (int ipos = S->ia[n] + 1; ipos < S->ia[n + 1]; ipos++) {
      int m = S->ja[ipos];
      if (m < n) continue;
      double hn = hnew[n], hm = hnew[m];
      double condnm = conn_cond(S, n, m, ipos, hn, hm);
      double hntemp = hn, hmtemp = hm;
      if (S->iperched != 0) {
        if (S->ihc[S->jas[ipos]] == 0) {
          if (n > m) {
            if (S->icelltype[n] != 0)
              if (hn < S->top[n]) hntemp = S->bot[m];
          } else {
            if (S->icelltype[m] != 0)
              if (hm < S->top[m]) hmtemp = S->bot[n];
          }
        }
      }
      double qnm = condnm * (hmtemp - hntemp);
      S->flowja[ipos] = qnm;
      S->flowja[S->isym[ipos]] = -qnm;
    }
  }
}

/* ---------------- STO ------------------------------------------------------ */
static double SsCapacity(int istor_coef, double top, double bot, double area,
                         double ss) {
  double thick = (istor_coef == 0) ? (top - bot) : 1.0;
  return ss * thick * area;
}

static void SsTerms(int iconvert, int iorig_ss, int iconf_ss, double top,
                    double bot, double rho1, double rho1old, double snnew,
                    double snold, double hnew, double hold, double *aterm,
                    double *rhsterm, double *rate) {
  *aterm = -rho1 * snnew;
  *rhsterm = 0.0;
  if (iconvert != 0) {
    if (iorig_ss == 0) {
      if (iconf_ss == 0) {
        double tthk = top - bot;
        double zold = bot + 0.5 * tthk * snold;
        double znew = bot + 0.5 * tthk * snnew;
        *rhsterm = -rho1old * snold * (hold - zold) - rho1 * snnew * znew;
      } else {
        if (snold == 1.0) *rhsterm = *rhsterm - rho1old * (hold - top);
        if (snnew == 1.0)
          *rhsterm = *rhsterm - rho1 * top;
        else
          *aterm = 0.0;
      }
    } else {
      *rhsterm = -rho1old * snold * hold;
    }
  } else {
    *rhsterm = -rho1old * snold * hold;
  }
  if (rate) *rate = *aterm * hnew - *rhsterm;
}

static void SyTerms(double top, double bot, double rho2, double rho2old,
                    double snnew, double snold, double *aterm, double *rhsterm,
                    double *rate) {
  *aterm = 0.0;
  double tthk = top - bot;
  if (snnew < 1.0) {
    if (snnew > 0.0) {
      *aterm = -rho2;
      *rhsterm = -rho2old * tthk * snold - rho2 * bot;
    } else {
      *rhsterm = tthk * (0.0 - rho2old * snold);
    }
  } else {
    *rhsterm = tthk * (rho2 * snnew - rho2old * snold);
  }
  if (rate) *rate = rho2old * tthk * snold - rho2 * tthk * snnew;
}

static void sto_fc(orc_solution *S) {
  if (S->iss != 0) return;
  double tled = 1.0 / S->delt;
  for (int n = 0; n < S->nodes; n++) {
    int idiag = S->ia[n];
    if (S->ibound[n] < 1) continue;
    double tp = S->top[n], bt = S->bot[n], snold, snnew;
    if (S->iconvert[n] == 0) {
      snold = 1.0;
      snnew = 1.0;
    } else {
      snold = sQuadraticSaturation(tp, bt, S->xold[n], S->satomega);
      snnew = sQuadraticSaturation(tp, bt, S->x[n], S->satomega);
    }
    double sc1 = SsCapacity(S->istor_coef, tp, bt, S->area[n], S->ss[n]);
    double rho1 = sc1 * tled, rho1old = rho1, aterm, rhsterm;
    SsTerms(S->iconvert[n], S->iorig_ss, S->iconf_ss, tp, bt, rho1, rho1old, snnew,
            snold, S->x[n], S->xold[n], &aterm, &rhsterm, NULL);
    S->amat[idiag] += aterm;
    S->rhs[n] = S->rhs[n] + rhsterm;
    if (S->iconvert[n] != 0) {
      rhsterm = 0.0;
      double sc2 = S->sy[n] * S->area[n];
      double rho2 = sc2 * tled, rho2old = rho2;
      SyTerms(tp, bt, rho2, rho2old, snnew, snold, &aterm, &rhsterm, NULL);
      S->amat[idiag] += aterm;
      S->rhs[n] = S->rhs[n] + rhsterm;
    }
  }
}

static void sto_fn(orc_solution *S) {
  if (S->iss != 0) return;
  double tled = 1.0 / S->delt;
  for (int n = 0; n < S->nodes; n++) {
    int idiag = S->ia[n];
    if (S->ibound[n] <= 0) continue;
    double tp = S->top[n], bt = S->bot[n], tthk = tp - bt, h = S->x[n];
    /* NB: sto_fn calls the smoothing functions WITHOUT eps => default 1e-6 */
    double snnew = sQuadraticSaturation(tp, bt, h, DEM6);
    double sc1 = SsCapacity(S->istor_coef, tp, bt, S->area[n], S->ss[n]);
    double sc2 = S->sy[n] * S->area[n];
    double rho1 = sc1 * tled, rho2 = sc2 * tled;
    if (S->iconvert[n] != 0) {
      double derv = sQuadraticSaturationDerivative(tp, bt, h, DEM6), drterm;
      if (S->iconf_ss == 0) {
        if (S->iorig_ss == 0)
          drterm = -rho1 * derv * (h - bt) + rho1 * tthk * snnew * derv;
        else
          drterm = -(rho1 * derv * h);
        S->amat[idiag] += drterm;
        S->rhs[n] = S->rhs[n] + drterm * h;
      }
      if (snnew < 1.0) {
        if (snnew > 0.0) {
          double rterm = -rho2 * tthk * snnew;
          drterm = -rho2 * tthk * derv;
          S->amat[idiag] += drterm + rho2;
          S->rhs[n] = S->rhs[n] - rterm + drterm * h + rho2 * bt;
        }
      }
    }
  }
}

static void sto_cq(orc_solution *S) {
  for (int n = 0; n < S->nodes; n++) S->strgss[n] = S->strgsy[n] = 0.0;
  if (S->iss != 0) return;
  double tled = 1.0 / S->delt;
  for (int n = 0; n < S->nodes; n++) {
    if (S->ibound[n] <= 0) continue;
    double tp = S->top[n], bt = S->bot[n], snold, snnew;
    if (S->iconvert[n] == 0) {
      snold = 1.0;
      snnew = 1.0;
    } else {
      snold = sQuadraticSaturation(tp, bt, S->xold[n], S->satomega);
      snnew = sQuadraticSaturation(tp, bt, S->x[n], S->satomega);
    }
    double sc1 = SsCapacity(S->istor_coef, tp, bt, S->area[n], S->ss[n]);
    double rho1 = sc1 * tled, rho1old = rho1, aterm, rhsterm, rate;
    SsTerms(S->iconvert[n], S->iorig_ss, S->iconf_ss, tp, bt, rho1, rho1old, snnew,
            snold, S->x[n], S->xold[n], &aterm, &rhsterm, &rate);
    S->strgss[n] = rate;
    int idiag = S->ia[n];
    S->flowja[idiag] = S->flowja[idiag] + rate;
    rate = 0.0;
    if (S->iconvert[n] != 0) {
      double sc2 = S->sy[n] * S->area[n];
      double rho2 = sc2 * tled, rho2old = rho2;
      SyTerms(tp, bt, rho2, rho2old, snnew, snold, &aterm, &rhsterm, &rate);
    }
    S->strgsy[n] = rate;
    S->flowja[idiag] = S->flowja[idiag] + rate;
  }
}

/* ---------------- boundary packages --------------------------------------- */
/* DiscretizationBase.f90:1077-1112: walk down the vertical connections (m > n, ihc == 0) to the first
 * cell that is not inactive, or to the bottom cell */
static int highest_active(const orc_solution *S, int n) {
  for (;;) {
    int below = -1;
    for (int ii = S->ia[n] + 1; ii < S->ia[n + 1]; ii++) {
      int m = S->ja[ii];
      if (S->ihc[S->jas[ii]] == 0 && m > n) {
        below = m;
        break;
      }
    }
    if (below < 0) return n; /* bottom cell */
    n = below;
    if (S->ibound[n] != 0) return n;
  }
}

static void bnd_cf(orc_solution *S, pkg_t *p) {
  for (int i = 0; i < p->nbound; i++) {
    int node = p->nodelist[i];
    switch (p->type) {
    case MF6GPU_PKG_CHD: /* chd_cf: nothing ; hcof/rhs stay 0 */
      p->hcof[i] = 0.0;
      p->rhs[i] = 0.0;
      break;
    case MF6GPU_PKG_WEL: { /* gwf-wel.f90:296-332 */
      p->hcof[i] = 0.0;
      if (S->ibound[node] <= 0) {
        p->rhs[i] = 0.0;
        break;
      }
      double q = p->b1[i];
      if (p->iflowred != 0 && q < 0.0) {
        if (S->icelltype[node] != 0) {
          double tp = S->top[node], bt = S->bot[node], thick = tp - bt;
          tp = bt + p->flowred * thick;
          q = q * sQSaturation(tp, bt, S->x[node]);
        }
      }
      p->rhs[i] = -q;
      break;
    }
    case MF6GPU_PKG_RIV: { /* gwf-riv.f90:270-299 */
      if (S->ibound[node] <= 0) {
        p->hcof[i] = 0.0;
        p->rhs[i] = 0.0;
        break;
      }
      double hriv = p->b1[i], criv = p->b2[i], rbot = p->b3[i];
      if (S->x[node] <= rbot) {
        p->rhs[i] = -criv * (hriv - rbot);
        p->hcof[i] = 0.0;
      } else {
        p->rhs[i] = -criv * hriv;
        p->hcof[i] = -criv;
      }
      break;
    }
    case MF6GPU_PKG_RCH: /* gwf-rch.f90:303-353; iflowred carries FIXED_CELL for this package type */
      if (p->iflowred == 0) {
        node = p->nodetop[i];
        if (S->ibound[node] == 0) node = highest_active(S, node);
        p->nodelist[i] = node;
      }
      p->hcof[i] = 0.0;
      p->rhs[i] = -p->b1[i] * S->area[node];
      if (S->ibound[node] <= 0) p->rhs[i] = 0.0;
      break;
    case MF6GPU_PKG_GHB: /* gwf-ghb.f90:245-265 */
      if (S->ibound[node] <= 0) {
        p->hcof[i] = 0.0;
        p->rhs[i] = 0.0;
        break;
      }
      p->hcof[i] = -p->b2[i];
      p->rhs[i] = -p->b2[i] * p->b1[i];
      break;
    case MF6GPU_PKG_DRN: { /* drn_cf + get_drain_factor, gwf-drn.f90:340-373, 534-574 */
      if (S->ibound[node] <= 0) {
        p->hcof[i] = 0.0;
        p->rhs[i] = 0.0;
        break;
      }
      double cdrn = p->b2[i], drndepth = p->b3[i], drntop, drnbot, fact;
      drain_elevations(p->b1[i], drndepth, &drntop, &drnbot);
      if (drndepth != 0.0) {
        if (p->iflowred != 0) /* DEV_CUBIC_SCALING */
          fact = sQSaturationC(drntop, drnbot, S->x[node], -1.0, 2.0);
        else
          fact = sQuadraticSaturation(drntop, drnbot, S->x[node], 0.0);
      } else {
        fact = (S->x[node] <= drnbot) ? 0.0 : 1.0;
      }
      p->rhs[i] = -fact * cdrn * drnbot;
      p->hcof[i] = -fact * cdrn;
      break;
    }
    }
  }
}

/* BoundaryPackage.f90:453-472 (chd_fc is a no-op, gwf-chd.f90:238-246) */
static void bnd_fc(orc_solution *S, pkg_t *p) {
  if (p->type == MF6GPU_PKG_CHD) return;
  for (int i = 0; i < p->nbound; i++) {
    int n = p->nodelist[i];
    S->rhs[n] = S->rhs[n] + p->rhs[i];
    S->amat[S->ia[n]] += p->hcof[i];
  }
}

/* gwf-wel.f90:378-424, gwf-drn.f90:420-470 */
static void bnd_fn(orc_solution *S, pkg_t *p) {
  if (p->type == MF6GPU_PKG_DRN) {
    for (int i = 0; i < p->nbound; i++) {
      int node = p->nodelist[i];
      if (S->ibound[node] <= 0) continue;
      double cdrn = p->b2[i], xnew = S->x[node], drndepth = p->b3[i], drntop, drnbot;
      drain_elevations(p->b1[i], drndepth, &drntop, &drnbot);
      if (drndepth != 0.0) {
        double drterm = sQSaturationDerivativeC(drntop, drnbot, xnew, -1.0, 2.0);
        drterm = drterm * cdrn * (drnbot - xnew);
        S->amat[S->ia[node]] += drterm;
        S->rhs[node] = S->rhs[node] + drterm * xnew;
      }
    }
    return;
  }
  if (p->type != MF6GPU_PKG_WEL) return;
  for (int i = 0; i < p->nbound; i++) {
    int node = p->nodelist[i];
    if (S->ibound[node] <= 0) continue;
    if (p->iflowred != 0 && S->icelltype[node] != 0) {
      double q = -p->rhs[i];
      if (q < 0.0) {
        double tp = S->top[node], bt = S->bot[node], thick = tp - bt;
        tp = bt + p->flowred * thick;
        double drterm = sQSaturationDerivative(tp, bt, S->x[node]);
        drterm = drterm * p->b1[i];
        S->amat[S->ia[node]] += drterm;
        S->rhs[node] = S->rhs[node] + drterm * S->x[node];
      }
    }
  }
}

/* BoundaryPackage.f90:583-619 */
static void bnd_cq_simrate(orc_solution *S, pkg_t *p) {
  for (int i = 0; i < p->nbound; i++) {
    int node = p->nodelist[i];
    double rrate = 0.0;
    int idiag = S->ia[node];
    if (S->ibound[node] > 0) rrate = p->hcof[i] * S->x[node] - p->rhs[i];
    S->flowja[idiag] = S->flowja[idiag] + rrate;
    p->simvals[i] = rrate;
  }
}

/* gwf-chd.f90:264-320 */
static void calc_chd_rate(orc_solution *S, pkg_t *p) {
  for (int i = 0; i < p->nbound; i++) {
    int node = p->nodelist[i];
    int idiag = S->ia[node];
    double rate = 0.0, ratein = 0.0, rateout = 0.0;
    for (int ipos = S->ia[node] + 1; ipos < S->ia[node + 1]; ipos++) {
      double q = S->flowja[ipos];
      rate = rate - q;
      int n2 = S->ja[ipos];
      if (S->ibound[n2] > 0) {
        if (q < 0.0)
          ratein = ratein - q;
        else
          rateout = rateout + q;
      }
    }
    p->rhs[i] = -rate;
    p->hcof[i] = 0.0;
    p->simvals[i] = rate;
    p->ratein[i] = ratein;
    p->rateout[i] = rateout;
    S->flowja[idiag] = S->flowja[idiag] + rate;
  }
}

/* Budget.f90:631-648 */
static void rate_accumulator(const double *flow, int n, double *rin, double *rout) {
  *rin = 0.0;
  *rout = 0.0;
  for (int i = 0; i < n; i++) {
    if (flow[i] < 0.0)
      *rout = *rout - flow[i];
    else
      *rin = *rin + flow[i];
  }
}

/* ---------------- solution ------------------------------------------------ */
static int *dup_idx(const int32_t *src, size_t n, int base) {
  int *d = (int *)malloc(sizeof(int) * (n ? n : 1));
  for (size_t i = 0; i < n; i++) d[i] = src[i] - base;
  return d;
}
static double *dup_d(const double *src, size_t n) {
  double *d = (double *)calloc(n ? n : 1, sizeof(double));
  if (src) memcpy(d, src, sizeof(double) * n);
  return d;
}
static int *dup_i(const int32_t *src, size_t n) {
  int *d = (int *)calloc(n ? n : 1, sizeof(int));
  if (src) memcpy(d, src, sizeof(int) * n);
  return d;
}

orc_solution *orc_sln_create(const mf6gpu_gwf_model *m,
                             const mf6gpu_sln_settings *ss,
                             const mf6gpu_ims_settings *ls, const int *perm) {
  orc_solution *S = (orc_solution *)calloc(1, sizeof(*S));
  const int base = m->index_base;
  size_t n = (size_t)m->nodes, nja = (size_t)m->nja, njas = (size_t)m->njas;
  S->nodes = m->nodes;
  S->nja = m->nja;
  S->njas = m->njas;
  S->ia = dup_idx(m->ia, n + 1, base);
  S->ja = dup_idx(m->ja, nja, base);
  S->jas = dup_idx(m->jas, nja, base);
  S->isym = dup_idx(m->isym, nja, base);
  S->ihc = dup_i(m->ihc, njas);
  S->cl1 = dup_d(m->cl1, njas);
  S->cl2 = dup_d(m->cl2, njas);
  S->hwva = dup_d(m->hwva, njas);
  S->top = dup_d(m->top, n);
  S->bot = dup_d(m->bot, n);
  S->area = dup_d(m->area, n);
  S->k11 = dup_d(m->k11, n);
  S->k33 = dup_d(m->k33 ? m->k33 : m->k11, n);
  S->ss = dup_d(m->ss, n);
  S->sy = dup_d(m->sy, n);
  S->icelltype = dup_i(m->icelltype, n);
  S->iconvert = dup_i(m->iconvert, n);
  S->ibound0 = dup_i(m->ibound, n);
  if (!m->ibound)
    for (size_t i = 0; i < n; i++) S->ibound0[i] = 1;
  S->ibound = dup_i(S->ibound0, n);
  if (m->ibotnode)
    S->ibotnode = dup_idx(m->ibotnode, n, base);
  else {
    S->ibotnode = (int *)malloc(sizeof(int) * n);
    for (size_t i = 0; i < n; i++) S->ibotnode[i] = (int)i;
  }
  S->icellavg = m->icellavg;
  S->inewton = m->inewton;
  S->inewtonur = m->inewtonur;
  S->iperched = m->iperched;
  S->ivarcv = m->ivarcv;
  S->idewatcv = m->idewatcv;
  S->insto = m->insto;
  S->istor_coef = m->istor_coef;
  S->iconf_ss = m->iconf_ss;
  S->iorig_ss = m->iorig_ss;
  S->satomega = (m->inewton > 0) ? DEM6 : 0.0; /* gwf-npf.f90:1395, gwf-sto.f90:827 */
  S->x = dup_d(m->strt, n);
  S->xold = dup_d(m->strt, n);
  S->sat = (double *)malloc(sizeof(double) * n);
  for (size_t i = 0; i < n; i++) S->sat[i] = 1.0;
  S->condsat = (double *)calloc(njas ? njas : 1, sizeof(double));
  S->amat = (double *)calloc(nja, sizeof(double));
  S->rhs = (double *)calloc(n, sizeof(double));
  S->xtemp = (double *)calloc(n, sizeof(double));
  S->dxold = (double *)calloc(n, sizeof(double));
  S->wsave = (double *)calloc(n, sizeof(double));
  S->hchold = (double *)calloc(n, sizeof(double));
  S->deold = (double *)calloc(n, sizeof(double));
  S->flowja = (double *)calloc(nja, sizeof(double));
  S->strgss = (double *)calloc(n, sizeof(double));
  S->strgsy = (double *)calloc(n, sizeof(double));
  S->resid = (double *)calloc(n, sizeof(double));
  S->ss_ = *ss;
  S->ims = orc_ims_create(S->nodes, S->nja, S->ia, S->ja, ls, perm);
  S->isymmetric = (ls->ilinmeth == 1) ? 1 : 0; /* NumericalSolution.f90:914-916 */
  calc_hyc(S, m);
  /* prepcheck (gwf-npf.f90:1838-1882): a negative ICELLTYPE means "convertible" without THICKSTRT; with THICKSTRT
   * the cell is confined with the saturated thickness of its STARTING head (calc_initial_sat :2046-2057) */
  S->sat0 = (double *)malloc(sizeof(double) * n);
  for (size_t i = 0; i < n; i++) {
    S->sat0[i] = 1.0;
    if (S->icelltype[i] < 0) {
      if (m->ithickstrt != 0) {
        if (S->ibound[i] != 0) S->sat0[i] = thksat(S, (int)i, m->strt[i]);
        S->sat[i] = S->sat0[i]; /* npf_cf never touches a confined cell: this%sat keeps the initial saturation */
        S->icelltype[i] = 0;
      } else {
        S->icelltype[i] = 1;
      }
    }
  }
  calc_condsat(S);
  S->wetdry = m->wetdry ? dup_d(m->wetdry, n) : NULL;
  S->irewet = (m->wetdry && m->irewet) ? 1 : 0;
  S->wetfct = m->wetfct;
  S->iwetit = m->iwetit > 0 ? m->iwetit : 1;
  S->ihdwet = m->ihdwet;
  /* prepcheck (gwf-npf.f90:1817-1821): without NEWTON the wet/dry routine runs once on the initial heads, kiter = 0 */
  if (!S->inewton) npf_wd(S, 0);
  return S;
}

static void free_pkgs(orc_solution *S) {
  for (int k = 0; k < S->npkg; k++) {
    pkg_t *p = &S->pkg[k];
    free(p->nodelist); free(p->nodetop); free(p->b1); free(p->b2); free(p->b3); free(p->hcof);
    free(p->rhs); free(p->simvals); free(p->ratein); free(p->rateout);
  }
  free(S->pkg);
  S->pkg = NULL;
  S->npkg = 0;
}

void orc_sln_set_blocks(orc_solution *S, const int *block) { orc_ims_set_blocks(S->ims, block); }

void orc_sln_destroy(orc_solution *S) {
  if (!S) return;
  free_pkgs(S);
  free(S->ia); free(S->ja); free(S->jas); free(S->isym); free(S->ihc);
  free(S->cl1); free(S->cl2); free(S->hwva); free(S->top); free(S->bot);
  free(S->area); free(S->k11); free(S->k33); free(S->ss); free(S->sy); free(S->hyc); free(S->wetdry); free(S->sat0);
  free(S->hfb_n); free(S->hfb_m); free(S->hfb_pos); free(S->hfb_hydchr); free(S->hfb_condsav); free(S->hfb_csatsav);
  free(S->gnc_n); free(S->gnc_m); free(S->gnc_j); free(S->gnc_pos); free(S->gnc_alpha); free(S->gnc_cond);
  free(S->icelltype); free(S->iconvert); free(S->ibound0); free(S->ibound);
  free(S->ibotnode); free(S->x); free(S->xold); free(S->sat); free(S->condsat);
  free(S->amat); free(S->rhs); free(S->xtemp); free(S->dxold); free(S->wsave);
  free(S->hchold); free(S->deold); free(S->flowja); free(S->strgss);
  free(S->strgsy); free(S->resid);
  orc_ims_destroy(S->ims);
  free(S);
}

void orc_sln_set_packages(orc_solution *S, int npkg, const mf6gpu_bnd_package *pk) {
  /* chd_rp (gwf-chd.f90:124-141): the cells of the previous constant-head list become ordinary active cells */
  for (int k = 0; k < S->npkg; k++)
    if (S->pkg[k].type == MF6GPU_PKG_CHD)
      for (int i = 0; i < S->pkg[k].nbound; i++) S->ibound0[S->pkg[k].nodelist[i]] = 1;
  free_pkgs(S);
  S->npkg = npkg;
  S->pkg = (pkg_t *)calloc((size_t)(npkg ? npkg : 1), sizeof(pkg_t));
  memcpy(S->ibound, S->ibound0, sizeof(int) * (size_t)S->nodes);
  for (int k = 0; k < npkg; k++) {
    pkg_t *p = &S->pkg[k];
    size_t nb = (size_t)pk[k].nbound;
    p->type = pk[k].type;
    p->nbound = pk[k].nbound;
    p->iflowred = pk[k].iflowred;
    p->flowred = pk[k].flowred;
    p->nodelist = dup_idx(pk[k].nodelist, nb, pk[k].index_base);
    p->nodetop = dup_idx(pk[k].nodelist, nb, pk[k].index_base);
    p->b1 = dup_d(pk[k].b1, nb);
    p->b2 = dup_d(pk[k].b2, nb);
    p->b3 = dup_d(pk[k].b3, nb);
    p->hcof = (double *)calloc(nb ? nb : 1, sizeof(double));
    p->rhs = (double *)calloc(nb ? nb : 1, sizeof(double));
    p->simvals = (double *)calloc(nb ? nb : 1, sizeof(double));
    p->ratein = (double *)calloc(nb ? nb : 1, sizeof(double));
    p->rateout = (double *)calloc(nb ? nb : 1, sizeof(double));
    if (p->type == MF6GPU_PKG_CHD) /* chd_rp gwf-chd.f90:143-155 */
      for (size_t i = 0; i < nb; i++) S->ibound[p->nodelist[i]] = -(k + 1);
  }
}

/* ---- HFB (gwf-hfb.f90) ------------------------------------------------------------------- */
static double hfb_faheight(const orc_solution *S, int n, int m, int jj, int use_heads) {
  double topn = S->top[n], topm = S->top[m], botn = S->bot[n], botm = S->bot[m];
  if (use_heads) {
    if (S->icelltype[n] != 0 && S->x[n] < topn) topn = S->x[n];
    if (S->icelltype[m] != 0 && S->x[m] < topm) topm = S->x[m];
  }
  if (S->ihc[jj] == 2) {
    double t = topn < topm ? topn : topm, b = botn > botm ? botn : botm;
    return t - b;
  }
  return 0.5 * ((topn - botn) + (topm - botm));
}

/* hfb_rp: condsat_reset + read_data + condsat_modify (:149-201, 770-832); nodes 0-based */
void orc_sln_set_hfb(orc_solution *S, int nhfb, const int *noden, const int *nodem, const double *hydchr) {
  for (int i = 0; i < S->nhfb; i++) S->condsat[S->jas[S->hfb_pos[i]]] = S->hfb_csatsav[i]; /* condsat_reset */
  free(S->hfb_n); free(S->hfb_m); free(S->hfb_pos); free(S->hfb_hydchr); free(S->hfb_condsav); free(S->hfb_csatsav);
  S->nhfb = nhfb;
  size_t c = (size_t)(nhfb ? nhfb : 1);
  S->hfb_n = (int *)calloc(c, sizeof(int));
  S->hfb_m = (int *)calloc(c, sizeof(int));
  S->hfb_pos = (int *)calloc(c, sizeof(int));
  S->hfb_hydchr = (double *)calloc(c, sizeof(double));
  S->hfb_condsav = (double *)calloc(c, sizeof(double));
  S->hfb_csatsav = (double *)calloc(c, sizeof(double));
  for (int i = 0; i < nhfb; i++) {
    int n = noden[i], m = nodem[i], pos = -1;
    for (int p = S->ia[n] + 1; p < S->ia[n + 1]; p++)
      if (S->ja[p] == m) pos = p;
    S->hfb_n[i] = n;
    S->hfb_m[i] = m;
    S->hfb_pos[i] = pos; /* idxloc: position of (n, m) in ja; < 0 = the cells are not connected (input error) */
    S->hfb_hydchr[i] = hydchr[i];
  }
  for (int i = 0; i < nhfb; i++) { /* condsat_modify */
    if (S->hfb_pos[i] < 0) continue;
    int n = S->hfb_n[i], m = S->hfb_m[i], jj = S->jas[S->hfb_pos[i]];
    double cond = S->condsat[jj];
    S->hfb_csatsav[i] = cond;
    if (S->inewton == 1 || (S->icelltype[n] == 0 && S->icelltype[m] == 0)) {
      if (S->hfb_hydchr[i] > 0.0) {
        double condhfb = S->hfb_hydchr[i] * S->hwva[jj] * hfb_faheight(S, n, m, jj, 0);
        cond = cond * condhfb / (cond + condhfb);
      } else {
        cond = -cond * S->hfb_hydchr[i];
      }
      S->condsat[jj] = cond;
    }
  }
}

/* hfb_fc without XT3D (:296-345): Picard with a convertible cell on either side */
static void hfb_fc(orc_solution *S) {
  if (S->inewton != 0) return;
  for (int i = 0; i < S->nhfb; i++) {
    int ipos = S->hfb_pos[i];
    if (ipos < 0) continue;
    double aterm = S->amat[ipos];
    int n = S->hfb_n[i], m = S->hfb_m[i];
    if (S->ibound[n] == 0 || S->ibound[m] == 0) continue;
    if (S->icelltype[n] != 0 || S->icelltype[m] != 0) {
      int jj = S->jas[ipos];
      double cond;
      if (S->hfb_hydchr[i] > 0.0) {
        double condhfb = S->hfb_hydchr[i] * 1.0 * S->hwva[jj] * hfb_faheight(S, n, m, jj, 1);
        cond = aterm * condhfb / (aterm + condhfb);
      } else {
        cond = -aterm * S->hfb_hydchr[i];
      }
      S->hfb_condsav[i] = cond;
      S->amat[S->ia[n]] += aterm - cond;
      S->amat[ipos] = cond;
      S->amat[S->ia[m]] += aterm - cond;
      S->amat[S->isym[ipos]] = cond;
    }
  }
}

/* hfb_cq without XT3D (:432-449) */
static void hfb_cq(orc_solution *S) {
  if (S->inewton != 0) return;
  for (int i = 0; i < S->nhfb; i++) {
    int ipos = S->hfb_pos[i];
    if (ipos < 0) continue;
    int n = S->hfb_n[i], m = S->hfb_m[i];
    if (S->ibound[n] == 0 || S->ibound[m] == 0) continue;
    if (S->icelltype[n] != 0 || S->icelltype[m] != 0) {
      double qnm = S->hfb_condsav[i] * (S->x[m] - S->x[n]);
      S->flowja[ipos] = qnm;
      S->flowja[S->isym[ipos]] = -qnm;
    }
  }
}

/* ---- GNC: ghost node correction, EXPLICIT (GhostNode.f90).  noden / nodem: the connected pair; nodesj[numj] per
 * entry: the contributing cells of noden's grid (< 0 = none) with weights alphasj.  Nodes 0-based. ---- */
void orc_sln_set_gnc(orc_solution *S, int ngnc, int numj, const int *noden, const int *nodem, const int *nodesj,
                     const double *alphasj) {
  free(S->gnc_n); free(S->gnc_m); free(S->gnc_j); free(S->gnc_pos); free(S->gnc_alpha); free(S->gnc_cond);
  S->ngnc = ngnc;
  S->gnc_numj = numj;
  size_t c = (size_t)(ngnc ? ngnc : 1), cj = c * (size_t)(numj ? numj : 1);
  S->gnc_n = (int *)calloc(c, sizeof(int));
  S->gnc_m = (int *)calloc(c, sizeof(int));
  S->gnc_pos = (int *)calloc(c, sizeof(int));
  S->gnc_cond = (double *)calloc(c, sizeof(double));
  S->gnc_j = (int *)calloc(cj, sizeof(int));
  S->gnc_alpha = (double *)calloc(cj, sizeof(double));
  for (int i = 0; i < ngnc; i++) {
    int n = noden[i], m = nodem[i], pos = -1;
    for (int p = S->ia[n] + 1; p < S->ia[n + 1]; p++)
      if (S->ja[p] == m) pos = p;
    S->gnc_n[i] = n;
    S->gnc_m[i] = m;
    S->gnc_pos[i] = pos; /* idxglo (gnc_mc :172-246); < 0 = not connected (input error there) */
    for (int k = 0; k < numj; k++) {
      S->gnc_j[i * numj + k] = nodesj[i * numj + k];
      S->gnc_alpha[i * numj + k] = alphasj[i * numj + k];
    }
  }
}

/* gnc_fmsav :251-273 + the explicit branch of gnc_fc :280-324 */
static void gnc_fc(orc_solution *S) {
  for (int i = 0; i < S->ngnc; i++) S->gnc_cond[i] = (S->gnc_pos[i] >= 0) ? S->amat[S->gnc_pos[i]] : 0.0;
  for (int i = 0; i < S->ngnc; i++) {
    int n = S->gnc_n[i], m = S->gnc_m[i];
    if (S->ibound[n] == 0 || S->ibound[m] == 0) continue;
    double cond = S->gnc_cond[i];
    for (int k = 0; k < S->gnc_numj; k++) {
      int j = S->gnc_j[i * S->gnc_numj + k];
      if (j < 0) continue;
      double alpha = S->gnc_alpha[i * S->gnc_numj + k];
      if (alpha == 0.0) continue;
      double aterm = alpha * cond;
      double rterm = aterm * (S->x[n] - S->x[j]);
      S->rhs[n] = S->rhs[n] - rterm;
      S->rhs[m] = S->rhs[m] + rterm;
    }
  }
}

/* gnc_fn :340-443 (single-model arguments of gwf_fc, gwf.f90:521-528) */
static void gnc_fn(orc_solution *S) {
  for (int i = 0; i < S->ngnc; i++) {
    int n = S->gnc_n[i], m = S->gnc_m[i], ipos = S->gnc_pos[i];
    if (S->ibound[n] == 0 || S->ibound[m] == 0 || ipos < 0) continue;
    int jj = S->jas[ipos], ihc = S->ihc[jj];
    double csat = S->condsat[jj];
    if (ihc == 0 && S->ivarcv == 0) continue;
    int iups = (S->x[m] > S->x[n]) ? 1 : 0;
    int up = iups ? m : n;
    double topup = S->top[up], botup = S->bot[up], xup = S->x[up];
    if (S->icelltype[up] == 0) continue;
    if (ihc == 2) {
      topup = S->top[n] < S->top[m] ? S->top[n] : S->top[m];
      botup = S->bot[n] > S->bot[m] ? S->bot[n] : S->bot[m];
    }
    for (int k = 0; k < S->gnc_numj; k++) {
      int j = S->gnc_j[i * S->gnc_numj + k];
      if (j < 0) continue;
      if (S->ibound[j] == 0) continue;
      double alpha = S->gnc_alpha[i * S->gnc_numj + k];
      if (alpha == 0.0) continue;
      double consterm = csat * alpha * (S->x[n] - S->x[j]);
      double derv = sQuadraticSaturationDerivative(topup, botup, xup, DEM6);
      double term = consterm * derv;
      if (iups == 0) {
        S->amat[S->ia[n]] += term;
        if (S->ibound[m] > 0) S->amat[S->isym[ipos]] += -term;
        S->rhs[n] = S->rhs[n] + term * S->x[n];
        S->rhs[m] = S->rhs[m] - term * S->x[n];
      } else {
        S->amat[S->ia[m]] += -term;
        if (S->ibound[n] > 0) S->amat[ipos] += term;
        S->rhs[n] = S->rhs[n] + term * S->x[m];
        S->rhs[m] = S->rhs[m] - term * S->x[m];
      }
    }
  }
}

/* gnc_cq :478-503 with deltaQgnc :509-542 */
static void gnc_cq(orc_solution *S) {
  for (int i = 0; i < S->ngnc; i++) {
    int n = S->gnc_n[i], m = S->gnc_m[i], ipos = S->gnc_pos[i];
    double dq = 0.0;
    if (S->ibound[n] != 0 && S->ibound[m] != 0) {
      double sigalj = 0.0, hd = 0.0;
      for (int k = 0; k < S->gnc_numj; k++) {
        int j = S->gnc_j[i * S->gnc_numj + k];
        if (j < 0) continue;
        if (S->ibound[j] == 0) continue;
        double alpha = S->gnc_alpha[i * S->gnc_numj + k];
        sigalj = sigalj + alpha;
        hd = hd + alpha * S->x[j];
      }
      double aterm = sigalj * S->x[n] - hd;
      dq = aterm * S->gnc_cond[i];
    }
    if (ipos < 0) continue;
    S->flowja[ipos] = S->flowja[ipos] + dq;
    S->flowja[S->isym[ipos]] = S->flowja[S->isym[ipos]] - dq;
  }
}

/* sln_buildsystem :1941-1991 (single model, no exchanges) */
static void buildsystem(orc_solution *S, int inewton) {
  memset(S->amat, 0, sizeof(double) * (size_t)S->nja);
  memset(S->rhs, 0, sizeof(double) * (size_t)S->nodes);
  /* gwf_cf */
  npf_cf(S);
  for (int k = 0; k < S->npkg; k++) bnd_cf(S, &S->pkg[k]);
  /* gwf_fc */
  npf_fc(S);
  if (S->nhfb > 0) hfb_fc(S); /* gwf_fc: right after npf_fc (gwf.f90:500-506) */
  if (S->ngnc > 0) gnc_fc(S);
  if (S->insto) sto_fc(S);
  for (int k = 0; k < S->npkg; k++) bnd_fc(S, &S->pkg[k]);
  if (inewton && S->inewton) {
    npf_fn(S);
    if (S->ngnc > 0) gnc_fn(S);
    if (S->insto) sto_fn(S);
    for (int k = 0; k < S->npkg; k++) bnd_fn(S, &S->pkg[k]);
  }
}

/* sln_calc_residual :2966-2982 */
static void calc_residual(orc_solution *S, double *r) {
  orc_amux(S->nodes, S->x, r, S->amat, S->ja, S->ia);
  for (int i = 0; i < S->nodes; i++) r[i] = r[i] + (-1.0) * S->rhs[i];
  for (int i = 0; i < S->nodes; i++)
    if (S->ibound[i] < 1) r[i] = 0.0;
}

static double l2norm_resid(orc_solution *S) {
  calc_residual(S, S->resid);
  double n2 = 0.0;
  for (int i = 0; i < S->nodes; i++) n2 = n2 + S->resid[i] * S->resid[i];
  return sqrt(n2);
}

/* sln_calc_ptc :2936-2962 + gwf_ptc gwf.f90:625-687 */
static void calc_ptc(orc_solution *S, int *iptc, double *ptcf) {
  *iptc = 0;
  *ptcf = 0.0;
  calc_residual(S, S->resid);
  int iptct = 0;
  if (S->iss > 0) iptct = S->inewton;
  if (iptct > 0) {
    for (int n = 0; n < S->nodes; n++) {
      if (S->ibound[n] < 1) continue;
      double v = S->area[n] * (S->top[n] - S->bot[n]); /* get_cell_volume(n, top) */
      double ptcdelem1 = fabs(S->resid[n]) / v;
      if (ptcdelem1 > *ptcf) *ptcf = ptcdelem1;
    }
    if (*ptcf == 0.0) *ptcf = 1.0 / (S->delt * 10.0);
  }
  if (*iptc == 0)
    if (iptct > 0) *iptc = 1;
}

/* pre-solve fix-ups of sln_ls :2434-2573 */
static void ls_fixups(orc_solution *S, int kiter, int kstp, int kper, int iptc,
                      double ptcf) {
  const int n = S->nodes;
  for (int ieq = 0; ieq < n; ieq++) {
    S->xtemp[ieq] = S->x[ieq];
    int id = S->ia[ieq];
    if (S->ibound[ieq] > 0) {
      double diagval = -1.0;
      double adiag = fabs(S->amat[id]);
      if (adiag < DEM15) {
        S->amat[id] = diagval;
        S->rhs[ieq] = S->rhs[ieq] + diagval * S->x[ieq];
      }
    } else {
      S->amat[id] = 1.0;
      for (int ipos = S->ia[ieq] + 1; ipos < S->ia[ieq + 1]; ipos++) S->amat[ipos] = 0.0;
      S->rhs[ieq] = S->x[ieq];
    }
  }
  if (S->isymmetric == 1) {
    for (int ieq = 0; ieq < n; ieq++) {
      if (S->ibound[ieq] > 0) {
        for (int ipos = S->ia[ieq]; ipos < S->ia[ieq + 1]; ipos++) {
          int jcol = S->ja[ipos];
          if (jcol == ieq) continue;
          if (S->ibound[jcol] < 0) {
            S->rhs[ieq] = S->rhs[ieq] - (S->amat[ipos] * S->x[jcol]);
            S->amat[ipos] = 0.0;
          }
        }
      }
    }
  }
  /* pseudo transient continuation */
  int iallowptc;
  if (S->ss_.iallowptc < 0)
    iallowptc = (kper > 1) ? 1 : 0;
  else
    iallowptc = S->ss_.iallowptc;
  int iptct = iptc * iallowptc;
  double l2norm = 0.0;
  if (iptct != 0) {
    l2norm = l2norm_resid(S);
    if (kiter == 1) {
      if (kper > 1 || kstp > 1)
        if (l2norm <= S->l2norm0) iptc = 0;
    } else {
      if (orc_is_close(l2norm, S->l2norm0)) iptc = 0;
    }
  }
  iptct = iptc * iallowptc;
  if (iptct != 0) {
    if (kiter == 1) {
      S->ptcdel = 1.0 / ptcf; /* ptcdel0 = 0, iptcopt = 0 defaults */
    } else {
      if (l2norm > 0.0)
        S->ptcdel = S->ptcdel * pow(S->l2norm0 / l2norm, 1.0); /* ptcexp = 1 */
      else
        S->ptcdel = 0.0;
    }
    double ptcval = (S->ptcdel > 0.0) ? 1.0 / S->ptcdel : 1.0;
    for (int ieq = 0; ieq < n; ieq++) {
      if (S->ibound[ieq] > 0) {
        S->amat[S->ia[ieq]] += -ptcval;
        S->rhs[ieq] = S->rhs[ieq] - ptcval * S->x[ieq];
      }
    }
    S->l2norm0 = l2norm;
  }
}

void orc_sln_formulate(orc_solution *S, int kiter, double delt, int iss) {
  S->delt = delt;
  S->iss = iss;
  S->kiter_cur = kiter;
  buildsystem(S, 1);
  int iptc;
  double ptcf;
  calc_ptc(S, &iptc, &ptcf);
  ls_fixups(S, kiter, 1, 1, iptc, ptcf);
}

/* sln_underrelax :2989-3114 */
static void underrelax(orc_solution *S, int kiter, double bigch) {
  const int n = S->nodes;
  const mf6gpu_sln_settings *c = &S->ss_;
  double *x = S->x, *xtemp = S->xtemp;
  if (c->nonmeth == 1) {
    for (int i = 0; i < n; i++) {
      if (S->ibound[i] < 1) continue;
      double delx = x[i] - xtemp[i];
      S->dxold[i] = delx;
      x[i] = xtemp[i] + c->gamma * delx;
    }
  } else if (c->nonmeth == 2) {
    double relax;
    S->bigch = bigch;
    if (kiter == 1) {
      relax = 1.0;
      S->relaxold = 1.0;
      S->bigchold = bigch;
    } else {
      double es = S->bigch / (S->bigchold * S->relaxold);
      double aes = fabs(es);
      if (es < -1.0)
        relax = 0.5 / aes;
      else
        relax = (3.0 + es) / (3.0 + aes);
    }
    S->relaxold = relax;
    S->bigchold = (1.0 - c->gamma) * S->bigch + c->gamma * S->bigchold;
    if (relax < 1.0) {
      for (int i = 0; i < n; i++) {
        if (S->ibound[i] < 1) continue;
        double delx = x[i] - xtemp[i];
        S->dxold[i] = delx;
        x[i] = xtemp[i] + relax * delx;
      }
    }
  } else if (c->nonmeth == 3) {
    for (int i = 0; i < n; i++) {
      if (S->ibound[i] < 1) continue;
      double delx = x[i] - xtemp[i];
      if (kiter == 1) {
        S->wsave[i] = 1.0;
        S->hchold[i] = DEM20;
        S->deold[i] = 0.0;
      }
      double ww;
      if (S->deold[i] * delx < 0.0)
        ww = c->theta * S->wsave[i];
      else
        ww = S->wsave[i] + c->akappa;
      if (ww > 1.0) ww = 1.0;
      S->wsave[i] = ww;
      if (kiter == 1)
        S->hchold[i] = delx;
      else
        S->hchold[i] = (1.0 - c->gamma) * delx + c->gamma * S->hchold[i];
      S->deold[i] = delx;
      S->dxold[i] = delx;
      double amom = 0.0;
      if (kiter > 4) amom = c->amomentum;
      delx = delx * ww + amom * S->hchold[i];
      x[i] = xtemp[i] + delx;
    }
  }
}

static void get_dxmax(orc_solution *S, double *hncg, int *lrch) {
  int nb = -1;
  double bigch = 0.0, abigch = 0.0;
  for (int i = 0; i < S->nodes; i++) {
    if (S->ibound[i] < 1) continue;
    double hdif = S->x[i] - S->xtemp[i];
    double ahdif = fabs(hdif);
    if (ahdif > abigch) {
      bigch = hdif;
      abigch = ahdif;
      nb = i;
    }
  }
  *hncg = bigch;
  *lrch = nb;
}

/* sln_backtracking :2680-2776 (+ get_backtracking_flag / apply_backtracking :2790-2842) */
static void backtracking(orc_solution *S, int kiter) {
  const mf6gpu_sln_settings *c = &S->ss_;
  S->kiter_cur = kiter;
  buildsystem(S, 0);
  if (kiter == 1) {
    S->res_prev = l2norm_resid(S);
  } else {
    S->res_new = l2norm_resid(S);
  }
  if (kiter > 1) {
    if (S->res_new > S->res_prev * c->btol) {
      for (int nb = 1; nb <= c->numtrack; nb++) {
        /* get_backtracking_flag */
        double dx_abs_max = 0.0;
        for (int i = 0; i < S->nodes; i++) {
          if (S->ibound[i] < 1) continue;
          double dx_abs = fabs(S->x[i] - S->xtemp[i]);
          if (dx_abs > dx_abs_max) dx_abs_max = dx_abs;
        }
        if (!(c->breduc * dx_abs_max >= c->dvclose)) break;
        /* apply_backtracking */
        for (int i = 0; i < S->nodes; i++) {
          if (S->ibound[i] < 1) continue;
          double delx = c->breduc * (S->x[i] - S->xtemp[i]);
          S->x[i] = S->xtemp[i] + delx;
        }
        S->nbacktracks++;
        buildsystem(S, 0);
        S->res_new = l2norm_resid(S);
        if (nb == c->numtrack) break;
        if (S->res_new < S->res_prev * c->btol) break;
        if (S->res_new < c->res_lim) break;
      }
    }
    S->res_prev = S->res_new;
  }
}

/* solve(kiter) :1482-1837 ; returns inner iterations */
static int solve_outer(orc_solution *S, int kiter, int kstp, int kper,
                       double *hncg, int *lrch) {
  double t0 = now_s();
  S->kiter_cur = kiter;
  if (S->ss_.numtrack > 0) backtracking(S, kiter);
  buildsystem(S, 1);
  int iptc;
  double ptcf;
  calc_ptc(S, &iptc, &ptcf);
  double t1 = now_s();
  S->t_form += t1 - t0;
  ls_fixups(S, kiter, kstp, kper, iptc, ptcf);
  int icnvg_lin = 0;
  int iter = orc_ims_apply(S->ims, S->amat, S->x, S->rhs, &icnvg_lin, kstp, kiter, NULL);
  S->t_ls += now_s() - t1;
  get_dxmax(S, hncg, lrch);
  S->icnvg = 0;
  if (fabs(*hncg) <= S->ss_.dvclose) S->icnvg = 1;
  if (S->icnvg != 1) {
    if (S->ss_.nonmeth > 0) {
      underrelax(S, kiter, *hncg);
    } else {
      for (int i = 0; i < S->nodes; i++) /* sln_calcdx */
        S->dxold[i] = (S->ibound[i] < 1) ? 0.0 : S->x[i] - S->xtemp[i];
    }
    int inewtonur = 0, locmax_nur = -1;
    double dxmax_nur = 0.0;
    if (S->inewton != 0 && S->inewtonur != 0) /* gwf_nur gwf.f90:696-734 */
      npf_nur(S, &inewtonur, &dxmax_nur, &locmax_nur);
    if (inewtonur != 0) {
      double dxold_max = 0.0; /* sln_maxval: largest |dxold| */
      for (int i = 0; i < S->nodes; i++)
        if (fabs(S->dxold[i]) > fabs(dxold_max)) dxold_max = S->dxold[i];
      if (fabs(dxold_max) <= S->ss_.dvclose && fabs(*hncg) <= S->ss_.dvclose) {
        S->icnvg = 1;
        get_dxmax(S, hncg, lrch);
      }
    }
  }
  return iter;
}

int orc_sln_timestep(orc_solution *S, int kper, int kstp, double delt, int iss,
                     mf6gpu_step_report *rep) {
  S->delt = delt;
  S->iss = iss;
  double tf0 = S->t_form, tl0 = S->t_ls;
  /* prepareSolve -> gwf_ad :396-442 ; chd_ad gwf-chd.f90:175-197 */
  for (int i = 0; i < S->nodes; i++) S->xold[i] = S->x[i];
  /* npf_ad (gwf-npf.f90:393-408): a dry wettable cell starts the step with hold = bottom, hnew = HDRY */
  if (S->irewet > 0)
    for (int i = 0; i < S->nodes; i++) {
      if (S->wetdry[i] == 0.0 || S->ibound[i] != 0) continue;
      S->xold[i] = S->bot[i];
      S->x[i] = DHDRY;
    }
  for (int k = 0; k < S->npkg; k++) {
    pkg_t *p = &S->pkg[k];
    if (p->type != MF6GPU_PKG_CHD) continue;
    for (int i = 0; i < p->nbound; i++) {
      int node = p->nodelist[i];
      S->x[node] = p->b1[i];
      S->xold[node] = S->x[node];
    }
  }
  int kiter, inner_total = 0, lrch = -1;
  double hncg = 0.0;
  S->icnvg = 0;
  S->nbacktracks = 0;
  for (kiter = 1; kiter <= S->ss_.mxiter; kiter++) {
    inner_total += solve_outer(S, kiter, kstp, kper, &hncg, &lrch);
    if (S->icnvg == 1) break;
  }
  if (kiter > S->ss_.mxiter) kiter = S->ss_.mxiter;
  /* finalizeSolve: gwf_cq :741-778 */
  memset(S->flowja, 0, sizeof(double) * (size_t)S->nja);
  npf_cq(S);
  if (S->nhfb > 0) hfb_cq(S);
  if (S->ngnc > 0) gnc_cq(S);
  if (S->insto) sto_cq(S);
  for (int k = 0; k < S->npkg; k++) {
    pkg_t *p = &S->pkg[k];
    bnd_cf(S, p);
    if (p->type == MF6GPU_PKG_CHD)
      ; /* chd_cq does nothing; rate comes in chd_bd */
    else
      bnd_cq_simrate(S, p);
  }
  /* gwf_bd :785-824 : csr_diagsum then budget entries */
  for (int n = 0; n < S->nodes; n++) {
    int idiag = S->ia[n];
    for (int ipos = S->ia[n] + 1; ipos < S->ia[n + 1]; ipos++)
      S->flowja[idiag] = S->flowja[idiag] + S->flowja[ipos];
  }
  if (rep) {
    memset(rep, 0, sizeof(*rep));
    int nt = 0;
    double totrin = 0.0, totrot = 0.0;
    if (S->insto) {
      double rin, rout;
      rate_accumulator(S->strgss, S->nodes, &rin, &rout);
      rep->term_id[nt] = 100; rep->term_in[nt] = rin; rep->term_out[nt] = rout; nt++;
      totrin += rin; totrot += rout;
      rate_accumulator(S->strgsy, S->nodes, &rin, &rout);
      rep->term_id[nt] = 101; rep->term_in[nt] = rin; rep->term_out[nt] = rout; nt++;
      totrin += rin; totrot += rout;
    }
    for (int k = 0; k < S->npkg && nt < MF6GPU_MAX_BUDGET_TERMS; k++) {
      pkg_t *p = &S->pkg[k];
      double rin, rout, dum;
      if (p->type == MF6GPU_PKG_CHD) {
        calc_chd_rate(S, p);
        rate_accumulator(p->ratein, p->nbound, &rin, &dum);
        rate_accumulator(p->rateout, p->nbound, &rout, &dum);
      } else {
        rate_accumulator(p->simvals, p->nbound, &rin, &rout);
      }
      rep->term_id[nt] = p->type; rep->term_in[nt] = rin; rep->term_out[nt] = rout; nt++;
      totrin += rin; totrot += rout;
    }
    rep->nterms = nt;
    rep->totrin = totrin;
    rep->totrot = totrot;
    double avgrat = (totrin + totrot) / 2.0;
    rep->pdiffr = (avgrat != 0.0) ? 100.0 * (totrin - totrot) / avgrat : 0.0;
    rep->converged = S->icnvg;
    rep->outer_iterations = kiter;
    rep->inner_iterations = inner_total;
    rep->max_dv = hncg;
    rep->max_dv_loc = lrch + 1;
    rep->npivot_fixes = S->ims->npivfix;
    rep->nbacktracks = S->nbacktracks;
    rep->t_formulate = S->t_form - tf0;
    rep->t_linsolve = S->t_ls - tl0;
  } else {
    for (int k = 0; k < S->npkg; k++)
      if (S->pkg[k].type == MF6GPU_PKG_CHD) calc_chd_rate(S, &S->pkg[k]);
  }
  return S->icnvg;
}

double *orc_sln_x(orc_solution *S) { return S->x; }
double *orc_sln_flowja(orc_solution *S) { return S->flowja; }
const double *orc_sln_amat(orc_solution *S) { return S->amat; }
const double *orc_sln_rhs(orc_solution *S) { return S->rhs; }
const double *orc_sln_condsat(orc_solution *S) { return S->condsat; }
/* simulated rates of package k (bnd_cq_simrate / calc_chd_rate) and the STO-SS / STO-SY rates per cell
 * (sto_cq) of the last time step: what the budget file records hold */
const double *orc_sln_simvals(orc_solution *S, int k) {
  return (k >= 0 && k < S->npkg) ? S->pkg[k].simvals : 0;
}
const int *orc_sln_nodes(orc_solution *S, int k) { return (k >= 0 && k < S->npkg) ? S->pkg[k].nodelist : 0; }
int orc_sln_dry_chd(const orc_solution *S) { return S->dry_chd; }
const double *orc_sln_strgss(orc_solution *S) { return S->strgss; }
const double *orc_sln_strgsy(orc_solution *S) { return S->strgsy; }
void orc_sln_timers(orc_solution *S, double *t2) {
  t2[0] = S->t_form;
  t2[1] = S->t_ls;
}
