/*
 * ims.c -- oracle restatement of the IMS linear accelerators.
 * TEST INFRASTRUCTURE ONLY (see mf6_oracle.h).
 *
 * Follows, routine by routine and loop by loop:
 *   src/Solution/LinearMethods/ImsLinearBase.f90
 *   src/Solution/LinearMethods/ImsLinear.f90:617-750
 *   src/Utilities/Libraries/sparsekit/sparsekit.f90:1-59   (amux)
 *   src/Utilities/Libraries/blas/blas1_d.f90:295-333, 387-480 (ddot, dnrm2)
 *   src/Utilities/MathUtil.f90:45-86 (is_close)
 * Indices are 0-based here.
 */
#include "mf6_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define DPREC DBL_EPSILON           /* Constants.f90:117 */
#define DSAME (100.0 * DBL_EPSILON) /* Constants.f90:119 */

static double dsign(double a, double b) { return copysign(fabs(a), b); }

/* sparsekit.f90:44-57 */
void orc_amux(int n, const double *x, double *y, const double *a, const int *ja,
              const int *ia) {
  for (int i = 0; i < n; i++) {
    double t = 0.0;
    for (int k = ia[i]; k < ia[i + 1]; k++) t = t + a[k] * x[ja[k]];
    y[i] = t;
  }
}

/* blas1_d.f90:370 -- dot_product intrinsic, restated as the plain sequential sum */
double orc_ddot(int n, const double *x, const double *y) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}

/* blas1_d.f90:449-475 */
double orc_dnrm2(int n, const double *x) {
  if (n < 1) return 0.0;
  if (n == 1) return fabs(x[0]);
  double scale = 0.0, ssq = 1.0;
  for (int i = 0; i < n; i++) {
    if (x[i] != 0.0) {
      double absxi = fabs(x[i]);
      if (scale < absxi) {
        double r = scale / absxi;
        ssq = 1.0 + ssq * (r * r);
        scale = absxi;
      } else {
        double r = absxi / scale;
        ssq = ssq + r * r;
      }
    }
  }
  return scale * sqrt(ssq);
}

/* MathUtil.f90:45-86 with defaults */
int orc_is_close(double a, double b) {
  if (a == b) return 1;
  double m = fmax(fabs(a), fabs(b));
  return fabs(a - b) <= fmax(DSAME * m, 0.0);
}

/* ImsLinearBase.f90:1207-1261 */
orc_ilu0 *orc_ilu0_create(int n, int nja, const int *ia, const int *ja) {
  orc_ilu0 *p = (orc_ilu0 *)calloc(1, sizeof(*p));
  p->n = n;
  p->nja = nja;
  p->iapc = (int *)malloc(sizeof(int) * (size_t)(n + 1));
  p->japc = (int *)malloc(sizeof(int) * (size_t)nja);
  p->apc = (double *)calloc((size_t)nja, sizeof(double));
  p->iw = (int *)calloc((size_t)n, sizeof(int));
  p->w = (double *)calloc((size_t)n, sizeof(double));
  int ip = n;
  for (int r = 0; r < n; r++) {
    int i0 = ia[r], i1 = ia[r + 1];
    p->iapc[r] = ip;
    int start = ip;
    for (int j = i0; j < i1; j++) {
      if (ja[j] == r) continue;
      p->japc[ip++] = ja[j];
    }
    /* ims_base_isort :1268-1284 (result: ascending) */
    for (int a = start; a < ip - 1; a++)
      for (int b = a + 1; b < ip; b++)
        if (p->japc[a] > p->japc[b]) {
          int t = p->japc[b];
          p->japc[b] = p->japc[a];
          p->japc[a] = t;
        }
  }
  p->iapc[n] = nja;
  for (int r = 0; r < n; r++) {
    int i0 = p->iapc[r], i1 = p->iapc[r + 1];
    p->japc[r] = i1;
    for (int j = i0; j < i1; j++)
      if (p->japc[j] > r) {
        p->japc[r] = j;
        break;
      }
  }
  return p;
}

void orc_ilu0_destroy(orc_ilu0 *p) {
  if (!p) return;
  free(p->iapc);
  free(p->japc);
  free(p->apc);
  free(p->iw);
  free(p->w);
  free(p);
}

/* ImsLinearBase.f90:928-1042 */
int orc_pcilu0(orc_ilu0 *p, const double *amat, const int *ia, const int *ja,
               double relax, int ipcflag, double delta) {
  const int n = p->n;
  int *iw = p->iw, *iapc = p->iapc, *japc = p->japc;
  double *w = p->w, *apc = p->apc;
  for (int r = 0; r < n; r++) {
    iw[r] = 0;
    w[r] = 0.0;
  }
  for (int r = 0; r < n; r++) {
    for (int j = ia[r]; j < ia[r + 1]; j++) {
      int jcol = ja[j];
      iw[jcol] = 1;
      w[jcol] = w[jcol] + amat[j];
    }
    int ic0 = iapc[r], ic1 = iapc[r + 1], iu = japc[r];
    double rs = 0.0;
    for (int j = ic0; j < iu; j++) {
      int jcol = japc[j];
      int iic1 = iapc[jcol + 1], iiu = japc[jcol];
      double tl = w[jcol] * apc[jcol];
      w[jcol] = tl;
      for (int jj = iiu; jj < iic1; jj++) {
        int jjcol = japc[jj];
        if (iw[jjcol] != 0)
          w[jjcol] = w[jjcol] - tl * apc[jj];
        else
          rs = rs + tl * apc[jj];
      }
    }
    double d = w[r];
    double tl = (1.0 + delta) * d - (relax * rs);
    double sd1 = dsign(d, tl);
    if (sd1 != d) {
      if (ipcflag > 1)
        tl = dsign(1.0e-6, d);
      else
        return 1; /* IPCFLAG = 1 ; EXIT MAIN */
    }
    if (fabs(tl) == 0.0) {
      if (ipcflag > 1)
        tl = dsign(1.0e-6, d);
      else
        return 1;
    }
    apc[r] = 1.0 / tl;
    iw[r] = 0;
    w[r] = 0.0;
    for (int j = ic0; j < ic1; j++) {
      int jcol = japc[j];
      apc[j] = w[jcol];
      iw[jcol] = 0;
      w[jcol] = 0.0;
    }
  }
  return 0;
}

/* ImsLinearBase.f90:808-858 */
int orc_pcu(orc_ilu0 *p, const double *amat, const int *ia, const int *ja,
            double relax) {
  int ipcflag = 0, icount = 0;
  double delta = 0.0;
  for (;;) {
    ipcflag = orc_pcilu0(p, amat, ia, ja, relax, ipcflag, delta);
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

/* ImsLinearBase.f90:1049-1092 */
void orc_ilu0a(const orc_ilu0 *p, const double *r, double *d) {
  const int n = p->n;
  const int *iapc = p->iapc, *japc = p->japc;
  const double *apc = p->apc;
  for (int i = 0; i < n; i++) {
    double tv = r[i];
    int ic0 = iapc[i], iu = japc[i];
    for (int j = ic0; j < iu; j++) tv = tv - apc[j] * d[japc[j]];
    d[i] = tv;
  }
  for (int i = n - 1; i >= 0; i--) {
    int ic1 = iapc[i + 1], iu = japc[i];
    double tv = d[i];
    for (int j = iu; j < ic1; j++) tv = tv - apc[j] * d[japc[j]];
    d[i] = tv * apc[i];
  }
}

/* ImsLinearBase.f90:1101-1146 */
void orc_testcnvg(int icnvgopt, int *icnvg, int iiter, double dvmax, double rmax,
                  double rmax0, double epfact, double dvclose, double rclose) {
  if (icnvgopt == 0) {
    if (fabs(dvmax) <= dvclose && fabs(rmax) <= rclose) *icnvg = 1;
  } else if (icnvgopt == 1) {
    if (fabs(dvmax) <= dvclose && fabs(rmax) <= rclose) {
      if (iiter == 1)
        *icnvg = 1;
      else
        *icnvg = -1;
    }
  } else if (icnvgopt == 2) {
    if (fabs(dvmax) <= dvclose || rmax <= rclose)
      *icnvg = 1;
    else if (rmax <= rmax0 * epfact)
      *icnvg = -1;
  } else if (icnvgopt == 3) {
    if (fabs(dvmax) <= dvclose)
      *icnvg = 1;
    else if (rmax <= rmax0 * rclose)
      *icnvg = -1;
  } else if (icnvgopt == 4) {
    if (fabs(dvmax) <= dvclose && rmax <= rclose)
      *icnvg = 1;
    else if (rmax <= rmax0 * epfact)
      *icnvg = -1;
  }
}

/* ImsLinearBase.f90:1316-1333.  NB: 0.01 and 0.10 are default-REAL (single
 * precision) literals in the Fortran source, so the DP value is (double)0.01f. */
double orc_epfact(int icnvgopt, int kstp) {
  if (icnvgopt == 2) return kstp == 1 ? (double)0.01f : (double)0.10f;
  if (icnvgopt == 4) return 1.0e-4;
  return 1.0;
}

/* ImsLinearBase.f90:619-754 */
void orc_scale(int iopt, int iscl, int n, const int *ia, const int *ja,
               double *amat, double *x, double *b, double *dscale,
               double *dscale2) {
  if (iopt == 0) {
    if (iscl == 1) {
      for (int r = 0; r < n; r++) {
        double v = amat[ia[r]];
        double c1 = 1.0 / sqrt(fabs(v));
        dscale[r] = c1;
        dscale2[r] = c1;
      }
      for (int r = 0; r < n; r++) {
        double c1 = dscale[r];
        for (int i = ia[r]; i < ia[r + 1]; i++) {
          double c2 = dscale2[ja[i]];
          amat[i] = c1 * amat[i] * c2;
        }
      }
    } else if (iscl == 2) {
      for (int r = 0; r < n; r++) {
        double c1 = 0.0;
        for (int i = ia[r]; i < ia[r + 1]; i++) c1 = c1 + amat[i] * amat[i];
        c1 = sqrt(c1);
        if (c1 == 0.0)
          c1 = 1.0;
        else
          c1 = 1.0 / c1;
        dscale[r] = c1;
        for (int i = ia[r]; i < ia[r + 1]; i++) amat[i] = c1 * amat[i];
      }
      for (int r = 0; r < n; r++) dscale2[r] = 0.0;
      for (int r = 0; r < n; r++)
        for (int i = ia[r]; i < ia[r + 1]; i++) {
          double c2 = amat[i];
          dscale2[ja[i]] = dscale2[ja[i]] + c2 * c2;
        }
      for (int r = 0; r < n; r++) {
        double c2 = dscale2[r];
        if (c2 == 0.0)
          c2 = 1.0;
        else
          c2 = 1.0 / sqrt(c2);
        dscale2[r] = c2;
      }
      for (int r = 0; r < n; r++)
        for (int i = ia[r]; i < ia[r + 1]; i++) amat[i] = dscale2[ja[i]] * amat[i];
    }
    for (int r = 0; r < n; r++) {
      x[r] = x[r] / dscale2[r];
      b[r] = b[r] * dscale[r];
    }
  } else {
    for (int r = 0; r < n; r++) {
      double c1 = dscale[r];
      for (int i = ia[r]; i < ia[r + 1]; i++) {
        double c2 = dscale2[ja[i]];
        amat[i] = (1.0 / c1) * amat[i] * (1.0 / c2);
      }
      double c2 = dscale2[r];
      x[r] = x[r] * c2;
      b[r] = b[r] / c1;
    }
  }
}

/* ImsLinearBase.f90:1291-1312 */
static void residual(int n, const double *x, const double *b, double *d,
                     const double *a, const int *ia, const int *ja) {
  orc_amux(n, x, d, a, ja, ia);
  for (int i = 0; i < n; i++) d[i] = b[i] - d[i];
}

static void sum_record(orc_summary *s, int iiter, double dv, int locdv, double r,
                       int locr, double alpha, double omega) {
  if (!s) return;
  int k = s->count; /* already incremented: 1-based count */
  if (s->cap > 0 && k <= s->cap) {
    s->itinner[k - 1] = iiter;
    s->dvmax[k - 1] = dv;
    s->locdv[k - 1] = locdv;
    s->rmax[k - 1] = r;
    s->locr[k - 1] = locr;
    s->alpha[k - 1] = alpha;
    s->omega[k - 1] = omega;
  }
}

/* per-model maxima of one inner iteration (ImsLinearBase.f90:143-176): the same strict ">" rule per model */
typedef struct {
  int nmod;
  const int *modid, *lorder;
  double dv[64], r[64];
  int locdv[64], locr[64];
} modtrack;

static void mt_begin(modtrack *t, const orc_summary *sum, const orc_imslinear *L) {
  t->nmod = (sum && sum->nmod > 0 && sum->nmod <= 64 && sum->modid) ? sum->nmod : 0;
  t->modid = sum ? sum->modid : NULL;
  t->lorder = L->use_perm ? L->lorder : NULL;
  for (int im = 0; im < t->nmod; im++) {
    t->dv[im] = t->r[im] = 0.0;
    t->locdv[im] = t->locr[im] = -1;
  }
}

static void mt_row(modtrack *t, int i, double tv, double rv) {
  if (!t->nmod) return;
  const int o = t->lorder ? t->lorder[i] : i, im = t->modid[o];
  if (fabs(tv) > fabs(t->dv[im])) {
    t->dv[im] = tv;
    t->locdv[im] = o;
  }
  if (fabs(rv) > fabs(t->r[im])) {
    t->r[im] = rv;
    t->locr[im] = o;
  }
}

static void mt_record(const modtrack *t, orc_summary *s) {
  if (!s || !t->nmod) return;
  const int k = s->count;
  if (s->cap > 0 && k <= s->cap)
    for (int im = 0; im < t->nmod; im++) {
      const size_t q = (size_t)(k - 1) * (size_t)t->nmod + (size_t)im;
      s->mdvmax[q] = t->dv[im];
      s->mlocdv[q] = t->locdv[im];
      s->mrmax[q] = t->r[im];
      s->mlocr[q] = t->locr[im];
    }
}

/* APPLY PRECONDITIONER (ImsLinearBase.f90:104-114): ILU0 / MILU0 or ILUT / MILUT */
static void precond(const orc_imslinear *L, const double *r, double *z) {
  if (L->pct)
    orc_lusol(L->pct, r, z);
  else
    orc_ilu0a(L->pc, r, z);
}

/* ImsLinearBase.f90:30-240 */
static int ims_cg(orc_imslinear *L, int *icnvg, int itmax, const int *ia,
                  const int *ja, const double *a, double *x, double *b,
                  orc_summary *sum) {
  const int n = L->n;
  double *d = L->d, *p = L->p, *q = L->q, *z = L->z;
  const mf6gpu_ims_settings *s = &L->s;
  double rho0 = 0.0, rho = 0.0, alpha, beta;
  int innerit = 0;
  for (int iiter = 1; iiter <= itmax; iiter++) {
    innerit++;
    if (sum) sum->count++;
    precond(L, d, z);
    rho = orc_ddot(n, d, z);
    if (iiter == 1) {
      for (int i = 0; i < n; i++) p[i] = z[i];
    } else {
      beta = rho / rho0;
      for (int i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
    }
    orc_amux(n, p, q, a, ja, ia);
    double denominator = orc_ddot(n, p, q);
    denominator = denominator + dsign(DPREC, denominator);
    alpha = rho / denominator;
    double deltax = 0.0, rmax = 0.0, l2norm = 0.0;
    int xloc = -1, rloc = -1;
    modtrack mt;
    mt_begin(&mt, sum, L);
    for (int i = 0; i < n; i++) {
      double tv = alpha * p[i];
      const double dvi = tv;
      x[i] = x[i] + tv;
      if (fabs(tv) > fabs(deltax)) {
        deltax = tv;
        xloc = i;
      }
      tv = d[i];
      tv = tv - alpha * q[i];
      d[i] = tv;
      if (fabs(tv) > fabs(rmax)) {
        rmax = tv;
        rloc = i;
      }
      mt_row(&mt, i, dvi, tv);
      l2norm = l2norm + tv * tv;
    }
    l2norm = sqrt(l2norm);
    sum_record(sum, iiter, deltax, xloc, rmax, rloc, alpha, 0.0);
    mt_record(&mt, sum);
    double rcnvg = (s->icnvgopt == 2 || s->icnvgopt == 3 || s->icnvgopt == 4) ? l2norm : rmax;
    orc_testcnvg(s->icnvgopt, icnvg, innerit, deltax, rcnvg, L->l2norm0,
                 L->epfact, s->dvclose, s->rclose);
    if (rcnvg == 0.0) *icnvg = 1;
    if (*icnvg != 0) break;
    if (orc_is_close(rho, rho0)) break;
    if (s->north > 0) {
      if ((iiter + 1) % s->north == 0) residual(n, x, b, d, a, ia, ja);
    }
    if (rho == 0.0) break;
    rho0 = rho;
  }
  if (*icnvg < 0) *icnvg = 0;
  return innerit;
}

/* ImsLinearBase.f90:249-549 */
static int ims_bcgs(orc_imslinear *L, int *icnvg, int itmax, const int *ia,
                    const int *ja, const double *a, double *x, double *b,
                    orc_summary *sum) {
  const int n = L->n;
  double *d = L->d, *p = L->p, *q = L->q, *t = L->t, *v = L->v;
  double *dhat = L->dhat, *phat = L->phat, *qhat = L->qhat;
  const mf6gpu_ims_settings *s = &L->s;
  const int iscl = s->iscl;
  const double *dscale = L->dscale;
  int innerit = 0;
  double alpha = 0.0, alpha0 = 0.0, beta = 0.0, rho = 0.0, rho0 = 0.0;
  double omega = 0.0, omega0 = 0.0;
  for (int i = 0; i < n; i++) dhat[i] = d[i];
  for (int iiter = 1; iiter <= itmax; iiter++) {
    innerit++;
    if (sum) sum->count++;
    rho = orc_ddot(n, dhat, d);
    if (iiter == 1) {
      for (int i = 0; i < n; i++) p[i] = d[i];
    } else {
      beta = (rho / rho0) * (alpha0 / omega0);
      for (int i = 0; i < n; i++) p[i] = d[i] + beta * (p[i] - omega0 * v[i]);
    }
    precond(L, p, phat);
    orc_amux(n, phat, v, a, ja, ia);
    double denominator = orc_ddot(n, dhat, v);
    denominator = denominator + dsign(DPREC, denominator);
    alpha = rho / denominator;
    for (int i = 0; i < n; i++) q[i] = d[i] - alpha * v[i];
    precond(L, q, qhat);
    orc_amux(n, qhat, t, a, ja, ia);
    double numerator = orc_ddot(n, t, q);
    denominator = orc_ddot(n, t, t);
    denominator = denominator + dsign(DPREC, denominator);
    omega = numerator / denominator;
    double deltax = 0.0, rmax = 0.0, l2norm = 0.0;
    int xloc = -1, rloc = -1;
    modtrack mt;
    mt_begin(&mt, sum, L);
    for (int i = 0; i < n; i++) {
      double tv = alpha * phat[i] + omega * qhat[i];
      x[i] = x[i] + tv;
      if (iscl != 0) tv = tv * dscale[i];
      const double dvi = tv;
      if (fabs(tv) > fabs(deltax)) {
        deltax = tv;
        xloc = i;
      }
      tv = q[i] - omega * t[i];
      d[i] = tv;
      if (iscl != 0) tv = tv / dscale[i];
      if (fabs(tv) > fabs(rmax)) {
        rmax = tv;
        rloc = i;
      }
      mt_row(&mt, i, dvi, tv);
      l2norm = l2norm + tv * tv;
    }
    l2norm = sqrt(l2norm);
    sum_record(sum, iiter, deltax, xloc, rmax, rloc, alpha, omega);
    mt_record(&mt, sum);
    double rcnvg = (s->icnvgopt == 2 || s->icnvgopt == 3 || s->icnvgopt == 4) ? l2norm : rmax;
    orc_testcnvg(s->icnvgopt, icnvg, innerit, deltax, rcnvg, L->l2norm0,
                 L->epfact, s->dvclose, s->rclose);
    if (rcnvg == 0.0) *icnvg = 1;
    if (*icnvg != 0) break;
    if (orc_is_close(rho, rho0)) break;
    if (orc_is_close(alpha, alpha0)) break;
    if (orc_is_close(omega, omega0)) break;
    if (s->north > 0) {
      if ((iiter + 1) % s->north == 0) residual(n, x, b, d, a, ia, ja);
    }
    if (rho * omega == 0.0) break;
    rho0 = rho;
    alpha0 = alpha;
    omega0 = omega;
  }
  if (*icnvg < 0) *icnvg = 0;
  return innerit;
}

/* ImsLinear.f90:111-339 */
orc_imslinear *orc_ims_create(int n, int nja, const int *ia, const int *ja,
                              const mf6gpu_ims_settings *s, const int *perm) {
  orc_imslinear *L = (orc_imslinear *)calloc(1, sizeof(*L));
  L->s = *s;
  if (L->s.iscl < 0) L->s.iscl = 0;
  L->n = n;
  L->nja = nja;
  L->ia = ia;
  L->ja = ja;
  size_t nb = sizeof(double) * (size_t)n;
  L->d = (double *)calloc(1, nb);
  L->p = (double *)calloc(1, nb);
  L->q = (double *)calloc(1, nb);
  L->z = (double *)calloc(1, nb);
  L->t = (double *)calloc(1, nb);
  L->v = (double *)calloc(1, nb);
  L->dhat = (double *)calloc(1, nb);
  L->phat = (double *)calloc(1, nb);
  L->qhat = (double *)calloc(1, nb);
  L->dscale = (double *)malloc(nb);
  L->dscale2 = (double *)malloc(nb);
  for (int i = 0; i < n; i++) L->dscale[i] = L->dscale2[i] = 1.0;
  if (perm) {
    /* symmetric permutation B = P A P^T with rows re-sorted "diagonal first,
     * then ascending" (what dperm + the solution's sort would give);
     * the ILU0 structure is then built on the PERMUTED pattern. */
    L->use_perm = 1;
    L->lorder = (int *)malloc(sizeof(int) * (size_t)n);
    L->iorder = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) {
      L->lorder[i] = perm[i];
      L->iorder[perm[i]] = i;
    }
    L->iaro = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    L->jaro = (int *)malloc(sizeof(int) * (size_t)nja);
    L->aro = (double *)malloc(sizeof(double) * (size_t)nja);
    L->xp = (double *)malloc(nb);
    L->bp = (double *)malloc(nb);
    int pos = 0;
    for (int r = 0; r < n; r++) {
      int o = perm[r];
      L->iaro[r] = pos;
      L->jaro[pos++] = r; /* diagonal first */
      int start = pos;
      for (int j = ia[o] + 1; j < ia[o + 1]; j++) L->jaro[pos++] = L->iorder[ja[j]];
      for (int a = start; a < pos - 1; a++)
        for (int b = a + 1; b < pos; b++)
          if (L->jaro[a] > L->jaro[b]) {
            int t = L->jaro[b];
            L->jaro[b] = L->jaro[a];
            L->jaro[a] = t;
          }
    }
    L->iaro[n] = pos;
    L->pc = orc_ilu0_create(n, nja, L->iaro, L->jaro);
  } else {
    L->pc = orc_ilu0_create(n, nja, ia, ja);
  }
  /* ImsLinear.f90:178-185: LEVEL > 0 or DROPTOL > 0 selects ILUT (MILUT with RELAX > 0) */
  if (L->s.level > 0 || L->s.droptol > 0.0)
    L->pct = orc_ilut_create(n, nja, L->use_perm ? L->iaro : ia, L->s.level);
  return L;
}

void orc_ims_set_blocks(orc_imslinear *L, const int *block) {
  const int n = L->n;
  const int *ia0 = L->use_perm ? L->iaro : L->ia;
  const int *ja0 = L->use_perm ? L->jaro : L->ja;
  L->use_blocks = 1;
  L->iaf = (int *)malloc(sizeof(int) * (size_t)(n + 1));
  L->jaf = (int *)malloc(sizeof(int) * (size_t)L->nja);
  L->fmap = (int *)malloc(sizeof(int) * (size_t)L->nja);
  L->af = (double *)malloc(sizeof(double) * (size_t)L->nja);
  int pos = 0;
  for (int r = 0; r < n; r++) {
    L->iaf[r] = pos;
    int br = block[L->use_perm ? L->lorder[r] : r];
    for (int k = ia0[r]; k < ia0[r + 1]; k++) {
      int c = ja0[k];
      int bc = block[L->use_perm ? L->lorder[c] : c];
      if (bc != br) continue;
      L->jaf[pos] = c;
      L->fmap[pos] = k;
      pos++;
    }
  }
  L->iaf[n] = pos;
  orc_ilu0_destroy(L->pc);
  L->pc = orc_ilu0_create(n, pos, L->iaf, L->jaf);
  if (L->pct) {
    orc_ilut_destroy(L->pct);
    L->pct = orc_ilut_create(n, pos, L->iaf, L->s.level);
  }
}

void orc_ims_destroy(orc_imslinear *L) {
  if (!L) return;
  orc_ilut_destroy(L->pct);
  free(L->iaf); free(L->jaf); free(L->fmap); free(L->af);
  free(L->d); free(L->p); free(L->q); free(L->z); free(L->t); free(L->v);
  free(L->dhat); free(L->phat); free(L->qhat); free(L->dscale); free(L->dscale2);
  free(L->lorder); free(L->iorder); free(L->iaro); free(L->jaro); free(L->aro);
  free(L->xp); free(L->bp);
  orc_ilu0_destroy(L->pc);
  free(L);
}

/* fill the permuted value array from the original one */
static void permute_values(orc_imslinear *L, const double *amat) {
  const int n = L->n;
  const int *ia = L->ia, *ja = L->ja;
  for (int r = 0; r < n; r++) {
    int o = L->lorder[r];
    for (int k = L->iaro[r]; k < L->iaro[r + 1]; k++) {
      int ocol = L->lorder[L->jaro[k]];
      double v = 0.0;
      for (int j = ia[o]; j < ia[o + 1]; j++)
        if (ja[j] == ocol) {
          v = amat[j];
          break;
        }
      L->aro[k] = v;
    }
  }
}

/* ImsLinear.f90:617-750 */
int orc_ims_apply(orc_imslinear *L, double *amat, double *x, double *rhs,
                  int *icnvg, int kstp, int kiter, orc_summary *sum) {
  const int n = L->n;
  const mf6gpu_ims_settings *s = &L->s;
  L->epfact = orc_epfact(s->icnvgopt, kstp);
  if (s->iscl != 0)
    orc_scale(0, s->iscl, n, L->ia, L->ja, amat, x, rhs, L->dscale, L->dscale2);
  const int *ia0 = L->ia, *ja0 = L->ja;
  const double *a0 = amat;
  double *x0 = x, *b0 = rhs;
  if (L->use_perm) {
    permute_values(L, amat);
    for (int i = 0; i < n; i++) {
      L->xp[i] = x[L->lorder[i]];
      L->bp[i] = rhs[L->lorder[i]];
    }
    ia0 = L->iaro;
    ja0 = L->jaro;
    a0 = L->aro;
    x0 = L->xp;
    b0 = L->bp;
  }
  if (L->use_blocks) {
    const int nf = L->iaf[n];
    for (int k = 0; k < nf; k++) L->af[k] = a0[L->fmap[k]];
    if (L->pct)
      L->npivfix = orc_pcu_ilut(L->pct, L->af, L->iaf, L->jaf, s->level, s->droptol, s->relax, &L->ilut_ierr);
    else
      L->npivfix = orc_pcu(L->pc, L->af, L->iaf, L->jaf, s->relax);
  } else if (L->pct) {
    L->npivfix = orc_pcu_ilut(L->pct, a0, ia0, ja0, s->level, s->droptol, s->relax, &L->ilut_ierr);
  } else {
    L->npivfix = orc_pcu(L->pc, a0, ia0, ja0, s->relax);
  }
  if (kiter == 1) {
    L->niterc = 0;
    if (sum) sum->count = 0;
  }
  *icnvg = 0;
  const int c0 = sum ? sum->count : 0;
  for (int i = 0; i < n; i++) L->d[i] = L->p[i] = L->q[i] = L->z[i] = 0.0;
  residual(n, x0, b0, L->d, a0, ia0, ja0);
  L->l2norm0 = orc_dnrm2(n, L->d);
  int itmax = s->iter1;
  if (L->l2norm0 == 0.0) {
    itmax = 0;
    *icnvg = 1;
  }
  int innerit = 0;
  if (s->ilinmeth == 1)
    innerit = ims_cg(L, icnvg, itmax, ia0, ja0, a0, x0, b0, sum);
  else if (s->ilinmeth == 2)
    innerit = ims_bcgs(L, icnvg, itmax, ia0, ja0, a0, x0, b0, sum);
  if (L->use_perm) {
    for (int i = 0; i < n; i++) x[L->lorder[i]] = L->xp[i];
    /* rhs is unchanged by the accelerators; nothing to back-permute */
    if (sum && sum->cap > 0) {
      /* report locations in original numbering */
      for (int k = c0; k < sum->count && k < sum->cap; k++) {
        if (sum->locdv[k] >= 0) sum->locdv[k] = L->lorder[sum->locdv[k]];
        if (sum->locr[k] >= 0) sum->locr[k] = L->lorder[sum->locr[k]];
      }
    }
  }
  if (s->iscl != 0)
    orc_scale(1, s->iscl, n, L->ia, L->ja, amat, x, rhs, L->dscale, L->dscale2);
  return innerit;
}
