/* ilut.c -- CPU ORACLE (test infrastructure only): ILUT / MILUT, the third-party SPARSKIT2 routines the
 * reference vendors as src/Utilities/Libraries/sparskit2/ilut.f90 (with the MODFLOW modifications: relaxation of
 * the dropped terms, diagonal scaling delta, sign-preserving pivot rescue):
 *   ilut   :48-428     lusol :431-481     qsplit :484-548
 * and their use by IMS: ims_base_pcu (ImsLinearBase.f90:761-864, IPC 3/4), ims_calc_pcdims (:1148-1197),
 * lusol calls in ims_base_cg / ims_base_bcgs (:113, 366, 417).
 * Arrays keep the Fortran 1-based indexing internally (element 0 unused) so that every index expression reads like
 * the reference's. */
#include "mf6_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* quick-sort split: on return |a(i)| >= |a(ncut)| for i < ncut, <= for i > ncut (a, ind 1-based) */
static void qsplit(int n, double *a, int *ind, int ncut) {
  int first = 1, last = n;
  if (ncut < first || ncut > last) return;
  for (;;) {
    int mid = first;
    double abskey = fabs(a[mid]);
    for (int j = first + 1; j <= last; j++) {
      if (fabs(a[j]) > abskey) {
        mid = mid + 1;
        double tmp = a[mid];
        int itmp = ind[mid];
        a[mid] = a[j];
        ind[mid] = ind[j];
        a[j] = tmp;
        ind[j] = itmp;
      }
    }
    double tmp = a[mid];
    a[mid] = a[first];
    a[first] = tmp;
    int itmp = ind[mid];
    ind[mid] = ind[first];
    ind[first] = itmp;
    if (mid == ncut) return;
    if (mid > ncut)
      last = mid - 1;
    else
      first = mid + 1;
  }
}

orc_ilut *orc_ilut_create(int n, int nja, const int *ia, int lfil) {
  orc_ilut *p = (orc_ilut *)calloc(1, sizeof(*p));
  p->n = n;
  long long iwk;
  if (lfil > 0) {
    iwk = (long long)n * (lfil * 2 + 1);
  } else {
    int mx = 0;
    for (int i = 0; i < n; i++)
      if (ia[i + 1] - ia[i] > mx) mx = ia[i + 1] - ia[i];
    iwk = (long long)n * mx;
  }
  if (iwk < n + 2) iwk = n + 2;
  p->iwk = (int)iwk;
  p->alu = (double *)calloc((size_t)iwk + 2, sizeof(double));
  p->jlu = (int *)calloc((size_t)iwk + 2, sizeof(int));
  p->ju = (int *)calloc((size_t)n + 2, sizeof(int));
  p->w = (double *)calloc((size_t)n + 3, sizeof(double));
  p->jw = (int *)calloc(2 * (size_t)n + 2, sizeof(int));
  (void)nja;
  return p;
}

void orc_ilut_destroy(orc_ilut *p) {
  if (!p) return;
  free(p->alu); free(p->jlu); free(p->ju); free(p->w); free(p->jw);
  free(p);
}

/* ilut.f90:48-428.  a/ja/ia: 0-based CSR of the caller; the row is read in storage order like the reference.
 * Returns ierr; *izero is the reference's in/out flag. */
int orc_ilut_factor(orc_ilut *P, const double *a, const int *ja, const int *ia, int lfil, double droptol,
                    double relax, int *izero, double delta) {
  const int n = P->n, iwk = P->iwk;
  double *alu = P->alu, *w = P->w;
  int *jlu = P->jlu, *ju = P->ju, *jw = P->jw;
  if (lfil < 0) return -4;
  int ju0 = n + 2;
  jlu[1] = ju0;
  for (int j = 1; j <= n; j++) jw[n + j] = 0;
  for (int ii = 1; ii <= n; ii++) {
    const int j1 = ia[ii - 1] + 1, j2 = ia[ii]; /* 1-based positions j1..j2 */
    double dropsum = 0.0, tnorm = 0.0;
    for (int k = j1; k <= j2; k++) tnorm = tnorm + fabs(a[k - 1]);
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
        lenl = lenl + 1;
        jw[lenl] = k;
        w[lenl] = t;
        jw[n + k] = lenl;
      } else if (k == ii) {
        w[ii] = t;
      } else {
        lenu = lenu + 1;
        const int jpos = ii + lenu - 1;
        jw[jpos] = k;
        w[jpos] = t;
        jw[n + k] = jpos;
      }
    }
    int jj = 0, ilen = 0;
    for (;;) {
      jj = jj + 1;
      if (jj > lenl) break;
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
        const double s = w[jj];
        w[jj] = w[k];
        w[k] = s;
      }
      jw[n + jrow] = 0;
      const double fact = w[jj] * alu[jrow];
      if (fabs(fact) <= droptol) {
        dropsum = dropsum + w[jj];
        continue;
      }
      for (int kk = ju[jrow]; kk <= jlu[jrow + 1] - 1; kk++) {
        const double s = fact * alu[kk];
        const int j = jlu[kk];
        const int jpos = jw[n + j];
        if (j >= ii) {
          if (jpos == 0) {
            lenu = lenu + 1;
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
            lenl = lenl + 1;
            if (lenl > n) return -1;
            jw[lenl] = j;
            jw[n + j] = lenl;
            w[lenl] = -s;
          } else {
            w[jpos] = w[jpos] - s;
          }
        }
      }
      ilen = ilen + 1;
      w[ilen] = fact;
      jw[ilen] = jrow;
    }
    for (int k = 1; k <= lenu; k++) jw[n + jw[ii + k - 1]] = 0;
    lenl = ilen;
    ilen = lenl < lfil ? lenl : lfil;
    qsplit(lenl, w, jw, ilen);
    for (int k = 1; k <= ilen; k++) {
      if (ju0 > iwk) return -2;
      alu[ju0] = w[k];
      jlu[ju0] = jw[k];
      ju0 = ju0 + 1;
    }
    ju[ii] = ju0;
    ilen = 0;
    for (int k = 1; k <= lenu - 1; k++) {
      if (fabs(w[ii + k]) > droptol * tnorm) {
        ilen = ilen + 1;
        w[ii + ilen] = w[ii + k];
        jw[ii + ilen] = jw[ii + k];
      } else {
        dropsum = dropsum + w[ii + k];
      }
    }
    lenu = ilen + 1;
    ilen = lenu < lfil ? lenu : lfil;
    qsplit(lenu - 1, w + ii, jw + ii, ilen); /* w(ii+1) is element 1 of the split */
    if (ilen + ju0 > iwk) return -3;
    for (int k = ii + 1; k <= ii + ilen - 1; k++) {
      jlu[ju0] = jw[k];
      alu[ju0] = w[k];
      ju0 = ju0 + 1;
    }
    const double diag = w[ii];
    double diag_working = (1.0 + delta) * diag + (relax * dropsum);
    const double sign_check = copysign(fabs(diag), diag_working);
    if (sign_check != diag) {
      if (*izero > 1) {
        diag_working = copysign(1.0, diag) * (1.0e-4 + droptol) * tnorm;
      } else {
        *izero = 1;
        return 0; /* exit main */
      }
    }
    if (fabs(diag_working) == 0.0) {
      if (*izero > 1) {
        diag_working = copysign(1.0, diag) * (1.0e-4 + droptol) * tnorm;
      } else {
        *izero = 1;
        return 0;
      }
    }
    w[ii] = diag_working;
    alu[ii] = 1.0 / w[ii];
    jlu[ii + 1] = ju0;
  }
  return 0;
}

/* ims_base_pcu with IPC 3/4 (ImsLinearBase.f90:808-858): the delta loop around ilut.  Returns icount. */
int orc_pcu_ilut(orc_ilut *P, const double *amat, const int *ia, const int *ja, int lfil, double droptol,
                 double relax, int *ierr_out) {
  int ipcflag = 0, icount = 0, ierr = 0;
  double delta = 0.0;
  for (;;) {
    ierr = orc_ilut_factor(P, amat, ja, ia, lfil, droptol, relax, &ipcflag, delta);
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
  if (ierr_out) *ierr_out = ierr;
  return icount;
}

/* lusol, ilut.f90:431-481 (y, x 0-based, distinct or identical) */
void orc_lusol(const orc_ilut *P, const double *y, double *x) {
  const int n = P->n;
  const double *alu = P->alu;
  const int *jlu = P->jlu, *ju = P->ju;
  for (int i = 1; i <= n; i++) {
    x[i - 1] = y[i - 1];
    for (int k = jlu[i]; k <= ju[i] - 1; k++) x[i - 1] = x[i - 1] - alu[k] * x[jlu[k] - 1];
  }
  for (int i = n; i >= 1; i--) {
    for (int k = ju[i]; k <= jlu[i + 1] - 1; k++) x[i - 1] = x[i - 1] - alu[k] * x[jlu[k] - 1];
    x[i - 1] = alu[i] * x[i - 1];
  }
}
