/*
 * mf6_oracle.h -- CPU oracle for the MODFLOW 6 IMS + GWF-assembly hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C, single-thread restatement of the
 * reference's Fortran algorithm, used as the checker for the CUDA product path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference).
 * Nothing under modflow6_b200/ may call, link or import it.
 *
 * Parity status: PINNED through the reference's own known-answer tests
 * (autotest/test_gwf_chd01.py:126-127, test_par_gwf01.py:200-212,
 * test_gwf_newton01.py:95-103, ...) -- see tests/test_oracle_known_answers.py.
 * The reference itself (Fortran) cannot be compiled in this image (no Fortran
 * compiler), so there is no oracle/_ref build.
 *
 * Conventions: all indices 0-based here (the Fortran reference is 1-based);
 * CSR rows are "diagonal first, then ascending columns"
 * (src/Utilities/Sparse.f90:217-239).  Compile with -O2 -ffp-contract=off so
 * that no FMA contraction happens (gfortran -O2 on baseline x86-64 emits none).
 */
#ifndef MF6_ORACLE_H
#define MF6_ORACLE_H

#include <stdint.h>
#include "../include/mf6gpu_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- BLAS-1 / SPARSKIT ------------------------------------------------- */
/* sparsekit.f90:1-59 */
void orc_amux(int n, const double *x, double *y, const double *a, const int *ja,
              const int *ia);
/* blas1_d.f90:295-333 (dot_product: sequential accumulation) */
double orc_ddot(int n, const double *x, const double *y);
/* blas1_d.f90:387-480 */
double orc_dnrm2(int n, const double *x);
/* MathUtil.f90:45-86 (default rtol = DSAME, atol = 0, symmetric) */
int orc_is_close(double a, double b);

/* ---- IMS linear (ImsLinearBase.f90) ------------------------------------ */
typedef struct {
  int n, nja;
  int *iapc;   /* [n+1]  (values are indices into japc/apc, first = n) */
  int *japc;   /* [nja]  0..n-1: first-upper pointer, n..: columns      */
  double *apc; /* [nja]  0..n-1: inverse pivots,     n..: L/U entries   */
  int *iw;     /* [n] */
  double *w;   /* [n] */
} orc_ilu0;

orc_ilu0 *orc_ilu0_create(int n, int nja, const int *ia, const int *ja); /* pccrs :1207-1261 */
void orc_ilu0_destroy(orc_ilu0 *p);
/* pcilu0 :928-1042 ; returns ipcflag (0 ok, 1 failed with this delta) */
int orc_pcilu0(orc_ilu0 *p, const double *amat, const int *ia, const int *ja,
               double relax, int ipcflag, double delta);
/* pcu :761-864 ; returns number of pivot corrections (icount) */
int orc_pcu(orc_ilu0 *p, const double *amat, const int *ia, const int *ja,
            double relax);
/* ilu0a :1049-1092 */
void orc_ilu0a(const orc_ilu0 *p, const double *r, double *d);
/* testcnvg :1101-1146 */
void orc_testcnvg(int icnvgopt, int *icnvg, int iiter, double dvmax, double rmax,
                  double rmax0, double epfact, double dvclose, double rclose);
/* epfact :1316-1333 */
double orc_epfact(int icnvgopt, int kstp);
/* scale :619-754 */
void orc_scale(int iopt, int iscl, int n, const int *ia, const int *ja,
               double *amat, double *x, double *b, double *dscale,
               double *dscale2);

/* per-inner-iteration record (ConvergenceSummary.f90:13-34, 1 model) */
/* ILUT / MILUT (SPARSKIT2 ilut.f90 as vendored and modified by the reference), see ilut.c */
typedef struct {
  int n, iwk;
  double *alu; /* [iwk+1] MSR values, 1-based: 1..n inverse pivots, n+2.. L and U rows */
  int *jlu;    /* [iwk+1] 1..n+1 row pointers, then columns (1-based) */
  int *ju;     /* [n+1] start of the U part of every row */
  double *w;
  int *jw;
} orc_ilut;
orc_ilut *orc_ilut_create(int n, int nja, const int *ia, int lfil);
void orc_ilut_destroy(orc_ilut *p);
int orc_ilut_factor(orc_ilut *P, const double *a, const int *ja, const int *ia, int lfil, double droptol,
                    double relax, int *izero, double delta);
int orc_pcu_ilut(orc_ilut *P, const double *amat, const int *ia, const int *ja, int lfil, double droptol,
                 double relax, int *ierr_out);
void orc_lusol(const orc_ilut *P, const double *y, double *x);

typedef struct {
  int cap;     /* capacity of the arrays below (0 = do not record) */
  int count;   /* iter_cnt */
  int *itinner;
  double *dvmax; /* signed value of largest |dx| */
  double *rmax;  /* signed value of largest |r|  */
  int *locdv;    /* 0-based row */
  int *locr;
  double *alpha;
  double *omega;
  /* per-model records (ConvergenceSummaryType convdvmax(im, n) ..., ImsLinearBase.f90:143-197): optional.
   * nmod models; modid[row] = model of every ORIGINAL row; arrays [cap * nmod], model index fastest */
  int nmod;
  const int *modid;
  double *mdvmax;
  double *mrmax;
  int *mlocdv;   /* 0-based original row, -1 = none */
  int *mlocr;
} orc_summary;

typedef struct {
  /* settings (ImsLinearSettings.f90:13-32) */
  mf6gpu_ims_settings s;
  int n, nja;
  const int *ia, *ja; /* borrowed, 0-based, original ordering */
  /* optional symmetric permutation (GPU-path "multicolour" ordering restated
   * as an IORD-style reordering, ImsLinear.f90:652-666): lorder[new] = old */
  int use_perm;
  int *lorder, *iorder;
  int *iaro, *jaro;
  double *aro;
  orc_ilu0 *pc;
  orc_ilut *pct; /* IPC 3/4 (PRECONDITIONER_LEVELS > 0 or DROP_TOLERANCE > 0, ImsLinear.f90:178-185) */
  int ilut_ierr;
  double *d, *p, *q, *z, *t, *v, *dhat, *phat, *qhat;
  double *dscale, *dscale2;
  double *xp, *bp; /* permuted x / rhs */
  double l2norm0, epfact;
  int niterc;
  int npivfix; /* last pcu icount */
  /* optional block-Jacobi preconditioner (the reference's parallel PC: PCBJACOBI + IMS ILU on
   * each rank's diagonal block, PetscSolver.F90:251-278): entries whose row and column lie in
   * different blocks are dropped from the ILU only */
  int use_blocks;
  int *iaf, *jaf, *fmap; /* filtered pattern (in the working ordering) + source position */
  double *af;
} orc_imslinear;

/* imslinear_ar :111-339 ; perm may be NULL (natural order).  perm[new]=old. */
orc_imslinear *orc_ims_create(int n, int nja, const int *ia, const int *ja,
                              const mf6gpu_ims_settings *s, const int *perm);
void orc_ims_destroy(orc_imslinear *L);
/* block[n] (original numbering) switches the preconditioner to block Jacobi; call right after create */
void orc_ims_set_blocks(orc_imslinear *L, const int *block);
/* imslinear_ap :617-750 ; amat/x/rhs in original ordering.  Returns innerit. */
int orc_ims_apply(orc_imslinear *L, double *amat, double *x, double *rhs,
                  int *icnvg, int kstp, int kiter, orc_summary *sum);

/* ---- GWF model + numerical solution (gwf.c, solution.c) ---------------- */
typedef struct orc_solution orc_solution;

void orc_sln_set_blocks(orc_solution *S, const int *block);
orc_solution *orc_sln_create(const mf6gpu_gwf_model *m,
                             const mf6gpu_sln_settings *ss,
                             const mf6gpu_ims_settings *ls, const int *perm);
void orc_sln_destroy(orc_solution *S);
/* set stress data for the coming period(s); arrays are copied */
void orc_sln_set_packages(orc_solution *S, int npkg, const mf6gpu_bnd_package *pk);
/* HFB list of the coming period(s) (hfb_rp, gwf-hfb.f90:149-201); 0-based cells */
void orc_sln_set_gnc(orc_solution *S, int ngnc, int numj, const int *noden, const int *nodem, const int *nodesj,
                     const double *alphasj);
void orc_sln_set_hfb(orc_solution *S, int nhfb, const int *noden, const int *nodem, const double *hydchr);
/* one time step: prepareSolve + outer loop + finalizeSolve
 * (NumericalSolution.f90:1287-1327, 1437-1470, 1844-1938).
 * iss = 1 steady state.  Returns 1 if converged. */
int orc_sln_timestep(orc_solution *S, int kper, int kstp, double delt, int iss,
                     mf6gpu_step_report *rep);
/* access to state */
double *orc_sln_x(orc_solution *S);
double *orc_sln_flowja(orc_solution *S);
const double *orc_sln_simvals(orc_solution *S, int k);
/* the cell every bound of package k acts on (0-based; RCH: the highest active cell, gwf-rch.f90:327-333) */
const int *orc_sln_nodes(orc_solution *S, int k);
/* 1 once a constant-head cell went dry (the reference aborts the simulation there) */
int orc_sln_dry_chd(const orc_solution *S);
const double *orc_sln_strgss(orc_solution *S);
const double *orc_sln_strgsy(orc_solution *S);
const double *orc_sln_amat(orc_solution *S);
const double *orc_sln_rhs(orc_solution *S);
const double *orc_sln_condsat(orc_solution *S);
/* formulate only (sln_buildsystem + sln_ls fix-ups, no linear solve); used by
 * assembly parity tests */
void orc_sln_formulate(orc_solution *S, int kiter, double delt, int iss);
/* timers (seconds, accumulated): [0] formulate, [1] linear solve */
void orc_sln_timers(orc_solution *S, double *t2);

#ifdef __cplusplus
}
#endif
#endif
