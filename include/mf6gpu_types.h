/*
 * mf6gpu_types.h -- plain-data types that cross the C ABI of libmf6gpu.
 *
 * Every struct is POD (pointers + scalars) so that it can be filled from
 * Fortran through ISO_C_BINDING, from C/C++ or from Python ctypes.
 * Arrays are BORROWED for the duration of the call that receives them unless
 * stated otherwise.  Index arrays use the base given by `index_base`
 * (1 = the Fortran arrays passed unchanged, 0 = C).  All reals are f64 and all
 * integers are i32, exactly as in the reference (KindModule DP / I4B).
 */
#ifndef MF6GPU_TYPES_H
#define MF6GPU_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- IMS LINEAR block ----------------------------------------------------
 * mirrors ImsLinearSettingsType (src/Solution/LinearMethods/ImsLinearSettings.f90:13-32) */
typedef struct mf6gpu_ims_settings {
  double dvclose;   /* INNER_DVCLOSE */
  double rclose;    /* INNER_RCLOSE */
  int32_t icnvgopt; /* 0 infinity norm, 1 STRICT, 2 L2NORM_RCLOSE, 3 RELATIVE_RCLOSE, 4 L2NORM_RELATIVE_RCLOSE */
  int32_t iter1;    /* INNER_MAXIMUM */
  int32_t ilinmeth; /* 1 CG, 2 BICGSTAB */
  int32_t iscl;     /* SCALING_METHOD 0 none, 1 diagonal, 2 L2NORM */
  int32_t iord;     /* REORDERING_METHOD 0 none, 1 RCM, 2 MD (1/2 downgraded with a warning, cf. PetscSolver.F90:138-142) */
  int32_t north;    /* NUMBER_ORTHOGONALIZATIONS */
  double relax;     /* RELAXATION_FACTOR (0 => ILU0, >0 => MILU0; ImsLinear.f90:178-185) */
  int32_t level;    /* PRECONDITIONER_LEVELS  (>0 => ILUT: not on the GPU path, rejected) */
  double droptol;   /* PRECONDITIONER_DROP_TOLERANCE */
  /* GPU-path extension (not an .ims keyword; chosen like a PETSc rc option):
   * ordering of the ILU0/MILU0 elimination.
   *   0 = NATURAL    exact reference order, level-scheduled wavefronts
   *   1 = MULTICOLOR greedy colouring of the matrix graph (few, wide levels)
   *   2 = BLOCK_MULTICOLOR colouring of the vertical cell columns (blocks), natural order inside a
   *       column: nlay+1 levels, convergence close to NATURAL (the strong vertical couplings are
   *       eliminated exactly); falls back to 1 when no blocks are known */
  int32_t gpu_ordering;
  int32_t reserved;
} mf6gpu_ims_settings;

#define MF6GPU_ORDER_NATURAL 0
#define MF6GPU_ORDER_MULTICOLOR 1
#define MF6GPU_ORDER_BLOCK_MULTICOLOR 2

/* ---- IMS NONLINEAR block + OPTIONS (NumericalSolution.f90:568-866) ------- */
typedef struct mf6gpu_sln_settings {
  double dvclose;   /* OUTER_DVCLOSE */
  int32_t mxiter;   /* OUTER_MAXIMUM */
  int32_t nonmeth;  /* UNDER_RELAXATION 0 NONE, 1 SIMPLE, 2 COOLEY, 3 DBD */
  double theta;     /* UNDER_RELAXATION_THETA */
  double akappa;    /* UNDER_RELAXATION_KAPPA */
  double gamma;     /* UNDER_RELAXATION_GAMMA */
  double amomentum; /* UNDER_RELAXATION_MOMENTUM */
  int32_t iallowptc; /* 1 default; 0 = NO_PTC ALL; -1 = NO_PTC FIRST */
  int32_t numtrack;  /* BACKTRACKING_NUMBER */
  double btol;       /* BACKTRACKING_TOLERANCE */
  double breduc;     /* BACKTRACKING_REDUCTION_FACTOR */
  double res_lim;    /* BACKTRACKING_RESIDUAL_LIMIT */
} mf6gpu_sln_settings;

/* ---- one GWF model: DIS/DISV connectivity + NPF + STO ------------------
 * the arrays are the ones ConnectionsType / GwfNpfType / GwfStoType hold
 * (Connections.f90:18-55, gwf-npf.f90, gwf-sto.f90) */
typedef struct mf6gpu_gwf_model {
  int32_t index_base; /* base of ia/ja/jas/isym/ibotnode below */
  int32_t nodes;
  int32_t nja;
  int32_t njas;
  const int32_t *ia;   /* [nodes+1] CSR row pointers, diagonal first */
  const int32_t *ja;   /* [nja] */
  const int32_t *jas;  /* [nja]  connection -> symmetric (upper-triangle) index; diag entry ignored */
  const int32_t *isym; /* [nja]  position of the transposed entry */
  const int32_t *ihc;  /* [njas] 0 vertical, 1 horizontal, 2 staggered horizontal */
  const double *cl1;   /* [njas] */
  const double *cl2;   /* [njas] */
  const double *hwva;  /* [njas] */
  const double *top;   /* [nodes] */
  const double *bot;   /* [nodes] */
  const double *area;  /* [nodes] */
  const int32_t *ibound;    /* [nodes] initial ibound (idomain>0 -> 1) */
  const double *strt;       /* [nodes] initial head */
  /* NPF */
  const double *k11;        /* [nodes] */
  const double *k33;        /* [nodes] */
  const int32_t *icelltype; /* [nodes] */
  int32_t icellavg;   /* 0 harmonic, 1 logarithmic, 2 AMT-LMK, 3 AMT-HMK */
  int32_t inewton;    /* NEWTON */
  int32_t inewtonur;  /* NEWTON UNDER_RELAXATION */
  int32_t iperched;   /* PERCHED */
  int32_t ivarcv;     /* VARIABLECV */
  int32_t idewatcv;   /* VARIABLECV DEWATERED */
  int32_t ithickstrt; /* THICKSTRT */
  int32_t insto;      /* 1 if a STO package is present */
  const int32_t *ibotnode;  /* [nodes] lowest cell of the column (NPF ibotnode), may be NULL => self */
  /* STO */
  const double *ss;         /* [nodes] */
  const double *sy;         /* [nodes] */
  const int32_t *iconvert;  /* [nodes] */
  int32_t istor_coef; /* STORAGECOEFFICIENT */
  int32_t iconf_ss;   /* SS_CONFINED_ONLY */
  int32_t iorig_ss;   /* 1 = original (pre 6.2.1) ss formulation */
  int32_t reserved;
  /* NPF anisotropy (gwf-npf.f90 hy_eff :2280-2355, HGeoUtil.f90 hyeff :29-108).  All NULL = isotropic in the
   * plane (K22 = K).  k22 [nodes]; angle1/2/3 [nodes] in RADIANS (the deck's degrees x pi/180); conn_nx / conn_ny
   * [njas] = x, y components of the unit normal of every connection pointing from its lower- to its
   * higher-numbered cell (DIS: exactly (1,0) or (0,-1), Dis.f90:1039-1085; DISV/DISU: cos / sin of ANGLDEGX,
   * Disv.f90:979-1018); needed when k22 or angle1 is given */
  const double *k22;
  const double *angle1;
  const double *angle2;
  const double *angle3;
  const double *conn_nx;
  const double *conn_ny;
  /* NPF REWET (gwf-npf.f90 sgwf_npf_wetdry / rewet_check :2061-2223): wetdry [nodes] may be NULL (no rewetting);
   * irewet 1 = REWET given; wetfct / iwetit / ihdwet as in the REWET record */
  const double *wetdry;
  double wetfct;
  int32_t irewet;
  int32_t iwetit;
  int32_t ihdwet;
  int32_t reserved2;
} mf6gpu_gwf_model;

/* ---- stress packages (BoundaryPackage.f90:47-166) ------------------------ */
enum {
  MF6GPU_PKG_CHD = 1, /* b1 = head                           gwf-chd.f90 */
  MF6GPU_PKG_WEL = 2, /* b1 = q                              gwf-wel.f90 */
  MF6GPU_PKG_RIV = 3, /* b1 = stage, b2 = cond, b3 = rbot    gwf-riv.f90 */
  MF6GPU_PKG_RCH = 4, /* b1 = recharge (list based, fixed_cell) gwf-rch.f90 */
  MF6GPU_PKG_GHB = 5, /* b1 = bhead, b2 = cond               gwf-ghb.f90 */
  MF6GPU_PKG_DRN = 6  /* b1 = elev,  b2 = cond               gwf-drn.f90 */
};

typedef struct mf6gpu_bnd_package {
  int32_t type;
  int32_t nbound;
  int32_t index_base;
  int32_t iflowred;   /* WEL: AUTO_FLOW_REDUCE on/off.  RCH: FIXED_CELL (0 = recharge goes to the highest active
                         cell below the listed one, gwf-rch.f90:315-333; 1 = it stays on the listed cell) */
  double flowred;     /* WEL AUTO_FLOW_REDUCE fraction */
  const int32_t *nodelist; /* [nbound] */
  const double *b1;
  const double *b2;
  const double *b3;
} mf6gpu_bnd_package;

#define MF6GPU_MAX_BUDGET_TERMS 16

/* ---- what one time step reports (mfsim.lst / budget table content) ------ */
typedef struct mf6gpu_step_report {
  int32_t converged;
  int32_t outer_iterations;
  int32_t inner_iterations;      /* total over the outer iterations */
  int32_t nterms;                /* budget terms filled below */
  double max_dv;                 /* hncg of the last outer iteration (signed) */
  int32_t max_dv_loc;            /* 1-based node */
  int32_t npivot_fixes;
  int32_t nbacktracks;           /* backtracking steps taken in this time step */
  int32_t reserved;
  double totrin, totrot, pdiffr; /* Budget.f90:259-267 */
  double term_in[MF6GPU_MAX_BUDGET_TERMS];
  double term_out[MF6GPU_MAX_BUDGET_TERMS];
  int32_t term_id[MF6GPU_MAX_BUDGET_TERMS]; /* 100 STO-SS, 101 STO-SY, else package type */
  double t_formulate;            /* seconds in "Formulate" */
  double t_linsolve;             /* seconds in "Linear solve" */
} mf6gpu_step_report;

#ifdef __cplusplus
}
#endif
#endif
