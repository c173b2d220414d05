/*
 * mf6gpu.h -- C ABI of libmf6gpu.so: the B200 (sm_100a) implementation of the
 * MODFLOW 6 per-outer-iteration hot path (GWF formulate + IMS linear solve).
 *
 * This is the drop-in boundary.  A Fortran host binds these entry points with
 * ISO_C_BINDING (see fortran/ and INTEGRATION.md) from three thin types that
 * extend the reference's own abstract interfaces, exactly the way the PETSc
 * backend does:
 *
 *   GpuMatrixType  extends MatrixBaseType        src/Utilities/Matrix/MatrixBase.f90:9-38
 *       (pattern: PetscMatrixType, src/Utilities/Matrix/PetscMatrix.F90)
 *   GpuVectorType  extends VectorBaseType        src/Utilities/Vector/VectorBase.f90:6-20
 *   GpuSolverType  extends LinearSolverBaseType  src/Solution/LinearSolverBase.f90:17-61
 *       (pattern: PetscSolverType, src/Solution/PETSc/PetscSolver.F90:82-362)
 *
 * and, for the device-resident formulate + outer iteration, from
 * NumericalSolutionType (src/Solution/NumericalSolution.f90) through the
 * mf6gpu_solution_* group.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on failure; the message is in
 *     mf6gpu_last_error().  Non-convergence is NOT an error (cf.
 *     PetscSolver.F90:346-360).  Nothing throws or exits across the ABI.
 *   - host arrays are borrowed for the duration of the call; the library owns
 *     every device buffer.  Structure (ia/ja) is uploaded once at create.
 *   - index arrays carry their base explicitly (1 = Fortran arrays unchanged).
 *   - calls are blocking; one handle is used by one host thread at a time.
 *   - there is NO CPU fallback: without a usable CUDA device every compute
 *     entry point fails with an error.
 */
#ifndef MF6GPU_H
#define MF6GPU_H

#include <stddef.h>
#include <stdint.h>
#include "mf6gpu_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MF6GPU_ABI_VERSION 1

typedef struct mf6gpu_matrix mf6gpu_matrix;
typedef struct mf6gpu_vector mf6gpu_vector;
typedef struct mf6gpu_solver mf6gpu_solver;
typedef struct mf6gpu_solution mf6gpu_solution;
typedef struct mf6gpu_comm mf6gpu_comm;

/* ---- library ------------------------------------------------------------ */
int mf6gpu_abi_version(void);
const char *mf6gpu_last_error(void);
/* sizeof() of the ABI structs, for binding self-checks: 0 ims_settings,
 * 1 sln_settings, 2 gwf_model, 3 bnd_package, 4 step_report */
size_t mf6gpu_sizeof(int which);
/* select the CUDA device of this process (one process per GPU); -1 keeps the
 * current device.  Fails if no device is usable. */
int mf6gpu_init(int device);
int mf6gpu_device_count(void);
/* Page-lock / release a host array that stays allocated across calls (the solution's amat, rhs, x): the host <->
 * device copies of mf6gpu_matrix_update / mf6gpu_solver_solve then run at full PCIe rate.  Optional. */
int mf6gpu_host_register(void *ptr, size_t bytes);
int mf6gpu_host_unregister(void *ptr);

/* ---- MatrixBaseType ------------------------------------------------------
 * SparseMatrixType%init (SparseMatrix.f90:54-74): CSR pattern, rows stored
 * diagonal first (Sparse.f90:217-239).  The values live on the device in a
 * level-sorted SELL-32 layout chosen by `gpu_ordering` (MF6GPU_ORDER_*). */
int mf6gpu_matrix_create(int32_t n, int32_t nja, const int32_t *ia,
                         const int32_t *ja, int32_t index_base,
                         int32_t gpu_ordering, mf6gpu_matrix **out);
int mf6gpu_matrix_destroy(mf6gpu_matrix *m);
/* PetscMatrixType%update analogue (PetscMatrix.F90:149-162): push the host
 * CSR values amat[nja] (original CSR order) to the device. */
int mf6gpu_matrix_update(mf6gpu_matrix *m, const double *amat);
/* spm_zero_entries (SparseMatrix.f90:251-260) */
int mf6gpu_matrix_zero_entries(mf6gpu_matrix *m);
/* read the device values back in original CSR order (get_aij analogue) */
int mf6gpu_matrix_get_values(mf6gpu_matrix *m, double *amat);
/* spm_multiply (SparseMatrix.f90:298-316 -> amux): y = A x, host vectors */
int mf6gpu_matrix_multiply(mf6gpu_matrix *m, const double *x, double *y);
/* split-model variant: n_own rows whose columns may name halo cells n_own..n_ext-1;
 * global_id[n_ext] (may be NULL) breaks arg-max ties like the unsplit model would */
int mf6gpu_matrix_create_ext(int32_t n_own, int32_t n_ext, int32_t nja, const int32_t *ia,
                             const int32_t *ja, int32_t index_base, int32_t gpu_ordering,
                             const int32_t *global_id, mf6gpu_matrix **out);
/* as above plus block_id[n_own] (may be NULL): the blocks of MF6GPU_ORDER_BLOCK_MULTICOLOR */
int mf6gpu_matrix_create_blocked(int32_t n_own, int32_t n_ext, int32_t nja, const int32_t *ia,
                                 const int32_t *ja, int32_t index_base, int32_t gpu_ordering,
                                 const int32_t *global_id, const int32_t *block_id, mf6gpu_matrix **out);
/* structure facts: 0 n, 1 nja, 2 number of ILU levels, 3 ordering, 4 SELL slots */
int64_t mf6gpu_matrix_info(const mf6gpu_matrix *m, int what);
/* elimination order of the ILU: perm[k] = row (0-based, original numbering) eliminated k-th; this is the
 * symmetric permutation under which the device ILU0 equals the reference algorithm (NATURAL: identity) */
int mf6gpu_matrix_get_permutation(const mf6gpu_matrix *m, int32_t *perm);
/* HOST ONLY (needs no device): the elimination order mf6gpu_matrix_create[_blocked] would choose for this
 * pattern, perm[k] = row eliminated k-th.  block_id may be NULL: MF6GPU_ORDER_BLOCK_MULTICOLOR then derives
 * chains from the pattern (dominant far stride = the vertical cell columns of a DIS / DISV numbering).
 * Lets a CPU checker run the reference algorithm on the identically permuted system anywhere. */
int mf6gpu_ordering_compute(int32_t n, int32_t n_ext, int32_t nja, const int32_t *ia, const int32_t *ja,
                            int32_t index_base, int32_t gpu_ordering, const int32_t *block_id, int32_t *perm);

/* ---- VectorBaseType (SeqVector.f90) ---------------------------------------- */
int mf6gpu_vector_create(int32_t n, mf6gpu_vector **out);
int mf6gpu_vector_destroy(mf6gpu_vector *v);
int mf6gpu_vector_set(mf6gpu_vector *v, const double *host);   /* set from host array */
int mf6gpu_vector_get(const mf6gpu_vector *v, double *host);   /* get_array analogue */
int mf6gpu_vector_zero_entries(mf6gpu_vector *v);              /* sqv_zero_entries :111-120 */
int mf6gpu_vector_axpy(mf6gpu_vector *y, double alpha, const mf6gpu_vector *x); /* sqv_axpy :135-148 */
int mf6gpu_vector_norm2(const mf6gpu_vector *v, double *result);               /* sqv_norm2 :152-164 */
int mf6gpu_vector_dot(const mf6gpu_vector *a, const mf6gpu_vector *b, double *result); /* ddot */

/* ---- LinearSolverBaseType ------------------------------------------------- */
/* create_matrix/initialize (LinearSolverBase.f90:31-41): the solver keeps a
 * reference to `m` (not owned).  summary_capacity = nitermax of
 * ConvergenceSummaryType (0 = do not record per-iteration data). */
int mf6gpu_solver_create(mf6gpu_matrix *m, const mf6gpu_ims_settings *settings,
                         int32_t summary_capacity, mf6gpu_solver **out);
int mf6gpu_solver_destroy(mf6gpu_solver *s);
/* solve (LinearSolverBase.f90:43-51) == imslinear_ap (ImsLinear.f90:617-750)
 * on the device: scale, ILU0/MILU0 factorisation (with the pivot rescue loop),
 * residual, CG or BiCGSTAB with the IMS stopping rules.  rhs/x are host arrays
 * of length n in ORIGINAL ordering; x is updated in place.
 * Outputs: iteration_number (inner iterations), is_converged (ICNVG: 1/0). */
int mf6gpu_solver_solve(mf6gpu_solver *s, int32_t kiter, int32_t kstp,
                        const double *rhs, double *x, int32_t *iteration_number,
                        int32_t *is_converged);
/* ConvergenceSummaryType side channel: copies min(count, cap) records of the
 * current time step; loc* are 1-based solution rows (original ordering).
 * Any output pointer may be NULL.  Returns the number of records. */
int mf6gpu_solver_get_summary(mf6gpu_solver *s, int32_t cap, int32_t *itinner,
                              double *dvmax, int32_t *locdv, double *rmax,
                              int32_t *locr, double *alpha, double *omega);
/* Per-model ConvergenceSummary (ImsLinearBase.f90:143-197, ConvergenceSummary.f90): convmodstart [nmod + 1] =
 * first row of every model (NumericalSolution.f90:409-416 builds it the same way), then the records come back
 * laid out like convdvmax(nmod, niter) / convlocdv / convrmax / convlocr (model index fastest; locations are
 * 1-based rows, 0 = none).  get_model_summary returns the number of iterations recorded (< 0 on error). */
int mf6gpu_solver_set_models(mf6gpu_solver *s, int32_t nmod, const int32_t *convmodstart, int32_t index_base);
int mf6gpu_solver_get_model_summary(mf6gpu_solver *s, int32_t cap, double *convdvmax, int32_t *convlocdv,
                                    double *convrmax, int32_t *convlocr);
/* facts of the last solve: 0 l2norm0, 1 pivot corrections, 2 device seconds in
 * factorisation, 3 device seconds in the Krylov loop, 4 kernel launches, 5 split-model path: 1 when the
 * fused peer-memory exchange (in-kernel pushes / waits over NVLink) was used, 0 for the NCCL transport */
double mf6gpu_solver_stat(const mf6gpu_solver *s, int what);
/* per-kernel-class device timing (CUDA events on the solver's stream):
 * classes 0 spmv, 1 ilu0 apply, 2 x/r update, 3 dot, 4 direction update, 5 factorisation */
int mf6gpu_solver_profile(mf6gpu_solver *s, int32_t enable);
int mf6gpu_solver_profile_get(mf6gpu_solver *s, int32_t cls, double *total_ms, int64_t *count);
/* preconditioner pieces exposed for parity tests (pcu + ilu0a):
 * factor the current matrix values, apply M^-1 to a host vector */
int mf6gpu_solver_factor(mf6gpu_solver *s, int32_t *npivot_fixes);
int mf6gpu_solver_apply_preconditioner(mf6gpu_solver *s, const double *r, double *z);

/* ---- one process per GPU: communicator of the split-model path ------------
 * replaces MpiRunControl / MpiRouter / PETSc's MPI use for the path in scope
 * (src/Distributed/MpiRouter.f90:238-343, src/Solution/ParallelSolution.f90:38-246).
 * Rank 0 calls mf6gpu_comm_unique_id, the 128 bytes are broadcast by the host
 * launcher (MPI, torch.distributed, ...), every rank calls mf6gpu_comm_create. */
int mf6gpu_comm_unique_id(void *out128);
int mf6gpu_comm_create(int32_t nranks, int32_t rank, const void *id128, mf6gpu_comm **out);
/* optional peer-memory (NVLink / NVSwitch) transport: each rank exports a mailbox (CUDA IPC handle,
 * 64 bytes; halo_doubles = largest halo message of any rank), the launcher all-gathers the handles,
 * every rank imports them.  Halo messages and the small all-gathers then move by direct st.global into
 * the consumer's memory + system-scope flags instead of NCCL calls. */
int mf6gpu_comm_p2p_export(mf6gpu_comm *c, int64_t halo_doubles, void *handle64);
int mf6gpu_comm_p2p_import(mf6gpu_comm *c, const void *handles);
int mf6gpu_comm_p2p_enabled(const mf6gpu_comm *c);
int mf6gpu_comm_p2p_disable(mf6gpu_comm *c); /* back to NCCL (e.g. when another rank could not map) */
int mf6gpu_comm_destroy(mf6gpu_comm *c);
int mf6gpu_comm_rank(const mf6gpu_comm *c);
int mf6gpu_comm_size(const mf6gpu_comm *c);

/* ---- NumericalSolutionType + GWF formulate, device resident --------------- */
int mf6gpu_solution_create(const mf6gpu_gwf_model *model,
                           const mf6gpu_sln_settings *sln,
                           const mf6gpu_ims_settings *ims,
                           mf6gpu_solution **out);
/* split-model path (GWF-GWF exchanges / interface model, exg-gwfgwf.f90:363-661,
 * SpatialModelConnection.f90:306-510): `model` describes this rank's submodel EXTENDED by its
 * halo cells -- cells 0..n_own-1 are owned, cells n_own..nodes-1 are the neighbour cells named by
 * the exchanges (their rows hold the diagonal only; exchange connections are ordinary
 * connections with their own jas / cl1 / cl2 / hwva / ihc).  Neighbour k (rank nbr_rank[k])
 * receives the owned cells send_idx[send_ptr[k]..send_ptr[k+1]) and fills the halo cells
 * n_own + recv_ptr[k] .. n_own + recv_ptr[k+1].  global_id[nodes] = cell numbers of the unsplit
 * model.  Every rank must pass the same list of package types to set_packages (possibly with
 * zero bounds).  Heads/ibound of halo cells move by NCCL send/recv, Krylov and outer-loop scalars
 * by small all-gathers; the preconditioner is ILU0/MILU0 of the rank's diagonal block
 * (block Jacobi, as PetscSolver.F90:251-278). */
int mf6gpu_solution_create_dist(const mf6gpu_gwf_model *model, const mf6gpu_sln_settings *sln,
                                const mf6gpu_ims_settings *ims, mf6gpu_comm *comm, int32_t n_own,
                                int32_t nnbr, const int32_t *nbr_rank, const int32_t *send_ptr,
                                const int32_t *send_idx, const int32_t *recv_ptr,
                                const int32_t *global_id, mf6gpu_solution **out);
int mf6gpu_solution_destroy(mf6gpu_solution *s);
/* GNC, ghost node correction of the connections of a locally refined grid (GhostNode.f90: read_data :739-864,
 * gnc_fc explicit branch :280-324, gnc_fn :340-443, gnc_cq :478-542).  Entry i corrects the connection between the
 * connected cells noden[i], nodem[i]; nodesj[i * numj + k] are the contributing cells of noden's grid (a value below
 * index_base = none) with weights alphasj[i * numj + k].  The correction is applied EXPLICITLY (right-hand side, one
 * outer iteration behind): the matrix keeps its pattern and its symmetry.  ngnc = 0 removes the corrections. */
int mf6gpu_solution_set_gnc(mf6gpu_solution *s, int32_t ngnc, int32_t numj, const int32_t *noden,
                            const int32_t *nodem, const int32_t *nodesj, const double *alphasj, int32_t index_base);
/* HFB, horizontal flow barriers of the coming stress period(s) (hfb_rp / condsat_modify / hfb_fc / hfb_cq,
 * gwf-hfb.f90:149-450, 770-832): barrier i lies between the connected cells noden[i], nodem[i] with hydraulic
 * characteristic hydchr[i] (negative = multiplier of the conductance).  nhfb = 0 removes the barriers. */
int mf6gpu_solution_set_hfb(mf6gpu_solution *s, int32_t nhfb, const int32_t *noden, const int32_t *nodem,
                            const double *hydchr, int32_t index_base);
/* bnd_rp: stress-period data of every package (copied to the device) */
int mf6gpu_solution_set_packages(mf6gpu_solution *s, int32_t npkg,
                                 const mf6gpu_bnd_package *pk);
/* sln_ca: prepareSolve + outer loop of solve(kiter) + finalizeSolve for one
 * time step (NumericalSolution.f90:1287-1327) */
int mf6gpu_solution_timestep(mf6gpu_solution *s, int32_t kper, int32_t kstp,
                             double delt, int32_t iss, mf6gpu_step_report *rep);
/* sln_buildsystem + the pre-solve fix-ups of sln_ls, no linear solve */
int mf6gpu_solution_formulate(mf6gpu_solution *s, int32_t kiter, double delt,
                              int32_t iss);
int mf6gpu_solution_get_x(mf6gpu_solution *s, double *x);        /* heads, original order */
int mf6gpu_solution_set_x(mf6gpu_solution *s, const double *x);
int mf6gpu_solution_reset_x(mf6gpu_solution *s);                 /* x = IC strt, device side */
int mf6gpu_solution_get_amat(mf6gpu_solution *s, double *amat);  /* CSR order */
int mf6gpu_solution_get_rhs(mf6gpu_solution *s, double *rhs);
int mf6gpu_solution_get_flowja(mf6gpu_solution *s, double *flowja);
int mf6gpu_solution_get_condsat(mf6gpu_solution *s, double *condsat);
/* simulated rate of every boundary of the last time step (bnd_cq_simrate, BoundaryPackage.f90:583-619;
 * calc_chd_rate, gwf-chd.f90:264-320), packages concatenated in set_packages order: what
 * save_print_model_flows (BoundaryPackage.f90:1753-1900) writes to the budget file.  *count = number
 * of boundaries; simvals may be NULL to query it */
int mf6gpu_solution_get_simvals(mf6gpu_solution *s, int32_t cap, double *simvals, int32_t *count);
/* the cell every boundary acted on in the last formulate, 0-based, packages concatenated like get_simvals.  Equal
 * to the input nodelist except for RCH without FIXED_CELL, whose recharge rch_cf hands down to the highest
 * active cell (gwf-rch.f90:315-333) -- the node the budget file records */
int mf6gpu_solution_get_nodes(mf6gpu_solution *s, int32_t cap, int32_t *nodes, int32_t *count);
/* STO-SS and STO-SY rates per cell of the last time step (sto_cq, gwf-sto.f90:447-564), original order */
int mf6gpu_solution_get_storage(mf6gpu_solution *s, double *strgss, double *strgsy);
/* elimination order of the owned cells (see mf6gpu_matrix_get_permutation) */
int mf6gpu_solution_get_permutation(mf6gpu_solution *s, int32_t *perm);
/* HOST ONLY: the elimination order mf6gpu_solution_create would use for this model (no device needed) */
int mf6gpu_model_elimination_order(const mf6gpu_gwf_model *m, int32_t gpu_ordering, int32_t *perm);
/* the linear solver owned by the solution (for stats / summary) */
mf6gpu_solver *mf6gpu_solution_solver(mf6gpu_solution *s);
/* facts: 0 kernel launches of the last time step, 1 ILU levels, 2 SELL slots,
 * 3 fraction of stencil-compressed (slice, slot) column groups, 4 fixed SELL width (0 = ragged),
 * 5 colours of the BLOCK_MULTICOLOR sweeps that need no row tables (-1: sweeps not applicable) */
double mf6gpu_solution_stat(const mf6gpu_solution *s, int what);

#ifdef __cplusplus
}
#endif
#endif
