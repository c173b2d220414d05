// solver.cuh -- device-resident IMS linear solver (LinearSolverBaseType replacement)
#pragma once
#include "ilu0.cuh"
#include "ilut.cuh"
#include <memory>
#include "comm.cuh"

namespace mf6 {

// one rank's contribution to a reduction of the split-model path (8 doubles)
struct RedRec {
  double s[2];
  MaxLocPOD mx, mr;
};

// Scalars of the Krylov recurrences; lives in device memory, never read by the
// host inside the inner loop (only `done`/`iter` are polled every few iterations).
struct KState {
  double rho, rho0, alpha, alpha0, omega, omega0, beta;
  double rho_acc;      // running r.z of the fused ILU-apply dot (0 between applies)
  double l2norm0, epfact, dvclose, rclose;
  double deltax, rmax, l2norm;
  int xloc, rloc;      // device-numbering rows of deltax / rmax
  int icnvgopt;
  int icnvg;           // ICNVG of the reference (1, 0, -1)
  int done;            // inner loop has exited
  int iter;            // INNERIT
  int sum_count;       // summary%iter_cnt
  int sum_cap;
  int iscl;
  int dist;            // split-model path: reductions are completed by global_finalize_kernel
  RedRec red;          // this rank's partial result of the reduction in flight
};

}  // namespace mf6

struct mf6gpu_solver {
  mf6gpu_matrix *A = nullptr;  // not owned
  mf6::HaloPlan *halo = nullptr;  // not owned; non-null on the split-model path
  mf6::DevBuf<double> red_all;    // [nranks * 8] gathered RedRec
  void reduce_finalize(int mode, double *out, int bcgs);  // all-gather + global finalize (no-op on 1 GPU)
  mf6gpu_ims_settings s{};
  int ipc = 1;                            // 1 ILU0, 2 MILU0, 3 ILUT, 4 MILUT (ImsLinear.f90:178-185)
  std::unique_ptr<mf6::IlutPlan> ilut;    // IPC 3 / 4
  // z = M^-1 r with the active preconditioner; `dot` (fused rho) only with ILU0 / MILU0
  int precond(const double *rin, double *d, cudaStream_t S, const mf6::IluDotArgs *dot = nullptr);
  cudaStream_t stream = 0;
  int n = 0;
  // work vectors (final numbering)
  mf6::DevBuf<double> lu;                 // factor values (SELL slots)
  mf6::DevBuf<double> x, b, d, p, q, z;   // CG
  mf6::DevBuf<double> t, v, dhat, phat, qhat;  // BiCGSTAB
  mf6::DevBuf<double> dscale, dscale2;
  mf6::DevBuf<double> hx, hb;             // staging in original numbering
  mf6::DevBuf<mf6::KState> st;
  mf6::DevBuf<double> partial;            // [4 * kMaxBlocks]
  mf6::DevBuf<double> ilu_partial;        // [n / kBlock + 2] fused ILU-apply dot partials
  mf6::DevBuf<mf6::MaxLoc> pmx, pmr;      // [kMaxBlocks]
  mf6::DevBuf<unsigned int> tickets;      // [8]
  mf6::DevBuf<int> failflag;
  // summary ring (device), capacity sum_cap
  int sum_cap = 0;
  mf6::DevBuf<int> sum_itinner, sum_locdv, sum_locr;
  mf6::DevBuf<double> sum_dvmax, sum_rmax, sum_alpha, sum_omega;
  // per-model records (ConvergenceSummaryType convdvmax(im, n) ...): optional, mf6gpu_solver_set_models
  int nmod = 0;
  mf6::DevBuf<int> modid;                 // [n] model of every final row
  mf6::DevBuf<double> msum_dvmax, msum_rmax;   // [sum_cap * nmod], model index fastest
  mf6::DevBuf<int> msum_locdv, msum_locr;
  mf6::DevBuf<mf6::MaxLoc> mpmx, mpmr;    // [nmod * kMaxBlocks]
  mf6::DevBuf<unsigned int> mtickets;     // [nmod]
  mf6::PinnedBuf<mf6::KState> h_st;
  mf6::PinnedBuf<int> h_flag;
  // stats of the last solve
  double l2norm0 = 0.0;
  int npivfix = 0;
  double t_factor = 0.0, t_krylov = 0.0;
  long long launches = 0;
  bool fused_exchange = false;  // last solve used the fused peer-memory exchange (split-model path)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t cap_stream = nullptr;     // capture-only stream of the inner-iteration graph (the work stream is the legacy default stream, which cannot capture)
  // optional per-kernel-class timing with CUDA events on the launching stream
  // (classes: 0 spmv, 1 ilu apply, 2 x/r update, 3 dot, 4 direction update, 5 factor)
  enum { PC_SPMV = 0, PC_ILU = 1, PC_UPD = 2, PC_DOT = 3, PC_PUPD = 4, PC_FACTOR = 5, PC_N = 6 };
  bool profiling = false;
  bool prof_on = true;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<int> ev_cls;
  double prof_ms[PC_N] = {0, 0, 0, 0, 0, 0};
  long long prof_cnt[PC_N] = {0, 0, 0, 0, 0, 0};
  void prof_begin(int cls);
  void prof_end();
  void prof_collect();

  // device-side entry: x_dev / b_dev already in FINAL numbering on the device.
  void solve_device(int kiter, int kstp, double *x_dev, double *b_dev, int *iters, int *icnvg);
  int factor();  // pcu: returns pivot corrections
};
