// ilu0.cuh -- level-scheduled ILU0/MILU0 (see ilu0.cu)
#pragma once
#include "matrix.cuh"
#include "comm.cuh"

namespace mf6 {
// one numeric factorisation pass with fixed (delta, ipcflag); *d_failflag is set
// to 1 by any row whose pivot fails the reference's checks.  Returns launches.
int ilu0_factor(const mf6gpu_matrix &A, const double *aval, double *lu, double relax,
                double delta, int ipcflag, int *d_failflag, cudaStream_t s);
// optional fused rho = rin . d (CG): accumulated by the launches that finalise d
struct IluDotArgs {
  double *partial;       // [(n / kBlock + nlevels + 1) * 8] scratch: one slot per WARP of every finalising launch
  double *cta_sums;      // [256] scratch of the reduce kernel
  unsigned int *ticket;
  double *rho_out;       // receives the dot product
  double *beta_out;      // receives rho / *rho0
  const double *rho0;
  DistPush push{};       // fused split-model path: the rank's partial rho goes straight to the peers' mailboxes
};
// d = (LU)^-1 rin.  `done` (device flag, may be null) turns the kernels into no-ops.
// rin and d must be different arrays.
int ilu0_apply(const mf6gpu_matrix &A, const double *lu, const double *rin, double *d,
               const int *done, cudaStream_t s, const IluDotArgs *dot = nullptr);
// number of rho-partial slots the BLOCK_MULTICOLOR sweeps need (0 without block tables)
size_t ilu0_block_dot_slots(const mf6gpu_matrix &A);
}  // namespace mf6
