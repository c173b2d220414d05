/* Minimal C host of libmf6gpu: what a Fortran / C / C++ MODFLOW 6 build links against.
 *
 *   gcc -std=c99 -Iinclude examples/host_cabi.c -Lmodflow6_b200 -lmf6gpu -Wl,-rpath,$PWD/modflow6_b200 -o host_cabi
 *
 * Solves the 1-D Laplace system of autotest/test_gwf_chd01.py (heads 1 ... 0 between two constant heads)
 * through the LinearSolverBase seam: mf6gpu_matrix_create / update, mf6gpu_solver_create / solve.
 * Without a GPU mf6gpu_init fails and the program says so (there is no CPU fallback); the CPU test suite
 * compiles and runs it only up to that point to prove that the header is plain C and the library links. */
#include <stdio.h>
#include <stdlib.h>

#include "mf6gpu.h"

#define N 100

int main(void) {
  static int32_t ia[N + 1], ja[3 * N];
  static double amat[3 * N], rhs[N], x[N];
  int nja = 0;
  printf("abi %d, sizeof(mf6gpu_ims_settings) %zu\n", mf6gpu_abi_version(), sizeof(mf6gpu_ims_settings));
  /* rows: diagonal first, then ascending columns (SparseMatrixType layout); first and last cell constant head */
  for (int i = 0; i < N; i++) {
    const int chd = (i == 0 || i == N - 1);
    ia[i] = nja;
    ja[nja] = i;
    amat[nja++] = chd ? 1.0 : -2.0;
    if (i > 0) { ja[nja] = i - 1; amat[nja++] = chd ? 0.0 : 1.0; }
    if (i < N - 1) { ja[nja] = i + 1; amat[nja++] = chd ? 0.0 : 1.0; }
    rhs[i] = (i == 0) ? 1.0 : 0.0;
    x[i] = 0.0;
  }
  ia[N] = nja;
  if (mf6gpu_init(0) != 0) {
    printf("no usable GPU: %s\n", mf6gpu_last_error());
    return 3;
  }
  mf6gpu_matrix *A = NULL;
  mf6gpu_solver *S = NULL;
  mf6gpu_ims_settings ims = {0};
  ims.dvclose = 1e-10; ims.rclose = 1e-8; ims.iter1 = 300; ims.ilinmeth = 2 /* BICGSTAB */;
  ims.gpu_ordering = MF6GPU_ORDER_NATURAL;
  int32_t iters = 0, cnvg = 0;
  if (mf6gpu_matrix_create(N, nja, ia, ja, 0, MF6GPU_ORDER_NATURAL, &A) != 0 ||
      mf6gpu_matrix_update(A, amat) != 0 || mf6gpu_solver_create(A, &ims, 0, &S) != 0 ||
      mf6gpu_solver_solve(S, 1, 1, rhs, x, &iters, &cnvg) != 0) {
    printf("error: %s\n", mf6gpu_last_error());
    return 1;
  }
  double err = 0.0;
  for (int i = 0; i < N; i++) {
    const double want = 1.0 - (double)i / (N - 1), e = x[i] > want ? x[i] - want : want - x[i];
    if (e > err) err = e;
  }
  printf("converged %d after %d inner iterations, max |h - linspace(1,0)| = %.3e\n", cnvg, iters, err);
  mf6gpu_solver_destroy(S);
  mf6gpu_matrix_destroy(A);
  return (cnvg == 1 && err < 1e-6) ? 0 : 2;
}
