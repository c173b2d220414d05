// ilut.cuh -- ILUT / MILUT preconditioner (IPC 3 / 4) of the IMS linear solver on the GPU path.
//
// Restates for the device path:
//   ilut, lusol, qsplit      src/Utilities/Libraries/sparskit2/ilut.f90:48-548 (SPARSKIT2 as vendored and modified
//                            by the reference: relaxed dropped terms, diagonal scaling, sign-preserving pivots)
//   ims_base_pcu (IPC 3/4)   src/Solution/LinearMethods/ImsLinearBase.f90:761-864
//   ims_calc_pcdims          :1148-1197 (capacity of the factor: neq * (2 * LEVEL + 1))
// Division of labour: the APPLY (lusol, once or twice per inner iteration -- the hot part) runs on the device,
// level-scheduled over the pattern the factorisation produced.  The FACTORISATION itself is inherently
// sequential -- the sparsity of row i of L and U depends on the numerical values of all rows before it (dual
// threshold dropping), so there is no static dependency graph to schedule -- and runs once per outer iteration on
// one host core inside this library, on the matrix values gathered from the device (elimination order, CSR).
#pragma once
#include "matrix.cuh"

namespace mf6 {

struct IlutPlan {
  int n = 0;
  int lfil = 0;
  double droptol = 0.0;
  // matrix in ELIMINATION numbering (row e = the row eliminated e-th), CSR, diagonal first then ascending columns
  // -- the layout the reference hands to ilut -- without halo columns
  std::vector<int> e_ia, e_ja;
  std::vector<int> row_of_e;          // final (device) row of elimination row e
  DevBuf<int> d_src;                  // [nnz_e] SELL slot of every entry
  DevBuf<double> d_val;               // [nnz_e] gathered values
  PinnedBuf<double> h_val;
  // host factor (1-based MSR like the reference)
  long long iwk = 0;
  std::vector<double> alu, w;
  std::vector<int> jlu, ju, jw;
  int ierr = 0;
  // device factor, rebuilt after every factorisation: rows in elimination numbering, columns = device rows
  DevBuf<int> lptr, lcol, uptr, ucol, erow;
  DevBuf<double> lval, uval, piv;
  DevBuf<int> flist, blist;           // rows sorted by forward / backward dependency level
  struct Group {
    int first, count;                 // range in flist / blist
    int nlev;                         // 1: one wide level (grid launch); > 1: a run of narrow levels (one CTA)
    int lev_off;                      // start of the run's level sizes in flev_sz / blev_sz
  };
  std::vector<Group> fgroups, bgroups;
  DevBuf<int> flev_sz, blev_sz;       // sizes of the levels inside the narrow runs
  int nflev = 0, nblev = 0;
  long long nnz_l = 0, nnz_u = 0;
  void build(const mf6gpu_matrix &A, int level, double droptol_);
  // pcu loop (delta / izero rescue); returns icount.  val = device matrix values (SELL)
  int factor(const mf6gpu_matrix &A, const double *val, double relax, cudaStream_t s);
  // d = (LU)^-1 rin on the device; returns launches
  int apply(const double *rin, double *d, const int *done, cudaStream_t s) const;
};

}  // namespace mf6
