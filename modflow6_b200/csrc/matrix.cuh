// matrix.cuh -- device-resident system matrix (MatrixBaseType replacement).
//
// HBM layout (all arrays in the FINAL numbering = rows sorted by ILU level):
//   SELL-32: rows are grouped in slices of 32 consecutive rows; slot k of row r
//   lives at  slice_ptr[r>>5] + 32*k + (r&31)  so that the 32 lanes of a warp
//   read 256 contiguous bytes of `val` and 128 of `col` per slot.
//   Slot 0 is the diagonal; slots 1..nlow are the lower (earlier-eliminated)
//   neighbours, slots nlow+1..rowlen-1 the upper ones, each group ascending in
//   the reference's elimination order so that per-row arithmetic is performed
//   in exactly the order of amux / ims_base_pcilu0 / ims_base_ilu0a.
//   The ILU0/MILU0 factor shares slice_ptr/col/rowlen/nlow and only adds a
//   second value array (slot 0 = inverse pivot = APC(n)).
#pragma once
#include "common.cuh"
#include "../../include/mf6gpu.h"

struct mf6gpu_matrix {
  int n = 0, nja = 0;   // n = OWNED rows
  int n_ext = 0;        // owned + halo columns (== n on a single GPU); vectors have this length
  int ordering = 0;
  int nlevels = 0;
  int nslices = 0;
  long long nslots = 0;
  int maxlen = 0;
  int uniform_w = 0;           // width shared by every slice (0 = ragged): enables the fixed-width kernels
  std::vector<int> elim;       // elimination order of the ILU: elim[k] = original row eliminated k-th
  std::vector<int> perm;       // perm[new] = old
  std::vector<int> iperm;      // iperm[old] = new
  std::vector<int> level_ptr;  // [nlevels+1] row ranges in final numbering
  mf6::DevBuf<int> d_perm, d_iperm;
  mf6::DevBuf<int> d_ord;      // elimination-order index of each final row (tie-breaks); empty => identity
  mf6::DevBuf<int> slice_ptr;  // [nslices+1]
  mf6::DevBuf<int> col;        // [nslots]
  mf6::DevBuf<double> val;     // [nslots]
  mf6::DevBuf<unsigned char> rowlen, nlow;  // [n]
  mf6::DevBuf<unsigned char> rowlen_loc_buf; // [n] entries without halo columns (ILU); empty => same as rowlen
  const unsigned char *rowlen_loc() const { return rowlen_loc_buf.n ? rowlen_loc_buf.p : rowlen.p; }
  // stencil compression of the column indices (fixed-width layout only): slot_off[s*W + k] = col - row
  // when that difference is the same for every real entry of slot k in slice s, else kNoOffset ->
  // the Krylov kernels then skip the 4-byte column loads of that slot for the whole warp
  mf6::DevBuf<int> slot_off;
  double slot_off_hit = 0.0;   // fraction of (slice, slot) pairs that are compressed
  // BLOCK_MULTICOLOR: the blocks (cell columns) of every colour, for the block-sweep triangular solves:
  // blk_rows[blk_off[c] + k * blk_nb[c] + q] = final row of the k-th cell (elimination order) of the q-th
  // block of colour c, or -1 past the end of a short block
  int blk_ncolors = 0;
  bool blk_chain_ok = false;   // every block is a chain: cell k couples only to cells k-1 / k+1 of its block
  std::vector<int> blk_off, blk_nb, blk_maxk;
  mf6::DevBuf<int> blk_rows;
  mf6::DevBuf<unsigned char> blk_nlow;   // nlow of blk_rows' rows, same layout (saves the sweeps one dependent load)
  mf6::DevBuf<unsigned char> blk_chain;  // same layout: SELL slot of the entry towards the previous (low nibble) / next
                                         // (high nibble) cell of the chain, 0 = none
  // regular colours: row of cell k of block q = blk_base[blk_base_off[c] + k] + q and the chain slots are nlow /
  // nlow + 1, so the sweeps need none of the three tables above
  std::vector<char> blk_affine;
  std::vector<int> blk_base_off;
  mf6::DevBuf<int> blk_base;
  std::vector<int> blk_base_h;  // host copy (short chains pass their bases as kernel parameters)
  std::vector<char> blk_has_lower, blk_has_upper;  // per colour: factor entries outside the chains in the L / U half
  mf6::DevBuf<int> csr2sell;   // [nja] slot of each original CSR entry
  mf6::DevBuf<double> stage;   // [nja] H2D/D2H staging of CSR values
  mf6::DevBuf<double> xs, ys;  // [n] staging vectors for host multiply
  cudaStream_t stream = 0;
  const int *ord_ptr() const { return d_ord.n ? d_ord.p : nullptr; }
};

namespace mf6 {

constexpr int kNoOffset = INT_MIN;

// y = A x (device vectors in final numbering); optional fused dot partial:
// if dot_with != nullptr accumulates sum_r dot_with[r]*y[r] into partial[blockIdx]
void launch_spmv(const mf6gpu_matrix &A, const double *val, const double *x, double *y,
                 cudaStream_t s);
// out[new] = in[perm[new]]
void launch_gather(int n, const int *perm, const double *in, double *out, cudaStream_t s);
// out[perm[new]] = in[new]
void launch_scatter(int n, const int *perm, const double *in, double *out, cudaStream_t s);

}  // namespace mf6
