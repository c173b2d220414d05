// matrix.cu -- build of the level-sorted SELL-32 system matrix, value upload /
// download, SpMV (amux) and permutation kernels.
//
// Reference behaviour restated here:
//   SparseMatrixType   src/Utilities/Matrix/SparseMatrix.f90:54-74, 251-260, 298-316
//   amux               src/Utilities/Libraries/sparsekit/sparsekit.f90:1-59
//   PetscMatrixType%update (host CSR -> backend)  src/Utilities/Matrix/PetscMatrix.F90:149-162
#include "matrix.cuh"
#include "spmv.cuh"
#include <algorithm>
#include <cstring>
#include <numeric>

namespace mf6 {

int max_blocks() {
  static int v = [] {
    const char *e = std::getenv("MF6GPU_GRID_CAP");
    int c = e ? std::atoi(e) : 0;
    return c > 0 ? c : 148 * 8;
  }();
  return v;
}

std::string &last_error() {
  static thread_local std::string e;
  return e;
}

// ---------------------------------------------------------------- kernels ---
// y = A x : one thread per row, slots read slice-coalesced.  Loads of a chunk
// of 8 slots are issued before any use so that each thread keeps up to 24
// independent loads in flight; the accumulation keeps the row's storage order
// (diagonal, then ascending) so the result is bit-identical to amux.
__global__ void __launch_bounds__(kBlock)
spmv_sell32_kernel(int n, const int *__restrict__ slice_ptr,
                   const unsigned char *__restrict__ rowlen,
                   const int *__restrict__ col, const double *__restrict__ val,
                   const double *__restrict__ x, double *__restrict__ y) {
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n;
       row += gridDim.x * blockDim.x) {
    const double t = sell_row_dot(row, slice_ptr, rowlen, col, val, x);
    y[row] = t;
  }
}

__global__ void gather_kernel(int n, const int *__restrict__ perm,
                              const double *__restrict__ in, double *__restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = in[perm[i]];
}

__global__ void scatter_kernel(int n, const int *__restrict__ perm,
                               const double *__restrict__ in, double *__restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[perm[i]] = in[i];
}

// sell[csr2sell[p]] = csr[p]
__global__ void csr_to_sell_kernel(int nja, const int *__restrict__ map,
                                   const double *__restrict__ csr, double *__restrict__ sell) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nja; p += gridDim.x * blockDim.x)
    sell[map[p]] = csr[p];
}

__global__ void sell_to_csr_kernel(int nja, const int *__restrict__ map,
                                   const double *__restrict__ sell, double *__restrict__ csr) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nja; p += gridDim.x * blockDim.x)
    csr[p] = sell[map[p]];
}

void launch_spmv(const mf6gpu_matrix &A, const double *val, const double *x, double *y,
                 cudaStream_t s) {
  spmv_sell32_kernel<<<grid_for(A.n), kBlock, 0, s>>>(A.n, A.slice_ptr.p, A.rowlen.p, A.col.p,
                                                      val, x, y);
}

void launch_gather(int n, const int *perm, const double *in, double *out, cudaStream_t s) {
  gather_kernel<<<grid_for(n), kBlock, 0, s>>>(n, perm, in, out);
}

void launch_scatter(int n, const int *perm, const double *in, double *out, cudaStream_t s) {
  scatter_kernel<<<grid_for(n), kBlock, 0, s>>>(n, perm, in, out);
}

// ------------------------------------------------------------- host build ---
// Blocks for BLOCK_MULTICOLOR when the caller of the LinearSolverBase seam gives none (it only has the
// sparsity pattern): chains along the dominant far stride of the pattern.  In the reference's DIS / DISV
// numbering the LAST entry of a row is the cell below (offset nrow*ncol resp. ncpl, the same for every
// row above the bottom layer), so the most frequent "last upper offset" S recovers the vertical cell
// columns; on a single-layer grid it recovers the grid lines.  Cells v and v+S are chained when they are
// connected; any partition into such chains is a valid block set (blocks are only required to be chains).
// Returns an empty vector when fewer than a quarter of the rows can be chained.
static std::vector<int32_t> derive_chain_blocks(int n, const std::vector<int> &ia, const std::vector<int> &ja) {
  std::vector<int> off(n, 0);
  std::vector<int> offs;
  offs.reserve(n);
  for (int v = 0; v < n; v++) {
    int best = 0;
    for (int p = ia[v] + 1; p < ia[v + 1]; p++)
      if (ja[p] < n && ja[p] - v > best) best = ja[p] - v;
    off[v] = best;
    if (best > 0) offs.push_back(best);
  }
  if (offs.empty()) return {};
  std::sort(offs.begin(), offs.end());
  int S = 0;
  size_t bestc = 0;
  for (size_t i = 0; i < offs.size();) {
    size_t j = i;
    while (j < offs.size() && offs[j] == offs[i]) j++;
    if (j - i > bestc || (j - i == bestc && offs[i] > S)) {
      bestc = j - i;
      S = offs[i];
    }
    i = j;
  }
  if (S <= 0 || bestc * 4 < (size_t)n) return {};
  std::vector<int32_t> block(n);
  for (int v = 0; v < n; v++) {
    block[v] = v;
    const int u = v - S;
    if (u >= 0 && off[u] == S) block[v] = block[u];  // u's last entry is v: chained
  }
  return block;
}

// Elimination order of the ILU for one of the gpu_ordering choices (host only, no device work):
// ordidx[old] = position in the reference-style elimination loop.  BLOCK_MULTICOLOR also returns the
// compact block of every row and the colour of every block.
struct ElimOrder {
  std::vector<int> ordidx;
  std::vector<int> blk_of, blk_color;
  int blk_count = 0, blk_colors = 0;
};

static ElimOrder compute_elimination(int n, const std::vector<int> &ia, const std::vector<int> &ja, int ordering,
                                     const int32_t *block_id) {
  ElimOrder E;
  std::vector<int> &ordidx = E.ordidx;
  ordidx.resize(n);
  std::vector<int> &blk_of = E.blk_of, &blk_color = E.blk_color;
  int &blk_count = E.blk_count, &blk_colors = E.blk_colors;
  auto is_halo = [&](int c) { return c >= n; };
  if (ordering == MF6GPU_ORDER_NATURAL) {
    std::iota(ordidx.begin(), ordidx.end(), 0);
  } else if (ordering == MF6GPU_ORDER_BLOCK_MULTICOLOR) {
    // Greedy colouring of the QUOTIENT graph of the blocks (e.g. the vertical cell columns of a layered
    // grid): blocks of one colour are mutually independent, inside a block the reference's natural order is
    // kept, so the strong intra-block couplings are eliminated exactly like in the reference.
    // Elimination order = (colour of the block, original index).
    std::vector<int> bmap(n);  // compact block numbers in order of first appearance
    int nblk = 0;
    {
      std::vector<int> seen;
      int maxb = 0;
      for (int v = 0; v < n; v++) maxb = std::max(maxb, (int)block_id[v]);
      seen.assign((size_t)maxb + 1, -1);
      for (int v = 0; v < n; v++) {
        MF6_REQUIRE(block_id[v] >= 0, "matrix_create: negative block id");
        int &sb = seen[block_id[v]];
        if (sb < 0) sb = nblk++;
        bmap[v] = sb;
      }
    }
    std::vector<int> bptr(nblk + 1, 0), bmem(n);
    for (int v = 0; v < n; v++) bptr[bmap[v] + 1]++;
    for (int b = 0; b < nblk; b++) bptr[b + 1] += bptr[b];
    {
      std::vector<int> cur(bptr.begin(), bptr.end() - 1);
      for (int v = 0; v < n; v++) bmem[cur[bmap[v]]++] = v;
    }
    std::vector<int> bcolor(nblk, -1);
    int ncolors = 0;
    for (int b = 0; b < nblk; b++) {
      unsigned long long mask = 0ull;
      for (int q = bptr[b]; q < bptr[b + 1]; q++) {
        const int v = bmem[q];
        for (int p = ia[v] + 1; p < ia[v + 1]; p++) {
          if (is_halo(ja[p])) continue;
          const int nbk = bmap[ja[p]];
          if (nbk == b) continue;
          const int c = bcolor[nbk];
          MF6_REQUIRE(c < 64, "matrix_create: more than 64 block colours");
          if (c >= 0) mask |= (1ull << c);
        }
      }
      int c = 0;
      while ((mask >> c) & 1ull) c++;
      bcolor[b] = c;
      if (c + 1 > ncolors) ncolors = c + 1;
    }
    std::vector<int> cnt(ncolors + 1, 0);
    for (int v = 0; v < n; v++) cnt[bcolor[bmap[v]] + 1]++;
    for (int c = 0; c < ncolors; c++) cnt[c + 1] += cnt[c];
    for (int v = 0; v < n; v++) ordidx[v] = cnt[bcolor[bmap[v]]]++;
    blk_of = bmap;
    blk_color = bcolor;
    blk_count = nblk;
    blk_colors = ncolors;
  } else {
    // greedy colouring in natural order
    std::vector<int> color(n, -1);
    int ncolors = 0;
    std::vector<unsigned long long> forb;
    for (int v = 0; v < n; v++) {
      unsigned long long mask = 0ull;
      bool big = false;
      for (int p = ia[v] + 1; p < ia[v + 1]; p++) {
        if (is_halo(ja[p])) continue;
        int c = color[ja[p]];
        if (c >= 64) big = true;
        else if (c >= 0) mask |= (1ull << c);
      }
      int c = 0;
      if (!big) {
        while (c < 64 && (mask >> c) & 1ull) c++;
      }
      if (big || c == 64) {  // rare: fall back to an explicit set
        std::vector<char> used(ncolors + 2, 0);
        for (int p = ia[v] + 1; p < ia[v + 1]; p++)
          if (!is_halo(ja[p]) && color[ja[p]] >= 0) used[color[ja[p]]] = 1;
        c = 0;
        while (used[c]) c++;
      }
      color[v] = c;
      if (c + 1 > ncolors) ncolors = c + 1;
    }
    std::vector<int> cnt(ncolors + 1, 0);
    for (int v = 0; v < n; v++) cnt[color[v] + 1]++;
    for (int c = 0; c < ncolors; c++) cnt[c + 1] += cnt[c];
    for (int v = 0; v < n; v++) ordidx[v] = cnt[color[v]]++;
  }
  return E;
}


// n = owned rows, n_ext >= n = owned + halo columns; gid (optional, [n_ext]) = global ids used
// for arg-max tie-breaks in the split-model path
static void build_matrix(mf6gpu_matrix &M, int n, int n_ext, int nja, const int32_t *ia_in,
                         const int32_t *ja_in, int base, int ordering, const int32_t *gid,
                         const int32_t *block_id = nullptr) {
  MF6_REQUIRE(n > 0 && nja >= n && n_ext >= n, "matrix_create: bad dimensions");
  MF6_REQUIRE(ordering == MF6GPU_ORDER_NATURAL || ordering == MF6GPU_ORDER_MULTICOLOR ||
                  ordering == MF6GPU_ORDER_BLOCK_MULTICOLOR,
              "matrix_create: unknown gpu_ordering");
  M.n = n;
  M.n_ext = n_ext;
  M.nja = nja;
  M.ordering = ordering;
  std::vector<int> ia(n + 1), ja(nja);
  for (int i = 0; i <= n; i++) ia[i] = ia_in[i] - base;
  for (int i = 0; i < nja; i++) ja[i] = ja_in[i] - base;
  MF6_REQUIRE(ia[0] == 0 && ia[n] == nja, "matrix_create: ia does not span nja (check index_base)");
  for (int r = 0; r < n; r++) {
    MF6_REQUIRE(ia[r + 1] > ia[r], "matrix_create: empty row");
    MF6_REQUIRE(ja[ia[r]] == r, "matrix_create: rows must store the diagonal first (Sparse.f90:217-239)");
    MF6_REQUIRE(ia[r + 1] - ia[r] <= 255, "matrix_create: more than 255 entries in a row");
  }
  for (int p = 0; p < nja; p++) MF6_REQUIRE(ja[p] >= 0 && ja[p] < n_ext, "matrix_create: column out of range");
  auto is_halo = [&](int c) { return c >= n; };
  std::vector<int32_t> derived;
  if (ordering == MF6GPU_ORDER_BLOCK_MULTICOLOR && !block_id) {
    derived = derive_chain_blocks(n, ia, ja);
    if (derived.empty())
      ordering = MF6GPU_ORDER_MULTICOLOR;
    else
      block_id = derived.data();
  }
  M.ordering = ordering;
  ElimOrder E = compute_elimination(n, ia, ja, ordering, block_id);
  std::vector<int> &ordidx = E.ordidx;
  std::vector<int> &blk_of = E.blk_of, &blk_color = E.blk_color;
  const int blk_count = E.blk_count, blk_colors = E.blk_colors;
  std::vector<int> byord(n);  // byord[ord] = old
  for (int v = 0; v < n; v++) byord[ordidx[v]] = v;
  M.elim = byord;
  // --- dependency levels of the lower-triangular solve in that order
  std::vector<int> level(n, 0);
  int nlevels = 0;
  for (int o = 0; o < n; o++) {
    int v = byord[o];
    int lv = 0;
    for (int p = ia[v] + 1; p < ia[v + 1]; p++) {
      int u = ja[p];
      if (is_halo(u)) continue;
      if (ordidx[u] < o && level[u] + 1 > lv) lv = level[u] + 1;
    }
    level[v] = lv;
    if (lv + 1 > nlevels) nlevels = lv + 1;
  }
  M.nlevels = nlevels;
  // --- final numbering: stable sort by (level, ordidx)
  M.level_ptr.assign(nlevels + 1, 0);
  for (int v = 0; v < n; v++) M.level_ptr[level[v] + 1]++;
  for (int l = 0; l < nlevels; l++) M.level_ptr[l + 1] += M.level_ptr[l];
  M.perm.resize(n_ext);
  M.iperm.resize(n_ext);
  for (int h = n; h < n_ext; h++) M.perm[h] = M.iperm[h] = h;  // halo columns keep their place
  {
    std::vector<int> cur(M.level_ptr.begin(), M.level_ptr.end() - 1);
    for (int o = 0; o < n; o++) {
      int v = byord[o];
      int r = cur[level[v]]++;
      M.perm[r] = v;
      M.iperm[v] = r;
    }
  }
  // --- SELL-32 structure
  M.nslices = (n + 31) / 32;
  std::vector<unsigned char> rowlen(n), nlow(n), rowlen_loc(n);
  std::vector<int> slice_ptr(M.nslices + 1, 0);
  int maxlen = 0;
  // slice widths: the max row length of each slice, or -- when that costs <= MF6GPU_UNIFORM_PAD_PCT
  // percent (default 12.5: a 5-layer DIS grid pads 6 %) extra slots -- the global max for every slice
  // (fixed-width kernels, stencil table, block sweeps)
  long long ragged_slots = 0;
  std::vector<int> sw(M.nslices, 0);
  for (int s = 0; s < M.nslices; s++) {
    int w = 0;
    for (int r = s * 32; r < std::min(n, s * 32 + 32); r++) {
      int v = M.perm[r];
      int len = ia[v + 1] - ia[v];
      rowlen[r] = (unsigned char)len;
      if (len > w) w = len;
    }
    sw[s] = w;
    ragged_slots += 32LL * w;
    if (w > maxlen) maxlen = w;
  }
  {
    const char *e = std::getenv("MF6GPU_UNIFORM_PAD_PCT");
    const double pct = e ? std::atof(e) : 12.5;
    const long long uni_slots = 32LL * maxlen * M.nslices;
    if ((double)(uni_slots - ragged_slots) <= 0.01 * pct * (double)ragged_slots)
      for (int s = 0; s < M.nslices; s++) sw[s] = maxlen;
  }
  for (int s = 0; s < M.nslices; s++) {
    long long next = (long long)slice_ptr[s] + 32LL * sw[s];
    MF6_REQUIRE(next < (long long)INT_MAX, "matrix_create: SELL storage exceeds 2^31 slots");
    slice_ptr[s + 1] = (int)next;
  }
  M.maxlen = maxlen;
  {
    bool uni = true;
    for (int sl = 0; sl < M.nslices; sl++)
      if (slice_ptr[sl + 1] - slice_ptr[sl] != 32 * maxlen) uni = false;
    M.uniform_w = uni ? maxlen : 0;
  }
  M.nslots = slice_ptr[M.nslices];
  std::vector<int> col((size_t)M.nslots);
  std::vector<int> csr2sell(nja);
  std::vector<std::pair<int, int>> tmp;  // (ordidx(col), csr position)
  for (int r = 0; r < n; r++) {
    int v = M.perm[r];
    long long base_slot = (long long)slice_ptr[r >> 5] + (r & 31);
    int w = (slice_ptr[(r >> 5) + 1] - slice_ptr[r >> 5]) / 32;
    tmp.clear();
    // local neighbours in elimination order, then halo columns (ascending): halo sorts last
    for (int p = ia[v] + 1; p < ia[v + 1]; p++)
      tmp.emplace_back(is_halo(ja[p]) ? (n + ja[p]) : ordidx[ja[p]], p);
    std::sort(tmp.begin(), tmp.end());
    col[base_slot] = r;
    csr2sell[ia[v]] = (int)base_slot;
    int lo = 0, nloc = 1;
    for (size_t k = 0; k < tmp.size(); k++) {
      int p = tmp[k].second;
      long long slot = base_slot + 32LL * (long long)(k + 1);
      col[slot] = M.iperm[ja[p]];
      csr2sell[p] = (int)slot;
      if (!is_halo(ja[p]) && tmp[k].first < ordidx[v]) lo++;
      if (!is_halo(ja[p])) nloc++;
      MF6_REQUIRE(ja[p] != v, "matrix_create: duplicate diagonal entry");
      if (k > 0) MF6_REQUIRE(tmp[k].first != tmp[k - 1].first, "matrix_create: duplicate column in a row");
    }
    nlow[r] = (unsigned char)lo;
    rowlen_loc[r] = (unsigned char)nloc;
    for (int k = (int)tmp.size() + 1; k < w; k++) col[base_slot + 32LL * k] = r;  // padding
  }
  // padding lanes of the last slice
  for (int r = n; r < M.nslices * 32; r++) {
    long long base_slot = (long long)slice_ptr[r >> 5] + (r & 31);
    int w = (slice_ptr[(r >> 5) + 1] - slice_ptr[r >> 5]) / 32;
    for (int k = 0; k < w; k++) col[base_slot + 32LL * k] = 0;
  }
  // --- block tables for the block-sweep triangular solves
  if (ordering == MF6GPU_ORDER_BLOCK_MULTICOLOR && blk_count > 0) {
    M.blk_ncolors = blk_colors;
    std::vector<int> bsize(blk_count, 0), bq(blk_count, 0);
    for (int v = 0; v < n; v++) bsize[blk_of[v]]++;
    M.blk_nb.assign(blk_colors, 0);
    M.blk_maxk.assign(blk_colors, 0);
    for (int b = 0; b < blk_count; b++) {  // blocks are numbered in order of first appearance
      const int c = blk_color[b];
      bq[b] = M.blk_nb[c]++;
      M.blk_maxk[c] = std::max(M.blk_maxk[c], bsize[b]);
    }
    M.blk_off.assign(blk_colors + 1, 0);
    for (int c = 0; c < blk_colors; c++) M.blk_off[c + 1] = M.blk_off[c] + M.blk_nb[c] * M.blk_maxk[c];
    std::vector<int> rows((size_t)M.blk_off[blk_colors], -1), fill(blk_count, 0);
    for (int o = 0; o < n; o++) {  // elimination order: cells of a block appear top to bottom
      const int v = byord[o], b = blk_of[v], c = blk_color[b];
      rows[(size_t)M.blk_off[c] + (size_t)fill[b]++ * M.blk_nb[c] + bq[b]] = M.iperm[v];
    }
    M.blk_rows.upload(rows);
    bool nibble_ok = false;
    {
      // per cell: nlow, and the SELL slots of the entries towards the previous / next cell of the chain
      // (low / high nibble, 0 = none); per colour: is there any factor entry outside the chains?
      std::vector<unsigned char> bl(rows.size(), 0), bc(rows.size(), 0);
      nibble_ok = true;
      M.blk_has_lower.assign(blk_colors, 0);
      M.blk_has_upper.assign(blk_colors, 0);
      for (int c = 0; c < blk_colors; c++) {
        const size_t nbc = (size_t)M.blk_nb[c];
        for (size_t i = (size_t)M.blk_off[c]; i < (size_t)M.blk_off[c + 1]; i++) {
          const int r = rows[i];
          if (r < 0) continue;
          bl[i] = nlow[r];
          const size_t k = (i - (size_t)M.blk_off[c]) / nbc;
          const int rprev = (k > 0) ? rows[i - nbc] : -1;
          const int rnext = ((int)k + 1 < M.blk_maxk[c]) ? rows[i + nbc] : -1;
          int slo = 0, sup = 0;
          for (int u = 1; u < rowlen_loc[r]; u++) {
            const int cc = col[(size_t)slice_ptr[r >> 5] + 32 * (size_t)u + (r & 31)];
            if (u <= nlow[r]) {
              if (cc == rprev && !slo) slo = u; else M.blk_has_lower[c] = 1;
            } else {
              if (cc == rnext && !sup) sup = u; else M.blk_has_upper[c] = 1;
            }
          }
          if (slo > 15 || sup > 15) nibble_ok = false;
          bc[i] = (unsigned char)(slo | (sup << 4));
        }
      }
      M.blk_nlow.upload(bl);
      M.blk_chain.upload(bc);
      // regular colours (structured grids): all blocks full, the k-th cells of consecutive blocks are
      // consecutive rows, the chain entries are the last lower / first upper slot -> no table loads
      std::vector<int> base;
      M.blk_affine.assign(blk_colors, 0);
      M.blk_base_off.assign(blk_colors + 1, 0);
      for (int c = 0; c < blk_colors; c++) {
        const size_t nbc = (size_t)M.blk_nb[c], off = (size_t)M.blk_off[c];
        bool aff = nbc > 0 && !std::getenv("MF6GPU_NO_AFFINE");
        for (int k = 0; k < M.blk_maxk[c] && aff; k++) {
          const int r0 = rows[off + k * nbc];
          for (size_t q = 0; q < nbc; q++) {
            const size_t i = off + k * nbc + q;
            const int want = (k > 0 ? bl[i] : 0) | ((k + 1 < M.blk_maxk[c] ? bl[i] + 1 : 0) << 4);
            if (r0 < 0 || rows[i] != r0 + (int)q || bc[i] != want) {
              aff = false;
              break;
            }
          }
        }
        M.blk_affine[c] = aff;
        for (int k = 0; k < M.blk_maxk[c]; k++) base.push_back(nbc ? rows[off + k * nbc] : -1);
        M.blk_base_off[c + 1] = (int)base.size();
      }
      M.blk_base.upload(base);
      M.blk_base_h = base;
    }
    // the block-sweep kernels assume chains: the only intra-block neighbours of the k-th cell are cells k-1, k+1
    bool chain = true;
    {
      std::vector<int> kpos(n, 0), fill2(blk_count, 0);
      for (int o = 0; o < n; o++) kpos[byord[o]] = fill2[blk_of[byord[o]]]++;
      for (int v = 0; v < n && chain; v++)
        for (int p = ia[v] + 1; p < ia[v + 1]; p++) {
          const int u = ja[p];
          if (is_halo(u) || blk_of[u] != blk_of[v]) continue;
          if (std::abs(kpos[u] - kpos[v]) != 1) chain = false;
        }
    }
    M.blk_chain_ok = chain && nibble_ok;
  }
  // --- stencil compression table (fixed-width layout)
  if (M.uniform_w > 0 && !std::getenv("MF6GPU_NO_STENCIL")) {
    const int W = M.uniform_w;
    std::vector<int> soff((size_t)M.nslices * W, 0);
    long long hit = 0;
    for (int sl = 0; sl < M.nslices; sl++) {
      for (int k = 0; k < W; k++) {
        bool have = false, uni = true;
        int off = 0;
        for (int r = sl * 32; r < std::min(n, sl * 32 + 32); r++) {
          if (k >= rowlen[r]) continue;  // padding slot: its value is 0, any in-range column will do
          const int o = col[(size_t)slice_ptr[sl] + 32 * k + (r & 31)] - r;
          if (!have) {
            off = o;
            have = true;
          } else if (o != off) {
            uni = false;
            break;
          }
        }
        soff[(size_t)sl * W + k] = uni ? off : kNoOffset;
        if (uni) hit++;
      }
    }
    M.slot_off_hit = (double)hit / (double)soff.size();
    M.slot_off.upload(soff);
  }
  // --- upload
  M.d_perm.upload(M.perm);
  M.d_iperm.upload(M.iperm);
  if (gid) {
    std::vector<int> o(n);
    for (int r = 0; r < n; r++) o[r] = gid[M.perm[r]] - base;  // global cell id of final row r
    M.d_ord.upload(o);
  } else if (ordering == MF6GPU_ORDER_BLOCK_MULTICOLOR) {
    std::vector<int> o(n);
    for (int r = 0; r < n; r++) o[r] = ordidx[M.perm[r]];  // position in the elimination order
    M.d_ord.upload(o);
  } else if (ordering == MF6GPU_ORDER_NATURAL && nlevels > 1) {
    std::vector<int> o(M.perm.begin(), M.perm.begin() + n);
    M.d_ord.upload(o);  // elimination index of final row r is its original index
  }
  if (n_ext > n) M.rowlen_loc_buf.upload(rowlen_loc);
  M.slice_ptr.upload(slice_ptr);
  M.col.upload(col);
  M.rowlen.upload(rowlen);
  M.nlow.upload(nlow);
  M.csr2sell.upload(csr2sell);
  M.val.alloc_zero((size_t)M.nslots);
  M.stage.alloc((size_t)nja);
  M.xs.alloc_zero((size_t)n_ext);
  M.ys.alloc_zero((size_t)n_ext);
}

}  // namespace mf6

using namespace mf6;

extern "C" {

int mf6gpu_matrix_create(int32_t n, int32_t nja, const int32_t *ia, const int32_t *ja,
                         int32_t index_base, int32_t gpu_ordering, mf6gpu_matrix **out) {
  return guard([&] {
    MF6_REQUIRE(out && ia && ja, "matrix_create: null argument");
    int dev;
    MF6_CK(cudaGetDevice(&dev));
    auto *M = new mf6gpu_matrix();
    try {
      build_matrix(*M, n, n, nja, ia, ja, index_base, gpu_ordering, nullptr);
    } catch (...) {
      delete M;
      throw;
    }
    *out = M;
  });
}

int mf6gpu_matrix_create_ext(int32_t n_own, int32_t n_ext, int32_t nja, const int32_t *ia,
                             const int32_t *ja, int32_t index_base, int32_t gpu_ordering,
                             const int32_t *global_id, mf6gpu_matrix **out) {
  return guard([&] {
    MF6_REQUIRE(out && ia && ja, "matrix_create_ext: null argument");
    auto *M = new mf6gpu_matrix();
    try {
      build_matrix(*M, n_own, n_ext, nja, ia, ja, index_base, gpu_ordering, global_id, nullptr);
    } catch (...) {
      delete M;
      throw;
    }
    *out = M;
  });
}

int mf6gpu_matrix_create_blocked(int32_t n_own, int32_t n_ext, int32_t nja, const int32_t *ia,
                                 const int32_t *ja, int32_t index_base, int32_t gpu_ordering,
                                 const int32_t *global_id, const int32_t *block_id, mf6gpu_matrix **out) {
  return guard([&] {
    MF6_REQUIRE(out && ia && ja, "matrix_create_blocked: null argument");
    auto *M = new mf6gpu_matrix();
    try {
      build_matrix(*M, n_own, n_ext, nja, ia, ja, index_base, gpu_ordering, global_id, block_id);
    } catch (...) {
      delete M;
      throw;
    }
    *out = M;
  });
}

// Host-only (no device needed): the elimination order build_matrix would choose for this pattern,
// perm[k] = original row eliminated k-th.  Lets a CPU checker apply the reference algorithm to the
// same symmetrically permuted system without a GPU (oracle runs at full size, tests/golden/).
int mf6gpu_ordering_compute(int32_t n, int32_t n_ext, int32_t nja, const int32_t *ia_in, const int32_t *ja_in,
                            int32_t index_base, int32_t gpu_ordering, const int32_t *block_id, int32_t *perm) {
  return guard([&] {
    MF6_REQUIRE(ia_in && ja_in && perm && n > 0 && nja >= n && n_ext >= n, "ordering_compute: bad argument");
    std::vector<int> ia(n + 1), ja(nja);
    for (int i = 0; i <= n; i++) ia[i] = ia_in[i] - index_base;
    for (int i = 0; i < nja; i++) ja[i] = ja_in[i] - index_base;
    int ordering = gpu_ordering;
    std::vector<int32_t> derived;
    if (ordering == MF6GPU_ORDER_BLOCK_MULTICOLOR && !block_id) {
      derived = derive_chain_blocks(n, ia, ja);
      if (derived.empty())
        ordering = MF6GPU_ORDER_MULTICOLOR;
      else
        block_id = derived.data();
    }
    ElimOrder E = compute_elimination(n, ia, ja, ordering, block_id);
    for (int v = 0; v < n; v++) perm[E.ordidx[v]] = v;
  });
}

int mf6gpu_matrix_destroy(mf6gpu_matrix *m) {
  return guard([&] { delete m; });
}

int mf6gpu_matrix_update(mf6gpu_matrix *m, const double *amat) {
  return guard([&] {
    MF6_REQUIRE(m && amat, "matrix_update: null argument");
    MF6_CK(cudaMemcpyAsync(m->stage.p, amat, sizeof(double) * (size_t)m->nja,
                           cudaMemcpyHostToDevice, m->stream));
    csr_to_sell_kernel<<<grid_for(m->nja), kBlock, 0, m->stream>>>(m->nja, m->csr2sell.p,
                                                                   m->stage.p, m->val.p);
    MF6_CK(cudaGetLastError());
    MF6_CK(cudaStreamSynchronize(m->stream));
  });
}

int mf6gpu_matrix_zero_entries(mf6gpu_matrix *m) {
  return guard([&] {
    MF6_REQUIRE(m, "matrix_zero_entries: null argument");
    m->val.zero(m->stream);
    MF6_CK(cudaStreamSynchronize(m->stream));
  });
}

int mf6gpu_matrix_get_values(mf6gpu_matrix *m, double *amat) {
  return guard([&] {
    MF6_REQUIRE(m && amat, "matrix_get_values: null argument");
    sell_to_csr_kernel<<<grid_for(m->nja), kBlock, 0, m->stream>>>(m->nja, m->csr2sell.p,
                                                                   m->val.p, m->stage.p);
    MF6_CK(cudaGetLastError());
    m->stage.download(amat, (size_t)m->nja, m->stream);
  });
}

int mf6gpu_matrix_multiply(mf6gpu_matrix *m, const double *x, double *y) {
  return guard([&] {
    MF6_REQUIRE(m && x && y, "matrix_multiply: null argument");
    // stage in original order, permute, multiply, permute back
    mf6::DevBuf<double> tmp;
    tmp.alloc((size_t)m->n);
    MF6_CK(cudaMemcpyAsync(tmp.p, x, sizeof(double) * (size_t)m->n, cudaMemcpyHostToDevice, m->stream));
    launch_gather(m->n, m->d_perm.p, tmp.p, m->xs.p, m->stream);
    launch_spmv(*m, m->val.p, m->xs.p, m->ys.p, m->stream);
    launch_scatter(m->n, m->d_perm.p, m->ys.p, tmp.p, m->stream);
    MF6_CK(cudaGetLastError());
    tmp.download(y, (size_t)m->n, m->stream);
  });
}

int64_t mf6gpu_matrix_info(const mf6gpu_matrix *m, int what) {
  if (!m) return -1;
  switch (what) {
    case 0: return m->n;
    case 1: return m->nja;
    case 2: return m->nlevels;
    case 3: return m->ordering;
    case 4: return m->nslots;
    case 5: return m->maxlen;
    case 6: return m->uniform_w;
    case 7: return (int64_t)(1000.0 * m->slot_off_hit);
  }
  return -1;
}

int mf6gpu_matrix_get_permutation(const mf6gpu_matrix *m, int32_t *perm) {
  return guard([&] {
    MF6_REQUIRE(m && perm, "matrix_get_permutation: null argument");
    std::memcpy(perm, m->elim.data(), sizeof(int) * (size_t)m->n);
  });
}

}  // extern "C"
