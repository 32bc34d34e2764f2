// Device-resident local sparse block (the reference's Matrix_lsr / Matrix_lsc,
// Source/Fortran/SMatrixModule.F90:15-30): CSC, 0-based, row ids ascending
// inside a column. nnz is always known on the host.
#pragma once
#include "device.cuh"
#include <memory>
#include <vector>
#include <type_traits>

namespace ntb {

// Chunked-tile form of a real local block (spgemm_tile.cu): 8x4 (left operand) or 4x8 (right operand) tiles in
// DMMA fragment order, grouped into 64x32 / 32x64 super-tiles whose present tiles are contiguous in memory.
struct ChunkTiles {
  int ncc = 0;                               // chunk columns: left form ceil(cols/32), right form ceil(cols/64)
  int nsuper = 0;
  long long ntiles = 0;
  DevBuf<int4> colmeta;                      // [ncc]   {first entry, entry count, first id, last id}
  DevBuf<int4> ent;                          // [nsuper] {id, first tile, mask lo, mask hi}; id = row block (left) / inner chunk (right)
  DevBuf<double> tval;                       // [ntiles*32] fragment-ordered values
  DevBuf<int4> kmeta;                        // left form only: per inner tile K {0, tile count, first row tile, last row tile}
  DevBuf<int> coltile;                       // left form only: [ncc+1] first tile of every chunk column (halo exchange)
  bool emitted = false;                      // written by a product (fixed 64-tile slots), not built from CSC
  // a gathered left form (halo exchange) addresses the rank's own tiles in place: tile indices below HALO_TILE_BIAS
  // are relative to tval_view (the own form's tiles), indices from HALO_TILE_BIAS on to `tval` (the received tiles)
  const double* tval_view = nullptr;
};
constexpr long long HALO_TILE_BIAS = 1ll << 30;
// One rank's left form as the other ranks of a column-split grid see it (peer.h): where its arrays lie in the owner's
// peer-visible slab. Exactly one PeerPayload; ntiles / nnz sit in the two words a kernel can patch in.
struct PeerLeftDesc {
  long long off_colmeta, off_ent, off_kmeta, off_tval;   // byte offsets in the owner's slab
  int ncc, nsuper;
  int ok;                                    // the form exists, is worth the tensor cores and lies in the slab
  int right_ok;                              // the same verdict on the publishing rank's RIGHT operand (decision exchange)
  long long ntiles;
  long long nnz;                             // entries of the block (panel fill for the rule table, algorithmic bytes)
};
static_assert(sizeof(PeerLeftDesc) == 64, "PeerLeftDesc is one peer payload");
// The left operand of a tile product as the kernels see it: one piece per rank of the process row (1 on a single
// rank). Chunk column q of the gathered operand is chunk column q % ncc_piece of piece q / ncc_piece; entries address
// tiles relative to their own piece. Passed by value as a kernel parameter.
struct LeftPieceView { const int4* colmeta; const int4* ent; const int4* kmeta; const double* tval; };
struct LeftView {
  int npieces = 1;
  int ncc_piece = 0;
  const double* tval2 = nullptr;             // copied halo tiles (NCCL fallback path): indices from HALO_TILE_BIAS on, piece 0
  LeftPieceView piece[PEER_MAX];
};
// Both forms of one matrix, built on first use as a product operand or emitted together with a product's result.
// Immutable once built, so copies of a matrix share them; any change of the entries drops them.
struct TileForms {
  ChunkTiles left, right;                    // as the A (Y) operand / as the B (X) operand
  int has_left = 0, has_right = 0;           // 0 not built, 1 built, -1 cannot be built (pattern reach overflow)
  // multi-GPU (column-split grids, peer.h): the left form's descriptors of EVERY rank once it has been published
  // (by the product that wrote it, or at its first use as a left operand); the collective verdict on the right forms
  std::vector<PeerLeftDesc> left_pub;
  std::vector<PeerLeftDesc> pub_landing;     // where the exchange of the publishing product lands (valid after a sync)
  bool pub_pending = false;                  // published, descriptors on their way (they arrive with the next stream_sync)
  bool left_needs_barrier = false;           // the tiles were changed in place after publication (ScaleMatrix)
  int right_all_ok = 0;                      // 0 unknown, 1 every rank's right form is usable, -1 not
};

template <typename T> struct LocalCsc;
// fills inner/val of a block whose entries were deferred (spgemm_tile.cu)
void tile_materialize_entries(const LocalCsc<double>& M);

template <typename T> struct LocalCsc {
  int rows = 0;
  int cols = 0;
  LazyCount nnz;      // known on the host, or on its way (deferred tile products: device.cuh LazyCount)
  DevBuf<int> outer;  // [cols+1]
  mutable DevBuf<int> inner;  // [nnz]
  mutable DevBuf<T> val;      // [nnz]
  mutable std::shared_ptr<TileForms> forms;   // cached tile forms of these very entries (real blocks only)
  // A product whose only consumer is another tile product (the drivers' intermediates, e.g. 3I - a^2 X^2 of the
  // sign iteration) is emitted as outer index + right tile form only: inner/val are DEFERRED and filled from the
  // form the first time anything asks for the entries (view(), ensure_entries()). nnz and outer are always valid.
  mutable bool deferred = false;

  void ensure_entries() const {
    if (!deferred) return;
    if constexpr (std::is_same<T, double>::value) tile_materialize_entries(*this);
    deferred = false;
    rt().deferred_materialized++;
  }
  void init_empty(int r, int c) {
    forms.reset();
    deferred = false;
    rows = r; cols = c; nnz = 0;
    outer.alloc((size_t)c + 1);
    outer.zero();
    inner.alloc(0);
    val.alloc(0);
  }
  void alloc_entries(long long count) {
    forms.reset();
    deferred = false;
    nnz = count;
    inner.alloc((size_t)count);
    val.alloc((size_t)count);
  }
  CscView<T> view() const { ensure_entries(); return CscView<T>{rows, cols, outer.get(), inner.get(), val.get()}; }
  void copy_from(const LocalCsc<T>& o) {
    rows = o.rows; cols = o.cols; nnz = o.nnz;
    outer.alloc((size_t)cols + 1);
    d2d(outer.get(), o.outer.get(), (size_t)cols + 1);
    forms = o.forms;
    deferred = o.deferred;                     // a deferred block is copied as such (the forms are shared)
    if (deferred) { inner.alloc(0); val.alloc(0); return; }
    inner.alloc((size_t)nnz);
    val.alloc((size_t)nnz);
    d2d(inner.get(), o.inner.get(), (size_t)nnz);
    d2d(val.get(), o.val.get(), (size_t)nnz);
  }
  void swap(LocalCsc<T>& o) {
    std::swap(rows, o.rows); std::swap(cols, o.cols); std::swap(nnz, o.nnz);
    std::swap(outer, o.outer); std::swap(inner, o.inner); std::swap(val, o.val);
    std::swap(forms, o.forms); std::swap(deferred, o.deferred);
  }
  size_t bytes() const {  // algorithmic bytes of this block
    return (size_t)nnz * (sizeof(T) + 4) + ((size_t)cols + 1) * 4;
  }
};

// algorithmic bytes of one local product: bytes(X) + bytes(Y, unless it is X) + bytes(Z kept), added to
// rt().alg_bytes now or - when a count is still on its way from the device - inside the next stream_sync()
inline void account_product_bytes(const LazyCount& x, int xcols, const LazyCount& y, int ycols, bool count_y,
                                  const LazyCount& z, int zcols, size_t elem_bytes) {
  auto fx = x.later(), fy = y.later(), fz = z.later();
  auto add = [fx, fy, fz, xcols, ycols, zcols, count_y, elem_bytes] {
    auto b = [elem_bytes](long long nnz, int cols) { return (double)nnz * (double)(elem_bytes + 4) + ((double)cols + 1) * 4; };
    rt().alg_bytes += b(fx(), xcols) + b(fz(), zcols) + (count_y ? b(fy(), ycols) : 0.0);
  };
  if (x.pending() || y.pending() || z.pending()) on_next_sync(add);
  else add();
}

// Threshold-rule table of the local product. The reference decides per local
// block pair whether the product runs through its dense branch, which tests
// |v| > thr BEFORE scaling by alpha, or its sparse branch, which tests
// |alpha*v| > thr (sparse_includes/GemmMatrix.f90:59-61, DenseBranch.f90:14-15,
// PruneList.f90:27). tbl[I * nJ + J] != 0 selects the dense rule for the output
// block (inner block I, outer block J); tbl == nullptr means sparse rule everywhere.
struct RuleView {
  const unsigned char* tbl = nullptr;
  int rb = 1;   // rows per inner block
  int cb = 1;   // columns per outer block
  int nJ = 1;
};

struct GemmStats {
  double flops = 0.0;        // 2 * sum over x-entries of len(Y[k]) (real flops; x4 for complex)
  long long tmp_entries = 0; // staging entries written before compaction
  int bins[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool shift_applied = false;  // the DiagShift was fused into the product (tile path only)
};

// Z = alpha * (Y-columns combined by X): for every outer index j of X,
// Z(:,j) = sum_k X(k,j) * Y(:,k), entries kept by the threshold rule, sorted.
// In NTPoly terms (C = A*B in CSC): X = B panel, Y = A panel, Z = C block.
// Optional fused diagonal shift (the drivers' "Gemm then IncrementMatrix(Identity, C, sigma)" with threshold 0,
// reference sparse_includes/AddSparseVectors.f90): after thresholding, sigma is added at local positions
// row == col + dd for col < ncols_diag; a shifted entry is kept iff it is non-zero.
struct DiagShift {
  double sigma = 0.0;
  int dd = 0;            // local row of the diagonal entry of local column 0
  int ncols_diag = 0;    // local columns whose global index is below the actual (unpadded) dimension
  // Optional second fusion (tile path only, one rank or a column-split grid): column sums of |Z - Y| where Y is the
  // LEFT operand of the product (same rows / local columns as Z) - the drivers' "IncrementMatrix(Xnew, X, -1);
  // MatrixNorm(X)" right after X*T (SignSolversModule.F90:230-234) without reading either iterate again: the epilogue of
  // the numeric kernel has the finished strip in registers and fetches the matching tiles of Y's left form (they
  // were this product's A operand a moment ago: L2). diff_colsum: device array [columns of Z]; diff_applied is set
  // by the product when it has filled it (the caller runs the separate norm kernel otherwise).
  double* diff_colsum = nullptr;
  mutable bool diff_applied = false;
};
// What the caller needs of a product (tile path only; every other path delivers plain CSC): the CSC entries and/or
// the tile forms of the result. Without WANT_CSC the entries are deferred (LocalCsc::deferred) and WANT_RIGHT is implied.
constexpr unsigned WANT_CSC = 1u, WANT_LEFT = 2u, WANT_RIGHT = 4u, WANT_ALL = 7u;
template <typename T>
void spgemm(const LocalCsc<T>& X, const LocalCsc<T>& Y, double alpha, double thr,
            const RuleView& rules, LocalCsc<T>& Z, GemmStats* stats, const DiagShift* shift = nullptr,
            unsigned want = WANT_ALL);
// ---- pieces of the tile path used by the distributed layer (spgemm_tile.cu)
// cached or freshly built tile form of a real block (nullptr: the pattern cannot be tiled)
const ChunkTiles* tile_operand_form(const LocalCsc<double>& M, bool left);
// product from tile forms alone: A = left form of the (possibly gathered) left operand, B = right form of the local
// right operand. force: never decline (the caller has already committed collectively to this path).
// publish: on a column-split multi-GPU grid the product ends with a peer exchange that hands the descriptor of the
// left form it has written to every rank (all ranks must pass the same value).
bool spgemm_tile_core(const LeftView& A, const ChunkTiles& B, int ncols, int nrows, double alpha, double thr,
                      const RuleView& rules, LocalCsc<double>& Z, double useful_products, const DiagShift* shift,
                      bool force, unsigned want = WANT_ALL, bool publish = false);
// the one-piece view of a local (or NCCL-gathered) left form
LeftView left_view_of(const ChunkTiles& A);
// descriptor of a rank's own left form for publication (ok = 0 when an array lies outside the peer-visible slab)
PeerLeftDesc left_desc_of(const ChunkTiles* L, long long nnz, bool usable);
// ---- tile-space helpers of the fused driver steps (spgemm_tile.cu, end): scalars and linear combinations straight
// from the right tile forms of iterates whose CSC entries are deferred. All return false when an operand has no
// usable right form (the caller then takes the reference's CSC call sequence).
// mode 0: d_out[0] = sum A.*B;  mode 1 (TRS4, A = X^2, B = X): d_out[0] = sum X2.*Fx, d_out[1] = sum X2.*Gx;
// mode 2: d_out[0] = trace(A) (B = nullptr). Identity / diagonal: local row of column c is c + dd, for c < ncols_diag.
bool tile_form_scalars(int mode, const LocalCsc<double>& A, const LocalCsc<double>* B, int dd, int ncols_diag, double* d_out);
// mode 0: Z = alpha*P + beta*Q with the sparse add's threshold on matched entries;  mode 1 (TRS4, P = X^2, Q = X):
// Z = Fx + sigma*Gx. Z is a tile-space result like a product's (forms per `want`, deferred entries).
// rb: height of the reference's local row blocks (the untested tail of the sparse add is per row block).
bool tile_combine(const LocalCsc<double>& P, const LocalCsc<double>& Q, int mode, double alpha, double beta, double thr,
                  double sigma, int dd, int ncols_diag, int rb, LocalCsc<double>& Z, unsigned want, bool publish);
// column sums of |alpha*A + B| from the right tile forms of both blocks; false when either has none (use the CSC kernel)
bool tile_diff_col_abs_sums(const LocalCsc<double>& A, const LocalCsc<double>& B, double alpha, double* d_colsum);
// assemble the left form of a row of column blocks from per-rank pieces (see psmatrix.cu: halo gather)
struct LeftPiece {          // one rank's contribution, all offsets in units of that rank / of the gathered arrays
  int ent_base;             // first entry of this rank in the gathered entry array
  int ncols_chunk;          // chunk columns per rank
  int a, b;                 // chunk columns [a, b) of this rank were received
  int tile_lo;              // first tile (rank-local numbering) that was received
  int recv_base;            // where that tile sits in the gathered value array
  int nent;                 // entries of this rank
};
void tile_fixup_gathered_left(ChunkTiles& G, const LeftPiece* pieces, int npieces);
// sum over the entries (k,j) of X of ylen[k]  (useful products when the left operand exists only as a tile form)
double useful_products_from_lengths(const LocalCsc<double>& X, const int* d_ylen);
// true when spgemm can apply a DiagShift to this product (real operands on the tile path decide at run time:
// the caller must check GemmStats::shift_applied)

// 1: locally dense real products may run on the DMMA tile path (default), 0: scalar kernels only
void set_tile_path(int on);
bool tile_path_on();

}  // namespace ntb
