// Host-side mirror of NTPoly's distributed layer: ProcessGrid_t, Matrix_ps,
// MatrixMemoryPool_p and the PSMatrixAlgebraModule operations, with every
// matrix DEVICE-RESIDENT (one CSC block per GPU) between calls.
// Reference: Source/Fortran/ProcessGridModule.F90, PSMatrixModule.F90,
// PSMatrixAlgebraModule.F90, PMatrixMemoryPoolModule.F90.
#pragma once
#include "comm.h"
#include "csc.cuh"
#include <vector>

namespace ntb {

// ---- ProcessGrid_t (ProcessGridModule.F90:15-56, 130-264) -----------------------
struct ProcessGrid {
  int R = 1, C = 1, S = 1;
  int rank = 0, size = 1, slice_size = 1;
  int my_slice = 0, my_row = 0, my_col = 0;
  int within_slice_rank = 0, between_slice_rank = 0;
  int block_multiplier = 1, nbr = 1, nbc = 1;  // number_of_blocks_rows / columns per rank
  CommHandle* global = nullptr;
  CommHandle* within_slice = nullptr;
  CommHandle* between_slice = nullptr;
  CommHandle* row = nullptr;      // same (slice,row): size C, my rank = my_col
  CommHandle* column = nullptr;   // same (slice,col): size R, my rank = my_row
  bool peer_ok = false;           // column-split grid whose ranks have mapped each other's slabs (peer.h)
  bool constructed = false;
};
void grid_construct(ProcessGrid& g, int rows, int cols, int slices);
void grid_construct_onlyslice(ProcessGrid& g, int slices);
void grid_construct_default(ProcessGrid& g);
void grid_destruct(ProcessGrid& g);
void grid_copy(const ProcessGrid& src, ProcessGrid& dst);
ProcessGrid& global_grid();
void compute_grid_size(int total, int slices, int* rows, int* cols);   // ProcessGridModule.F90:576-601
int compute_num_slices(int total);                                    // ProcessGridModule.F90:606-638

// ---- host triplet lists (TripletModule.F90:13-26; TripletListModule.F90) ---------
struct Triplet_r { int index_column; int index_row; double point_value; };
struct Triplet_c { int index_column; int index_row; double re; double im; };
struct TripletList_r { std::vector<Triplet_r> data; };
struct TripletList_c { std::vector<Triplet_c> data; };

// ---- Matrix_ps (PSMatrixModule.F90:33-51) -----------------------------------------
struct Matrix {
  int actual_dim = 0;
  int logical_dim = 0;
  ProcessGrid* grid = nullptr;
  bool is_complex = false;
  int local_rows = 0, local_cols = 0;
  int start_row = 0, start_col = 0;   // 0-based first global row / column held here
  LocalCsc<double> r;
  LocalCsc<cplx> c;
  bool constructed = false;
  long long local_nnz() const { return is_complex ? c.nnz : r.nnz; }
  int row_block() const { return local_rows / grid->nbr; }
  int col_block() const { return local_cols / grid->nbc; }
};

// MatrixMemoryPool_p (PMatrixMemoryPoolModule.F90:12-49). On the GPU the scratch of the
// multiply lives in the stream-ordered allocator's cache; the handle keeps the API and
// remembers the shape it was built for (CheckMemoryPoolValidity).
struct MemoryPool {
  int rows = 0, cols = 0;
  bool is_complex = false;
  bool constructed = false;
};

int scaled_dimension(const ProcessGrid& g, int n);                      // PSMatrixModule.F90:1596-1618
void mat_construct_empty(Matrix& M, int n, ProcessGrid* grid, bool is_complex);
void mat_construct_like(Matrix& M, const Matrix& ref);
void mat_destruct(Matrix& M);
void mat_copy(const Matrix& A, Matrix& B);
void mat_fill_identity(Matrix& M);
void mat_fill_permutation(Matrix& M, const int* index_lookup_1based, bool permute_rows);
// global 1-based triplets; mirrors FillMatrixFromTripletList (preduplicated / prepartitioned flags)
void mat_fill_from_triplets(Matrix& M, const int* rows, const int* cols, const double* vals_r,
                            const cplx* vals_c, long long n, bool preduplicated, bool prepartitioned);
// local block as global 1-based triplets, column-major order. Call with nullptrs to get the count.
long long mat_get_triplets(const Matrix& M, int* rows, int* cols, double* vals_r, cplx* vals_c);
// the same for a real matrix without waiting: host buffers (pinned) are complete after mat_egress_wait()
long long mat_get_triplets_async(const Matrix& M, int* rows, int* cols, double* vals_r);
void mat_egress_wait();
// Staged ingest (real matrices): stage_triplets only enqueues the host-to-device copies of a global 1-based list on a
// copy stream and returns; mat_fill_from_staged later turns the staged list into the matrix (the library stream waits
// for the copies). The host arrays (pinned memory, or the copies are not asynchronous) must stay valid until then.
struct StagedTriplets {
  DevBuf<int> row, col;
  DevBuf<double> val;
  long long n = 0;
  cudaEvent_t ready = nullptr;
  ~StagedTriplets();
};
void stage_triplets(StagedTriplets& S, const int* rows, const int* cols, const double* vals, long long n);
void mat_fill_from_staged(Matrix& M, StagedTriplets& S);
void mat_transpose(const Matrix& A, Matrix& out);
// out(map[r], map[c]) = A(r, c), exact zeros dropped; map has logical_dim 0-based entries (host)
void mat_relabel(const Matrix& A, Matrix& out, const int* h_map0);
void mat_conjugate(Matrix& M);
void mat_to_complex(const Matrix& in, Matrix& out);
void mat_to_real(const Matrix& in, Matrix& out);
void mat_filter(Matrix& M, double threshold);
// container utilities of PSMatrixModule / MatrixConversionModule that sit either side of the path
void mat_fill_dense(Matrix& M);
void mat_resize(Matrix& M, int new_size);
void mat_get_slice(const Matrix& M, Matrix& sub, int start_row, int end_row, int start_col, int end_col);
long long mat_get_block(const Matrix& M, int start_row, int end_row, int start_col, int end_col, std::vector<int>& rows,
                        std::vector<int>& cols, std::vector<double>& vals_interleaved);
void mat_snap_to_pattern(Matrix& mat, const Matrix& pattern);
long long mat_global_nnz(const Matrix& M);
bool mat_is_identity(const Matrix& M);

// PSMatrixAlgebraModule
// want (csc.cuh: WANT_*): what the caller will read of C. A driver intermediate that only feeds the next product as
// its right operand asks for WANT_RIGHT alone; C is valid for every later use either way (entries are materialized
// on demand).
void mat_multiply(const Matrix& A, const Matrix& B, Matrix& C, double alpha, double beta, double threshold,
                  MemoryPool* pool, unsigned want = WANT_ALL);
void mat_multiply_shift(const Matrix& A, const Matrix& B, Matrix& C, double alpha, double threshold, double sigma,
                        const Matrix& Identity, MemoryPool* pool, unsigned want = WANT_ALL);
void set_fused_shift(int on);
void set_halo_path(int on);
void mat_increment(const Matrix& A, Matrix& B, double alpha, double threshold);
void mat_scale(Matrix& M, double c);
void mat_scale_c(Matrix& M, cplx c);
double mat_trace(const Matrix& M);
double mat_norm(const Matrix& M);
double mat_diff_norm(const Matrix& A, const Matrix& B, double alpha);   // MatrixNorm(alpha*A + B), sum not formed
// C = alpha*A*B and MatrixNorm(C - A) from the product's own epilogue when it runs on the tile path (true: *norm_out set)
bool mat_multiply_diffnorm(const Matrix& A, const Matrix& B, Matrix& C, double alpha, double threshold, MemoryPool* pool,
                           unsigned want, double* norm_out);
void set_fused_norm(int on);
// ---- tile-space helpers of the fused driver steps (csc.cuh: tile_form_scalars / tile_combine) on distributed
// matrices; false (on every rank alike) when the operands do not live as tile forms - the caller then issues the
// reference's own call sequence. mode as in csc.cuh; out receives 1 (modes 0, 2) or 2 (mode 1) reduced scalars.
bool mat_tile_scalars(int mode, const Matrix& A, const Matrix* B, double* out);
bool mat_tile_combine(const Matrix& P, const Matrix& Q, int mode, double alpha, double beta, double thr, double sigma,
                      Matrix& Out, unsigned want);
double mat_sigma(const Matrix& M);
void mat_dot(const Matrix& A, const Matrix& B, double* re, double* im);
void mat_pairwise(const Matrix& A, const Matrix& B, Matrix& C);
void mat_gershgorin(const Matrix& M, double* e_min, double* e_max);
double mat_measure_asymmetry(const Matrix& M);
void mat_symmetrize(Matrix& M);
void mat_similarity_transform(const Matrix& A, const Matrix& P, const Matrix& PInv, Matrix& Res, MemoryPool* pool,
                              double threshold);

}  // namespace ntb
