// extern "C" surface of the LOCAL sparse matrix layer: NTPoly's `*_lsr_wrp` / `*_lsc_wrp` / `*_lr_wrp` / `*_lc_wrp`
// symbols (reference Source/C/SMatrix_c.h, MatrixMemoryPool_c.h; Source/Wrapper/SMatrixModule_wrp.F90,
// SMatrixAlgebraModule_wrp.F90, MatrixMemoryPoolModule_wrp.F90) over device-resident CSC blocks.
#include "c_api_common.h"
#include "psmatrix.h"
#include "smatrix.h"

using namespace ntb;
using namespace ntb::capi;

namespace {
void construct_from_file(int* ih, const char* file_name, int name_size, bool want_complex) {
  const MMData d = read_matrix_market(std::string(file_name, (size_t)name_size), 0, 1);
  const long long n = (long long)d.rows.size();
  auto* M = new LocalMatrix();
  if (want_complex) {
    std::vector<double> v((size_t)n * 2);
    for (long long i = 0; i < n; ++i) { v[2 * i] = d.re[i]; v[2 * i + 1] = d.im[i]; }
    lmat_from_triplets(*M, d.n, d.ncols, d.rows.data(), d.cols.data(), v.data(), n, true);
  } else {
    NTB_CHECK(!d.is_complex, "ConstructMatrixFromFile_lsr: the file holds a complex matrix");
    lmat_from_triplets(*M, d.n, d.ncols, d.rows.data(), d.cols.data(), d.re.data(), n, false);
  }
  put(ih, M);
}

// sparse_includes/PrintMatrix.f90: MatrixMarket coordinate format, column-major entry order
void print_matrix(const LocalMatrix& M, const char* path) {
  const long long nnz = M.nnz();
  std::vector<int> rows((size_t)nnz), cols((size_t)nnz);
  std::vector<double> vals((size_t)nnz * (M.is_complex ? 2 : 1));
  lmat_to_triplets(M, rows.data(), cols.data(), vals.data());
  std::FILE* f = path ? std::fopen(path, "w") : stdout;
  NTB_CHECK(f != nullptr, "PrintMatrix: cannot open the output file");
  std::fprintf(f, "%%%%MatrixMarket matrix coordinate %s general\n%%\n", M.is_complex ? "complex" : "real");
  std::fprintf(f, "%d %d %lld\n", M.rows(), M.cols(), nnz);
  for (long long i = 0; i < nnz; ++i) {
    if (M.is_complex) std::fprintf(f, "%d %d %.17e %.17e\n", rows[i], cols[i], vals[2 * i], vals[2 * i + 1]);
    else std::fprintf(f, "%d %d %.17e\n", rows[i], cols[i], vals[i]);
  }
  if (path) std::fclose(f); else std::fflush(f);
}

void from_triplet_list_r(int* ih, const int* ih_tl, int rows, int cols) {
  const auto& d = get<TripletList_r>(ih_tl)->data;
  std::vector<int> r(d.size()), c(d.size());
  std::vector<double> v(d.size());
  for (size_t i = 0; i < d.size(); ++i) { r[i] = d[i].index_row; c[i] = d[i].index_column; v[i] = d[i].point_value; }
  auto* M = new LocalMatrix();
  lmat_from_triplets(*M, rows, cols, r.data(), c.data(), v.data(), (long long)d.size(), false);
  put(ih, M);
}
void from_triplet_list_c(int* ih, const int* ih_tl, int rows, int cols) {
  const auto& d = get<TripletList_c>(ih_tl)->data;
  std::vector<int> r(d.size()), c(d.size());
  std::vector<double> v(d.size() * 2);
  for (size_t i = 0; i < d.size(); ++i) { r[i] = d[i].index_row; c[i] = d[i].index_column; v[2 * i] = d[i].re; v[2 * i + 1] = d[i].im; }
  auto* M = new LocalMatrix();
  lmat_from_triplets(*M, rows, cols, r.data(), c.data(), v.data(), (long long)d.size(), true);
  put(ih, M);
}
LocalMemoryPool* opt_pool(int* ih) { LocalMemoryPool* p = nullptr; if (ih) std::memcpy(&p, ih, sizeof(p)); return p; }

// TripletListModule.F90 SortTripletList: by column, then by row (stable)
template <typename L> void sort_list(const L& in, L& out) {
  out.data = in.data;
  std::stable_sort(out.data.begin(), out.data.end(), [](const auto& a, const auto& b) {
    return a.index_column != b.index_column ? a.index_column < b.index_column : a.index_row < b.index_row;
  });
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------- real local matrices (Source/C/SMatrix_c.h:3-41)
void ConstructMatrixFromFile_lsr_wrp(int* ih, const char* file_name, const int* name_size) { construct_from_file(ih, file_name, *name_size, false); }
void ConstructMatrixFromTripletList_lsr_wrp(int* ih, const int* ih_tl, const int* rows, const int* columns) { from_triplet_list_r(ih, ih_tl, *rows, *columns); }
void ConstructZeroMatrix_lsr_wrp(int* ih, const int* rows, const int* columns) { auto* M = new LocalMatrix(); lmat_construct_zero(*M, *rows, *columns, false); put(ih, M); }
void DestructMatrix_lsr_wrp(int* ih) { delete get<LocalMatrix>(ih); clear(ih); }
void CopyMatrix_lsr_wrp(const int* ih_a, int* ih_b) { lmat_copy(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_b)); }
void GetMatrixRows_lsr_wrp(const int* ih, int* rows) { *rows = get<LocalMatrix>(ih)->rows(); }
void GetMatrixColumns_lsr_wrp(const int* ih, int* columns) { *columns = get<LocalMatrix>(ih)->cols(); }
// the reference shim ALLOCATEs a fresh object for the output and overwrites the handle (SMatrixModule_wrp.F90:157-189)
void ExtractMatrixRow_lsr_wrp(const int* ih, int* row_number, int* ih_row_out) { auto* o = new LocalMatrix(); lmat_extract_row(*get<LocalMatrix>(ih), *row_number, *o); put(ih_row_out, o); }
void ExtractMatrixColumn_lsr_wrp(const int* ih, int* column_number, int* ih_column_out) { auto* o = new LocalMatrix(); lmat_extract_column(*get<LocalMatrix>(ih), *column_number, *o); put(ih_column_out, o); }
void ScaleMatrix_lsr_wrp(int* ih, const double* constant) { lmat_scale(*get<LocalMatrix>(ih), *constant); }
void IncrementMatrix_lsr_wrp(const int* ih_a, int* ih_b, const double* alpha, const double* threshold) { lmat_increment(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_b), *alpha, *threshold); }
void DotMatrix_lsr_wrp(const int* ih_a, const int* ih_b, double* product) { lmat_dot(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_b), product, nullptr); }
void PairwiseMultiplyMatrix_lsr_wrp(const int* ih_a, const int* ih_b, int* ih_c) { lmat_pairwise(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_b), *get<LocalMatrix>(ih_c)); }
void MatrixMultiply_lsr_wrp(const int* ih_a, const int* ih_b, int* ih_c, const bool* ta, const bool* tb, const double* alpha,
                            const double* beta, const double* threshold, int* ih_pool) {
  lmat_gemm(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_b), *get<LocalMatrix>(ih_c), *ta, *tb, *alpha, *beta, *threshold, opt_pool(ih_pool));
}
void TransposeMatrix_lsr_wrp(const int* ih_a, int* ih_at) { lmat_transpose(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_at)); }
void PrintMatrix_lsr_wrp(const int* ih) { print_matrix(*get<LocalMatrix>(ih), nullptr); }
void PrintMatrixF_lsr_wrp(const int* ih, const char* file_name, const int* name_size) { print_matrix(*get<LocalMatrix>(ih), std::string(file_name, (size_t)*name_size).c_str()); }
void MatrixToTripletList_lsr_wrp(const int* ih, int* ih_tl) {
  const LocalMatrix& M = *get<LocalMatrix>(ih);
  NTB_CHECK(!M.is_complex, "MatrixToTripletList_lsr on a complex matrix");
  const long long n = M.nnz();
  std::vector<int> r((size_t)n), c((size_t)n);
  std::vector<double> v((size_t)n);
  lmat_to_triplets(M, r.data(), c.data(), v.data());
  // the caller's (constructed) list is filled in place. NB the reference shim allocates a list of its own and then
  // drops it (`ih_triplet_list = TRANSFER(ih_triplet_list, ...)`, SMatrixModule_wrp.F90:241-245), so its callers see
  // their list unchanged; the documented behaviour is implemented here instead of that slip.
  auto& out = get<TripletList_r>(ih_tl)->data;
  out.resize((size_t)n);
  for (long long i = 0; i < n; ++i) out[i] = Triplet_r{c[i], r[i], v[i]};
}
void MatrixDiagonalScale_lsr_wrp(int* ih, const int* ih_tl) {
  const auto& d = get<TripletList_r>(ih_tl)->data;
  std::vector<int> c(d.size());
  std::vector<double> v(d.size());
  for (size_t i = 0; i < d.size(); ++i) { c[i] = d[i].index_column; v[i] = d[i].point_value; }
  lmat_diagonal_scale(*get<LocalMatrix>(ih), c.data(), v.data(), (long long)d.size());
}

// ---------------------------------------------------------------- complex local matrices (Source/C/SMatrix_c.h:42-83)
void ConstructMatrixFromFile_lsc_wrp(int* ih, const char* file_name, const int* name_size) { construct_from_file(ih, file_name, *name_size, true); }
void ConstructMatrixFromTripletList_lsc_wrp(int* ih, const int* ih_tl, const int* rows, const int* columns) { from_triplet_list_c(ih, ih_tl, *rows, *columns); }
void ConstructZeroMatrix_lsc_wrp(int* ih, const int* rows, const int* columns) { auto* M = new LocalMatrix(); lmat_construct_zero(*M, *rows, *columns, true); put(ih, M); }
void DestructMatrix_lsc_wrp(int* ih) { delete get<LocalMatrix>(ih); clear(ih); }
void CopyMatrix_lsc_wrp(const int* ih_a, int* ih_b) { lmat_copy(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_b)); }
void GetMatrixRows_lsc_wrp(const int* ih, int* rows) { *rows = get<LocalMatrix>(ih)->rows(); }
void GetMatrixColumns_lsc_wrp(const int* ih, int* columns) { *columns = get<LocalMatrix>(ih)->cols(); }
void ExtractMatrixRow_lsc_wrp(const int* ih, int* row_number, int* ih_row_out) { auto* o = new LocalMatrix(); lmat_extract_row(*get<LocalMatrix>(ih), *row_number, *o); put(ih_row_out, o); }
void ExtractMatrixColumn_lsc_wrp(const int* ih, int* column_number, int* ih_column_out) { auto* o = new LocalMatrix(); lmat_extract_column(*get<LocalMatrix>(ih), *column_number, *o); put(ih_column_out, o); }
void ScaleMatrix_lsc_wrp(int* ih, const double* constant) { lmat_scale(*get<LocalMatrix>(ih), *constant); }
void IncrementMatrix_lsc_wrp(const int* ih_a, int* ih_b, const double* alpha, const double* threshold) { lmat_increment(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_b), *alpha, *threshold); }
void DotMatrix_lsc_wrp(const int* ih_a, const int* ih_b, double* product_real, double* product_complex) { lmat_dot(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_b), product_real, product_complex); }
void PairwiseMultiplyMatrix_lsc_wrp(const int* ih_a, const int* ih_b, int* ih_c) { lmat_pairwise(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_b), *get<LocalMatrix>(ih_c)); }
void MatrixMultiply_lsc_wrp(const int* ih_a, const int* ih_b, int* ih_c, const bool* ta, const bool* tb, const double* alpha,
                            const double* beta, const double* threshold, int* ih_pool) {
  lmat_gemm(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_b), *get<LocalMatrix>(ih_c), *ta, *tb, *alpha, *beta, *threshold, opt_pool(ih_pool));
}
void TransposeMatrix_lsc_wrp(const int* ih_a, int* ih_at) { lmat_transpose(*get<LocalMatrix>(ih_a), *get<LocalMatrix>(ih_at)); }
void ConjugateMatrix_lsc_wrp(int* ih) { lmat_conjugate(*get<LocalMatrix>(ih)); }
void PrintMatrix_lsc_wrp(const int* ih) { print_matrix(*get<LocalMatrix>(ih), nullptr); }
void PrintMatrixF_lsc_wrp(const int* ih, const char* file_name, const int* name_size) { print_matrix(*get<LocalMatrix>(ih), std::string(file_name, (size_t)*name_size).c_str()); }
void MatrixToTripletList_lsc_wrp(const int* ih, int* ih_tl) {
  const LocalMatrix& M = *get<LocalMatrix>(ih);
  NTB_CHECK(M.is_complex, "MatrixToTripletList_lsc on a real matrix");
  const long long n = M.nnz();
  std::vector<int> r((size_t)n), c((size_t)n);
  std::vector<double> v((size_t)n * 2);
  lmat_to_triplets(M, r.data(), c.data(), v.data());
  auto& out = get<TripletList_c>(ih_tl)->data;
  out.resize((size_t)n);
  for (long long i = 0; i < n; ++i) out[i] = Triplet_c{c[i], r[i], v[2 * i], v[2 * i + 1]};
}
void MatrixDiagonalScale_lsc_wrp(int* ih, const int* ih_tl) {
  const auto& d = get<TripletList_c>(ih_tl)->data;
  std::vector<int> c(d.size());
  std::vector<double> v(d.size() * 2);
  for (size_t i = 0; i < d.size(); ++i) { c[i] = d[i].index_column; v[2 * i] = d[i].re; v[2 * i + 1] = d[i].im; }
  lmat_diagonal_scale(*get<LocalMatrix>(ih), c.data(), v.data(), (long long)d.size());
}

// ---------------------------------------------------------------- local memory pools (Source/C/MatrixMemoryPool_c.h)
void ConstructMatrixMemoryPool_lr_wrp(int* ih, const int* columns, const int* rows) { auto* p = new LocalMemoryPool(); p->rows = *rows; p->cols = *columns; put(ih, p); }
void DestructMatrixMemoryPool_lr_wrp(int* ih) { delete get<LocalMemoryPool>(ih); clear(ih); }
void ConstructMatrixMemoryPool_lc_wrp(int* ih, const int* columns, const int* rows) { auto* p = new LocalMemoryPool(); p->rows = *rows; p->cols = *columns; p->is_complex = true; put(ih, p); }
void DestructMatrixMemoryPool_lc_wrp(int* ih) { delete get<LocalMemoryPool>(ih); clear(ih); }

// ---------------------------------------------------------------- triplet list sort
// Signature of the reference's C header and of its C++ caller (Source/C/TripletList_c.h:15-16,
// Source/CPlusPlus/TripletList.cc:88-96): three arguments. (The Fortran shim behind them takes four - columns, rows,
// sorted - TripletListModule_wrp.F90:135-151; the callers of the C ABI pass three, so three it is.) Like the shim, a
// fresh list is allocated for the result and its handle written to ih_sorted.
void SortTripletList_r_wrp(const int* ih, const int*, int* ih_sorted) { auto* t = new TripletList_r(); sort_list(*get<TripletList_r>(ih), *t); put(ih_sorted, t); }
void SortTripletList_c_wrp(const int* ih, const int*, int* ih_sorted) { auto* t = new TripletList_c(); sort_list(*get<TripletList_c>(ih), *t); put(ih_sorted, t); }

// ---------------------------------------------------------------- column scaling of a distributed matrix
// (PSMatrixAlgebraModule.F90:507-532, distributed_algebra_includes/ScaleDiagonal.f90): every rank passes the whole
// list and scales the columns it holds
void MatrixDiagonalScale_psr_wrp(int* ih, const int* ih_tl) {
  Matrix& M = *get<Matrix>(ih);
  NTB_CHECK(!M.is_complex, "MatrixDiagonalScale_psr on a complex matrix");
  const auto& d = get<TripletList_r>(ih_tl)->data;
  std::vector<int> c;
  std::vector<double> v;
  for (const auto& t : d) {
    const int col = t.index_column - 1;
    if (col >= M.start_col && col < M.start_col + M.local_cols) { c.push_back(col - M.start_col); v.push_back(t.point_value); }
  }
  csc_scale_columns<double>(M.r, c.data(), v.data(), (long long)c.size());
}
void MatrixDiagonalScale_psc_wrp(int* ih, const int* ih_tl) {
  Matrix& M = *get<Matrix>(ih);
  NTB_CHECK(M.is_complex, "MatrixDiagonalScale_psc on a real matrix");
  const auto& d = get<TripletList_c>(ih_tl)->data;
  std::vector<int> c;
  std::vector<cplx> v;
  for (const auto& t : d) {
    const int col = t.index_column - 1;
    if (col >= M.start_col && col < M.start_col + M.local_cols) { c.push_back(col - M.start_col); v.push_back(cplx{t.re, t.im}); }
  }
  csc_scale_columns<cplx>(M.c, c.data(), v.data(), (long long)c.size());
}

}  // extern "C"
