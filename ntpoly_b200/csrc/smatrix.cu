// NTPoly's local sparse matrix layer on the GPU: Matrix_lsr / Matrix_lsc as one device-resident CSC block each,
// with the SMatrixAlgebraModule operations routed to the same kernels the distributed layer uses per local block
// (spgemm.cu / spgemm_tile.cu for GemmMatrix, ops.cu for the merge-based helpers).
// Reference: Source/Fortran/SMatrixModule.F90, SMatrixAlgebraModule.F90, sparse_includes/*.f90.
#include "smatrix.h"
#include "ops.cuh"
#include <algorithm>
#include <vector>

namespace ntb {

namespace {
template <typename T> LocalCsc<T>& blk(LocalMatrix& M);
template <> LocalCsc<double>& blk<double>(LocalMatrix& M) { return M.r; }
template <> LocalCsc<cplx>& blk<cplx>(LocalMatrix& M) { return M.c; }
template <typename T> const LocalCsc<T>& blk(const LocalMatrix& M);
template <> const LocalCsc<double>& blk<double>(const LocalMatrix& M) { return M.r; }
template <> const LocalCsc<cplx>& blk<cplx>(const LocalMatrix& M) { return M.c; }

template <typename T> T scalar_from(double v);
template <> double scalar_from<double>(double v) { return v; }
template <> cplx scalar_from<cplx>(double v) { return cplx{v, 0.0}; }

__global__ void __launch_bounds__(256) k_minus_one2(int* __restrict__ a, int* __restrict__ b, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    a[i] -= 1;
    b[i] -= 1;
  }
}
__global__ void __launch_bounds__(256) k_plus_one2(int* __restrict__ a, int* __restrict__ b, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    a[i] += 1;
    b[i] += 1;
  }
}

// position of `row` in column j (row ids ascending), or -1
template <typename T> __device__ __forceinline__ int find_row(const CscView<T>& M, int j, int row) {
  int lo = M.outer[j], hi = M.outer[j + 1];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (M.inner[mid] < row) lo = mid + 1; else hi = mid;
  }
  return (lo < M.outer[j + 1] && M.inner[lo] == row) ? lo : -1;
}
template <typename T>
__global__ void __launch_bounds__(256) k_row_flags(CscView<T> M, int row, int* __restrict__ flags) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < M.cols) flags[j] = find_row(M, j, row) >= 0 ? 1 : 0;
}
// the extracted row is a 1 x cols matrix: one entry (row id 0) in every column that holds `row`
template <typename T>
__global__ void __launch_bounds__(256) k_row_fill(CscView<T> M, int row, const int* __restrict__ outer_out,
                                                  int* __restrict__ inner_out, T* __restrict__ val_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M.cols) return;
  const int p = find_row(M, j, row);
  if (p >= 0) { inner_out[outer_out[j]] = 0; val_out[outer_out[j]] = M.val[p]; }
}

// one warp per column: every value is multiplied by the column's factors in list order
template <typename T>
__global__ void __launch_bounds__(256) k_scale_columns(int cols, const int* __restrict__ outer, T* __restrict__ val,
                                                       const int* __restrict__ foff, const T* __restrict__ fac) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int j = gw; j < cols; j += nw) {
    const int f0 = foff[j], f1 = foff[j + 1];
    if (f0 == f1) continue;
    for (int p = outer[j] + lane; p < outer[j + 1]; p += 32) {
      T v = val[p];
      for (int q = f0; q < f1; ++q) v = s_mul(fac[q], v);
      val[p] = v;
    }
  }
}
}  // namespace

template <typename T> void csc_scale_columns(LocalCsc<T>& M, const int* h_cols0, const T* h_fac, long long n) {
  if (n == 0 || M.nnz == 0 || M.cols == 0) return;
  M.ensure_entries();
  M.forms.reset();
  // bucket the factors by column, keeping the list order inside a column
  std::vector<int> off((size_t)M.cols + 1, 0);
  for (long long i = 0; i < n; ++i) {
    NTB_CHECK(h_cols0[i] >= 0 && h_cols0[i] < M.cols, "MatrixDiagonalScale: column index outside the matrix");
    off[(size_t)h_cols0[i] + 1]++;
  }
  for (int j = 0; j < M.cols; ++j) off[(size_t)j + 1] += off[(size_t)j];
  std::vector<int> cur(off.begin(), off.end() - 1);
  std::vector<T> fac((size_t)n);
  for (long long i = 0; i < n; ++i) fac[(size_t)cur[(size_t)h_cols0[i]]++] = h_fac[i];
  DevBuf<int> d_off((size_t)M.cols + 1);
  DevBuf<T> d_fac((size_t)n);
  h2d(d_off.get(), off.data(), (size_t)M.cols + 1);
  h2d(d_fac.get(), fac.data(), (size_t)n);
  NTB_LAUNCH((k_scale_columns<T>), std::max(1, std::min(div_up((long long)M.cols * 32, 256), kNumSMs * 16)), 256, 0, M.cols,
             M.outer.get(), M.val.get(), d_off.get(), d_fac.get());
  stream_sync();                                 // the host vectors are the copies' sources
}
template void csc_scale_columns<double>(LocalCsc<double>&, const int*, const double*, long long);
template void csc_scale_columns<cplx>(LocalCsc<cplx>&, const int*, const cplx*, long long);

void lmat_construct_zero(LocalMatrix& M, int rows, int cols, bool is_complex) {
  NTB_CHECK(rows >= 0 && cols >= 0, "ConstructZeroMatrix: negative dimension");
  ensure_init();
  M = LocalMatrix();
  M.is_complex = is_complex;
  if (is_complex) M.c.init_empty(rows, cols); else M.r.init_empty(rows, cols);
}

template <typename T>
static void from_triplets_t(LocalMatrix& M, int rows, int cols, const int* h_rows, const int* h_cols, const T* h_vals, long long n) {
  for (long long i = 0; i < n; ++i)
    NTB_CHECK(h_rows[i] >= 1 && h_rows[i] <= rows && h_cols[i] >= 1 && h_cols[i] <= cols,
              "ConstructMatrixFromTripletList: index outside the matrix");
  DevBuf<int> d_row((size_t)n), d_col((size_t)n);
  DevBuf<T> d_val((size_t)n);
  if (n) {
    h2d(d_row.get(), h_rows, (size_t)n); h2d(d_col.get(), h_cols, (size_t)n); h2d(d_val.get(), h_vals, (size_t)n);
    NTB_LAUNCH(k_minus_one2, std::min(div_up(n, 256), kNumSMs * 16), 256, 0, d_row.get(), d_col.get(), n);
  }
  csc_from_device_triplets<T>(rows, cols, d_row.get(), d_col.get(), d_val.get(), n, blk<T>(M));
  stream_sync();
}
void lmat_from_triplets(LocalMatrix& M, int rows, int cols, const int* h_rows, const int* h_cols, const double* h_vals,
                        long long n, bool is_complex) {
  ensure_init();
  M = LocalMatrix();
  M.is_complex = is_complex;
  if (is_complex) from_triplets_t<cplx>(M, rows, cols, h_rows, h_cols, reinterpret_cast<const cplx*>(h_vals), n);
  else from_triplets_t<double>(M, rows, cols, h_rows, h_cols, h_vals, n);
}

template <typename T> static void to_triplets_t(const LocalMatrix& M, int* h_rows, int* h_cols, T* h_vals) {
  const LocalCsc<T>& B = blk<T>(M);
  const long long nnz = B.nnz;
  if (nnz == 0) return;
  const CscView<T> v = B.view();
  DevBuf<int> d_row((size_t)nnz), d_col((size_t)nnz);
  csc_to_device_triplets<T>(v, nnz, d_row.get(), d_col.get());
  NTB_LAUNCH(k_plus_one2, std::min(div_up(nnz, 256), kNumSMs * 16), 256, 0, d_row.get(), d_col.get(), nnz);
  CUDA_CHECK(cudaMemcpyAsync(h_rows, d_row.get(), nnz * sizeof(int), cudaMemcpyDeviceToHost, rt().stream));
  CUDA_CHECK(cudaMemcpyAsync(h_cols, d_col.get(), nnz * sizeof(int), cudaMemcpyDeviceToHost, rt().stream));
  CUDA_CHECK(cudaMemcpyAsync(h_vals, v.val, nnz * sizeof(T), cudaMemcpyDeviceToHost, rt().stream));
  stream_sync();
}
void lmat_to_triplets(const LocalMatrix& M, int* h_rows, int* h_cols, double* h_vals) {
  if (M.is_complex) to_triplets_t<cplx>(M, h_rows, h_cols, reinterpret_cast<cplx*>(h_vals));
  else to_triplets_t<double>(M, h_rows, h_cols, h_vals);
}

void lmat_copy(const LocalMatrix& A, LocalMatrix& B) {
  if (&A == &B) return;
  LocalMatrix res;
  res.is_complex = A.is_complex;
  if (A.is_complex) res.c.copy_from(A.c); else res.r.copy_from(A.r);
  B = std::move(res);
}

template <typename T> static void extract_row_t(const LocalMatrix& M, int row, LocalMatrix& out) {
  const LocalCsc<T>& B = blk<T>(M);
  NTB_CHECK(row >= 0 && row < B.rows, "ExtractMatrixRow: row outside the matrix");
  LocalCsc<T>& O = blk<T>(out);
  O.init_empty(1, B.cols);
  if (B.nnz == 0 || B.cols == 0) return;
  const CscView<T> v = B.view();
  DevBuf<int> flags((size_t)B.cols);
  NTB_LAUNCH((k_row_flags<T>), div_up(B.cols, 256), 256, 0, v, row, flags.get());
  exclusive_scan(flags.get(), O.outer.get(), B.cols);
  int found = 0;
  d2h(&found, O.outer.get() + B.cols, 1);
  O.alloc_entries(found);
  if (found) NTB_LAUNCH((k_row_fill<T>), div_up(B.cols, 256), 256, 0, v, row, O.outer.get(), O.inner.get(), O.val.get());
}
void lmat_extract_row(const LocalMatrix& M, int row_1based, LocalMatrix& out) {
  out = LocalMatrix();
  out.is_complex = M.is_complex;
  if (M.is_complex) extract_row_t<cplx>(M, row_1based - 1, out); else extract_row_t<double>(M, row_1based - 1, out);
}

template <typename T> static void extract_column_t(const LocalMatrix& M, int col, LocalMatrix& out) {
  const LocalCsc<T>& B = blk<T>(M);
  NTB_CHECK(col >= 0 && col < B.cols, "ExtractMatrixColumn: column outside the matrix");
  LocalCsc<T>& O = blk<T>(out);
  O.init_empty(B.rows, 1);
  if (B.nnz == 0) return;
  const CscView<T> v = B.view();
  int se[2] = {0, 0};
  d2h(se, v.outer + col, 2);
  const int cnt = se[1] - se[0];
  O.alloc_entries(cnt);
  const int h_outer[2] = {0, cnt};
  h2d(O.outer.get(), h_outer, 2);
  d2d(O.inner.get(), v.inner + se[0], (size_t)cnt);
  d2d(O.val.get(), v.val + se[0], (size_t)cnt);
  stream_sync();                                 // h_outer is the copy's source
}
void lmat_extract_column(const LocalMatrix& M, int col_1based, LocalMatrix& out) {
  out = LocalMatrix();
  out.is_complex = M.is_complex;
  if (M.is_complex) extract_column_t<cplx>(M, col_1based - 1, out); else extract_column_t<double>(M, col_1based - 1, out);
}

void lmat_scale(LocalMatrix& M, double c) {
  if (M.is_complex) csc_scale<cplx>(M.c, cplx{c, 0.0}); else csc_scale<double>(M.r, c);
}

void lmat_increment(const LocalMatrix& A, LocalMatrix& B, double alpha, double threshold) {
  NTB_CHECK(A.is_complex == B.is_complex, "IncrementMatrix: real and complex local matrices mixed");
  NTB_CHECK(A.rows() == B.rows() && A.cols() == B.cols(), "IncrementMatrix: shape mismatch");
  // one block: the untested tail of the merge is per whole column (sparse_includes/AddSparseVectors.f90:57-68)
  if (A.is_complex) csc_increment<cplx>(A.c.view(), B.c, alpha, threshold, std::max(1, B.rows()));
  else csc_increment<double>(A.r.view(), B.r, alpha, threshold, std::max(1, B.rows()));
}

void lmat_dot(const LocalMatrix& A, const LocalMatrix& B, double* re, double* im) {
  NTB_CHECK(A.is_complex == B.is_complex, "DotMatrix: real and complex local matrices mixed");
  NTB_CHECK(A.rows() == B.rows() && A.cols() == B.cols(), "DotMatrix: shape mismatch");
  double h[2] = {0.0, 0.0};
  if (A.cols() > 0 && A.nnz() > 0 && B.nnz() > 0) {
    DevBuf<double> d(2);
    if (A.is_complex) csc_dot<cplx>(A.c.view(), B.c.view(), d.get());
    else csc_dot<double>(A.r.view(), B.r.view(), d.get());
    d2h(h, d.get(), 2);
  }
  *re = h[0];
  if (im) *im = h[1];
}

void lmat_pairwise(const LocalMatrix& A, const LocalMatrix& B, LocalMatrix& C) {
  NTB_CHECK(A.is_complex == B.is_complex, "PairwiseMultiplyMatrix: real and complex local matrices mixed");
  NTB_CHECK(A.rows() == B.rows() && A.cols() == B.cols(), "PairwiseMultiplyMatrix: shape mismatch");
  LocalMatrix res;
  res.is_complex = A.is_complex;
  if (A.nnz() == 0 || B.nnz() == 0 || A.cols() == 0) {
    if (A.is_complex) res.c.init_empty(A.rows(), A.cols()); else res.r.init_empty(A.rows(), A.cols());
  } else if (A.is_complex) {
    csc_pairwise<cplx>(A.c.view(), B.c.view(), res.c);
  } else {
    csc_pairwise<double>(A.r.view(), B.r.view(), res.r);
  }
  C = std::move(res);
}

void lmat_transpose(const LocalMatrix& A, LocalMatrix& AT) {
  LocalMatrix res;
  res.is_complex = A.is_complex;
  if (A.is_complex) csc_transpose<cplx>(A.c.view(), res.c); else csc_transpose<double>(A.r.view(), res.r);
  AT = std::move(res);
}

void lmat_conjugate(LocalMatrix& M) { if (M.is_complex) csc_conjugate<cplx>(M.c); }

template <typename T>
static void gemm_t(const LocalMatrix& A, const LocalMatrix& B, LocalMatrix& C, bool ta, bool tb, double alpha, double beta,
                   double threshold) {
  const LocalCsc<T>& Al = blk<T>(A);
  const LocalCsc<T>& Bl = blk<T>(B);
  // spgemm computes Z = Y * X from CSC operands: Y = op(A), X = op(B)
  LocalCsc<T> At, Bt;
  const LocalCsc<T>* Y = &Al;
  const LocalCsc<T>* X = &Bl;
  if (ta) { csc_transpose<T>(Al.view(), At); Y = &At; }
  if (tb) { csc_transpose<T>(Bl.view(), Bt); X = &Bt; }
  NTB_CHECK(Y->cols == X->rows, "MatrixMultiply: inner dimensions differ");
  const int c_rows = Y->rows, c_cols = X->cols;
  LocalMatrix AB;
  AB.is_complex = scalar_traits<T>::is_complex;
  LocalCsc<T>& Z = blk<T>(AB);
  if (c_cols == 0 || c_rows == 0 || Y->cols == 0 || Al.nnz == 0 || Bl.nnz == 0) {
    Z.init_empty(c_rows, c_cols);
  } else {
    // dense or sparse threshold rule from the fills of the operands as passed (GemmMatrix.f90:49-61)
    const double fa = (double)Al.nnz / ((double)Al.rows * (double)Al.cols);
    const double fb = (double)Bl.nnz / ((double)Bl.rows * (double)Bl.cols);
    DevBuf<unsigned char> d_rule;
    RuleView rv;
    if (std::min(fa, fb) > 0.1) {
      const unsigned char one = 1;
      d_rule.alloc(1);
      h2d(d_rule.get(), &one, 1);
      stream_sync();
      rv.tbl = d_rule.get(); rv.rb = std::max(1, c_rows); rv.cb = std::max(1, c_cols); rv.nJ = 1;
      rt().dense_rule_blocks++;
    }
    GemmStats st;
    spgemm<T>(*X, *Y, alpha, threshold, rv, Z, &st, nullptr, WANT_ALL);
    rt().flops_useful += st.flops;
    rt().multiplies++;
  }
  // "the add part of GEMM" (GemmMatrix.f90:91-100): beta always arrives through the C interface
  LocalCsc<T>& Cl = blk<T>(C);
  if (std::fabs(beta) > 0.0 && C.is_complex == AB.is_complex && Cl.rows == c_rows && Cl.cols == c_cols) {
    csc_scale<T>(Cl, scalar_from<T>(beta));
    csc_increment<T>(Z.view(), Cl, 1.0, 0.0, std::max(1, c_rows));
  } else {
    NTB_CHECK(!(std::fabs(beta) > 0.0), "MatrixMultiply: beta /= 0 needs a C of the product's shape and type");
    C = std::move(AB);
  }
}

void lmat_gemm(const LocalMatrix& A, const LocalMatrix& B, LocalMatrix& C, bool a_transposed, bool b_transposed,
               double alpha, double beta, double threshold, LocalMemoryPool* pool) {
  NTB_CHECK(A.is_complex == B.is_complex, "MatrixMultiply: real and complex local matrices mixed");
  if (pool) {                                    // CheckMemoryPoolValidity: the pool follows the product's shape
    pool->rows = a_transposed ? A.cols() : A.rows();
    pool->cols = b_transposed ? B.rows() : B.cols();
    pool->is_complex = A.is_complex;
  }
  if (A.is_complex) gemm_t<cplx>(A, B, C, a_transposed, b_transposed, alpha, beta, threshold);
  else gemm_t<double>(A, B, C, a_transposed, b_transposed, alpha, beta, threshold);
}

void lmat_diagonal_scale(LocalMatrix& M, const int* h_cols, const double* h_vals, long long n) {
  std::vector<int> c0((size_t)n);
  for (long long i = 0; i < n; ++i) c0[(size_t)i] = h_cols[i] - 1;
  if (M.is_complex) csc_scale_columns<cplx>(M.c, c0.data(), reinterpret_cast<const cplx*>(h_vals), n);
  else csc_scale_columns<double>(M.r, c0.data(), h_vals, n);
}

}  // namespace ntb
