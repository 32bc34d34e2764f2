// Host-side mirror of NTPoly's LOCAL sparse matrix layer (Matrix_lsr / Matrix_lsc, MatrixMemoryPool_lr / _lc and
// the SMatrixAlgebraModule operations) with every matrix device-resident as one CSC block.
// Reference: Source/Fortran/SMatrixModule.F90:15-30, SMatrixAlgebraModule.F90:116-287,
// MatrixMemoryPoolModule.F90:13-53, Source/C/SMatrix_c.h, Source/Wrapper/SMatrixModule_wrp.F90.
#pragma once
#include "csc.cuh"

namespace ntb {

// Matrix_lsr / Matrix_lsc (SMatrixModule.F90:15-30); one handle type, the element type is a flag
struct LocalMatrix {
  bool is_complex = false;
  LocalCsc<double> r;
  LocalCsc<cplx> c;
  int rows() const { return is_complex ? c.rows : r.rows; }
  int cols() const { return is_complex ? c.cols : r.cols; }
  long long nnz() const { return is_complex ? c.nnz : r.nnz; }
};
// MatrixMemoryPool_lr / _lc (MatrixMemoryPoolModule.F90:13-53): the scratch of the local product lives in the device
// arena; the handle keeps the API and the shape it was made for
struct LocalMemoryPool {
  int rows = 0, cols = 0;
  bool is_complex = false;
};

void lmat_construct_zero(LocalMatrix& M, int rows, int cols, bool is_complex);
// 1-based (row, col) host triplets in any order; duplicates are summed
void lmat_from_triplets(LocalMatrix& M, int rows, int cols, const int* h_rows, const int* h_cols, const double* h_vals,
                        long long n, bool is_complex);
// column-major 1-based triplets of the matrix; vals has 2 doubles per entry for a complex matrix
void lmat_to_triplets(const LocalMatrix& M, int* h_rows, int* h_cols, double* h_vals);
void lmat_copy(const LocalMatrix& A, LocalMatrix& B);
void lmat_extract_row(const LocalMatrix& M, int row_1based, LocalMatrix& out);        // sparse_includes/ExtractMatrixRow.f90
void lmat_extract_column(const LocalMatrix& M, int col_1based, LocalMatrix& out);     // sparse_includes/ExtractMatrixColumn.f90
void lmat_scale(LocalMatrix& M, double c);
void lmat_increment(const LocalMatrix& A, LocalMatrix& B, double alpha, double threshold);   // B <- alpha*A + B
void lmat_dot(const LocalMatrix& A, const LocalMatrix& B, double* re, double* im);          // sum conj(a_ij) b_ij
void lmat_pairwise(const LocalMatrix& A, const LocalMatrix& B, LocalMatrix& C);
// C = alpha*op(A)*op(B) + beta*C, op(X) = X^T when the flag is set (sparse_includes/GemmMatrix.f90)
void lmat_gemm(const LocalMatrix& A, const LocalMatrix& B, LocalMatrix& C, bool a_transposed, bool b_transposed,
               double alpha, double beta, double threshold, LocalMemoryPool* pool);
void lmat_transpose(const LocalMatrix& A, LocalMatrix& AT);
void lmat_conjugate(LocalMatrix& M);
// values of column c multiplied by every listed factor of that column, in list order (sparse_includes/DiagonalScale.f90);
// h_cols 1-based, vals 2 doubles per factor for a complex matrix
void lmat_diagonal_scale(LocalMatrix& M, const int* h_cols, const double* h_vals, long long n);

// the same column scaling on a raw block with 0-based local column ids (used by MatrixDiagonalScale_ps*)
template <typename T> void csc_scale_columns(LocalCsc<T>& M, const int* h_cols0, const T* h_fac, long long n);

}  // namespace ntb
