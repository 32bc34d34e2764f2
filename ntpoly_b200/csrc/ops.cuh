// O(nnz) helpers every solver calls per iteration, on device-resident CSC blocks.
#pragma once
#include "csc.cuh"

namespace ntb {

// B <- alpha*A + B with NTPoly's threshold rules (AddSparseVectors.f90:21-70);
// rb = height of the reference's local row blocks (tail rule is per column segment).
template <typename T> void csc_increment(const CscView<T>& A, LocalCsc<T>& B, double alpha, double thr, int rb);
// C = A .* B (pattern intersection)
template <typename T> void csc_pairwise(const CscView<T>& A, const CscView<T>& B, LocalCsc<T>& C);
// values *= c
template <typename T> void csc_scale(LocalCsc<T>& M, T c);
template <typename T> void csc_conjugate(LocalCsc<T>& M);
// keep |v| > thr
template <typename T> void csc_filter(LocalCsc<T>& M, double thr);
// out = transpose(M) (CSC of M^T)
template <typename T> void csc_transpose(const CscView<T>& M, LocalCsc<T>& out);
void csc_to_complex(const LocalCsc<double>& in, LocalCsc<cplx>& out);
void csc_to_real(const LocalCsc<cplx>& in, LocalCsc<double>& out);

// select the outer ranges / inner rows of the blocks congruent to `s` mod S
// (slice-restricted panels, MatrixMultiply.f90:75-80,98-105,158-164)
template <typename T> void csc_select_col_blocks(const CscView<T>& M, int cb, int S, int s, LocalCsc<T>& out);
template <typename T> void csc_select_row_blocks(const CscView<T>& M, int rb, int S, int s, LocalCsc<T>& out);
// stack `n` blocks with identical column count on top of each other (row offsets given)
template <typename T> void csc_stack_rows(const CscView<T>* parts, const int* row_offsets, int n, int total_rows,
                                          LocalCsc<T>& out);

// ---- scalar reductions; results are written to device doubles -----------------
// d_out[0] = sum_j Re(M(grow(j), j)) for the local block starting at (start_row, start_col) (0-based)
template <typename T> void csc_trace(const CscView<T>& M, int start_row, int start_col, double* d_out);
// d_colsum[j] = sum_i |M(i,j)|
template <typename T> void csc_col_abs_sums(const CscView<T>& M, double* d_colsum);
template <typename T> void csc_diff_col_abs_sums(const CscView<T>& A, const CscView<T>& B, double alpha, double* d_colsum);
// d_min[j] / d_max[j] Gershgorin column contributions (solver_includes/GershgorinBounds.f90)
template <typename T> void csc_gershgorin_cols(const CscView<T>& M, int start_row, int start_col, double* d_min, double* d_max);
// d_out[0..1] = sum conj(a_ij) * b_ij  (re, im)
template <typename T> void csc_dot(const CscView<T>& A, const CscView<T>& B, double* d_out2);
// d_out[0]=ok(1/0: all entries on the diagonal & equal 1), d_out[1]=count of unit diagonal entries
template <typename T> void csc_identity_check(const CscView<T>& M, int start_row, int start_col, double* d_out2);

void reduce_max(const double* d_in, int n, double* d_out);
void reduce_min(const double* d_in, int n, double* d_out);
void reduce_sum(const double* d_in, int n, double* d_out);

// ---- ingest / egress (host triplets <-> device CSC) ---------------------------
// rows/cols are LOCAL 0-based indices; duplicates are summed. d_* are device arrays.
template <typename T> void csc_from_device_triplets(int rows, int cols, const int* d_row, const int* d_col,
                                                    const T* d_val, long long n, LocalCsc<T>& out);
template <typename T> void csc_to_device_triplets(const CscView<T>& M, long long nnz, int* d_row, int* d_col);

}  // namespace ntb
