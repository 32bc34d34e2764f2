// Device-resident local sparse block (the reference's Matrix_lsr / Matrix_lsc,
// Source/Fortran/SMatrixModule.F90:15-30): CSC, 0-based, row ids ascending
// inside a column. nnz is always known on the host.
#pragma once
#include "device.cuh"

namespace ntb {

template <typename T> struct LocalCsc {
  int rows = 0;
  int cols = 0;
  long long nnz = 0;
  DevBuf<int> outer;  // [cols+1]
  DevBuf<int> inner;  // [nnz]
  DevBuf<T> val;      // [nnz]

  void init_empty(int r, int c) {
    rows = r; cols = c; nnz = 0;
    outer.alloc((size_t)c + 1);
    outer.zero();
    inner.alloc(0);
    val.alloc(0);
  }
  void alloc_entries(long long count) {
    nnz = count;
    inner.alloc((size_t)count);
    val.alloc((size_t)count);
  }
  CscView<T> view() const { return CscView<T>{rows, cols, outer.get(), inner.get(), val.get()}; }
  void copy_from(const LocalCsc<T>& o) {
    rows = o.rows; cols = o.cols; nnz = o.nnz;
    outer.alloc((size_t)cols + 1);
    d2d(outer.get(), o.outer.get(), (size_t)cols + 1);
    inner.alloc((size_t)nnz);
    val.alloc((size_t)nnz);
    d2d(inner.get(), o.inner.get(), (size_t)nnz);
    d2d(val.get(), o.val.get(), (size_t)nnz);
  }
  void swap(LocalCsc<T>& o) {
    std::swap(rows, o.rows); std::swap(cols, o.cols); std::swap(nnz, o.nnz);
    std::swap(outer, o.outer); std::swap(inner, o.inner); std::swap(val, o.val);
  }
  size_t bytes() const {  // algorithmic bytes of this block
    return (size_t)nnz * (sizeof(T) + 4) + ((size_t)cols + 1) * 4;
  }
};

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
};

// Z = alpha * (Y-columns combined by X): for every outer index j of X,
// Z(:,j) = sum_k X(k,j) * Y(:,k), entries kept by the threshold rule, sorted.
// In NTPoly terms (C = A*B in CSC): X = B panel, Y = A panel, Z = C block.
template <typename T>
void spgemm(const CscView<T>& X, const CscView<T>& Y, double alpha, double thr,
            const RuleView& rules, LocalCsc<T>& Z, GemmStats* stats);

// 1: locally dense real products may run on the DMMA tile path (default), 0: scalar kernels only
void set_tile_path(int on);

}  // namespace ntb
