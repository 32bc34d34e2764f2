/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
 *
 * CPU restatement of NTPoly's local (per-process) sparse kernels, written
 * from the algorithm description in the reference sources cited per function.
 * This file is a "template": it is included twice by ntpoly_oracle.c, once
 * with real (double) and once with complex (double _Complex) scalars.
 *
 * Layout = the reference's Matrix_lsr/lsc (Source/Fortran/SMatrixModule.F90:15-30):
 * CSC, outer[cols+1] offsets, inner[nnz] row ids ascending inside a column.
 * All indices here are 0-based.
 */

/* ---- transpose: count rows, prefix, scatter ---------------------------------
 * follows Source/Fortran/sparse_includes/TransposeMatrix.f90:19-44 */
void FN(orc_transpose)(int rows, int cols, const int *outer, const int *inner,
                       const SCALAR *val, int *outerT, int *innerT,
                       SCALAR *valT) {
  int nnz = outer[cols];
  int *off = (int *)calloc((size_t)rows + 1, sizeof(int));
  for (int p = 0; p < nnz; ++p) off[inner[p] + 1]++;
  for (int i = 0; i < rows; ++i) off[i + 1] += off[i];
  for (int i = 0; i <= rows; ++i) outerT[i] = off[i];
  for (int j = 0; j < cols; ++j)
    for (int p = outer[j]; p < outer[j + 1]; ++p) {
      int q = off[inner[p]]++;
      innerT[q] = j;
      valT[q] = val[p];
    }
  free(off);
}

/* ---- sparse vector add with the reference's threshold quirks ----------------
 * follows Source/Fortran/sparse_includes/AddSparseVectors.f90:21-70:
 *   matched indices  -> keep iff |alpha*a + b| > thr
 *   unmatched (while both lists still have entries) -> keep iff |.| > thr
 *   remainder after one list is exhausted -> copied WITHOUT any test        */
static int FN(add_vectors)(const int *ia, const SCALAR *va, int na,
                           const int *ib, const SCALAR *vb, int nb, int *ic,
                           SCALAR *vc, double alpha, double thr) {
  int a = 0, b = 0, c = 0;
  while (a < na && b < nb) {
    SCALAR wa = alpha * va[a];
    SCALAR wb = vb[b];
    if (ia[a] == ib[b]) {
      SCALAR s = wa + wb;
      if (ABSF(s) > thr) { ic[c] = ia[a]; vc[c] = s; ++c; }
      ++a; ++b;
    } else if (ia[a] > ib[b]) {
      if (ABSF(wb) > thr) { ic[c] = ib[b]; vc[c] = wb; ++c; }
      ++b;
    } else {
      if (ABSF(wa) > thr) { ic[c] = ia[a]; vc[c] = wa; ++c; }
      ++a;
    }
  }
  while (a < na) { ic[c] = ia[a]; vc[c] = va[a] * alpha; ++a; ++c; }
  while (b < nb) { ic[c] = ib[b]; vc[c] = vb[b]; ++b; ++c; }
  return c;
}

/* ---- B <- alpha*A + B, column by column -------------------------------------
 * follows Source/Fortran/sparse_includes/IncrementMatrix.f90:36-65.
 * Output arrays must have room for nnz(A)+nnz(B). Returns nnz(C).           */
int FN(orc_increment)(int cols, const int *oa, const int *ia, const SCALAR *va,
                      const int *ob, const int *ib, const SCALAR *vb, int *oc,
                      int *ic, SCALAR *vc, double alpha, double thr) {
  int total = 0;
  oc[0] = 0;
  for (int j = 0; j < cols; ++j) {
    int n = FN(add_vectors)(ia + oa[j], va + oa[j], oa[j + 1] - oa[j],
                            ib + ob[j], vb + ob[j], ob[j + 1] - ob[j],
                            ic + total, vc + total, alpha, thr);
    total += n;
    oc[j + 1] = total;
  }
  return total;
}

/* ---- Hadamard product (pattern intersection) --------------------------------
 * follows Source/Fortran/sparse_includes/PairwiseMultiplyVectors.f90 and
 * PairwiseMultiplyMatrix.f90. Output room: min(nnzA, nnzB).                  */
int FN(orc_pairwise)(int cols, const int *oa, const int *ia, const SCALAR *va,
                     const int *ob, const int *ib, const SCALAR *vb, int *oc,
                     int *ic, SCALAR *vc) {
  int total = 0;
  oc[0] = 0;
  for (int j = 0; j < cols; ++j) {
    int a = oa[j], b = ob[j];
    while (a < oa[j + 1] && b < ob[j + 1]) {
      if (ia[a] == ib[b]) {
        ic[total] = ia[a];
        vc[total] = va[a] * vb[b];
        ++total; ++a; ++b;
      } else if (ia[a] > ib[b]) ++b;
      else ++a;
    }
    oc[j + 1] = total;
  }
  return total;
}

/* ---- local GEMM -------------------------------------------------------------
 * C = alpha * op(A) * op(B) with drop threshold, operands handed over in the
 * form the distributed multiply uses them (already "transposed"):
 *   AT : CSC of op(A)^T  -> column i of AT lists (k, a_ik) of row i of op(A)
 *   BT : CSC of op(B)^T  -> column k of BT lists (j, b_kj) of row k of op(B)
 * C has c_rows = AT.cols, c_cols = BT.rows; returned as CSC (malloc'd).
 *
 * follows Source/Fortran/sparse_includes/GemmMatrix.f90:47-61 (branch rule),
 * MultiplyBlock.f90:9-36 (Gustavson into a dense accumulator with dirty flags
 * and per-bucket index lists), PruneList.f90:8-38 (strict |alpha*v| > thr,
 * emit, sort, build CSC), DenseBranch.f90:1-18 + ConstructMatrixSFromD.f90
 * (dense product, |v| > thr tested BEFORE scaling by alpha).
 *
 * pool_mode: 0 = one accumulator column per worker thread (same arithmetic,
 *                bounded memory; used when rows*cols*36 B would not fit),
 *            1 = the reference's full (c_cols x c_rows) pool.
 * force_branch: 0 = reference rule, 1 = always sparse, 2 = always dense.
 * Returns nnz(C); *branch_taken = 1 sparse / 2 dense.                        */
typedef struct { int row; int col; SCALAR v; } FN(trip_t);

long FN(orc_gemm)(int at_rows, int at_cols, const int *at_outer,
                  const int *at_inner, const SCALAR *at_val, int bt_rows,
                  int bt_cols, const int *bt_outer, const int *bt_inner,
                  const SCALAR *bt_val, double alpha, double thr, int pool_mode,
                  int force_branch, int **c_outer_out, int **c_inner_out,
                  SCALAR **c_val_out, int *branch_taken) {
  const int c_rows = at_cols;
  const int c_cols = bt_rows;
  const int inner_dim = at_rows; /* == bt_cols */
  (void)inner_dim;
  double nnz_a = at_outer[at_cols], nnz_b = bt_outer[bt_cols];
  double sp_a = nnz_a / ((double)at_rows * (double)at_cols);
  double sp_b = nnz_b / ((double)bt_rows * (double)bt_cols);
  double est = 4.0 * (sp_a > sp_b ? sp_a : sp_b);
  if (est > 1.0) est = 1.0; else if (est < 1e-8) est = 1e-8;
  int dense = (sp_a < sp_b ? sp_a : sp_b) > 0.1;
  if (force_branch == 1) dense = 0;
  if (force_branch == 2) dense = 1;
  *branch_taken = dense ? 2 : 1;

  int *c_outer = (int *)calloc((size_t)c_cols + 1, sizeof(int));

  if (dense) {
    /* densify both (ConstructMatrixDFromS), multiply, re-sparsify */
    size_t sa = (size_t)c_rows * at_rows, sb = (size_t)at_rows * c_cols;
    SCALAR *DA = (SCALAR *)calloc(sa ? sa : 1, sizeof(SCALAR)); /* op(A): c_rows x K, row-major */
    SCALAR *DB = (SCALAR *)calloc(sb ? sb : 1, sizeof(SCALAR)); /* op(B): K x c_cols, row-major */
    for (int i = 0; i < at_cols; ++i)
      for (int p = at_outer[i]; p < at_outer[i + 1]; ++p)
        DA[(size_t)i * at_rows + at_inner[p]] = at_val[p];
    for (int k = 0; k < bt_cols; ++k)
      for (int p = bt_outer[k]; p < bt_outer[k + 1]; ++p)
        DB[(size_t)k * c_cols + bt_inner[p]] = bt_val[p];
    SCALAR *DC = (SCALAR *)calloc((size_t)c_rows * c_cols + 1, sizeof(SCALAR));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < c_rows; ++i) {
      SCALAR *ci = DC + (size_t)i * c_cols;
      for (int k = 0; k < at_rows; ++k) {
        SCALAR a = DA[(size_t)i * at_rows + k];
        if (a == 0) continue; /* skipping exact zeros does not change any sum */
        const SCALAR *bk = DB + (size_t)k * c_cols;
        for (int j = 0; j < c_cols; ++j) ci[j] += a * bk[j];
      }
    }
    long nnz = 0;
    for (int j = 0; j < c_cols; ++j) {
      for (int i = 0; i < c_rows; ++i)
        if (ABSF(DC[(size_t)i * c_cols + j]) > thr) ++nnz;
      c_outer[j + 1] = (int)nnz;
    }
    int *ci = (int *)malloc((nnz ? nnz : 1) * sizeof(int));
    SCALAR *cv = (SCALAR *)malloc((nnz ? nnz : 1) * sizeof(SCALAR));
    long q = 0;
    for (int j = 0; j < c_cols; ++j)
      for (int i = 0; i < c_rows; ++i) {
        SCALAR v = DC[(size_t)i * c_cols + j];
        if (ABSF(v) > thr) { ci[q] = i; cv[q] = alpha * v; ++q; }
      }
    free(DA); free(DB); free(DC);
    *c_outer_out = c_outer; *c_inner_out = ci; *c_val_out = cv;
    return nnz;
  }

  /* ---- sparse branch ---- */
  int hash_size = (int)(1.0 / est);
  if (hash_size > c_cols) hash_size = c_cols;
  if (hash_size < 1) hash_size = 1;
  int nbuckets = c_cols > 0 ? (c_cols - 1) / hash_size + 1 : 0;

  /* per-row emitted lists; rows are independent so they can run in parallel,
   * the emit order inside a row is the reference's bucket order. */
  long *row_cnt = (long *)calloc((size_t)c_rows + 1, sizeof(long));
  FN(trip_t) **row_list = (FN(trip_t) **)calloc((size_t)c_rows + 1, sizeof(void *));

  SCALAR *full_val = NULL; char *full_dirty = NULL; int *full_hidx = NULL, *full_ins = NULL;
  if (pool_mode == 1) {
    size_t e = (size_t)c_cols * c_rows;
    full_val = (SCALAR *)calloc(e ? e : 1, sizeof(SCALAR));
    full_dirty = (char *)calloc(e ? e : 1, 4); /* Fortran LOGICAL = 4 bytes */
    full_hidx = (int *)calloc(e ? e : 1, sizeof(int));
    full_ins = (int *)calloc(e ? e : 1, sizeof(int));
  }
#pragma omp parallel
  {
    SCALAR *acc = NULL; int *dirty = NULL, *hidx = NULL, *ins = NULL;
    if (pool_mode != 1) {
      acc = (SCALAR *)calloc((size_t)c_cols + 1, sizeof(SCALAR));
      dirty = (int *)calloc((size_t)c_cols + 1, sizeof(int));
      hidx = (int *)calloc((size_t)c_cols + 1, sizeof(int));
      ins = (int *)calloc((size_t)nbuckets + 1, sizeof(int));
    }
#pragma omp for schedule(dynamic, 16)
    for (int i = 0; i < c_rows; ++i) {
      if (pool_mode == 1) {
        acc = full_val + (size_t)i * c_cols;
        dirty = (int *)(full_dirty + (size_t)i * c_cols * 4);
        hidx = full_hidx + (size_t)i * c_cols;
        ins = full_ins + (size_t)i * c_cols;
      }
      long touched = 0;
      for (int pa = at_outer[i]; pa < at_outer[i + 1]; ++pa) {
        SCALAR a = at_val[pa];
        int k = at_inner[pa];
        for (int pb = bt_outer[k]; pb < bt_outer[k + 1]; ++pb) {
          int j = bt_inner[pb];
          SCALAR cur = acc[j];
          if (!dirty[j]) {
            dirty[j] = 1;
            int h = j / hash_size;
            hidx[ins[h] + h * hash_size] = j;
            ins[h]++;
            ++touched;
          }
          acc[j] = cur + a * bt_val[pb];
        }
      }
      FN(trip_t) *lst = (FN(trip_t) *)malloc((touched ? touched : 1) * sizeof(FN(trip_t)));
      long n = 0;
      for (int h = 0; h < nbuckets; ++h) {
        int cnt = ins[h];
        ins[h] = 0;
        for (int t = 0; t < cnt; ++t) {
          int j = hidx[t + h * hash_size];
          SCALAR v = acc[j];
          acc[j] = 0;
          dirty[j] = 0;
          if (ABSF(alpha * v) > thr) { lst[n].row = i; lst[n].col = j; lst[n].v = alpha * v; ++n; }
        }
      }
      row_list[i] = lst;
      row_cnt[i] = n;
    }
    if (pool_mode != 1) { free(acc); free(dirty); free(hidx); free(ins); }
  }
  if (pool_mode == 1) { free(full_val); free(full_dirty); free(full_hidx); free(full_ins); }

  /* SortTripletList (bucket by column, stable) + ConstructMatrixFromTripletList */
  long nnz = 0;
  for (int i = 0; i < c_rows; ++i) {
    nnz += row_cnt[i];
    for (long t = 0; t < row_cnt[i]; ++t) c_outer[row_list[i][t].col + 1]++;
  }
  for (int j = 0; j < c_cols; ++j) c_outer[j + 1] += c_outer[j];
  int *ci = (int *)malloc((nnz ? nnz : 1) * sizeof(int));
  SCALAR *cv = (SCALAR *)malloc((nnz ? nnz : 1) * sizeof(SCALAR));
  int *fill = (int *)malloc(((size_t)c_cols + 1) * sizeof(int));
  memcpy(fill, c_outer, ((size_t)c_cols + 1) * sizeof(int));
  for (int i = 0; i < c_rows; ++i) { /* rows ascending => row ids ascending per column */
    for (long t = 0; t < row_cnt[i]; ++t) {
      int q = fill[row_list[i][t].col]++;
      ci[q] = i;
      cv[q] = row_list[i][t].v;
    }
    free(row_list[i]);
  }
  free(fill); free(row_list); free(row_cnt);
  *c_outer_out = c_outer; *c_inner_out = ci; *c_val_out = cv;
  return nnz;
}
