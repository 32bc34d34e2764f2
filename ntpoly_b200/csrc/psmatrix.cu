#include "psmatrix.h"
#include "peer.h"
#include <chrono>
#include <functional>
#include "ops.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace ntb {

// ===========================================================================
// process grid
// ===========================================================================
static ProcessGrid g_global_grid;
ProcessGrid& global_grid() { return g_global_grid; }

void compute_grid_size(int total, int slices, int* rows, int* cols) {
  *rows = 1; *cols = 1;
  const int slice_size = total / slices;
  for (int ii = (int)std::floor(std::sqrt((float)slice_size)); ii >= 1; --ii) {
    if (slice_size % ii == 0) { *rows = ii; *cols = slice_size / ii; break; }
  }
}

int compute_num_slices(int total) {
  for (int slices = std::min(4, total); slices >= 2; --slices) {
    const int slice_size = total / slices;
    if (slice_size * slices != total) continue;
    int d = (int)std::floor(std::sqrt((float)slice_size));
    if (d * d == slice_size) return slices;
    d = (int)std::floor(std::sqrt((float)(slice_size / 2)));
    if (d * d * 2 == slice_size) return slices;
  }
  return 1;
}

void grid_destruct(ProcessGrid& g) {
  if (!g.constructed) return;
  comm_free(g.row); comm_free(g.column); comm_free(g.within_slice); comm_free(g.between_slice);
  // g.global aliases the world communicator: not freed here
  g = ProcessGrid();
}

void grid_construct(ProcessGrid& g, int rows, int cols, int slices) {
  world_init_from_env();
  grid_destruct(g);
  World& w = world();
  g.R = rows; g.C = cols; g.S = slices;
  g.size = w.size; g.rank = w.rank;
  NTB_CHECK(rows * cols * slices == g.size, "you did not specify a consistent process grid size");
  if (slices > 1)
    NTB_CHECK(std::max(rows, cols) % std::min(rows, cols) == 0,
              "if slices >1, either rows or columns must be a multiple of the other.");
  g.slice_size = g.size / slices;
  g.my_slice = g.rank / g.slice_size;                       // ProcessGridModule.F90:180-183
  g.my_row = (g.rank % g.slice_size) / cols;
  g.my_col = g.rank % cols;
  g.within_slice_rank = g.rank % g.slice_size;
  g.between_slice_rank = g.my_slice;
  // blocking (ProcessGridModule.F90:200-235) with one host thread driving the GPU
  int cbm = (rows / cols) * slices; if (cbm == 0) cbm = slices;
  int rbm = (cols / rows) * slices; if (rbm == 0) rbm = slices;
  g.block_multiplier = 1;
  g.nbc = cbm; g.nbr = rbm;
  g.global = w.comm;
  if (w.comm) {
    g.within_slice = comm_split(w.comm, g.my_slice, g.rank);
    comm_set_shape(g.within_slice, g.slice_size, g.within_slice_rank);
    g.between_slice = comm_split(w.comm, g.within_slice_rank, g.rank);
    comm_set_shape(g.between_slice, slices, g.my_slice);
    g.row = comm_split(w.comm, g.my_slice * rows + g.my_row, g.rank);
    comm_set_shape(g.row, cols, g.my_col);
    g.column = comm_split(w.comm, g.my_slice * cols + g.my_col, g.rank);
    comm_set_shape(g.column, rows, g.my_row);
    // column-split grids read the other ranks' operand tiles in place over NVLink (peer.h); collective, once per process
    if (rows == 1 && slices == 1 && cols == w.size) g.peer_ok = peer_setup(w.comm);
  }
  g.constructed = true;
}

void grid_construct_onlyslice(ProcessGrid& g, int slices) {
  world_init_from_env();
  int r, c;
  compute_grid_size(world().size, slices, &r, &c);
  grid_construct(g, r, c, slices);
}
void grid_construct_default(ProcessGrid& g) {
  world_init_from_env();
  grid_construct_onlyslice(g, compute_num_slices(world().size));
}
void grid_copy(const ProcessGrid& src, ProcessGrid& dst) { grid_construct(dst, src.R, src.C, src.S); }

// ===========================================================================
// construction
// ===========================================================================
int scaled_dimension(const ProcessGrid& g, int n) {
  const int lcm = g.block_multiplier * g.S * g.C * g.R;
  const int q = n / lcm;
  return (q * lcm == n) ? n : (q + 1) * lcm;
}

void mat_destruct(Matrix& M) {
  M.r = LocalCsc<double>();
  M.c = LocalCsc<cplx>();
  M.constructed = false;
}

void mat_construct_empty(Matrix& M, int n, ProcessGrid* grid, bool is_complex) {
  mat_destruct(M);
  ensure_init();
  if (!grid) grid = &global_grid();
  if (!grid->constructed) grid_construct_default(*grid);
  M.grid = grid;
  M.is_complex = is_complex;
  M.actual_dim = n;
  M.logical_dim = scaled_dimension(*grid, n);
  M.local_rows = M.logical_dim / grid->R;
  M.local_cols = M.logical_dim / grid->C;
  M.start_row = M.local_rows * grid->my_row;
  M.start_col = M.local_cols * grid->my_col;
  if (is_complex) M.c.init_empty(M.local_rows, M.local_cols);
  else M.r.init_empty(M.local_rows, M.local_cols);
  M.constructed = true;
}

void mat_construct_like(Matrix& M, const Matrix& ref) {
  mat_construct_empty(M, ref.actual_dim, ref.grid, ref.is_complex);
}

void mat_copy(const Matrix& A, Matrix& B) {
  if (&A == &B) return;
  B.actual_dim = A.actual_dim; B.logical_dim = A.logical_dim; B.grid = A.grid; B.is_complex = A.is_complex;
  B.local_rows = A.local_rows; B.local_cols = A.local_cols; B.start_row = A.start_row; B.start_col = A.start_col;
  if (A.is_complex) { B.c.copy_from(A.c); B.r = LocalCsc<double>(); }
  else { B.r.copy_from(A.r); B.c = LocalCsc<cplx>(); }
  B.constructed = true;
}

void mat_to_complex(const Matrix& in, Matrix& out) {
  if (in.is_complex) { mat_copy(in, out); return; }
  Matrix res;
  mat_construct_empty(res, in.actual_dim, in.grid, true);
  csc_to_complex(in.r, res.c);
  out = std::move(res);
}
void mat_to_real(const Matrix& in, Matrix& out) {
  if (!in.is_complex) { mat_copy(in, out); return; }
  Matrix res;
  mat_construct_empty(res, in.actual_dim, in.grid, false);
  csc_to_real(in.c, res.r);
  out = std::move(res);
}

// ---------------------------------------------------------------------------
// gathers of variable-size CSC blocks over a communicator
// ---------------------------------------------------------------------------
static std::vector<long long> exchange_counts(CommHandle* comm, long long mine) {
  const int n = comm_size(comm);
  std::vector<long long> all(n, mine);
  if (n == 1) return all;
  DevBuf<long long> d_mine(1), d_all((size_t)n);
  h2d(d_mine.get(), &mine, 1);
  comm_allgather_bytes(comm, d_mine.get(), d_all.get(), sizeof(long long));
  d2h(all.data(), d_all.get(), (size_t)n);
  return all;
}

// every rank of `comm` contributes one block of identical shape; result = the blocks
// as separate CSC objects (own block is NOT copied: parts[me] stays empty, views[me] aliases local)
template <typename T>
static void gather_parts(const LocalCsc<T>& local, CommHandle* comm, std::vector<LocalCsc<T>>& parts,
                         std::vector<CscView<T>>& views) {
  const int n = comm_size(comm), me = comm_rank(comm);
  parts.clear(); parts.resize(n);
  views.assign(n, local.view());
  if (n == 1) return;
  std::vector<long long> nnz = exchange_counts(comm, local.nnz);
  for (int q = 0; q < n; ++q) {
    if (q == me) continue;
    parts[q].rows = local.rows; parts[q].cols = local.cols;
    parts[q].outer.alloc((size_t)local.cols + 1);
    parts[q].alloc_entries(nnz[q]);
    views[q] = parts[q].view();
  }
  comm_group_start();
  for (int q = 0; q < n; ++q) {
    const LocalCsc<T>& dst = (q == me) ? local : parts[q];
    comm_broadcast_bytes(comm, local.outer.get(), dst.outer.get(), ((size_t)local.cols + 1) * sizeof(int), q);
    comm_broadcast_bytes(comm, local.inner.get(), dst.inner.get(), (size_t)nnz[q] * sizeof(int), q);
    comm_broadcast_bytes(comm, local.val.get(), dst.val.get(), (size_t)nnz[q] * sizeof(T), q);
  }
  comm_group_end();
}

__global__ void __launch_bounds__(256) k_concat_outer(const int* __restrict__ all_outer, int cols, int nparts,
                                                      const long long* __restrict__ nnz_off, int* __restrict__ out_outer) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)cols * nparts;
  if (i < total) {
    const int q = (int)(i / cols), j = (int)(i - (long long)q * cols);
    out_outer[i] = all_outer[(size_t)q * (cols + 1) + j] + (int)nnz_off[q];
  } else if (i == total) {
    out_outer[i] = (int)nnz_off[nparts];
  }
}

// blocks side by side (ComposeMatrixColumns over the row communicator): entries land
// directly at their final offset, only the outer index needs a fix-up pass
// (reference comm_includes/ReduceAndComposeMatrixCleanup.f90:6-13).
template <typename T>
static void gather_concat_cols(const LocalCsc<T>& local, CommHandle* comm, LocalCsc<T>& out) {
  const int n = comm_size(comm);
  NTB_CHECK(n > 1, "gather_concat_cols on a single rank");
  local.ensure_entries();
  std::vector<long long> nnz = exchange_counts(comm, local.nnz);
  std::vector<long long> off(n + 1, 0);
  for (int q = 0; q < n; ++q) off[q + 1] = off[q] + nnz[q];
  NTB_CHECK(off[n] < (1ll << 31), "gathered panel exceeds 2^31 entries");
  out.rows = local.rows; out.cols = local.cols * n;
  out.outer.alloc((size_t)out.cols + 1);
  out.alloc_entries(off[n]);
  DevBuf<int> all_outer((size_t)n * (local.cols + 1));
  comm_allgather_bytes(comm, local.outer.get(), all_outer.get(), ((size_t)local.cols + 1) * sizeof(int));
  comm_group_start();
  for (int q = 0; q < n; ++q) {
    comm_broadcast_bytes(comm, local.inner.get(), out.inner.get() + off[q], (size_t)nnz[q] * sizeof(int), q);
    comm_broadcast_bytes(comm, local.val.get(), out.val.get() + off[q], (size_t)nnz[q] * sizeof(T), q);
  }
  comm_group_end();
  DevBuf<long long> d_off((size_t)n + 1);
  h2d(d_off.get(), off.data(), (size_t)n + 1);
  const long long total = (long long)local.cols * n + 1;
  NTB_LAUNCH(k_concat_outer, div_up(total, 256), 256, 0, all_outer.get(), local.cols, n, d_off.get(), out.outer.get());
  stream_sync();  // off (host vector) was the h2d source
}

// between-slice "gather and sum" (comm_includes/ReduceAndSumMatrixCleanup.f90:11-32)
template <typename T>
static void reduce_and_sum(LocalCsc<T>& block, CommHandle* comm, double threshold, int rb) {
  const int n = comm_size(comm);
  if (n == 1) return;
  std::vector<LocalCsc<T>> parts;
  std::vector<CscView<T>> views;
  gather_parts(block, comm, parts, views);
  LocalCsc<T> sum;
  sum.init_empty(block.rows, block.cols);
  for (int q = 0; q < n; ++q) csc_increment<T>(views[q], sum, 1.0, (q == n - 1) ? threshold : 0.0, rb);
  block.swap(sum);
}

template <typename T> static LocalCsc<T>& loc(Matrix& M);
template <> LocalCsc<double>& loc<double>(Matrix& M) { return M.r; }
template <> LocalCsc<cplx>& loc<cplx>(Matrix& M) { return M.c; }
template <typename T> static const LocalCsc<T>& loc(const Matrix& M);
template <> const LocalCsc<double>& loc<double>(const Matrix& M) { return M.r; }
template <> const LocalCsc<cplx>& loc<cplx>(const Matrix& M) { return M.c; }

// ===========================================================================
// ingest / egress
// ===========================================================================
// Bulk host<->device copies are issued in pieces: a copy engine serves its queue in order, so a small read-back of
// the library stream (task counts, nnz, norms: ~10 per solver step) would otherwise wait behind a whole multi-hundred-MB
// transfer of a copy stream. With 4 MB pieces it waits for at most one piece (< 0.1 ms).
static void copy_in_pieces(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream) {
  constexpr size_t PIECE = 4u << 20;
  readback_flush();
  for (size_t off = 0; off < bytes; off += PIECE)
    CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off,
                               std::min(PIECE, bytes - off), kind, stream));
}
template <typename T>
__global__ void __launch_bounds__(256) k_local_filter(const int* __restrict__ row, const int* __restrict__ col, long long n,
                                                      int r0, int r1, int c0, int c1, int* __restrict__ flag) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    flag[i] = (row[i] >= r0 && row[i] < r1 && col[i] >= c0 && col[i] < c1) ? 1 : 0;
}
template <typename T>
__global__ void __launch_bounds__(256) k_local_take(const int* __restrict__ row, const int* __restrict__ col,
                                                    const T* __restrict__ val, long long n, const int* __restrict__ flag,
                                                    const int* __restrict__ pos, int r0, int c0, int* __restrict__ orow,
                                                    int* __restrict__ ocol, T* __restrict__ oval) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (flag[i]) { const int o = pos[i]; orow[o] = row[i] - r0; ocol[o] = col[i] - c0; oval[o] = val[i]; }
}

// d_row/d_col: GLOBAL 0-based indices on device. Keeps what belongs to this rank's block.
template <typename T>
static void build_local_from_global_triplets(Matrix& M, const int* d_row, const int* d_col, const T* d_val, long long n) {
  LocalCsc<T>& L = loc<T>(M);
  if (n == 0) { L.init_empty(M.local_rows, M.local_cols); return; }
  NTB_CHECK(n < (1ll << 31), "too many triplets");
  DevBuf<int> flag((size_t)n), pos((size_t)n + 1);
  const int g = std::min(div_up(n, 256), kNumSMs * 16);
  NTB_LAUNCH((k_local_filter<T>), g, 256, 0, d_row, d_col, n, M.start_row, M.start_row + M.local_rows, M.start_col,
             M.start_col + M.local_cols, flag.get());
  exclusive_scan(flag.get(), pos.get(), (int)n);
  int m = 0;
  d2h(&m, pos.get() + n, 1);
  DevBuf<int> lrow((size_t)m), lcol((size_t)m);
  DevBuf<T> lval((size_t)m);
  NTB_LAUNCH((k_local_take<T>), g, 256, 0, d_row, d_col, d_val, n, flag.get(), pos.get(), M.start_row, M.start_col,
             lrow.get(), lcol.get(), lval.get());
  csc_from_device_triplets<T>(M.local_rows, M.local_cols, lrow.get(), lcol.get(), lval.get(), m, L);
}

// all ranks of `comm` end up with the concatenation of everyone's (row,col,val) arrays
template <typename T>
static long long allgather_triplets(CommHandle* comm, DevBuf<int>& row, DevBuf<int>& col, DevBuf<T>& val, long long n) {
  const int np = comm_size(comm);
  if (np == 1) return n;
  std::vector<long long> cnt = exchange_counts(comm, n);
  std::vector<long long> off(np + 1, 0);
  for (int q = 0; q < np; ++q) off[q + 1] = off[q] + cnt[q];
  DevBuf<int> arow((size_t)off[np]), acol((size_t)off[np]);
  DevBuf<T> aval((size_t)off[np]);
  comm_group_start();
  for (int q = 0; q < np; ++q) {
    comm_broadcast_bytes(comm, row.get(), arow.get() + off[q], (size_t)cnt[q] * sizeof(int), q);
    comm_broadcast_bytes(comm, col.get(), acol.get() + off[q], (size_t)cnt[q] * sizeof(int), q);
    comm_broadcast_bytes(comm, val.get(), aval.get() + off[q], (size_t)cnt[q] * sizeof(T), q);
  }
  comm_group_end();
  row = std::move(arow); col = std::move(acol); val = std::move(aval);
  return off[np];
}

__global__ void __launch_bounds__(256) k_add_const2(int* __restrict__ a, int* __restrict__ b, long long n, int da, int db) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    a[i] += da;
    b[i] += db;
  }
}

// ---- fast ingest of lists that are already this rank's block in column-major order (what GetMatrixTripletList
// hands out and what host codes that keep CSC/CSR data produce): no sort, no gather. One pass decides, the rest is
// two streaming copies and a binary search per column.
// bad[0] != 0 unless every (global 0-based) entry lies in [r0,r1) x [c0,c1) and (col,row) is strictly ascending
__global__ void __launch_bounds__(256) k_check_sorted_local(const int* __restrict__ row, const int* __restrict__ col, long long n,
                                                            int r0, int r1, int c0, int c1, int* __restrict__ bad) {
  bool ok = true;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = row[i], c = col[i];
    ok = ok && r >= r0 && r < r1 && c >= c0 && c < c1;
    if (i > 0) {
      const int pr = row[i - 1], pc = col[i - 1];
      ok = ok && (pc < c || (pc == c && pr < r));
    }
  }
  if (!ok) *bad = 1;
}
__global__ void __launch_bounds__(256) k_sub_const(const int* __restrict__ in, int* __restrict__ out, long long n, int d) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = in[i] - d;
}
// outer[j] = first position whose column is >= c0 + j (col ascending)
__global__ void __launch_bounds__(256) k_outer_from_sorted_cols(const int* __restrict__ col, int n, int cols, int c0,
                                                                int* __restrict__ outer) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > cols) return;
  const int key = c0 + j;
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid] < key) lo = mid + 1; else hi = mid; }
  outer[j] = lo;
}
// collective over the slice: true when every rank's list is its own block, sorted and duplicate free
static bool lists_are_local_blocks(const Matrix& M, const int* d_row, const int* d_col, long long n) {
  DevBuf<int> bad(1);
  bad.zero();
  if (n >= (1ll << 31)) { const int one = 1; h2d(bad.get(), &one, 1); stream_sync(); }
  else if (n > 0)
    NTB_LAUNCH(k_check_sorted_local, std::min(div_up(n, 256), kNumSMs * 16), 256, 0, d_row, d_col, n, M.start_row,
               M.start_row + M.local_rows, M.start_col, M.start_col + M.local_cols, bad.get());
  int h = 0;
  d2h(&h, bad.get(), 1);
  double v = h ? 1.0 : 0.0;
  if (comm_size(M.grid->within_slice) > 1) {
    DevBuf<double> d(1);
    h2d(d.get(), &v, 1);
    comm_allreduce_f64(M.grid->within_slice, d.get(), 1, RedOp::Max);
    d2h(&v, d.get(), 1);
  }
  return v == 0.0;
}
template <typename T>
static void build_local_from_sorted_block(Matrix& M, const int* d_row, const int* d_col, const T* d_val, long long n) {
  LocalCsc<T>& L = loc<T>(M);
  if (n == 0) { L.init_empty(M.local_rows, M.local_cols); return; }
  L.rows = M.local_rows; L.cols = M.local_cols;
  L.outer.alloc((size_t)L.cols + 1);
  L.alloc_entries(n);
  NTB_LAUNCH(k_outer_from_sorted_cols, div_up(L.cols + 1, 256), 256, 0, d_col, (int)n, L.cols, M.start_col, L.outer.get());
  NTB_LAUNCH(k_sub_const, std::min(div_up(n, 256), kNumSMs * 16), 256, 0, d_row, L.inner.get(), n, M.start_row);
  d2d(L.val.get(), d_val, (size_t)n);
}

// d_*: this rank's list on the device, GLOBAL 0-based indices
template <typename T>
static void ingest_device_triplets(Matrix& M, DevBuf<int>& d_row, DevBuf<int>& d_col, DevBuf<T>& d_val, long long n,
                                   bool preduplicated, bool prepartitioned) {
  if (lists_are_local_blocks(M, d_row.get(), d_col.get(), n)) {
    build_local_from_sorted_block<T>(M, d_row.get(), d_col.get(), d_val.get(), n);
    rt().sorted_ingests++;
  } else {
    long long total = n;
    if (!prepartitioned) total = allgather_triplets<T>(M.grid->within_slice, d_row, d_col, d_val, n);
    build_local_from_global_triplets<T>(M, d_row.get(), d_col.get(), d_val.get(), total);
  }
  if (!prepartitioned && !preduplicated && M.grid->S > 1)   // FillMatrixFromTripletList.f90:37-42
    reduce_and_sum<T>(loc<T>(M), M.grid->between_slice, 0.0, M.row_block());
}

template <typename T>
static void fill_from_triplets_t(Matrix& M, const int* rows, const int* cols, const T* vals, long long n,
                                 bool preduplicated, bool prepartitioned) {
  // the caller's arrays go to the device as they are; 1-based global -> 0-based global happens there
  DevBuf<int> d_row((size_t)n), d_col((size_t)n);
  DevBuf<T> d_val((size_t)n);
  if (n) {
    h2d(d_row.get(), rows, (size_t)n); h2d(d_col.get(), cols, (size_t)n); h2d(d_val.get(), vals, (size_t)n);
    NTB_LAUNCH(k_add_const2, std::min(div_up(n, 256), kNumSMs * 16), 256, 0, d_row.get(), d_col.get(), n, -1, -1);
  }
  stream_sync();
  ingest_device_triplets<T>(M, d_row, d_col, d_val, n, preduplicated, prepartitioned);
}

// ---- staged (asynchronous) ingest: the host-to-device copies of a list run on a copy stream of their own while the
// library stream computes on an earlier matrix; the list becomes a matrix later, with mat_fill_from_staged
namespace {
cudaStream_t g_h2d_stream = nullptr;
}
StagedTriplets::~StagedTriplets() {
  if (!ready) return;
  if (row.get()) cudaEventSynchronize(ready);    // never consumed: the copies must be over before the buffers are recycled
  cudaEventDestroy(ready);
}
void stage_triplets(StagedTriplets& S, const int* rows, const int* cols, const double* vals, long long n) {
  ensure_init();
  if (!g_h2d_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&g_h2d_stream, cudaStreamNonBlocking));
  S.n = n;
  S.row.alloc((size_t)n); S.col.alloc((size_t)n); S.val.alloc((size_t)n);
  if (!S.ready) CUDA_CHECK(cudaEventCreateWithFlags(&S.ready, cudaEventDisableTiming));
  // arena blocks are recycled in library-stream order: the copies must not start before the work already enqueued there
  cudaEvent_t now;
  CUDA_CHECK(cudaEventCreateWithFlags(&now, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(now, rt().stream));
  CUDA_CHECK(cudaStreamWaitEvent(g_h2d_stream, now, 0));
  CUDA_CHECK(cudaEventDestroy(now));
  if (n) {
    copy_in_pieces(S.row.get(), rows, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, g_h2d_stream);
    copy_in_pieces(S.col.get(), cols, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, g_h2d_stream);
    copy_in_pieces(S.val.get(), vals, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, g_h2d_stream);
  }
  CUDA_CHECK(cudaEventRecord(S.ready, g_h2d_stream));
}
void mat_fill_from_staged(Matrix& M, StagedTriplets& S) {
  NTB_CHECK(M.constructed, "FillMatrixFromStaged on an unconstructed matrix");
  NTB_CHECK(S.ready != nullptr, "FillMatrixFromStaged: nothing was staged");
  if (M.is_complex) { Matrix t; mat_construct_empty(t, M.actual_dim, M.grid, false); M = std::move(t); }
  CUDA_CHECK(cudaStreamWaitEvent(rt().stream, S.ready, 0));
  if (S.n) NTB_LAUNCH(k_add_const2, std::min(div_up(S.n, 256), kNumSMs * 16), 256, 0, S.row.get(), S.col.get(), S.n, -1, -1);
  ingest_device_triplets<double>(M, S.row, S.col, S.val, S.n, false, false);
  S.row.release(); S.col.release(); S.val.release();       // back to the arena, in library-stream order
  S.n = 0;
}

void mat_fill_from_triplets(Matrix& M, const int* rows, const int* cols, const double* vals_r, const cplx* vals_c,
                            long long n, bool preduplicated, bool prepartitioned) {
  NTB_CHECK(M.constructed, "FillMatrixFromTripletList on an unconstructed matrix");
  if (M.is_complex) {
    std::vector<cplx> tmp;
    if (!vals_c) { tmp.resize((size_t)n); for (long long i = 0; i < n; ++i) tmp[i] = cplx{vals_r[i], 0.0}; vals_c = tmp.data(); }
    fill_from_triplets_t<cplx>(M, rows, cols, vals_c, n, preduplicated, prepartitioned);
  } else {
    NTB_CHECK(vals_r != nullptr || n == 0, "complex triplets into a real matrix");
    fill_from_triplets_t<double>(M, rows, cols, vals_r, n, preduplicated, prepartitioned);
  }
}

long long mat_get_triplets(const Matrix& M, int* rows, int* cols, double* vals_r, cplx* vals_c) {
  const long long nnz = M.local_nnz();
  if (!rows) return nnz;
  if (nnz == 0) return 0;
  DevBuf<int> d_row((size_t)nnz), d_col((size_t)nnz);
  if (M.is_complex) csc_to_device_triplets<cplx>(M.c.view(), nnz, d_row.get(), d_col.get());
  else csc_to_device_triplets<double>(M.r.view(), nnz, d_row.get(), d_col.get());
  // local 0-based -> global 1-based on the device
  NTB_LAUNCH(k_add_const2, std::min(div_up(nnz, 256), kNumSMs * 16), 256, 0, d_row.get(), d_col.get(), nnz, M.start_row + 1,
             M.start_col + 1);
  CUDA_CHECK(cudaMemcpyAsync(rows, d_row.get(), nnz * sizeof(int), cudaMemcpyDeviceToHost, rt().stream));
  CUDA_CHECK(cudaMemcpyAsync(cols, d_col.get(), nnz * sizeof(int), cudaMemcpyDeviceToHost, rt().stream));
  if (M.is_complex) {
    NTB_CHECK(vals_c != nullptr, "complex matrix needs a complex value buffer");
    CUDA_CHECK(cudaMemcpyAsync(vals_c, M.c.val.get(), nnz * sizeof(cplx), cudaMemcpyDeviceToHost, rt().stream));
  } else {
    NTB_CHECK(vals_r != nullptr, "real matrix needs a real value buffer");
    CUDA_CHECK(cudaMemcpyAsync(vals_r, M.r.val.get(), nnz * sizeof(double), cudaMemcpyDeviceToHost, rt().stream));
  }
  stream_sync();
  return nnz;
}

// ---- asynchronous egress: the device side (entries -> global triplets) runs on the library stream, the device-to-host
// copies on a second stream behind an event, so they overlap whatever the caller enqueues next (e.g. the host-to-device
// copy of the next input: PCIe is full duplex). The staged device arrays live until mat_egress_wait().
namespace {
struct PendingEgress { DevBuf<int> row, col; DevBuf<double> val; };
// (never destroyed: at process exit the blocks of an egress nobody waited for would otherwise be handed back to a
// device arena whose own static destructor may already have run)
std::vector<PendingEgress>& g_pending_egress = *new std::vector<PendingEgress>();
cudaStream_t g_copy_stream = nullptr;
cudaEvent_t g_egress_ready = nullptr;
}  // namespace
long long mat_get_triplets_async(const Matrix& M, int* rows, int* cols, double* vals_r) {
  NTB_CHECK(!M.is_complex, "asynchronous egress: real matrices only");
  const long long nnz = M.local_nnz();
  if (nnz == 0) return 0;
  if (!g_copy_stream) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&g_egress_ready, cudaEventDisableTiming));
  }
  PendingEgress pe;
  pe.row.alloc((size_t)nnz); pe.col.alloc((size_t)nnz); pe.val.alloc((size_t)nnz);
  const CscView<double> v = M.r.view();
  csc_to_device_triplets<double>(v, nnz, pe.row.get(), pe.col.get());
  NTB_LAUNCH(k_add_const2, std::min(div_up(nnz, 256), kNumSMs * 16), 256, 0, pe.row.get(), pe.col.get(), nnz, M.start_row + 1,
             M.start_col + 1);
  d2d(pe.val.get(), v.val, (size_t)nnz);       // the matrix may be overwritten before the copy has run
  CUDA_CHECK(cudaEventRecord(g_egress_ready, rt().stream));
  CUDA_CHECK(cudaStreamWaitEvent(g_copy_stream, g_egress_ready, 0));
  copy_in_pieces(rows, pe.row.get(), nnz * sizeof(int), cudaMemcpyDeviceToHost, g_copy_stream);
  copy_in_pieces(cols, pe.col.get(), nnz * sizeof(int), cudaMemcpyDeviceToHost, g_copy_stream);
  copy_in_pieces(vals_r, pe.val.get(), nnz * sizeof(double), cudaMemcpyDeviceToHost, g_copy_stream);
  g_pending_egress.push_back(std::move(pe));
  return nnz;
}
void mat_egress_wait() {
  if (g_copy_stream) CUDA_CHECK(cudaStreamSynchronize(g_copy_stream));
  g_pending_egress.clear();                    // blocks go back to the arena only now
}

void mat_fill_identity(Matrix& M) {
  // distributed_includes/FillMatrixIdentity.f90:9-22: only indices <= actual dimension
  std::vector<int> rows, cols;
  std::vector<double> vals;
  for (int jj = M.start_row; jj < M.start_row + M.local_rows; ++jj)
    if (jj >= M.start_col && jj < M.start_col + M.local_cols && jj < M.actual_dim) {
      rows.push_back(jj + 1); cols.push_back(jj + 1); vals.push_back(1.0);
    }
  mat_fill_from_triplets(M, rows.data(), cols.data(), vals.data(), nullptr, (long long)rows.size(), true, true);
}

void mat_fill_permutation(Matrix& M, const int* lookup, bool permute_rows) {
  // distributed_includes/FillMatrixPermutation.f90 (lookup is 1-based, over the logical dimension)
  std::vector<int> rows, cols;
  std::vector<double> vals;
  if (permute_rows) {
    for (int ii = M.start_row; ii < M.start_row + M.local_rows; ++ii) {
      const int pc = lookup[ii] - 1;
      if (pc >= M.start_col && pc < M.start_col + M.local_cols) { rows.push_back(ii + 1); cols.push_back(pc + 1); vals.push_back(1.0); }
    }
  } else {
    for (int ii = M.start_col; ii < M.start_col + M.local_cols; ++ii) {
      const int pr = lookup[ii] - 1;
      if (pr >= M.start_row && pr < M.start_row + M.local_rows) { rows.push_back(pr + 1); cols.push_back(ii + 1); vals.push_back(1.0); }
    }
  }
  mat_fill_from_triplets(M, rows.data(), cols.data(), vals.data(), nullptr, (long long)rows.size(), true, true);
}

__global__ void __launch_bounds__(256) k_shift_swap(int* __restrict__ row, int* __restrict__ col, long long n, int roff, int coff) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = row[i] + roff, c = col[i] + coff;
    row[i] = c; col[i] = r;   // transposed global position
  }
}

template <typename T> static void transpose_t(const Matrix& A, Matrix& out) {
  Matrix res;
  mat_construct_empty(res, A.actual_dim, A.grid, A.is_complex);
  const LocalCsc<T>& L = loc<T>(A);
  if (comm_size(A.grid->within_slice) == 1) {
    csc_transpose<T>(L.view(), loc<T>(res));
  } else {
    // every slice transposes its own replica: global triplets, swap, regroup inside the slice
    // (the reference routes this through triplet redistribution as well,
    //  distributed_includes/TransposeMatrix.f90)
    const long long n = L.nnz;
    DevBuf<int> d_row((size_t)n), d_col((size_t)n);
    DevBuf<T> d_val((size_t)n);
    csc_to_device_triplets<T>(L.view(), n, d_row.get(), d_col.get());
    d2d(d_val.get(), L.val.get(), (size_t)n);
    if (n) NTB_LAUNCH(k_shift_swap, std::min(div_up(n, 256), kNumSMs * 16), 256, 0, d_row.get(), d_col.get(), n,
                      A.start_row, A.start_col);
    const long long total = allgather_triplets<T>(A.grid->within_slice, d_row, d_col, d_val, n);
    build_local_from_global_triplets<T>(res, d_row.get(), d_col.get(), d_val.get(), total);
  }
  out = std::move(res);
}
void mat_transpose(const Matrix& A, Matrix& out) {
  if (A.is_complex) transpose_t<cplx>(A, out); else transpose_t<double>(A, out);
}

// ---- symmetric relabelling out(map[r], map[c]) = A(r, c): what PermuteMatrix / UndoPermuteMatrix compute with two
// products by permutation matrices (LoadBalancerModule.F90:38-47, 77-86), done as an index gather and a re-sort
// (SURVEY 8f row 3). Every result entry of those products is 1*v*1 with a single term and the products keep |v| > 0,
// so relabelling and dropping exact zeros gives the same matrix bit for bit.
__global__ void __launch_bounds__(256) k_relabel(int* __restrict__ row, int* __restrict__ col, long long n, int roff, int coff,
                                                 const int* __restrict__ map) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    row[i] = map[row[i] + roff];
    col[i] = map[col[i] + coff];
  }
}
template <typename T> static void relabel_t(const Matrix& A, Matrix& out, const int* h_map0) {
  Matrix res;
  mat_construct_empty(res, A.actual_dim, A.grid, A.is_complex);
  const LocalCsc<T>& L = loc<T>(A);
  const long long n = L.nnz;
  DevBuf<int> d_row((size_t)n), d_col((size_t)n), d_map((size_t)A.logical_dim);
  DevBuf<T> d_val((size_t)n);
  h2d(d_map.get(), h_map0, (size_t)A.logical_dim);
  if (n) {
    const CscView<T> v = L.view();
    csc_to_device_triplets<T>(v, n, d_row.get(), d_col.get());
    d2d(d_val.get(), v.val, (size_t)n);
    NTB_LAUNCH(k_relabel, std::min(div_up(n, 256), kNumSMs * 16), 256, 0, d_row.get(), d_col.get(), n, A.start_row,
               A.start_col, d_map.get());
  }
  stream_sync();                                 // h_map0 is the copy's source
  ingest_device_triplets<T>(res, d_row, d_col, d_val, n, true, false);
  csc_filter<T>(loc<T>(res), 0.0);               // the products would not keep an exact zero
  out = std::move(res);
}
void mat_relabel(const Matrix& A, Matrix& out, const int* h_map0) {
  NTB_CHECK(A.constructed, "PermuteMatrix on an unconstructed matrix");
  if (A.is_complex) relabel_t<cplx>(A, out, h_map0); else relabel_t<double>(A, out, h_map0);
}

void mat_conjugate(Matrix& M) { if (M.is_complex) csc_conjugate<cplx>(M.c); }

void mat_filter(Matrix& M, double threshold) {
  if (M.is_complex) csc_filter<cplx>(M.c, threshold); else csc_filter<double>(M.r, threshold);
}

// Reduction of n <= 8 device doubles over `comm`, result on the host. On a column-split peer grid (peer.h) whose
// communicator spans all ranks this is ONE small kernel - the barrier/exchange - riding on the read-back the caller
// needs anyway, reduced on the host in rank order (run-to-run and rank-to-rank identical); NCCL otherwise.
static void reduce_to_host(const ProcessGrid& g, CommHandle* comm, double* d_vals, int n, const RedOp* ops, double* h_out) {
  const int np = comm_size(comm);
  if (np > 1 && g.peer_ok && peer().ok && np == peer().n) {
    std::vector<PeerPayload> all((size_t)np);
    PeerPayload z{};
    peer_exchange(z, all.data(), d_vals, n);
    stream_sync();
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int p = 0; p < np; ++p) {
        double v;
        std::memcpy(&v, &all[p].w[i], sizeof(double));
        if (p == 0) acc = v;
        else if (ops[i] == RedOp::Sum) acc += v;
        else if (ops[i] == RedOp::Max) acc = std::max(acc, v);
        else acc = std::min(acc, v);
      }
      h_out[i] = acc;
    }
    return;
  }
  for (int i = 0; i < n; ++i) comm_allreduce_f64(comm, d_vals + i, 1, ops[i]);
  for (int i = 0; i < n; ++i) readback_async(&h_out[i], d_vals + i, sizeof(double));
  stream_sync();
}
static double reduce_to_host(const ProcessGrid& g, CommHandle* comm, double* d_val, RedOp op) {
  double h = 0.0;
  reduce_to_host(g, comm, d_val, 1, &op, &h);
  return h;
}

static void allreduce_host(CommHandle* comm, double* vals, int n, RedOp op) {
  if (comm_size(comm) == 1) return;
  DevBuf<double> d((size_t)n);
  h2d(d.get(), vals, (size_t)n);
  comm_allreduce_f64(comm, d.get(), (size_t)n, op);
  d2h(vals, d.get(), (size_t)n);
}

long long mat_global_nnz(const Matrix& M) {
  double v = (double)M.local_nnz();
  allreduce_host(M.grid->within_slice, &v, 1, RedOp::Sum);
  return (long long)(v + 0.5);
}

bool mat_is_identity(const Matrix& M) {
  DevBuf<double> d(2);
  if (M.is_complex) csc_identity_check<cplx>(M.c.view(), M.start_row, M.start_col, d.get());
  else csc_identity_check<double>(M.r.view(), M.start_row, M.start_col, d.get());
  comm_allreduce_f64(M.grid->within_slice, d.get(), 2, RedOp::Sum);
  double h[2];
  d2h(h, d.get(), 2);
  return h[0] == 0.0 && (long long)(h[1] + 0.5) == (long long)M.actual_dim;
}

// ===========================================================================
// algebra
// ===========================================================================
void mat_scale(Matrix& M, double c) {
  if (M.is_complex) csc_scale<cplx>(M.c, cplx{c, 0.0}); else csc_scale<double>(M.r, c);
}
void mat_scale_c(Matrix& M, cplx c) {
  if (!M.is_complex) { Matrix t; mat_to_complex(M, t); M = std::move(t); }   // PSMatrixAlgebraModule.F90:483-504
  csc_scale<cplx>(M.c, c);
}

void mat_increment(const Matrix& A, Matrix& B, double alpha, double threshold) {
  NTB_CHECK(A.constructed && B.constructed, "IncrementMatrix on an unconstructed matrix");
  NTB_CHECK(A.logical_dim == B.logical_dim, "IncrementMatrix: dimension mismatch");
  if (A.is_complex && !B.is_complex) { Matrix t; mat_to_complex(B, t); B = std::move(t); }  // :436-439
  if (!A.is_complex && B.is_complex) {
    Matrix t;
    mat_to_complex(A, t);
    csc_increment<cplx>(t.c.view(), B.c, alpha, threshold, B.row_block());
    return;
  }
  if (A.is_complex) csc_increment<cplx>(A.c.view(), B.c, alpha, threshold, B.row_block());
  else csc_increment<double>(A.r.view(), B.r, alpha, threshold, B.row_block());
}

double mat_trace(const Matrix& M) {
  if (!M.is_complex && M.r.deferred) {            // an iterate that lives in tile space: no CSC round trip
    double t;
    if (mat_tile_scalars(2, M, nullptr, &t)) return t;
  }
  DevBuf<double> d(1);
  if (M.is_complex) csc_trace<cplx>(M.c.view(), M.start_row, M.start_col, d.get());
  else csc_trace<double>(M.r.view(), M.start_row, M.start_col, d.get());
  return reduce_to_host(*M.grid, M.grid->within_slice, d.get(), RedOp::Sum);
}

double mat_norm(const Matrix& M) {
  // distributed_algebra_includes/MatrixNorm.f90: column sums over the process column, max over the row
  DevBuf<double> colsum((size_t)M.local_cols), d(1);
  if (M.is_complex) csc_col_abs_sums<cplx>(M.c.view(), colsum.get());
  else csc_col_abs_sums<double>(M.r.view(), colsum.get());
  comm_allreduce_f64(M.grid->column, colsum.get(), (size_t)M.local_cols, RedOp::Sum);
  reduce_max(colsum.get(), M.local_cols, d.get());
  return reduce_to_host(*M.grid, M.grid->row, d.get(), RedOp::Max);
}

// MatrixNorm(alpha*A + B) without forming the sum (see k_diff_col_abs); operands of equal type and layout
double mat_diff_norm(const Matrix& A, const Matrix& B, double alpha) {
  NTB_CHECK(A.constructed && B.constructed && A.logical_dim == B.logical_dim && A.grid == B.grid, "diff norm: operands differ in layout");
  if (A.is_complex != B.is_complex) {           // mixed types: the two reference calls on a scratch copy
    Matrix t;
    mat_copy(B, t);
    mat_increment(A, t, alpha, 0.0);
    return mat_norm(t);
  }
  PhaseScope ph(3);
  DevBuf<double> colsum((size_t)B.local_cols), d(1);
  if (A.is_complex) csc_diff_col_abs_sums<cplx>(A.c.view(), B.c.view(), alpha, colsum.get());
  else if (!(tile_path_on() && tile_diff_col_abs_sums(A.r, B.r, alpha, colsum.get())))   // iterates that live as tile forms
    csc_diff_col_abs_sums<double>(A.r.view(), B.r.view(), alpha, colsum.get());
  comm_allreduce_f64(B.grid->column, colsum.get(), (size_t)B.local_cols, RedOp::Sum);
  reduce_max(colsum.get(), B.local_cols, d.get());
  return reduce_to_host(*B.grid, B.grid->row, d.get(), RedOp::Max);
}

// ---- fused driver steps in tile space -----------------------------------------------------------------------------
static bool tile_space_operand(const Matrix& M) {
  return M.constructed && !M.is_complex && M.r.forms && M.r.forms->has_right == 1;
}
bool mat_tile_scalars(int mode, const Matrix& A, const Matrix* B, double* out) {
  if (!tile_path_on() || !tile_space_operand(A) || (B && (!tile_space_operand(*B) || B->grid != A.grid || B->logical_dim != A.logical_dim)))
    return false;
  // only worth it (and only collective-safe) for iterates that came out of tile products: those exist on every rank
  if (!A.r.forms->right.emitted && !(B && B->r.forms->right.emitted)) return false;
  // (ranks may disagree on a form built from CSC; harmless where the reduction is ONE peer exchange whatever each rank
  // contributes, not where it is a sequence of NCCL calls)
  if (A.grid->size > 1 && !(A.grid->peer_ok && peer().ok)) return false;
  PhaseScope ph(3);
  const int nout = (mode == 1) ? 2 : 1;
  DevBuf<double> d(2);
  const int dd = A.start_col - A.start_row;
  const int ncd = std::max(0, std::min(A.local_cols, A.actual_dim - A.start_col));
  if (!tile_form_scalars(mode, A.r, B ? &B->r : nullptr, dd, ncd, d.get())) return false;
  const RedOp ops[2] = {RedOp::Sum, RedOp::Sum};
  reduce_to_host(*A.grid, A.grid->within_slice, d.get(), nout, ops, out);
  return true;
}
bool mat_tile_combine(const Matrix& P, const Matrix& Q, int mode, double alpha, double beta, double thr, double sigma,
                      Matrix& Out, unsigned want) {
  if (!tile_path_on() || !tile_space_operand(P) || P.grid != Q.grid || P.logical_dim != Q.logical_dim) return false;
  if (!P.r.forms->right.emitted) return false;          // P is the product of the step (X^2, T_k): always a tile-space result
  // Q may still be a plain CSC block on a single rank (the first steps of a recurrence: Identity, the input): its
  // right form is built and cached like a product operand's
  if (!tile_space_operand(Q) && Q.constructed && !Q.is_complex && P.grid->size == 1 && Q.r.nnz > 0) tile_operand_form(Q.r, false);
  if (!tile_space_operand(Q)) return false;
  ProcessGrid& g = *P.grid;
  const bool multi = g.size > 1;
  if (multi && !(g.peer_ok && peer().ok)) return false;  // (general grids keep the reference's call sequence)
  PhaseScope ph(4);
  Matrix res;
  mat_construct_empty(res, P.actual_dim, P.grid, false);
  const int dd = P.start_col - P.start_row;
  const int ncd = std::max(0, std::min(P.local_cols, P.actual_dim - P.start_col));
  if (!tile_combine(P.r, Q.r, mode, alpha, beta, thr, sigma, dd, ncd, P.row_block(), res.r, want, multi)) return false;
  Out = std::move(res);
  return true;
}

double mat_sigma(const Matrix& M) {
  const double n = mat_norm(M);
  return 1.0 / (n * n);
}

void mat_gershgorin(const Matrix& M, double* e_min, double* e_max) {
  DevBuf<double> dmin((size_t)M.local_cols), dmax((size_t)M.local_cols), d(2);
  if (M.is_complex) csc_gershgorin_cols<cplx>(M.c.view(), M.start_row, M.start_col, dmin.get(), dmax.get());
  else csc_gershgorin_cols<double>(M.r.view(), M.start_row, M.start_col, dmin.get(), dmax.get());
  comm_allreduce_f64(M.grid->column, dmin.get(), (size_t)M.local_cols, RedOp::Sum);
  comm_allreduce_f64(M.grid->column, dmax.get(), (size_t)M.local_cols, RedOp::Sum);
  reduce_min(dmin.get(), M.local_cols, d.get());
  reduce_max(dmax.get(), M.local_cols, d.get() + 1);
  const RedOp ops[2] = {RedOp::Min, RedOp::Max};
  double h[2];
  reduce_to_host(*M.grid, M.grid->row, d.get(), 2, ops, h);
  *e_min = h[0];
  *e_max = h[1];
}

void mat_dot(const Matrix& A, const Matrix& B, double* re, double* im) {
  if (!A.is_complex && !B.is_complex && (A.r.deferred || B.r.deferred) && A.r.nnz.maybe_nonzero() && B.r.nnz.maybe_nonzero()) {
    // an iterate that lives in tile space (e.g. Tr(X H) of a purification): from the right forms, the other
    // operand's form is built once and cached
    if (!A.r.forms) A.r.forms = std::make_shared<TileForms>();
    if (!B.r.forms) B.r.forms = std::make_shared<TileForms>();
    const bool ok = tile_operand_form(A.r, false) != nullptr && tile_operand_form(B.r, false) != nullptr;
    double t;
    if (ok && mat_tile_scalars(0, A, &B, &t)) { *re = t; if (im) *im = 0.0; return; }
  }
  DevBuf<double> d(2);
  if (A.is_complex || B.is_complex) {
    Matrix ta, tb;
    const Matrix* pa = &A; const Matrix* pb = &B;
    if (!A.is_complex) { mat_to_complex(A, ta); pa = &ta; }
    if (!B.is_complex) { mat_to_complex(B, tb); pb = &tb; }
    csc_dot<cplx>(pa->c.view(), pb->c.view(), d.get());
  } else {
    csc_dot<double>(A.r.view(), B.r.view(), d.get());
  }
  const RedOp ops[2] = {RedOp::Sum, RedOp::Sum};
  double h[2];
  reduce_to_host(*A.grid, A.grid->within_slice, d.get(), 2, ops, h);
  *re = h[0];
  if (im) *im = h[1];
}

void mat_pairwise(const Matrix& A, const Matrix& B, Matrix& C) {
  Matrix res;
  const bool cx = A.is_complex || B.is_complex;
  mat_construct_empty(res, A.actual_dim, A.grid, cx);
  if (cx) {
    Matrix ta, tb;
    const Matrix* pa = &A; const Matrix* pb = &B;
    if (!A.is_complex) { mat_to_complex(A, ta); pa = &ta; }
    if (!B.is_complex) { mat_to_complex(B, tb); pb = &tb; }
    csc_pairwise<cplx>(pa->c.view(), pb->c.view(), res.c);
  } else {
    csc_pairwise<double>(A.r.view(), B.r.view(), res.r);
  }
  C = std::move(res);
}

double mat_measure_asymmetry(const Matrix& M) {       // PSMatrixAlgebraModule.F90:569-583
  Matrix t;
  mat_transpose(M, t);
  mat_conjugate(t);
  mat_increment(M, t, -1.0, 0.0);
  return mat_norm(t);
}

void mat_symmetrize(Matrix& M) {                       // PSMatrixAlgebraModule.F90:586-599
  Matrix t;
  mat_transpose(M, t);
  mat_conjugate(t);
  mat_increment(t, M, 1.0, 0.0);
  mat_scale(M, 0.5);
}

// ---------------------------------------------------------------------------
// the distributed multiply (distributed_algebra_includes/MatrixMultiply.f90)
// ---------------------------------------------------------------------------
// entries per row block of a CSC panel. Consecutive entries of a column fall into the same block almost always, so
// each thread counts runs locally and flushes a run with one shared-memory atomic; one global atomic per block and bin.
constexpr int RB_BINS = 1024;
__global__ void __launch_bounds__(256) k_rowblock_nnz(const int* __restrict__ inner, long long nnz, int rb, int nbins,
                                                      unsigned long long* __restrict__ counts) {
  __shared__ unsigned int sh[RB_BINS];
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  // contiguous chunk per thread so that runs are long
  const long long per = (nnz + (long long)gridDim.x * blockDim.x - 1) / ((long long)gridDim.x * blockDim.x);
  const long long lo = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * per, hi = min(nnz, lo + per);
  int cur = -1;
  unsigned int run = 0;
  for (long long i = lo; i < hi; ++i) {
    const int b = inner[i] / rb;
    if (b != cur) { if (run) atomicAdd(&sh[cur], run); cur = b; run = 0; }
    ++run;
  }
  if (run) atomicAdd(&sh[cur], run);
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += blockDim.x)
    if (sh[i]) atomicAdd(&counts[i], (unsigned long long)sh[i]);
}


// ---------------------------------------------------------------------------
// Tile-form product on a 1 x C x 1 grid (column split): every rank owns all rows of its columns, so
//   * the right operand B needs no gather at all (process column of size 1): its cached right form is used as is;
//   * the left operand A is needed only for the inner indices that occur as ROWS of the local B block. The left form
//     is stored per chunk column (32 matrix columns) with its tiles contiguous, so "gathering A" is: all-gather the
//     small index arrays, then fetch from each peer ONE contiguous run of tiles — the halo — with ncclSend/ncclRecv.
//     For banded/short-range matrices the halo is a few chunk columns instead of the reference's whole row panel
//     (comm_includes/ReduceAndComposeMatrix*.f90), and nothing is ever converted back to CSC or re-tiled.
// All decisions are taken from all-gathered records, so every rank takes the same branch.
__global__ void __launch_bounds__(256) k_right_form_range(const int4* __restrict__ colmeta, int ncc, int* __restrict__ out2) {
  int lo = INT_MAX, hi = -1;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ncc; q += gridDim.x * blockDim.x) {
    const int4 m = colmeta[q];
    if (m.y > 0) { lo = min(lo, m.z); hi = max(hi, m.w); }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&out2[0], lo); atomicMax(&out2[1], hi); }
}
__global__ void __launch_bounds__(256) k_col_len(const int* __restrict__ outer, int cols, int* __restrict__ len) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < cols) len[j] = outer[j + 1] - outer[j];
}

struct HaloRecord { long long ok, nsuper, ntiles, nnzA, qlo, qhi, nnzB, ntilesB; };
static long long halo_tile_limit() {
  static const long long lim = [] { const char* e = std::getenv("NTB_HALO_TILE_LIMIT"); return e ? std::atoll(e) : (1ll << 26); }();
  return lim;
}
__global__ void k_range_into_record(const int* __restrict__ r2, long long* __restrict__ qlo, long long* __restrict__ qhi) {
  if (threadIdx.x == 0) { *qlo = r2[0]; *qhi = r2[1]; }
}

// returns false (on every rank alike) when this product does not qualify; then the caller takes the CSC gather path
static bool halo_tile_product(const Matrix& A, const Matrix& B, double alpha, double wthr, const RuleView& rv,
                              const std::function<void()>& full_rules, const DiagShift* ds, LocalCsc<double>& out,
                              GemmStats& st, unsigned want) {
  ProcessGrid& g = *A.grid;
  if (!(g.R == 1 && g.S == 1 && g.C > 1) || !tile_path_on() || A.local_cols % 64 != 0 || !(wthr >= 0.0)) return false;
  static const bool timing = std::getenv("NTB_HALO_TIMING") != nullptr;      // developer probe: wall time per phase
  auto now = [&]() { if (timing) stream_sync(); return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = now();
  const int C = g.C, me = comm_rank(g.row);
  const LocalCsc<double>& Al = A.r;
  const LocalCsc<double>& Bl = B.r;
  const int lcols = A.local_cols, nccl = lcols / 32, nkl = lcols / 4;
  const ChunkTiles* Lf = tile_operand_form(Al, true);
  const ChunkTiles* Rf = tile_operand_form(Bl, false);
  // ---- records: who can play, sizes, and which global chunk columns of A each rank needs (the needed range is
  // written into the device copy of the record by a kernel: one all-gather, one read-back)
  HaloRecord mine{};
  // forms built from CSC must be reasonably full (mostly-padding tiles are the scalar kernels' business); forms written
  // by a product have fixed 64-tile slots, their tile count says nothing about the fill
  auto dense_enough = [](const ChunkTiles* f, long long nnz) {
    return f->emitted || (double)nnz >= 0.20 * 32.0 * (double)f->ntiles;
  };
  mine.ok = (Lf && Rf && dense_enough(Lf, Al.nnz) && dense_enough(Rf, Bl.nnz)) ? 1 : 0;
  mine.qlo = INT_MAX; mine.qhi = -1;
  if (mine.ok) { mine.nsuper = Lf->nsuper; mine.ntiles = Lf->ntiles; mine.nnzA = Al.nnz; mine.nnzB = Bl.nnz; mine.ntilesB = Rf->ntiles; }
  std::vector<HaloRecord> rec(C);
  {
    DevBuf<HaloRecord> d_mine(1), d_all((size_t)C);
    DevBuf<int> d2(2);
    const int init[2] = {INT_MAX, -1};
    h2d(d_mine.get(), &mine, 1);
    h2d(d2.get(), init, 2);
    if (mine.ok && Rf->ncc > 0) {
      NTB_LAUNCH(k_right_form_range, std::min(div_up(Rf->ncc, 256), kNumSMs), 256, 0, Rf->colmeta.get(), Rf->ncc, d2.get());
      NTB_LAUNCH(k_range_into_record, 1, 32, 0, d2.get(), &d_mine.get()->qlo, &d_mine.get()->qhi);
    }
    comm_allgather_bytes(g.row, d_mine.get(), d_all.get(), sizeof(HaloRecord));
    d2h(rec.data(), d_all.get(), (size_t)C);
    mine = rec[me];
  }
  const auto t1 = now();
  long long nnzA = 0, ntilesA = 0, nnzB = 0, ntilesB = 0;
  for (int p = 0; p < C; ++p) {
    if (!rec[p].ok) return false;
    nnzA += rec[p].nnzA; ntilesA += rec[p].ntiles; nnzB += rec[p].nnzB; ntilesB += rec[p].ntilesB;
  }
  if (nnzA == 0 || nnzB == 0) return false;
  // A row block of A's panel cannot be more than 10 % full when the whole panel holds fewer entries than 10 % of
  // ONE row block: only past that bound is the per-block histogram (and its all-reduce) needed for the rule table
  if ((double)nnzA > 0.1 * (double)A.row_block() * (double)Bl.rows) full_rules();

  // ---- small index arrays of every rank
  ChunkTiles G;
  G.ncc = C * nccl;
  G.colmeta.alloc((size_t)C * nccl);
  G.kmeta.alloc((size_t)C * nkl);
  const bool count = rt().count_flops;        // a process-wide setting: every rank takes the same branch
  DevBuf<int> coltile_g((size_t)C * (nccl + 1)), ylen, ylen_g;
  if (count) {
    ylen.alloc((size_t)lcols); ylen_g.alloc((size_t)C * lcols);
    NTB_LAUNCH(k_col_len, div_up(lcols, 256), 256, 0, Al.outer.get(), lcols, ylen.get());
  }
  std::vector<int> ent_base(C + 1, 0);
  for (int p = 0; p < C; ++p) ent_base[p + 1] = ent_base[p] + (int)rec[p].nsuper;
  G.nsuper = ent_base[C];
  G.ent.alloc((size_t)std::max(G.nsuper, 1));
  comm_group_start();
  comm_allgather_bytes(g.row, Lf->colmeta.get(), G.colmeta.get(), (size_t)nccl * sizeof(int4));
  comm_allgather_bytes(g.row, Lf->kmeta.get(), G.kmeta.get(), (size_t)nkl * sizeof(int4));
  comm_allgather_bytes(g.row, Lf->coltile.get(), coltile_g.get(), ((size_t)nccl + 1) * sizeof(int));
  if (count) comm_allgather_bytes(g.row, ylen.get(), ylen_g.get(), (size_t)lcols * sizeof(int));
  for (int p = 0; p < C; ++p)
    if (rec[p].nsuper > 0)
      comm_broadcast_bytes(g.row, Lf->ent.get(), G.ent.get() + ent_base[p], (size_t)rec[p].nsuper * sizeof(int4), p);
  comm_group_end();
  std::vector<int> ct((size_t)C * (nccl + 1));
  d2h(ct.data(), coltile_g.get(), ct.size());
  const auto t2 = now();

  // ---- the halo: tiles of chunk columns [a,b) of rank p, one contiguous run
  auto run_of = [&](int need_lo, int need_hi, int p, int& a, int& b, int& tl, int& th) {
    a = std::max(need_lo - p * nccl, 0);
    b = std::min(need_hi + 1 - p * nccl, nccl);
    if (a >= b) { a = b = 0; tl = th = 0; return; }
    tl = ct[(size_t)p * (nccl + 1) + a];
    th = ct[(size_t)p * (nccl + 1) + b];
  };
  // The rank's own tiles are used where they are; only the tiles received from the peers go into a new buffer, which
  // the entries address with indices from HALO_TILE_BIAS on (csc.cuh).
  std::vector<LeftPiece> pieces(C);
  long long halo_tiles = 0, total_tiles = 0;
  for (int p = 0; p < C; ++p) {
    int a, b, tl, th;
    run_of((int)mine.qlo, (int)mine.qhi, p, a, b, tl, th);
    total_tiles += th - tl;
    if (p != me) halo_tiles += th - tl;
  }
  // the halo buffer is addressed with 32-bit tile indices from HALO_TILE_BIAS on: a left operand or a halo that does
  // not fit DECLINES the path - on every rank alike, every rank evaluates every rank's sizes from the gathered records
  for (int r = 0; r < C; ++r) {
    long long ht = 0;
    for (int p = 0; p < C; ++p) {
      int a, b, tl, th;
      run_of((int)rec[r].qlo, (int)rec[r].qhi, p, a, b, tl, th);
      if (p != r) ht += th - tl;
    }
    if (ht >= halo_tile_limit() || rec[r].ntiles >= HALO_TILE_BIAS - 64) return false;
  }
  G.ntiles = total_tiles;
  G.tval.alloc((size_t)std::max(halo_tiles, 1ll) * 32);
  G.tval_view = Lf->tval.get();
  const long long halo_origin = HALO_TILE_BIAS;
  {
    long long at = 0;
    for (int p = 0; p < C; ++p) {
      int a, b, tl, th;
      run_of((int)mine.qlo, (int)mine.qhi, p, a, b, tl, th);
      const int base = (p == me) ? tl : (int)(halo_origin + at);
      pieces[p] = LeftPiece{ent_base[p], nccl, a, b, tl, base, (int)rec[p].nsuper};
      if (p != me) at += th - tl;
    }
  }
  comm_group_start();
  for (int p = 0; p < C; ++p) {
    if (p == me) continue;
    int a, b, tl, th;
    run_of((int)mine.qlo, (int)mine.qhi, p, a, b, tl, th);                        // what I take from p
    const size_t rbytes = (size_t)(th - tl) * 256;
    if (rbytes) comm_recv_bytes(g.row, G.tval.get() + ((long long)pieces[p].recv_base - halo_origin) * 32, rbytes, p);
    run_of((int)rec[p].qlo, (int)rec[p].qhi, me, a, b, tl, th);                   // what p takes from me
    const size_t sbytes = (size_t)(th - tl) * 256;
    if (sbytes) comm_send_bytes(g.row, Lf->tval.get() + (size_t)tl * 32, sbytes, p);
  }
  comm_group_end();
  tile_fixup_gathered_left(G, pieces.data(), C);
  const auto t3 = now();

  // ---- product from the forms
  double useful = -1.0;
  if (count) { Bl.ensure_entries(); useful = useful_products_from_lengths(Bl, ylen_g.get()); }
  const auto t4 = now();
  const bool done = spgemm_tile_core(left_view_of(G), *Rf, B.local_cols, A.local_rows, alpha, wthr, rv, out, useful, ds, true, want);
  const auto t5 = now();
  if (timing && me == 0)
    std::fprintf(stderr, "[halo] records %.3f  index gather %.3f  tiles %.3f (%.1f MB)  flops %.3f  product %.3f ms\n",
                 ms(t0, t1), ms(t1, t2), ms(t2, t3), (double)halo_tiles * 256.0 / 1e6, ms(t3, t4), ms(t4, t5));
  NTB_CHECK(done, "forced tile product declined");
  st.flops = count ? 2.0 * useful : 0.0;
  st.shift_applied = ds && ds->sigma != 0.0;
  // compulsory bytes: B block, kept C block, and the share of A that was actually needed (by tile count)
  double a_bytes = 0.0;
  for (int p = 0; p < C; ++p) {
    int a, b, tl, th;
    run_of((int)mine.qlo, (int)mine.qhi, p, a, b, tl, th);
    if (rec[p].ntiles > 0) a_bytes += (double)rec[p].nnzA * 12.0 * (double)(th - tl) / (double)rec[p].ntiles;
  }
  rt().alg_bytes += a_bytes + (double)Bl.bytes() + (double)out.bytes();
  rt().halo_products++;
  rt().halo_bytes += (double)halo_tiles * 256.0;
  return true;
}


// ---------------------------------------------------------------------------
// Tile-form product on a 1 x C x 1 grid WITHOUT any copy of the left operand (peer.h): every rank has mapped every
// other rank's slab, the left forms live there, and the numeric kernel's bulk copies fetch the few A super-tiles of
// the neighbouring ranks straight from the owner's HBM over NVLink, stage by stage, overlapped with the DMMAs of the
// stages before - the reference's row-panel gather (comm_includes/ReduceAndComposeMatrix*.f90) fused into the product.
// What the ranks need of each other is 64 bytes per left form: where its arrays lie (PeerLeftDesc). A product hands
// out the descriptor of the left form it has written with the same small kernel that acts as the barrier between its
// tile stores and the peers' reads (spgemm_tile_core, publish); a form built from CSC is published at its first use
// here. Steady state of a solver loop: no NCCL call, no host decision, nothing but the product's own two read-backs.
// Returns false ON EVERY RANK ALIKE when the product does not qualify (all verdicts come from exchanged records).
static bool peer_tile_product(const Matrix& A, const Matrix& B, double alpha, double wthr, const RuleView& rv,
                              const std::function<void()>& full_rules, const DiagShift* ds, LocalCsc<double>& out,
                              GemmStats& st, unsigned want) {
  ProcessGrid& g = *A.grid;
  if (!g.peer_ok || !peer().ok || peer().n != g.C || !tile_path_on() || A.local_cols % 64 != 0 || !(wthr >= 0.0)) return false;
  const int C = g.C;
  const LocalCsc<double>& Al = A.r;
  const LocalCsc<double>& Bl = B.r;
  const int lcols = A.local_cols;
  const ChunkTiles* Lf = tile_operand_form(Al, true);
  const ChunkTiles* Rf = tile_operand_form(Bl, false);
  TileForms& fa = *Al.forms;
  TileForms& fb = *Bl.forms;
  // forms built from CSC must be reasonably full (mostly-padding tiles are the scalar kernels' business); forms written
  // by a product have fixed 64-tile slots, their tile count says nothing about the fill
  auto dense_enough = [](const ChunkTiles* f, long long nnz) {
    return f->emitted || (double)nnz >= 0.20 * 32.0 * (double)f->ntiles;
  };
  // a right form written by a (collective) product exists on every rank; one built from CSC needs a collective verdict
  const bool right_known = (Rf && Rf->emitted) || fb.right_all_ok != 0;
  if (fa.left_pub.empty() && fa.pub_pending) stream_sync();      // published by the product that wrote it; descriptors on their way
  if (fa.left_pub.empty() || !right_known) {
    PeerLeftDesc mine = left_desc_of(Lf, Al.nnz, Lf && dense_enough(Lf, Al.nnz));
    mine.right_ok = (Rf && dense_enough(Rf, Bl.nnz)) ? 1 : 0;
    std::vector<PeerLeftDesc> all((size_t)C);
    PeerPayload pl;
    std::memcpy(&pl, &mine, sizeof(pl));
    peer_exchange(pl, reinterpret_cast<PeerPayload*>(all.data()));
    stream_sync();
    bool rok = true;
    for (int p = 0; p < C; ++p) rok = rok && all[p].right_ok != 0;
    if (!(Rf && Rf->emitted)) fb.right_all_ok = rok ? 1 : -1;
    else if (!rok) return false;
    if (fa.left_pub.empty()) fa.left_pub = all;
    fa.left_needs_barrier = false;                // the exchange was a barrier
  }
  if (!(Rf && Rf->emitted) && fb.right_all_ok != 1) return false;
  long long nnzA = 0;
  for (int p = 0; p < C; ++p) {
    if (!fa.left_pub[p].ok) return false;
    nnzA += fa.left_pub[p].nnz;
  }
  if (nnzA == 0) return false;
  if (fa.left_needs_barrier) { peer_barrier(); fa.left_needs_barrier = false; }
  // A row block of A's panel cannot be more than 10 % full when the whole panel holds fewer entries than 10 % of
  // ONE row block: only past that bound is the per-block histogram (and its all-reduce) needed for the rule table
  if ((double)nnzA > 0.1 * (double)A.row_block() * (double)Bl.rows) full_rules();

  LeftView V;
  V.npieces = C;
  V.ncc_piece = lcols / 32;
  for (int p = 0; p < C; ++p) {
    const PeerLeftDesc& d = fa.left_pub[p];
    const unsigned char* base = peer().base[p];
    V.piece[p] = LeftPieceView{reinterpret_cast<const int4*>(base + d.off_colmeta), reinterpret_cast<const int4*>(base + d.off_ent),
                               reinterpret_cast<const int4*>(base + d.off_kmeta), reinterpret_cast<const double*>(base + d.off_tval)};
  }
  // instrumentation (off in timed regions): useful products need the column lengths of the whole row panel of A
  double useful = -1.0;
  if (rt().count_flops) {
    DevBuf<int> ylen((size_t)lcols), ylen_g((size_t)C * lcols);
    NTB_LAUNCH(k_col_len, div_up(lcols, 256), 256, 0, Al.outer.get(), lcols, ylen.get());
    comm_allgather_bytes(g.row, ylen.get(), ylen_g.get(), (size_t)lcols * sizeof(int));
    Bl.ensure_entries();
    useful = useful_products_from_lengths(Bl, ylen_g.get());
  }
  if (!spgemm_tile_core(V, *Rf, B.local_cols, A.local_rows, alpha, wthr, rv, out, useful, ds, true, want, true)) return false;
  st.flops = rt().count_flops ? 2.0 * useful : 0.0;
  st.shift_applied = ds && ds->sigma != 0.0;
  // compulsory bytes: the rank's share of A (the few halo super-tiles read from the neighbours come on top), its B
  // block and the kept C block
  account_product_bytes(Bl.nnz, Bl.cols, Al.nnz, Al.cols, true, out.nnz, out.cols, sizeof(double));
  rt().halo_products++;
  rt().peer_products++;
  return true;
}

static int g_halo = -1;
void set_halo_path(int on) { g_halo = on ? 1 : 0; }
static bool halo_enabled() {
  if (g_halo < 0) { const char* e = std::getenv("NTB_HALO"); g_halo = (e && e[0] == '0') ? 0 : 1; }
  return g_halo == 1;
}
static int g_fused_shift = -1;
void set_fused_shift(int on) { g_fused_shift = on ? 1 : 0; }
static bool fused_shift_enabled() {
  if (g_fused_shift < 0) { const char* e = std::getenv("NTB_FUSED_SHIFT"); g_fused_shift = (e && e[0] == '0') ? 0 : 1; }
  return g_fused_shift == 1;
}

// returns true when the optional diagonal shift `sigma` (C = alpha*A*B + sigma*I, see DiagShift) was fused
template <typename T>
static bool multiply_t(const Matrix& A, const Matrix& B, Matrix& C, double alpha, double beta, double threshold,
                       double sigma = 0.0, unsigned want = WANT_ALL, double* diff_colsum = nullptr, bool* diff_applied = nullptr) {
  ProcessGrid& g = *A.grid;
  const int S = g.S;
  const double wthr = (S > 1) ? threshold / (S * 1000) : threshold;      // MatrixMultiply.f90:25-29
  const LocalCsc<T>& Al = loc<T>(A);
  const LocalCsc<T>& Bl = loc<T>(B);
  const int rb = A.row_block(), cb = A.col_block();

  const int nI = g.nbr, nJ = g.nbc;
  const bool want_shift = sigma != 0.0 && S == 1 && std::fabs(beta) < 2.2250738585072014e-308;
  DiagShift ds;
  if (want_shift) {
    ds.sigma = sigma;
    ds.dd = A.start_col - A.start_row;        // same block coordinates for A, B and the product
    ds.ncols_diag = std::max(0, std::min(A.local_cols, A.actual_dim - A.start_col));
  }
  // fused |AB - A| column sums (DiagShift::diff_colsum): whole columns must be local (one process row), one slice,
  // plain replacement of C
  const bool want_diff = diff_colsum != nullptr && S == 1 && g.R == 1 && std::fabs(beta) < 2.2250738585072014e-308;
  if (want_diff) ds.diff_colsum = diff_colsum;
  const DiagShift* const ds_p = (want_shift || want_diff) ? &ds : nullptr;
  DiagShift ds_plain = ds;                    // (paths that cannot take the second fusion)
  ds_plain.diff_colsum = nullptr;
  const DiagShift* const ds_plain_p = want_shift ? &ds_plain : nullptr;
  // per local block pair: dense or sparse threshold rule (GemmMatrix.f90:49-61) from the panel fills
  DevBuf<unsigned char> d_rule;
  RuleView rv;
  auto set_rules = [&](const std::vector<double>& fa, const std::vector<double>& fb) {
    std::vector<unsigned char> rule((size_t)nI * nJ, 0);
    bool any = false;
    for (int i = 0; i < nI; ++i)
      for (int j = 0; j < nJ; ++j) {
        rule[(size_t)i * nJ + j] = (std::min(fa[i], fb[j]) > 0.1) ? 1 : 0;
        if (rule[(size_t)i * nJ + j]) { rt().dense_rule_blocks++; any = true; }
      }
    if (any) {
      d_rule.alloc(rule.size());
      h2d(d_rule.get(), rule.data(), rule.size());
      stream_sync();                           // rule (host vector) is the h2d source
      rv.tbl = d_rule.get(); rv.rb = rb; rv.cb = cb; rv.nJ = nJ;
    }
  };
  auto rowblock_counts = [&](const LocalCsc<T>& Ypan, std::vector<double>& cnt_out) {
    cnt_out.assign(nI, 0.0);
    if (nI == 1) { cnt_out[0] = (double)Ypan.nnz; return; }
    Ypan.ensure_entries();
    DevBuf<unsigned long long> cnt((size_t)nI);
    cnt.zero();
    NTB_CHECK(nI <= RB_BINS, "more than 1024 row blocks per rank");
    if (Ypan.nnz) NTB_LAUNCH(k_rowblock_nnz, std::min(div_up(Ypan.nnz, 256 * 64), kNumSMs * 8), 256, 0, Ypan.inner.get(), Ypan.nnz, rb, nI, cnt.get());
    std::vector<unsigned long long> h(nI);
    d2h(h.data(), cnt.get(), (size_t)nI);
    for (int i = 0; i < nI; ++i) cnt_out[i] = (double)h[i];
  };
  auto colblock_fills = [&](const LocalCsc<T>& Xpan, double inner_dim, std::vector<double>& fb) {
    fb.assign(nJ, 0.0);
    if (nJ == 1) { fb[0] = (double)Xpan.nnz / ((double)cb * inner_dim); return; }
    std::vector<int> marks(nJ + 1);
    for (int j = 0; j <= nJ; ++j) readback_async(&marks[j], Xpan.outer.get() + (size_t)j * cb, sizeof(int));
    stream_sync();
    for (int j = 0; j < nJ; ++j) fb[j] = (double)(marks[j + 1] - marks[j]) / ((double)cb * inner_dim);
  };

  Matrix AB;
  mat_construct_empty(AB, A.actual_dim, A.grid, scalar_traits<T>::is_complex);
  GemmStats st;
  bool product_done = false;

  // ---- column-split grids: tile forms + halo exchange, no CSC gather (real matrices)
  if constexpr (!scalar_traits<T>::is_complex) {
    if (g.R == 1 && S == 1 && g.C > 1 && tile_path_on() && A.local_cols % 64 == 0 && halo_enabled()) {
      // panel fills for the rule table: A's row panel is the union of the ranks' blocks, B's column panel is local
      auto full_rules = [&]() {
        const double inner_dim = (double)Bl.rows;
        std::vector<double> cnt, fb;
        rowblock_counts(Al, cnt);
        {
          DevBuf<double> d((size_t)nI);
          h2d(d.get(), cnt.data(), (size_t)nI);
          comm_allreduce_f64(g.row, d.get(), (size_t)nI, RedOp::Sum);
          d2h(cnt.data(), d.get(), (size_t)nI);
        }
        std::vector<double> fa(nI);
        for (int i = 0; i < nI; ++i) fa[i] = cnt[i] / ((double)rb * inner_dim);
        colblock_fills(Bl, inner_dim, fb);
        set_rules(fa, fb);
      };
      product_done = peer_tile_product(A, B, alpha, wthr, rv, full_rules, ds_p, loc<double>(AB), st, want);
      if (!product_done) {
        rv = RuleView();
        // without a peer space (GPUs that cannot map each other, NTB_P2P=0): the tile halo is copied with NCCL
        if (!(g.peer_ok && peer().ok))
          product_done = halo_tile_product(A, B, alpha, wthr, rv, full_rules, ds_plain_p, loc<double>(AB), st, want);
        if (!product_done) { rv = RuleView(); }
      }
    }
  }

  if (!product_done) {
    // every path below except the single-rank tile product reads the CSC entries of both operands
    if (S > 1 || comm_size(g.row) > 1 || comm_size(g.column) > 1 || nI > 1) { Al.ensure_entries(); Bl.ensure_entries(); }
    // ---- A task: my slice's column blocks, gathered along the process row (:94-145)
    std::unique_ptr<PhaseScope> ph_gather(new PhaseScope(5));     // slice selection, panel gathers, stacking
    LocalCsc<T> Asel, Ypanel;
    const LocalCsc<T>* Ysrc = &Al;
    if (S > 1) { csc_select_col_blocks<T>(Al.view(), cb, S, g.my_slice, Asel); Ysrc = &Asel; }
    if (comm_size(g.row) > 1) { gather_concat_cols<T>(*Ysrc, g.row, Ypanel); Ysrc = &Ypanel; }
    // ---- B task: my slice's row blocks, gathered along the process column (:154-193)
    LocalCsc<T> Bsel, Xpanel;
    const LocalCsc<T>* Xsrc = &Bl;
    if (S > 1) { csc_select_row_blocks<T>(Bl.view(), rb, S, g.my_slice, Bsel); Xsrc = &Bsel; }
    if (comm_size(g.column) > 1) {
      std::vector<LocalCsc<T>> parts;
      std::vector<CscView<T>> views;
      gather_parts<T>(*Xsrc, g.column, parts, views);
      std::vector<int> roff(views.size());
      for (size_t q = 0; q < views.size(); ++q) roff[q] = (int)q * Xsrc->rows;
      csc_stack_rows<T>(views.data(), roff.data(), (int)views.size(), Xsrc->rows * (int)views.size(), Xpanel);
      Xsrc = &Xpanel;
    }
    NTB_CHECK(Ysrc->cols == Xsrc->rows, "multiply: gathered panels disagree on the inner dimension");
    ph_gather.reset();
    {
      const double inner_dim = (double)Xsrc->rows;
      std::vector<double> cnt, fa(nI), fb;
      rowblock_counts(*Ysrc, cnt);
      bool any_a = false;
      for (int i = 0; i < nI; ++i) { fa[i] = cnt[i] / ((double)rb * inner_dim); any_a = any_a || fa[i] > 0.1; }
      // the dense rule needs BOTH panels more than 10 % full: when no row block of A's panel is, B's fills are not
      // needed (and B's entry count - a deferred product's may still be on its way - is not waited for)
      if (any_a) {
        colblock_fills(*Xsrc, inner_dim, fb);
        set_rules(fa, fb);
      }
    }
    // ---- local product
    // deferred entries only for a plain replacement of C on a single slice
    const unsigned w = (S == 1 && (std::fabs(beta) < 2.2250738585072014e-308 || !C.constructed)) ? want : WANT_ALL;
    // (the second fusion only where the left operand is the rank's own block, not a gathered panel)
    const bool local_operands = comm_size(g.row) == 1 && comm_size(g.column) == 1;
    spgemm<T>(*Xsrc, *Ysrc, alpha, wthr, rv, loc<T>(AB), &st, local_operands ? ds_p : ds_plain_p, w);
  }
  if (diff_applied) *diff_applied = ds.diff_applied;
  rt().flops_useful += st.flops;
  rt().multiplies++;

  // ---- between-slice sum (:234-261)
  if (S > 1) {
    PhaseScope ph_sum(7);
    reduce_and_sum<T>(loc<T>(AB), g.between_slice, threshold, rb);
  }

  // ---- C = AB  or  C = beta*C + AB (:324-329)
  if (std::fabs(beta) < 2.2250738585072014e-308 || !C.constructed) {
    C = std::move(AB);
  } else {
    mat_scale(C, beta);
    mat_increment(AB, C, 1.0, 0.0);
  }
  return st.shift_applied;
}

void mat_multiply(const Matrix& A, const Matrix& B, Matrix& C, double alpha, double beta, double threshold,
                  MemoryPool* pool, unsigned want) {
  NTB_CHECK(A.constructed && B.constructed, "MatrixMultiply on an unconstructed matrix");
  NTB_CHECK(A.logical_dim == B.logical_dim && A.grid == B.grid, "MatrixMultiply: operands live on different grids/sizes");
  if (pool) { pool->rows = A.local_rows; pool->cols = A.local_cols; pool->is_complex = A.is_complex || B.is_complex; pool->constructed = true; }
  // A, B may alias C: results are built in a temporary and moved in at the end.
  if (A.is_complex || B.is_complex) {                                    // PSMatrixAlgebraModule.F90:171-205
    Matrix ta, tb;
    const Matrix* pa = &A; const Matrix* pb = &B;
    if (!A.is_complex) { mat_to_complex(A, ta); pa = &ta; }
    if (!B.is_complex) { mat_to_complex(B, tb); pb = &tb; }
    if (C.constructed && !C.is_complex && std::fabs(beta) > 0) { Matrix tc; mat_to_complex(C, tc); C = std::move(tc); }
    multiply_t<cplx>(*pa, *pb, C, alpha, beta, threshold);
  } else {
    multiply_t<double>(A, B, C, alpha, beta, threshold, 0.0, want);
  }
}

// C = alpha*A*B (thresholded) followed by IncrementMatrix(Identity, C, sigma) with threshold 0 — the pair every
// Newton-Schulz style driver issues (e.g. SignSolversModule.F90:226-229). The shift is fused into the product's
// emit pass when the product runs on the tile path; otherwise the two reference calls are made.
void mat_multiply_shift(const Matrix& A, const Matrix& B, Matrix& C, double alpha, double threshold, double sigma,
                        const Matrix& Identity, MemoryPool* pool, unsigned want) {
  NTB_CHECK(A.constructed && B.constructed, "MatrixMultiply on an unconstructed matrix");
  if (!A.is_complex && !B.is_complex && sigma != 0.0 && fused_shift_enabled()) {
    NTB_CHECK(A.logical_dim == B.logical_dim && A.grid == B.grid, "MatrixMultiply: operands live on different grids/sizes");
    if (pool) { pool->rows = A.local_rows; pool->cols = A.local_cols; pool->is_complex = false; pool->constructed = true; }
    if (multiply_t<double>(A, B, C, alpha, 0.0, threshold, sigma, want)) return;
  } else {
    mat_multiply(A, B, C, alpha, 0.0, threshold, pool);
  }
  mat_increment(Identity, C, sigma, 0.0);
}

// C = alpha*A*B (thresholded) and, in the same pass when the product runs on the tile path, MatrixNorm(C - A): the pair
// "Gemm; IncrementMatrix(C, A, -1); MatrixNorm(A)" of the sign / polar iteration (SignSolversModule.F90:230-234) with
// the difference taken in the product's epilogue (csc.cuh: DiagShift::diff_colsum). Always computes C; returns true
// and the norm when the fusion applied, false when the caller still has to call mat_diff_norm(C, A, -1).
static int g_fused_norm = -1;
void set_fused_norm(int on) { g_fused_norm = on ? 1 : 0; }
static bool fused_norm_enabled() {
  if (g_fused_norm < 0) { const char* e = std::getenv("NTB_FUSED_NORM"); g_fused_norm = (e && e[0] == '0') ? 0 : 1; }
  return g_fused_norm == 1;
}
bool mat_multiply_diffnorm(const Matrix& A, const Matrix& B, Matrix& C, double alpha, double threshold, MemoryPool* pool,
                           unsigned want, double* norm_out) {
  NTB_CHECK(A.constructed && B.constructed, "MatrixMultiply on an unconstructed matrix");
  if (A.is_complex || B.is_complex || !fused_norm_enabled() || !tile_path_on() || &A == &C) {
    mat_multiply(A, B, C, alpha, 0.0, threshold, pool, want);
    return false;
  }
  NTB_CHECK(A.logical_dim == B.logical_dim && A.grid == B.grid, "MatrixMultiply: operands live on different grids/sizes");
  if (pool) { pool->rows = A.local_rows; pool->cols = A.local_cols; pool->is_complex = false; pool->constructed = true; }
  DevBuf<double> colsum((size_t)std::max(A.local_cols, 1)), d(1);
  bool applied = false;
  ProcessGrid* grid = A.grid;
  const int lcols = A.local_cols;
  multiply_t<double>(A, B, C, alpha, 0.0, threshold, 0.0, want, colsum.get(), &applied);
  if (!applied) return false;
  PhaseScope ph(3);
  reduce_max(colsum.get(), lcols, d.get());
  *norm_out = reduce_to_host(*grid, grid->row, d.get(), RedOp::Max);
  rt().fused_norms++;
  return true;
}

void mat_similarity_transform(const Matrix& A, const Matrix& P, const Matrix& PInv, Matrix& Res, MemoryPool* pool,
                              double threshold) {
  if (mat_is_identity(P)) { mat_copy(A, Res); return; }       // PSMatrixAlgebraModule.F90:633-634
  Matrix tmp;
  mat_multiply(P, A, tmp, 1.0, 0.0, threshold, pool);
  mat_multiply(tmp, PInv, Res, 1.0, 0.0, threshold, pool);
}

// ===========================================================================
// container utilities either side of the path (host-side logic over the ingest / egress primitives; not per-iteration)
// ===========================================================================
namespace {
struct HostTriplets { std::vector<int> rows, cols; std::vector<double> re; std::vector<cplx> cx; };
// local block as global 1-based host triplets
HostTriplets local_host_triplets(const Matrix& M) {
  HostTriplets t;
  const long long n = M.local_nnz();
  t.rows.resize((size_t)n); t.cols.resize((size_t)n);
  if (M.is_complex) { t.cx.resize((size_t)n); mat_get_triplets(M, t.rows.data(), t.cols.data(), nullptr, t.cx.data()); }
  else { t.re.resize((size_t)n); mat_get_triplets(M, t.rows.data(), t.cols.data(), t.re.data(), nullptr); }
  return t;
}
// keep the entries for which pred(row, col) holds, renumbered by (row - dr, col - dc)
template <typename P> HostTriplets select_triplets(const HostTriplets& in, bool cx, int dr, int dc, P pred) {
  HostTriplets o;
  for (size_t i = 0; i < in.rows.size(); ++i)
    if (pred(in.rows[i], in.cols[i])) {
      o.rows.push_back(in.rows[i] - dr); o.cols.push_back(in.cols[i] - dc);
      if (cx) o.cx.push_back(in.cx[i]); else o.re.push_back(in.re[i]);
    }
  return o;
}
void fill_host_triplets(Matrix& M, const HostTriplets& t, bool preduplicated, bool prepartitioned) {
  mat_fill_from_triplets(M, t.rows.data(), t.cols.data(), M.is_complex ? nullptr : t.re.data(),
                         M.is_complex ? t.cx.data() : nullptr, (long long)t.rows.size(), preduplicated, prepartitioned);
}
}  // namespace

// distributed_includes/FillMatrixDense.f90: 1.0 at every position of the local block inside the actual dimension
void mat_fill_dense(Matrix& M) {
  NTB_CHECK(M.constructed, "FillMatrixDense on an unconstructed matrix");
  HostTriplets t;
  for (int c = M.start_col; c < M.start_col + M.local_cols && c < M.actual_dim; ++c)
    for (int r = M.start_row; r < M.start_row + M.local_rows && r < M.actual_dim; ++r) {
      t.rows.push_back(r + 1); t.cols.push_back(c + 1);
      if (M.is_complex) t.cx.push_back(cplx{1.0, 0.0}); else t.re.push_back(1.0);
    }
  fill_host_triplets(M, t, true, true);
}

// distributed_includes/ResizeMatrix.f90: entries inside the new size survive; the result lives on the GLOBAL grid
// (the reference calls ConstructEmptyMatrix without a grid there)
void mat_resize(Matrix& M, int new_size) {
  NTB_CHECK(M.constructed, "ResizeMatrix on an unconstructed matrix");
  const bool cx = M.is_complex;
  const HostTriplets kept = select_triplets(local_host_triplets(M), cx, 0, 0,
                                            [&](int r, int c) { return r <= new_size && c <= new_size; });
  Matrix res;
  mat_construct_empty(res, new_size, nullptr, cx);
  fill_host_triplets(res, kept, true, false);
  M = std::move(res);
}

// distributed_includes/SliceMatrix.f90: rows [start_row, end_row] x columns [start_column, end_column] (1-based,
// inclusive) as a new matrix of dimension max(height, width) on the same grid
void mat_get_slice(const Matrix& M, Matrix& sub, int start_row, int end_row, int start_col, int end_col) {
  NTB_CHECK(M.constructed, "GetMatrixSlice on an unconstructed matrix");
  const bool cx = M.is_complex;
  const HostTriplets kept = select_triplets(local_host_triplets(M), cx, start_row - 1, start_col - 1, [&](int r, int c) {
    return r >= start_row && r <= end_row && c >= start_col && c <= end_col; });
  const int new_dim = std::max(end_row - start_row + 1, end_col - start_col + 1);
  Matrix res;
  mat_construct_empty(res, new_dim, M.grid, cx);
  fill_host_triplets(res, kept, true, false);
  sub = std::move(res);
}

// every entry held by the ranks of M's slice as global 1-based host triplets (the same on every rank of the slice)
template <typename T>
static void slice_triplets_to_host(const Matrix& M, const LocalCsc<T>& L, std::vector<int>& rows, std::vector<int>& cols,
                                   std::vector<T>& vals) {
  const long long n = L.nnz;
  DevBuf<int> d_row((size_t)n), d_col((size_t)n);
  DevBuf<T> d_val((size_t)n);
  if (n) {
    const CscView<T> v = L.view();
    csc_to_device_triplets<T>(v, n, d_row.get(), d_col.get());
    d2d(d_val.get(), v.val, (size_t)n);
    NTB_LAUNCH(k_add_const2, std::min(div_up(n, 256), kNumSMs * 16), 256, 0, d_row.get(), d_col.get(), n, M.start_row + 1,
               M.start_col + 1);
  }
  const long long total = allgather_triplets<T>(M.grid->within_slice, d_row, d_col, d_val, n);
  rows.resize((size_t)total); cols.resize((size_t)total); vals.resize((size_t)total);
  if (total) {
    d2h(rows.data(), d_row.get(), (size_t)total);
    d2h(cols.data(), d_col.get(), (size_t)total);
    d2h(vals.data(), d_val.get(), (size_t)total);
  }
}

// distributed_includes/GetMatrixBlock.f90: every rank names a block [start_row, end_row) x [start_column, end_column)
// (1-based, end exclusive) and receives the entries of the whole matrix that fall into it; an entry wanted by several
// ranks of a slice goes to the first of them
long long mat_get_block(const Matrix& M, int start_row, int end_row, int start_col, int end_col, std::vector<int>& rows,
                        std::vector<int>& cols, std::vector<double>& vals_interleaved) {
  NTB_CHECK(M.constructed, "GetMatrixBlock on an unconstructed matrix");
  CommHandle* comm = M.grid->within_slice;
  const int np = comm_size(comm), me = comm_rank(comm);
  std::vector<int> ranges((size_t)np * 4);
  {
    const int mine[4] = {start_row, end_row, start_col, end_col};
    if (np == 1) { std::copy(mine, mine + 4, ranges.begin()); }
    else {
      DevBuf<int> d_mine(4), d_all((size_t)np * 4);
      h2d(d_mine.get(), mine, 4);
      stream_sync();
      comm_allgather_bytes(comm, d_mine.get(), d_all.get(), 4 * sizeof(int));
      d2h(ranges.data(), d_all.get(), (size_t)np * 4);
    }
  }
  // all entries of the slice, global 1-based, on the host
  HostTriplets all;
  const bool cx = M.is_complex;
  if (np == 1) all = local_host_triplets(M);
  else if (cx) slice_triplets_to_host<cplx>(M, M.c, all.rows, all.cols, all.cx);
  else slice_triplets_to_host<double>(M, M.r, all.rows, all.cols, all.re);
  auto wanted_by = [&](int r, int c) {
    for (int p = 0; p < np; ++p) {
      const int* q = &ranges[(size_t)p * 4];
      if (r >= q[0] && r < q[1] && c >= q[2] && c < q[3]) return p;
    }
    return -1;
  };
  rows.clear(); cols.clear(); vals_interleaved.clear();
  for (size_t i = 0; i < all.rows.size(); ++i)
    if (wanted_by(all.rows[i], all.cols[i]) == me) {
      rows.push_back(all.rows[i]); cols.push_back(all.cols[i]);
      if (cx) { vals_interleaved.push_back(all.cx[i].x); vals_interleaved.push_back(all.cx[i].y); }
      else vals_interleaved.push_back(all.re[i]);
    }
  return (long long)rows.size();
}

// MatrixConversionModule.F90:21-61: mat takes the sparsity pattern of `pattern` - positions of the pattern that mat
// lacks become explicit zeros (an addition with a negative threshold never drops anything), values outside are removed
void mat_snap_to_pattern(Matrix& mat, const Matrix& pattern) {
  NTB_CHECK(mat.constructed && pattern.constructed, "SnapMatrixToSparsityPattern on an unconstructed matrix");
  Matrix ones, zeros, filtered;
  {
    HostTriplets t = local_host_triplets(pattern);
    t.cx.clear();
    t.re.assign(t.rows.size(), 1.0);
    mat_construct_empty(ones, pattern.actual_dim, pattern.grid, false);
    fill_host_triplets(ones, t, true, true);
  }
  mat_copy(ones, zeros);
  mat_scale(zeros, 0.0);
  mat_increment(zeros, mat, 1.0, -1.0);
  mat_copy(mat, filtered);
  mat_pairwise(ones, filtered, mat);
}

}  // namespace ntb
