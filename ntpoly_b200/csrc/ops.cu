// Per-iteration helpers on device CSC blocks: sparse add with NTPoly's threshold
// rules, scale, dot, trace, norms, Gershgorin, filter, transpose, slice selection.
// All kernels are column-parallel (one warp per CSC column, lanes over entries)
// so index/value loads are coalesced; the two sorted lists of a column are
// combined by rank (binary search) instead of a serial merge.
#include "ops.cuh"
#include "peer.h"
#include <cub/device/device_radix_sort.cuh>

namespace ntb {

__device__ __forceinline__ int lower_bound_dev(const int* __restrict__ a, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

static inline int warp_grid(int ncols) {
  return max(1, min(div_up((long long)ncols * 32, 256), kNumSMs * 32));
}

#define WARP_COL_LOOP(ncols)                                                   \
  const int lane = threadIdx.x & 31;                                           \
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;                 \
  const int nw = (gridDim.x * blockDim.x) >> 5;                                \
  for (int j = gw; j < (ncols); j += nw)

// ---------------------------------------------------------------------------
// increment:  B <- alpha*A + B      (reference AddSparseVectors.f90:21-70)
//   matched index           keep iff |alpha*a + b| > thr
//   unmatched, not in tail  keep iff |alpha*a| > thr  /  |b| > thr
//   tail (the other list has no entry at or after this index inside the same
//   local row block)        always kept
// Merged position (A before B on ties) = own rank + rank in the other list.
// ---------------------------------------------------------------------------
template <typename T, bool FILL>
__global__ void __launch_bounds__(256)
k_increment(CscView<T> A, CscView<T> B, double alpha, double thr, int rb, int* __restrict__ flags,
            const int* __restrict__ opos, int* __restrict__ out_outer, int* __restrict__ out_inner,
            T* __restrict__ out_val) {
  WARP_COL_LOOP(A.cols) {
    const int a0 = A.outer[j], na = A.outer[j + 1] - a0;
    const int b0 = B.outer[j], nb = B.outer[j + 1] - b0;
    const int base = a0 + b0;
    const int* ai = A.inner + a0;
    const int* bi = B.inner + b0;
    if (FILL && lane == 0) out_outer[j] = opos[base];
    for (int t = lane; t < na; t += 32) {
      const int ia = ai[t];
      const int pb = lower_bound_dev(bi, nb, ia);
      const bool matched = pb < nb && bi[pb] == ia;
      const T wa = s_scale(alpha, A.val[a0 + t]);
      T v = wa;
      bool keep;
      if (matched) {
        v = s_add(wa, B.val[b0 + pb]);
        keep = s_abs(v) > thr;
      } else {
        const int seg_end = (ia / rb + 1) * rb;
        const int pe = (pb < nb && bi[nb - 1] >= seg_end) ? pb + lower_bound_dev(bi + pb, nb - pb, seg_end) : nb;
        const bool tail = (pb == pe);
        keep = tail || s_abs(wa) > thr;
      }
      const int pos = base + t + pb;
      if (!FILL) flags[pos] = keep ? 1 : 0;
      else if (keep) { const int o = opos[pos]; out_inner[o] = ia; out_val[o] = v; }
    }
    for (int t = lane; t < nb; t += 32) {
      const int ib = bi[t];
      const int pa = lower_bound_dev(ai, na, ib);
      const bool matched = pa < na && ai[pa] == ib;
      if (matched) {
        if (!FILL) flags[base + t + pa + 1] = 0;
        continue;
      }
      const T wb = B.val[b0 + t];
      const int seg_end = (ib / rb + 1) * rb;
      const int pe = (pa < na && ai[na - 1] >= seg_end) ? pa + lower_bound_dev(ai + pa, na - pa, seg_end) : na;
      const bool tail = (pa == pe);
      const bool keep = tail || s_abs(wb) > thr;
      const int pos = base + t + pa;
      if (!FILL) flags[pos] = keep ? 1 : 0;
      else if (keep) { const int o = opos[pos]; out_inner[o] = ib; out_val[o] = wb; }
    }
  }
}

template <typename T>
void csc_increment(const CscView<T>& A, LocalCsc<T>& B, double alpha, double thr, int rb) {
  NTB_CHECK(A.cols == B.cols && A.rows == B.rows, "increment: shape mismatch");
  const int cols = A.cols;
  long long nnzA = 0;
  {
    int h = 0;
    d2h(&h, A.outer + cols, 1);
    nnzA = h;
  }
  // adding into an EMPTY block is a copy: every entry of A is an untested tail of the merge (kept whatever the
  // threshold) and 1.0 * a is a bit for bit - the first of the S adds of the between-slice sum, FillFromTripletList
  // style accumulations, Increment into a freshly constructed matrix
  if (B.nnz == 0 && alpha == 1.0) {
    LocalCsc<T> out;
    out.rows = A.rows; out.cols = cols;
    out.outer.alloc((size_t)cols + 1);
    out.alloc_entries(nnzA);
    d2d(out.outer.get(), A.outer, (size_t)cols + 1);
    if (nnzA > 0) { d2d(out.inner.get(), A.inner, (size_t)nnzA); d2d(out.val.get(), A.val, (size_t)nnzA); }
    B.swap(out);
    return;
  }
  const long long total = nnzA + B.nnz;
  NTB_CHECK(total < (1ll << 31), "increment: more than 2^31 entries in a local block");
  if (rb <= 0) rb = A.rows > 0 ? A.rows : 1;
  DevBuf<int> flags((size_t)total), opos((size_t)total + 1);
  CscView<T> Bv = B.view();
  NTB_LAUNCH((k_increment<T, false>), warp_grid(cols), 256, 0, A, Bv, alpha, thr, rb, flags.get(),
             (const int*)nullptr, (int*)nullptr, (int*)nullptr, (T*)nullptr);
  exclusive_scan(flags.get(), opos.get(), (int)total);
  int h_nnz = 0;
  d2h(&h_nnz, opos.get() + total, 1);
  LocalCsc<T> out;
  out.rows = A.rows; out.cols = cols;
  out.outer.alloc((size_t)cols + 1);
  out.alloc_entries(h_nnz);
  NTB_LAUNCH((k_increment<T, true>), warp_grid(cols), 256, 0, A, Bv, alpha, thr, rb, (int*)nullptr, opos.get(),
             out.outer.get(), out.inner.get(), out.val.get());
  readback_flush();
  CUDA_CHECK(cudaMemcpyAsync(out.outer.get() + cols, opos.get() + total, sizeof(int), cudaMemcpyDeviceToDevice, rt().stream));
  B.swap(out);
}

// ---------------------------------------------------------------------------
template <typename T, bool FILL>
__global__ void __launch_bounds__(256)
k_pairwise(CscView<T> A, CscView<T> B, int* __restrict__ flags, const int* __restrict__ opos,
           int* __restrict__ out_outer, int* __restrict__ out_inner, T* __restrict__ out_val) {
  WARP_COL_LOOP(A.cols) {
    const int a0 = A.outer[j], na = A.outer[j + 1] - a0;
    const int b0 = B.outer[j], nb = B.outer[j + 1] - b0;
    if (FILL && lane == 0) out_outer[j] = opos[a0];
    for (int t = lane; t < na; t += 32) {
      const int ia = A.inner[a0 + t];
      const int pb = lower_bound_dev(B.inner + b0, nb, ia);
      const bool matched = pb < nb && B.inner[b0 + pb] == ia;
      if (!FILL) flags[a0 + t] = matched ? 1 : 0;
      else if (matched) {
        const int o = opos[a0 + t];
        out_inner[o] = ia;
        out_val[o] = s_mul(A.val[a0 + t], B.val[b0 + pb]);
      }
    }
  }
}

template <typename T> void csc_pairwise(const CscView<T>& A, const CscView<T>& B, LocalCsc<T>& C) {
  const int cols = A.cols;
  int nnzA = 0;
  d2h(&nnzA, A.outer + cols, 1);
  DevBuf<int> flags((size_t)nnzA), opos((size_t)nnzA + 1);
  NTB_LAUNCH((k_pairwise<T, false>), warp_grid(cols), 256, 0, A, B, flags.get(), (const int*)nullptr,
             (int*)nullptr, (int*)nullptr, (T*)nullptr);
  exclusive_scan(flags.get(), opos.get(), nnzA);
  int h_nnz = 0;
  d2h(&h_nnz, opos.get() + nnzA, 1);
  C.rows = A.rows; C.cols = cols;
  C.outer.alloc((size_t)cols + 1);
  C.alloc_entries(h_nnz);
  NTB_LAUNCH((k_pairwise<T, true>), warp_grid(cols), 256, 0, A, B, (int*)nullptr, opos.get(), C.outer.get(),
             C.inner.get(), C.val.get());
  readback_flush();
  CUDA_CHECK(cudaMemcpyAsync(C.outer.get() + cols, opos.get() + nnzA, sizeof(int), cudaMemcpyDeviceToDevice, rt().stream));
}

// ---------------------------------------------------------------------------
template <typename T> __global__ void __launch_bounds__(256) k_scale(T* __restrict__ v, long long n, T c) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    v[i] = s_mul(v[i], c);
}
template <typename T> void csc_scale(LocalCsc<T>& M, T c) {
  if (M.nnz == 0) return;
  M.ensure_entries();
  NTB_LAUNCH((k_scale<T>), min(div_up(M.nnz, 256), kNumSMs * 16), 256, 0, M.val.get(), M.nnz, c);
  // cached tile forms: scaled in place when this matrix is their only owner, dropped otherwise
  if constexpr (!scalar_traits<T>::is_complex) {
    if (M.forms && M.forms.use_count() == 1) {
      // a published left form is read in place by the other ranks (peer.h): they must have finished the products
      // enqueued so far before the tiles change, and must not start the next one before the change is complete
      const bool published = !M.forms->left_pub.empty() && M.forms->has_left == 1 && peer().ok;
      if (published) { peer_barrier(); M.forms->left_needs_barrier = true; }
      for (ChunkTiles* t : {&M.forms->left, &M.forms->right}) {
        const int has = (t == &M.forms->left) ? M.forms->has_left : M.forms->has_right;
        const long long n = t->ntiles * 32;
        if (has == 1 && n > 0)
          NTB_LAUNCH((k_scale<double>), min(div_up(n, 256), kNumSMs * 16), 256, 0, t->tval.get(), n, c);
      }
      return;
    }
  }
  M.forms.reset();
}
__global__ void __launch_bounds__(256) k_conj(cplx* __restrict__ v, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    v[i].y = -v[i].y;
}
template <> void csc_conjugate<double>(LocalCsc<double>&) {}
template <> void csc_conjugate<cplx>(LocalCsc<cplx>& M) {
  if (M.nnz == 0) return;
  NTB_LAUNCH(k_conj, min(div_up(M.nnz, 256), kNumSMs * 16), 256, 0, M.val.get(), M.nnz);
  M.forms.reset();
}

// ---------------------------------------------------------------------------
// generic per-entry selection (filter, slice row selection): flags -> scan -> fill
template <typename T, int MODE, bool FILL>   // MODE 0: |v|>thr   MODE 1: (row/rb)%S==s with row renumbering
__global__ void __launch_bounds__(256)
k_select(CscView<T> M, double thr, int rb, int S, int s, int* __restrict__ flags, const int* __restrict__ opos,
         int* __restrict__ out_outer, int* __restrict__ out_inner, T* __restrict__ out_val) {
  WARP_COL_LOOP(M.cols) {
    const int m0 = M.outer[j], m1 = M.outer[j + 1];
    if (FILL && lane == 0) out_outer[j] = opos[m0];
    for (int p = m0 + lane; p < m1; p += 32) {
      const int r = M.inner[p];
      bool keep;
      int rn = r;
      if (MODE == 0) keep = s_abs(M.val[p]) > thr;
      else { const int blk = r / rb; keep = (blk % S) == s; rn = (blk / S) * rb + (r - blk * rb); }
      if (!FILL) flags[p] = keep ? 1 : 0;
      else if (keep) { const int o = opos[p]; out_inner[o] = rn; out_val[o] = M.val[p]; }
    }
  }
}

template <typename T, int MODE>
static void select_impl(const CscView<T>& M, long long nnz, int out_rows, double thr, int rb, int S, int s,
                        LocalCsc<T>& out) {
  const int cols = M.cols;
  DevBuf<int> flags((size_t)nnz), opos((size_t)nnz + 1);
  NTB_LAUNCH((k_select<T, MODE, false>), warp_grid(cols), 256, 0, M, thr, rb, S, s, flags.get(),
             (const int*)nullptr, (int*)nullptr, (int*)nullptr, (T*)nullptr);
  exclusive_scan(flags.get(), opos.get(), (int)nnz);
  int h_nnz = 0;
  d2h(&h_nnz, opos.get() + nnz, 1);
  LocalCsc<T> res;
  res.rows = out_rows; res.cols = cols;
  res.outer.alloc((size_t)cols + 1);
  res.alloc_entries(h_nnz);
  NTB_LAUNCH((k_select<T, MODE, true>), warp_grid(cols), 256, 0, M, thr, rb, S, s, (int*)nullptr, opos.get(),
             res.outer.get(), res.inner.get(), res.val.get());
  readback_flush();
  CUDA_CHECK(cudaMemcpyAsync(res.outer.get() + cols, opos.get() + nnz, sizeof(int), cudaMemcpyDeviceToDevice, rt().stream));
  out.swap(res);
}

template <typename T> void csc_filter(LocalCsc<T>& M, double thr) {
  LocalCsc<T> res;
  select_impl<T, 0>(M.view(), M.nnz, M.rows, thr, 1, 1, 0, res);
  M.swap(res);
}

template <typename T> void csc_select_row_blocks(const CscView<T>& M, int rb, int S, int s, LocalCsc<T>& out) {
  int nnz = 0;
  d2h(&nnz, M.outer + M.cols, 1);
  select_impl<T, 1>(M, nnz, M.rows / S, 0.0, rb, S, s, out);
}

// column-block selection: whole column ranges are copied
template <typename T>
__global__ void __launch_bounds__(256) k_colsel_count(CscView<T> M, int cb, int S, int s, int ncols_out, int* __restrict__ cnt) {
  int jn = blockIdx.x * blockDim.x + threadIdx.x;
  if (jn >= ncols_out) return;
  const int t = jn / cb, off = jn - t * cb;
  const int jo = (s + S * t) * cb + off;
  cnt[jn] = M.outer[jo + 1] - M.outer[jo];
}
template <typename T>
__global__ void __launch_bounds__(256) k_colsel_fill(CscView<T> M, int cb, int S, int s, int ncols_out,
                                                     const int* __restrict__ out_outer, int* __restrict__ out_inner,
                                                     T* __restrict__ out_val) {
  WARP_COL_LOOP(ncols_out) {
    const int t = j / cb, off = j - t * cb;
    const int jo = (s + S * t) * cb + off;
    const int m0 = M.outer[jo], n = M.outer[jo + 1] - m0, o0 = out_outer[j];
    for (int p = lane; p < n; p += 32) { out_inner[o0 + p] = M.inner[m0 + p]; out_val[o0 + p] = M.val[m0 + p]; }
  }
}
template <typename T> void csc_select_col_blocks(const CscView<T>& M, int cb, int S, int s, LocalCsc<T>& out) {
  const int ncols_out = M.cols / S;
  DevBuf<int> cnt((size_t)ncols_out);
  NTB_LAUNCH((k_colsel_count<T>), div_up(ncols_out, 256), 256, 0, M, cb, S, s, ncols_out, cnt.get());
  out.rows = M.rows; out.cols = ncols_out;
  out.outer.alloc((size_t)ncols_out + 1);
  exclusive_scan(cnt.get(), out.outer.get(), ncols_out);
  int h_nnz = 0;
  d2h(&h_nnz, out.outer.get() + ncols_out, 1);
  out.alloc_entries(h_nnz);
  NTB_LAUNCH((k_colsel_fill<T>), warp_grid(ncols_out), 256, 0, M, cb, S, s, ncols_out, out.outer.get(),
             out.inner.get(), out.val.get());
}

// ---------------------------------------------------------------------------
// stack blocks on top of each other (B column panel assembled from the process column)
constexpr int MAX_PARTS = 16;
template <typename T> struct PartList {
  CscView<T> part[MAX_PARTS];
  int row_off[MAX_PARTS];
  int n;
};
// the same list in device memory: any number of parts (a process column of more than MAX_PARTS ranks)
template <typename T> struct PartRef {
  const CscView<T>* part;
  const int* row_off;
  int n;
};
template <typename T, typename L>
__global__ void __launch_bounds__(256) k_stack_count(L P, int cols, int* __restrict__ cnt) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  int c = 0;
  for (int q = 0; q < P.n; ++q) c += P.part[q].outer[j + 1] - P.part[q].outer[j];
  cnt[j] = c;
}
template <typename T, typename L>
__global__ void __launch_bounds__(256) k_stack_fill(L P, int cols, const int* __restrict__ out_outer,
                                                    int* __restrict__ out_inner, T* __restrict__ out_val) {
  WARP_COL_LOOP(cols) {
    int o = out_outer[j];
    for (int q = 0; q < P.n; ++q) {
      const int m0 = P.part[q].outer[j], n = P.part[q].outer[j + 1] - m0;
      for (int p = lane; p < n; p += 32) {
        out_inner[o + p] = P.part[q].inner[m0 + p] + P.row_off[q];
        out_val[o + p] = P.part[q].val[m0 + p];
      }
      o += n;
    }
  }
}
template <typename T>
void csc_stack_rows(const CscView<T>* parts, const int* row_offsets, int n, int total_rows, LocalCsc<T>& out) {
  NTB_CHECK(n >= 1, "stack_rows: no parts");
  const int cols = parts[0].cols;
  DevBuf<int> cnt((size_t)cols);
  out.rows = total_rows; out.cols = cols;
  out.outer.alloc((size_t)cols + 1);
  auto run = [&](auto P) {
    using L = decltype(P);
    NTB_LAUNCH((k_stack_count<T, L>), div_up(cols, 256), 256, 0, P, cols, cnt.get());
    exclusive_scan(cnt.get(), out.outer.get(), cols);
    int h_nnz = 0;
    d2h(&h_nnz, out.outer.get() + cols, 1);
    out.alloc_entries(h_nnz);
    NTB_LAUNCH((k_stack_fill<T, L>), warp_grid(cols), 256, 0, P, cols, out.outer.get(), out.inner.get(), out.val.get());
  };
  // (NTB_STACK_INLINE_PARTS: tests force the device-memory list on grids with two process rows)
  static const int inline_parts = [] { const char* e = std::getenv("NTB_STACK_INLINE_PARTS"); return e ? std::min(std::atoi(e), MAX_PARTS) : MAX_PARTS; }();
  if (n <= inline_parts) {                    // the list travels as a kernel argument
    PartList<T> P;
    P.n = n;
    for (int q = 0; q < n; ++q) { P.part[q] = parts[q]; P.row_off[q] = row_offsets[q]; }
    run(P);
  } else {                                    // more parts than fit into the argument: the list lives in device memory
    DevBuf<CscView<T>> d_parts((size_t)n);
    DevBuf<int> d_off((size_t)n);
    h2d(d_parts.get(), parts, (size_t)n);
    h2d(d_off.get(), row_offsets, (size_t)n);
    stream_sync();                            // (the caller's host arrays are the h2d sources)
    run(PartRef<T>{d_parts.get(), d_off.get(), n});
    stream_sync();                            // d_parts / d_off go back to the arena behind the kernels (stream order), nothing else to wait for
  }
}

// ---------------------------------------------------------------------------
// real <-> complex
__global__ void __launch_bounds__(256) k_r2c(const double* __restrict__ in, cplx* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = cplx{in[i], 0.0};
}
__global__ void __launch_bounds__(256) k_c2r(const cplx* __restrict__ in, double* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = in[i].x;
}
void csc_to_complex(const LocalCsc<double>& in, LocalCsc<cplx>& out) {
  in.ensure_entries();
  out.rows = in.rows; out.cols = in.cols;
  out.outer.alloc((size_t)in.cols + 1);
  d2d(out.outer.get(), in.outer.get(), (size_t)in.cols + 1);
  out.alloc_entries(in.nnz);
  d2d(out.inner.get(), in.inner.get(), (size_t)in.nnz);
  if (in.nnz) NTB_LAUNCH(k_r2c, min(div_up(in.nnz, 256), kNumSMs * 16), 256, 0, in.val.get(), out.val.get(), in.nnz);
}
void csc_to_real(const LocalCsc<cplx>& in, LocalCsc<double>& out) {
  out.rows = in.rows; out.cols = in.cols;
  out.outer.alloc((size_t)in.cols + 1);
  d2d(out.outer.get(), in.outer.get(), (size_t)in.cols + 1);
  out.alloc_entries(in.nnz);
  d2d(out.inner.get(), in.inner.get(), (size_t)in.nnz);
  if (in.nnz) NTB_LAUNCH(k_c2r, min(div_up(in.nnz, 256), kNumSMs * 16), 256, 0, in.val.get(), out.val.get(), in.nnz);
}

// ---------------------------------------------------------------------------
// reductions
template <int OP>  // 0 sum, 1 max, 2 min
__global__ void __launch_bounds__(1024) k_reduce(const double* __restrict__ in, int n, double* __restrict__ out) {
  __shared__ double sw[32];
  double acc = (OP == 0) ? 0.0 : (OP == 1 ? -INFINITY : INFINITY);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double v = in[i];
    acc = (OP == 0) ? acc + v : (OP == 1 ? fmax(acc, v) : fmin(acc, v));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    double o = __shfl_xor_sync(0xffffffffu, acc, d);
    acc = (OP == 0) ? acc + o : (OP == 1 ? fmax(acc, o) : fmin(acc, o));
  }
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = (threadIdx.x < (blockDim.x >> 5)) ? sw[threadIdx.x] : ((OP == 0) ? 0.0 : (OP == 1 ? -INFINITY : INFINITY));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      double o = __shfl_xor_sync(0xffffffffu, acc, d);
      acc = (OP == 0) ? acc + o : (OP == 1 ? fmax(acc, o) : fmin(acc, o));
    }
    if (threadIdx.x == 0) out[0] = acc;
  }
}
template <int OP>
__global__ void __launch_bounds__(1024) k_reduce_chunks(const double* __restrict__ in, int n, double* __restrict__ part) {
  __shared__ double sw[32];
  const int per = (n + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(n, lo + per);
  double acc = (OP == 0) ? 0.0 : (OP == 1 ? -INFINITY : INFINITY);
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const double v = in[i];
    acc = (OP == 0) ? acc + v : (OP == 1 ? fmax(acc, v) : fmin(acc, v));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, acc, d);
    acc = (OP == 0) ? acc + o : (OP == 1 ? fmax(acc, o) : fmin(acc, o));
  }
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = sw[threadIdx.x];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, acc, d);
      acc = (OP == 0) ? acc + o : (OP == 1 ? fmax(acc, o) : fmin(acc, o));
    }
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
  }
}
// two deterministic stages: up to 148 blocks reduce contiguous chunks, one block reduces the partials
template <int OP> static void reduce_impl(const double* d_in, int n, double* d_out) {
  const int chunk = 8192;
  const int nb = min(div_up(max(n, 1), chunk), kNumSMs);
  if (nb <= 1) { NTB_LAUNCH((k_reduce<OP>), 1, 1024, 0, d_in, n, d_out); return; }
  DevBuf<double> part((size_t)nb);
  NTB_LAUNCH((k_reduce_chunks<OP>), nb, 1024, 0, d_in, n, part.get());
  NTB_LAUNCH((k_reduce<OP>), 1, 1024, 0, (const double*)part.get(), nb, d_out);
}
void reduce_sum(const double* d_in, int n, double* d_out) { reduce_impl<0>(d_in, n, d_out); }
void reduce_max(const double* d_in, int n, double* d_out) { reduce_impl<1>(d_in, n, d_out); }
void reduce_min(const double* d_in, int n, double* d_out) { reduce_impl<2>(d_in, n, d_out); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256) k_trace_cols(CscView<T> M, int start_row, int start_col, double* __restrict__ part) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M.cols) return;
  const int r = start_col + j - start_row;  // local row of the global diagonal entry
  double v = 0.0;
  if (r >= 0 && r < M.rows) {
    const int m0 = M.outer[j], n = M.outer[j + 1] - m0;
    const int p = lower_bound_dev(M.inner + m0, n, r);
    if (p < n && M.inner[m0 + p] == r) v = s_real(M.val[m0 + p]);
  }
  part[j] = v;
}
template <typename T> void csc_trace(const CscView<T>& M, int start_row, int start_col, double* d_out) {
  DevBuf<double> part((size_t)M.cols);
  NTB_LAUNCH((k_trace_cols<T>), div_up(M.cols, 256), 256, 0, M, start_row, start_col, part.get());
  reduce_sum(part.get(), M.cols, d_out);
}

template <typename T>
__global__ void __launch_bounds__(256) k_col_abs(CscView<T> M, double* __restrict__ colsum) {
  WARP_COL_LOOP(M.cols) {
    double s = 0.0;
    for (int p = M.outer[j] + lane; p < M.outer[j + 1]; p += 32) s += s_abs(M.val[p]);
    s = warp_sum(s);
    if (lane == 0) colsum[j] = s;
  }
}
template <typename T> void csc_col_abs_sums(const CscView<T>& M, double* d_colsum) {
  NTB_LAUNCH((k_col_abs<T>), warp_grid(M.cols), 256, 0, M, d_colsum);
}

// column sums of |alpha*A + B| without forming the sum (the drivers' "IncrementMatrix then MatrixNorm" on a matrix
// that is overwritten right after, e.g. SignSolversModule.F90:230-232): entries dropped by the add are exact zeros
// and contribute nothing, so the value is the reference's up to summation order
template <typename T>
__global__ void __launch_bounds__(256) k_diff_col_abs(CscView<T> A, CscView<T> B, double alpha, double* __restrict__ colsum) {
  WARP_COL_LOOP(A.cols) {
    const int a0 = A.outer[j], na = A.outer[j + 1] - a0;
    const int b0 = B.outer[j], nb = B.outer[j + 1] - b0;
    const int* ai = A.inner + a0;
    const int* bi = B.inner + b0;
    double s = 0.0;
    // the two patterns are usually almost equal (successive iterates): try the position suggested by the
    // previous chunk's offset before falling back to the binary search
    int delta = 0;
    for (int t0 = 0; t0 < na; t0 += 32) {
      const int t = t0 + lane;
      int pb = 0;
      if (t < na) {
        const int ia = ai[t];
        const int g = t + delta;
        if (g >= 0 && g < nb && bi[g] == ia) pb = g;
        else pb = lower_bound_dev(bi, nb, ia);
        T v = s_scale(alpha, A.val[a0 + t]);
        if (pb < nb && bi[pb] == ia) v = s_add(v, B.val[b0 + pb]);
        s += s_abs(v);
      }
      delta = __shfl_sync(0xffffffffu, pb - t, min(31, na - 1 - t0));
    }
    delta = 0;
    for (int t0 = 0; t0 < nb; t0 += 32) {
      const int t = t0 + lane;
      int pa = 0;
      if (t < nb) {
        const int ib = bi[t];
        const int g = t + delta;
        bool matched;
        if (g >= 0 && g < na && ai[g] == ib) { pa = g; matched = true; }
        else { pa = lower_bound_dev(ai, na, ib); matched = pa < na && ai[pa] == ib; }
        if (!matched) s += s_abs(B.val[b0 + t]);
      }
      delta = __shfl_sync(0xffffffffu, pa - t, min(31, nb - 1 - t0));
    }
    s = warp_sum(s);
    if (lane == 0) colsum[j] = s;
  }
}
template <typename T> void csc_diff_col_abs_sums(const CscView<T>& A, const CscView<T>& B, double alpha, double* d_colsum) {
  NTB_CHECK(A.cols == B.cols && A.rows == B.rows, "diff norm: shape mismatch");
  NTB_LAUNCH((k_diff_col_abs<T>), warp_grid(A.cols), 256, 0, A, B, alpha, d_colsum);
}

template <typename T>
__global__ void __launch_bounds__(256) k_gersh(CscView<T> M, int start_row, int start_col, double* __restrict__ dmin,
                                               double* __restrict__ dmax) {
  WARP_COL_LOOP(M.cols) {
    const int rdiag = start_col + j - start_row;
    double off = 0.0, diag = 0.0;
    for (int p = M.outer[j] + lane; p < M.outer[j + 1]; p += 32) {
      if (M.inner[p] == rdiag) diag += s_real(M.val[p]);
      else off += s_abs(M.val[p]);
    }
    off = warp_sum(off);
    diag = warp_sum(diag);
    if (lane == 0) { dmin[j] = diag - off; dmax[j] = diag + off; }
  }
}
template <typename T>
void csc_gershgorin_cols(const CscView<T>& M, int start_row, int start_col, double* d_min, double* d_max) {
  NTB_LAUNCH((k_gersh<T>), warp_grid(M.cols), 256, 0, M, start_row, start_col, d_min, d_max);
}

template <typename T>
__global__ void __launch_bounds__(256) k_dot_cols(CscView<T> A, CscView<T> B, double* __restrict__ re, double* __restrict__ im) {
  WARP_COL_LOOP(A.cols) {
    const int a0 = A.outer[j], na = A.outer[j + 1] - a0;
    const int b0 = B.outer[j], nb = B.outer[j + 1] - b0;
    double sr = 0.0, si = 0.0;
    for (int t = lane; t < na; t += 32) {
      const int ia = A.inner[a0 + t];
      const int pb = lower_bound_dev(B.inner + b0, nb, ia);
      if (pb < nb && B.inner[b0 + pb] == ia) {
        if constexpr (scalar_traits<T>::is_complex) {
          const cplx pr = s_mul(s_conj(A.val[a0 + t]), B.val[b0 + pb]);
          sr += pr.x; si += pr.y;
        } else {
          sr = fma(A.val[a0 + t], B.val[b0 + pb], sr);
        }
      }
    }
    sr = warp_sum(sr); si = warp_sum(si);
    if (lane == 0) { re[j] = sr; im[j] = si; }
  }
}
template <typename T> void csc_dot(const CscView<T>& A, const CscView<T>& B, double* d_out2) {
  DevBuf<double> re((size_t)A.cols), im((size_t)A.cols);
  NTB_LAUNCH((k_dot_cols<T>), warp_grid(A.cols), 256, 0, A, B, re.get(), im.get());
  reduce_sum(re.get(), A.cols, d_out2);
  reduce_sum(im.get(), A.cols, d_out2 + 1);
}

template <typename T>
__global__ void __launch_bounds__(256) k_ident_cols(CscView<T> M, int start_row, int start_col, double* __restrict__ bad,
                                                    double* __restrict__ ones) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M.cols) return;
  const int rdiag = start_col + j - start_row;
  double b = 0.0, o = 0.0;
  for (int p = M.outer[j]; p < M.outer[j + 1]; ++p) {
    if (M.inner[p] != rdiag) { b = 1.0; continue; }
    T v = M.val[p];
    double dev;
    if constexpr (scalar_traits<T>::is_complex) dev = hypot(v.x - 1.0, v.y); else dev = fabs(v - 1.0);
    if (dev > 2.2250738585072014e-308) b = 1.0; else o += 1.0;
  }
  bad[j] = b; ones[j] = o;
}
template <typename T> void csc_identity_check(const CscView<T>& M, int start_row, int start_col, double* d_out2) {
  DevBuf<double> bad((size_t)M.cols), ones((size_t)M.cols);
  NTB_LAUNCH((k_ident_cols<T>), div_up(M.cols, 256), 256, 0, M, start_row, start_col, bad.get(), ones.get());
  reduce_sum(bad.get(), M.cols, d_out2);
  reduce_sum(ones.get(), M.cols, d_out2 + 1);
}

// ---------------------------------------------------------------------------
// sort-based construction (ingest, transpose). Radix sort = CUB (plumbing, not hot path).
__global__ void __launch_bounds__(256) k_make_keys(const int* __restrict__ major, const int* __restrict__ minor,
                                                   long long n, unsigned long long* __restrict__ keys, int* __restrict__ perm) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    keys[i] = ((unsigned long long)(unsigned)major[i] << 32) | (unsigned)minor[i];
    perm[i] = (int)i;
  }
}
__global__ void __launch_bounds__(256) k_heads(const unsigned long long* __restrict__ keys, long long n, int* __restrict__ head) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
template <typename T>
__global__ void __launch_bounds__(256) k_unique_fill(const unsigned long long* __restrict__ keys, const int* __restrict__ perm,
                                                     const int* __restrict__ head, const int* __restrict__ upos,
                                                     const T* __restrict__ val, long long n, int* __restrict__ out_inner,
                                                     T* __restrict__ out_val, unsigned long long* __restrict__ ukeys) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (!head[i]) continue;
    T acc = val[perm[i]];
    for (long long q = i + 1; q < n && !head[q]; ++q) acc = s_add(acc, val[perm[q]]);  // duplicates are summed
    const int o = upos[i];
    out_inner[o] = (int)(keys[i] & 0xffffffffull);
    out_val[o] = acc;
    ukeys[o] = keys[i];
  }
}
__global__ void __launch_bounds__(256) k_outer_from_keys(const unsigned long long* __restrict__ ukeys, int nu, int cols,
                                                         int* __restrict__ outer) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > cols) return;
  const unsigned long long key = (unsigned long long)(unsigned)j << 32;
  int lo = 0, hi = nu;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (ukeys[mid] < key) lo = mid + 1; else hi = mid; }
  outer[j] = lo;
}

template <typename T>
void csc_from_device_triplets(int rows, int cols, const int* d_row, const int* d_col, const T* d_val, long long n,
                              LocalCsc<T>& out) {
  out.rows = rows; out.cols = cols;
  if (n == 0) { out.init_empty(rows, cols); return; }
  NTB_CHECK(n < (1ll << 31), "too many triplets for one local block");
  DevBuf<unsigned long long> keys((size_t)n), keys_sorted((size_t)n);
  DevBuf<int> perm((size_t)n), perm_sorted((size_t)n);
  const int g = min(div_up(n, 256), kNumSMs * 16);
  NTB_LAUNCH(k_make_keys, g, 256, 0, d_col, d_row, n, keys.get(), perm.get());
  size_t tmp_bytes = 0;
  readback_flush();
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.get(), keys_sorted.get(), perm.get(), perm_sorted.get(),
                                  (int)n, 0, 64, rt().stream);
  DevBuf<unsigned char> tmp(tmp_bytes);
  cub::DeviceRadixSort::SortPairs(tmp.get(), tmp_bytes, keys.get(), keys_sorted.get(), perm.get(), perm_sorted.get(),
                                  (int)n, 0, 64, rt().stream);
  DevBuf<int> head((size_t)n), upos((size_t)n + 1);
  NTB_LAUNCH(k_heads, g, 256, 0, keys_sorted.get(), n, head.get());
  exclusive_scan(head.get(), upos.get(), (int)n);
  int nu = 0;
  d2h(&nu, upos.get() + n, 1);
  out.outer.alloc((size_t)cols + 1);
  out.alloc_entries(nu);
  DevBuf<unsigned long long> ukeys((size_t)nu);
  NTB_LAUNCH((k_unique_fill<T>), g, 256, 0, keys_sorted.get(), perm_sorted.get(), head.get(), upos.get(), d_val, n,
             out.inner.get(), out.val.get(), ukeys.get());
  NTB_LAUNCH(k_outer_from_keys, div_up(cols + 1, 256), 256, 0, ukeys.get(), nu, cols, out.outer.get());
}

template <typename T>
__global__ void __launch_bounds__(256) k_expand(CscView<T> M, int* __restrict__ row, int* __restrict__ col) {
  WARP_COL_LOOP(M.cols) {
    for (int p = M.outer[j] + lane; p < M.outer[j + 1]; p += 32) { row[p] = M.inner[p]; col[p] = j; }
  }
}
template <typename T> void csc_to_device_triplets(const CscView<T>& M, long long nnz, int* d_row, int* d_col) {
  if (nnz == 0) return;
  NTB_LAUNCH((k_expand<T>), warp_grid(M.cols), 256, 0, M, d_row, d_col);
}

template <typename T> void csc_transpose(const CscView<T>& M, LocalCsc<T>& out) {
  int nnz = 0;
  d2h(&nnz, M.outer + M.cols, 1);
  DevBuf<int> row((size_t)nnz), col((size_t)nnz);
  csc_to_device_triplets(M, nnz, row.get(), col.get());
  // transposed block: new column = old row, new row = old column
  LocalCsc<T> res;
  csc_from_device_triplets<T>(M.cols, M.rows, col.get(), row.get(), M.val, nnz, res);
  out.swap(res);
}

// ---------------------------------------------------------------------------
#define INSTANTIATE(T)                                                                                         \
  template void csc_increment<T>(const CscView<T>&, LocalCsc<T>&, double, double, int);                        \
  template void csc_pairwise<T>(const CscView<T>&, const CscView<T>&, LocalCsc<T>&);                           \
  template void csc_scale<T>(LocalCsc<T>&, T);                                                                 \
  template void csc_filter<T>(LocalCsc<T>&, double);                                                           \
  template void csc_transpose<T>(const CscView<T>&, LocalCsc<T>&);                                             \
  template void csc_select_col_blocks<T>(const CscView<T>&, int, int, int, LocalCsc<T>&);                      \
  template void csc_select_row_blocks<T>(const CscView<T>&, int, int, int, LocalCsc<T>&);                      \
  template void csc_stack_rows<T>(const CscView<T>*, const int*, int, int, LocalCsc<T>&);                      \
  template void csc_trace<T>(const CscView<T>&, int, int, double*);                                            \
  template void csc_col_abs_sums<T>(const CscView<T>&, double*);                                               \
  template void csc_diff_col_abs_sums<T>(const CscView<T>&, const CscView<T>&, double, double*);               \
  template void csc_gershgorin_cols<T>(const CscView<T>&, int, int, double*, double*);                         \
  template void csc_dot<T>(const CscView<T>&, const CscView<T>&, double*);                                     \
  template void csc_identity_check<T>(const CscView<T>&, int, int, double*);                                   \
  template void csc_from_device_triplets<T>(int, int, const int*, const int*, const T*, long long, LocalCsc<T>&); \
  template void csc_to_device_triplets<T>(const CscView<T>&, long long, int*, int*);
INSTANTIATE(double)
INSTANTIATE(cplx)

}  // namespace ntb
