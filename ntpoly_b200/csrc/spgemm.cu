// Local thresholded SpGEMM for sm_100a — the B200 replacement of NTPoly's
// MultiplyBlock + PruneList (reference Source/Fortran/sparse_includes/
// MultiplyBlock.f90:9-36, PruneList.f90:8-38) and of the memory pool scratch
// (dense_includes/ConstructMatrixMemoryPool.f90:10-30).
//
// Formulation. NTPoly stores CSC; C = A*B column by column is
//     C(:,j) = sum_k  B(k,j) * A(:,k)          (k ascending, like the reference)
// so with X := B and Y := A every output column is a Gustavson "row" and no
// operand ever has to be transposed (the reference transposes both, every call).
//
// Pipeline (all on the library stream):
//   k_bounds     per output column: product count ub, reachable row window [lo,lo+w)
//   k_binlists   columns -> bins by window size (accumulator kind), chunk-contiguous
//   numeric      bin 1-4: one WARP per column, dense window accumulator in shared memory
//                bin 5  : one CTA per column, shared-memory window up to ~200 KB
//                bin 6  : one CTA per column, window in a per-CTA global slab (L2 resident)
//                bin 7  : one warp per column, SHARED-MEMORY HASH accumulator: scattered columns whose row window
//                         is wide but whose product count is small (graph-like patterns, permuted / load-balanced
//                         matrices) - work proportional to the products, not to the window
//                each: accumulate in k order, then sweep the window in row order applying
//                the drop rule (|alpha*v|>thr or dense-branch |v|>thr) -> sorted, filtered,
//                alpha-scaled entries written to a staging area at a bound-derived offset
//   scan + k_compact   exact CSC of the kept entries
// The window sweep makes sort and filter free: no hash tables, no atomics, and the
// summation order per element is the reference's (k ascending).
#include "csc.cuh"

namespace ntb {

bool spgemm_tile(const LocalCsc<double>& X, const LocalCsc<double>& Y, double alpha, double thr, const RuleView& rules,
                 LocalCsc<double>& Z, double useful_products, const DiagShift* shift, unsigned want);

static int g_tile_mode = -1;   // -1: read NTB_TILE once; 0 off; 1 on
void set_tile_path(int on) { g_tile_mode = on ? 1 : 0; }
static bool tile_path_enabled() {
  if (g_tile_mode < 0) { const char* e = std::getenv("NTB_TILE"); g_tile_mode = (e && e[0] == '0') ? 0 : 1; }
  return g_tile_mode == 1;
}

bool tile_path_on() { return tile_path_enabled(); }

constexpr int NBINS = 8;       // 0: empty column, 1..4 warp windows, 5 CTA window, 6 global slab, 7 warp hash
constexpr int HASH_UB = 1024;  // bin 7: at most this many products per column ...
constexpr int HASH_H = 2048;   // ... in a table of this many slots per warp (load factor <= 1/2)
constexpr int HASH_WARPS = 4;  // warps (= columns in flight) per CTA of the hash kernel
constexpr int WARPS = 8;       // warps per CTA in the warp-window kernels
constexpr int CTA_T = 256;     // threads per CTA in the CTA-window kernels
constexpr size_t SMEM_BUDGET = 200 * 1024;

struct BinCfg { int wmax[NBINS]; };

__device__ __forceinline__ bool keep_entry(double mag_scaled, double mag_raw, double thr, bool dense_rule) {
  return (dense_rule ? mag_raw : mag_scaled) > thr;
}

__device__ __forceinline__ bool rule_for(const RuleView& r, int inner_idx, int outer_idx) {
  if (r.tbl == nullptr) return false;
  return r.tbl[(inner_idx / r.rb) * r.nJ + (outer_idx / r.cb)] != 0;
}

// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_bounds(CscView<T> X, CscView<T> Y, int* __restrict__ lo,
                                                int* __restrict__ wid, int* __restrict__ cap,
                                                int* __restrict__ binid, BinCfg cfg,
                                                unsigned long long* __restrict__ flops) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  unsigned long long my_flops = 0;
  for (int j = gw; j < X.cols; j += nw) {
    unsigned long long ub = 0;
    int mn = INT_MAX, mx = -1;
    for (int p = X.outer[j] + lane; p < X.outer[j + 1]; p += 32) {
      int k = X.inner[p];
      int s = Y.outer[k], e = Y.outer[k + 1];
      if (e > s) {
        ub += (unsigned long long)(e - s);
        mn = min(mn, Y.inner[s]);
        mx = max(mx, Y.inner[e - 1]);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      ub += __shfl_xor_sync(0xffffffffu, ub, d);
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if (lane == 0) {
      int w = (ub > 0) ? (mx - mn + 1) : 0;
      lo[j] = (ub > 0) ? mn : 0;
      wid[j] = w;
      cap[j] = (int)min((unsigned long long)w, ub);
      int b = 0;
      if (w > 0) {
        b = 6;
#pragma unroll
        for (int t = 5; t >= 1; --t)
          if (w <= cfg.wmax[t]) b = t;
        // a wide window with few products: the hash accumulator (sweeping the window would cost more than the products)
        if (cfg.wmax[7] > 0 && w > cfg.wmax[4] && ub <= (unsigned long long)HASH_UB && ub * 4ull < (unsigned long long)w) b = 7;
      }
      binid[j] = b;
      my_flops += ub;
    }
  }
  if (lane == 0 && my_flops) atomicAdd(flops, my_flops);
}

// chunk-contiguous bin lists: 32 consecutive columns stay adjacent inside a bin so
// that the columns one CTA works on share their Y columns in L1/L2.
__global__ void __launch_bounds__(256) k_binlists(const int* __restrict__ binid, int n, int* __restrict__ lists,
                                                  int* __restrict__ bin_count) {
  const int lane = threadIdx.x & 31;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int b = (j < n) ? binid[j] : -1;
#pragma unroll
  for (int t = 1; t < NBINS; ++t) {
    unsigned m = __ballot_sync(0xffffffffu, b == t);
    if (m == 0) continue;
    int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&bin_count[t], __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (b == t) lists[(size_t)t * n + base + __popc(m & ((1u << lane) - 1))] = j;
  }
}

// ---------------------------------------------------------------------------
// bins 1..4: one warp per output column, window accumulator in shared memory
template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
k_numeric_warp(CscView<T> X, CscView<T> Y, const int* __restrict__ list, const int* __restrict__ nlist_p,
               const int* __restrict__ lo, const int* __restrict__ wid,
               const long long* __restrict__ tmp_off, double alpha, double thr, RuleView rules,
               int* __restrict__ tmp_idx, T* __restrict__ tmp_val, int* __restrict__ cnt, int wmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* acc = reinterpret_cast<T*>(smem_raw) + (size_t)warp * wmax;
  const int nlist = *nlist_p;
  for (int li = blockIdx.x * WARPS + warp; li < nlist; li += gridDim.x * WARPS) {
    const int j = list[li];
    const int base = lo[j];
    const int w = wid[j];
    for (int t = lane; t < w; t += 32) acc[t] = zero_of<T>();
    __syncwarp();
    const int xs = X.outer[j], xe = X.outer[j + 1];
    for (int p0 = xs; p0 < xe; p0 += 32) {
      const int p = p0 + lane;
      int ys = 0, ye = 0;
      T xv = zero_of<T>();
      if (p < xe) {
        const int k = X.inner[p];
        xv = X.val[p];
        ys = Y.outer[k];
        ye = Y.outer[k + 1];
      }
      const int nv = min(32, xe - p0);
      for (int t = 0; t < nv; ++t) {
        const int s = shfl(ys, t), e = shfl(ye, t);
        const T b = shfl(xv, t);
        for (int q = s + lane; q < e; q += 32) {
          const int c = Y.inner[q] - base;
          acc[c] = s_fma(Y.val[q], b, acc[c]);
        }
        __syncwarp();
      }
    }
    // ordered sweep: filter + alpha + emit (sorted by construction)
    const long long off = tmp_off[j];
    int count = 0;
    for (int t0 = 0; t0 < w; t0 += 32) {
      const int t = t0 + lane;
      bool keep = false;
      T v = zero_of<T>();
      if (t < w) {
        v = acc[t];
        const T sv = s_scale(alpha, v);
        keep = keep_entry(s_abs(sv), s_abs(v), thr, rule_for(rules, base + t, j));
        v = sv;
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const long long pos = off + count + __popc(m & ((1u << lane) - 1));
        tmp_idx[pos] = base + t;
        tmp_val[pos] = v;
      }
      count += __popc(m);
    }
    if (lane == 0) cnt[j] = count;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// bins 5/6: one CTA per output column; window in shared memory (bin 5) or in a
// per-CTA global slab kept all-zero between columns (bin 6).
template <typename T, bool GLOBAL_SLAB>
__global__ void __launch_bounds__(CTA_T)
k_numeric_cta(CscView<T> X, CscView<T> Y, const int* __restrict__ list, const int* __restrict__ nlist_p,
              const int* __restrict__ lo, const int* __restrict__ wid,
              const long long* __restrict__ tmp_off, double alpha, double thr, RuleView rules,
              int* __restrict__ tmp_idx, T* __restrict__ tmp_val, int* __restrict__ cnt, T* __restrict__ slab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_warp_cnt[CTA_T / 32];
  __shared__ int s_running;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nlist = *nlist_p;
  for (int li = blockIdx.x; li < nlist; li += gridDim.x) {
    const int j = list[li];
    const int base = lo[j];
    const int w = wid[j];
    T* acc = GLOBAL_SLAB ? (slab + (size_t)blockIdx.x * Y.rows + base) : reinterpret_cast<T*>(smem_raw);
    if (!GLOBAL_SLAB) {
      for (int t = threadIdx.x; t < w; t += CTA_T) acc[t] = zero_of<T>();
    }
    __syncthreads();
    const int xs = X.outer[j], xe = X.outer[j + 1];
    for (int p = xs; p < xe; ++p) {
      const int k = X.inner[p];
      const T b = X.val[p];
      const int s = Y.outer[k], e = Y.outer[k + 1];
      for (int q = s + threadIdx.x; q < e; q += CTA_T) {
        const int c = Y.inner[q] - base;
        acc[c] = s_fma(Y.val[q], b, acc[c]);
      }
      __syncthreads();
    }
    const long long off = tmp_off[j];
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (int t0 = 0; t0 < w; t0 += CTA_T) {
      const int t = t0 + threadIdx.x;
      bool keep = false;
      T v = zero_of<T>();
      if (t < w) {
        v = acc[t];
        if (GLOBAL_SLAB) acc[t] = zero_of<T>();
        const T sv = s_scale(alpha, v);
        keep = keep_entry(s_abs(sv), s_abs(v), thr, rule_for(rules, base + t, j));
        v = sv;
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_warp_cnt[warp] = __popc(m);
      __syncthreads();
      int before = s_running;
      for (int ww = 0; ww < warp; ++ww) before += s_warp_cnt[ww];
      if (keep) {
        const long long pos = off + before + __popc(m & ((1u << lane) - 1));
        tmp_idx[pos] = base + t;
        tmp_val[pos] = v;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int tot = 0;
        for (int ww = 0; ww < CTA_T / 32; ++ww) tot += s_warp_cnt[ww];
        s_running += tot;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) cnt[j] = s_running;
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------
// bin 7: one warp per output column, hash accumulator in shared memory. Four phases, all deterministic:
//   A  symbolic: the rows of every Y(:,k), k in X(:,j), are inserted into an open-addressing table (keys only);
//   B  the distinct rows are compacted, sorted (bitonic, in shared memory) and every table slot learns the RANK of
//      its row in that order;
//   C  numeric: k ascending like the reference; the rows of one Y(:,k) are distinct, so the lanes of the warp update
//      distinct accumulators acc[rank] - no atomics, the summation order per entry is the reference's;
//   D  ordered sweep over the ranks: threshold rule, alpha, emit - sorted output for free, as in the window kernels.
// Replaces the O(window) sweep per column of bins 5/6 for scattered patterns (the reference's bucketed hash_index,
// sparse_includes/MultiplyBlock.f90:20-33, has the same purpose).
template <typename T> struct HashSmem {
  int keys[HASH_H];
  int rank[HASH_H];
  int sorted[HASH_UB];
  T acc[HASH_UB];
};
__device__ __forceinline__ unsigned hash_slot(int r) { return ((unsigned)r * 2654435761u) >> (32 - 11); }
static_assert(HASH_H == (1 << 11), "hash_slot assumes 2^11 slots");

template <typename T>
__global__ void __launch_bounds__(HASH_WARPS * 32)
k_numeric_hash(CscView<T> X, CscView<T> Y, const int* __restrict__ list, const int* __restrict__ nlist_p,
               const long long* __restrict__ tmp_off, double alpha, double thr, RuleView rules,
               int* __restrict__ tmp_idx, T* __restrict__ tmp_val, int* __restrict__ cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  HashSmem<T>& S = reinterpret_cast<HashSmem<T>*>(smem_raw)[warp];
  const int nlist = *nlist_p;
  for (int li = blockIdx.x * HASH_WARPS + warp; li < nlist; li += gridDim.x * HASH_WARPS) {
    const int j = list[li];
    for (int t = lane; t < HASH_H; t += 32) S.keys[t] = -1;
    __syncwarp();
    const int xs = X.outer[j], xe = X.outer[j + 1];
    // ---- A: keys
    for (int p = xs; p < xe; ++p) {
      const int k = X.inner[p];
      const int s = Y.outer[k], e = Y.outer[k + 1];
      for (int q = s + lane; q < e; q += 32) {
        const int r = Y.inner[q];
        unsigned slot = hash_slot(r);
        for (;;) {
          const int old = atomicCAS(&S.keys[slot], -1, r);
          if (old == -1 || old == r) break;
          slot = (slot + 1u) & (HASH_H - 1);
        }
      }
    }
    __syncwarp();
    // ---- B: compact, sort, ranks
    int nd = 0;
    for (int t0 = 0; t0 < HASH_H; t0 += 32) {
      const int key = S.keys[t0 + lane];
      const unsigned m = __ballot_sync(0xffffffffu, key >= 0);
      if (key >= 0) S.sorted[nd + __popc(m & ((1u << lane) - 1u))] = key;
      nd += __popc(m);
    }
    int np2 = 32;
    while (np2 < nd) np2 <<= 1;
    for (int t = nd + lane; t < np2; t += 32) S.sorted[t] = INT_MAX;
    __syncwarp();
    for (int kk = 2; kk <= np2; kk <<= 1)
      for (int jj = kk >> 1; jj > 0; jj >>= 1) {
        for (int t = lane; t < np2; t += 32) {
          const int partner = t ^ jj;
          if (partner > t) {
            const int a = S.sorted[t], b = S.sorted[partner];
            const bool up = (t & kk) == 0;
            if ((a > b) == up) { S.sorted[t] = b; S.sorted[partner] = a; }
          }
        }
        __syncwarp();
      }
    for (int t = lane; t < nd; t += 32) {
      const int r = S.sorted[t];
      unsigned slot = hash_slot(r);
      while (S.keys[slot] != r) slot = (slot + 1u) & (HASH_H - 1);
      S.rank[slot] = t;
      S.acc[t] = zero_of<T>();
    }
    __syncwarp();
    // ---- C: numeric, k ascending
    for (int p = xs; p < xe; ++p) {
      const int k = X.inner[p];
      const T b = X.val[p];
      const int s = Y.outer[k], e = Y.outer[k + 1];
      for (int q = s + lane; q < e; q += 32) {
        const int r = Y.inner[q];
        unsigned slot = hash_slot(r);
        while (S.keys[slot] != r) slot = (slot + 1u) & (HASH_H - 1);
        const int a = S.rank[slot];
        S.acc[a] = s_fma(Y.val[q], b, S.acc[a]);
      }
      __syncwarp();
    }
    // ---- D: ordered sweep
    const long long off = tmp_off[j];
    int count = 0;
    for (int t0 = 0; t0 < nd; t0 += 32) {
      const int t = t0 + lane;
      bool keep = false;
      T v = zero_of<T>();
      int row = 0;
      if (t < nd) {
        v = S.acc[t];
        row = S.sorted[t];
        const T sv = s_scale(alpha, v);
        keep = keep_entry(s_abs(sv), s_abs(v), thr, rule_for(rules, row, j));
        v = sv;
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const long long pos = off + count + __popc(m & ((1u << lane) - 1));
        tmp_idx[pos] = row;
        tmp_val[pos] = v;
      }
      count += __popc(m);
    }
    if (lane == 0) cnt[j] = count;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_compact(int ncols, const long long* __restrict__ tmp_off,
                                                 const int* __restrict__ cnt, const int* __restrict__ outer,
                                                 const int* __restrict__ tmp_idx, const T* __restrict__ tmp_val,
                                                 int* __restrict__ inner, T* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int j = gw; j < ncols; j += nw) {
    const long long src = tmp_off[j];
    const int dst = outer[j];
    const int n = cnt[j];
    for (int t = lane; t < n; t += 32) {
      inner[dst + t] = tmp_idx[src + t];
      val[dst + t] = tmp_val[src + t];
    }
  }
}

// useful products of the local product: sum over the entries (k,j) of X of len(Y(:,k))  (F = 2 * this)
template <typename T>
__global__ void __launch_bounds__(256) k_useful_products(CscView<T> X, CscView<T> Y, long long nnzX,
                                                         unsigned long long* __restrict__ total) {
  unsigned long long s = 0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nnzX; p += (long long)gridDim.x * blockDim.x) {
    const int k = X.inner[p];
    s += (unsigned long long)(Y.outer[k + 1] - Y.outer[k]);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

__global__ void __launch_bounds__(256) k_useful_products_len(const int* __restrict__ xinner, const int* __restrict__ ylen,
                                                             long long nnzX, unsigned long long* __restrict__ total) {
  unsigned long long s = 0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nnzX; p += (long long)gridDim.x * blockDim.x)
    s += (unsigned long long)ylen[xinner[p]];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}
double useful_products_from_lengths(const LocalCsc<double>& X, const int* d_ylen) {
  if (X.nnz == 0) return 0.0;
  DevBuf<unsigned long long> fl(1);
  fl.zero();
  NTB_LAUNCH(k_useful_products_len, min(div_up(X.nnz, 256 * 8), kNumSMs * 16), 256, 0, X.inner.get(), d_ylen, X.nnz, fl.get());
  unsigned long long h = 0;
  d2h(&h, fl.get(), 1);
  return (double)h;
}

// ---------------------------------------------------------------------------
template <typename T>
void spgemm(const LocalCsc<T>& Xl, const LocalCsc<T>& Yl, double alpha, double thr, const RuleView& rules,
            LocalCsc<T>& Z, GemmStats* stats, const DiagShift* shift, unsigned want) {
  NTB_CHECK(Xl.rows == Yl.cols, "spgemm: inner dimensions differ");
  if (stats) stats->shift_applied = false;
  const int ncols = Xl.cols;
  const int nrows = Yl.rows;

  auto csc_bytes = [](long long nnz, int cols) { return (double)nnz * (sizeof(T) + 4) + ((double)cols + 1) * 4; };
  auto account_bytes = [&](long long nnz_out) {
    double b = csc_bytes(Xl.nnz, Xl.cols) + csc_bytes(nnz_out, ncols);
    if (Yl.outer.get() != Xl.outer.get()) b += csc_bytes(Yl.nnz, Yl.cols);   // A counted once when A == B
    rt().alg_bytes += b;
  };

  // ---- operands that already carry their tile forms (results of earlier tile products): with the flop accounting
  // switched off nothing needs the CSC entries, which may even be deferred (LocalCsc::deferred)
  if constexpr (!scalar_traits<T>::is_complex) {
    if (tile_path_enabled() && !rt().count_flops && ncols > 0 && Xl.nnz.maybe_nonzero() && Yl.nnz.maybe_nonzero() && Xl.forms && Yl.forms &&
        Xl.forms->has_right == 1 && Yl.forms->has_left == 1 &&
        spgemm_tile_core(left_view_of(Yl.forms->left), Xl.forms->right, ncols, nrows, alpha, thr, rules, Z, -1.0, shift, false, want)) {
      if (stats) { stats->shift_applied = shift && shift->sigma != 0.0; stats->flops = 0.0; stats->tmp_entries = 0; }
      // (the counts of a deferred product, or of deferred operands, may still be on their way: accounted when they land)
      account_product_bytes(Xl.nnz, Xl.cols, Yl.nnz, Yl.cols, Yl.outer.get() != Xl.outer.get(), Z.nnz, ncols, sizeof(T));
      return;
    }
  }

  const CscView<T> X = Xl.view(), Y = Yl.view();
  Z.rows = nrows;
  Z.cols = ncols;
  Z.outer.alloc((size_t)ncols + 1);
  if (ncols == 0) { Z.alloc_entries(0); return; }

  // bin configuration by scalar width (bytes of shared memory per window)
  BinCfg cfg;
  const int per_warp_elems[5] = {0, (int)(4096 / sizeof(T)), (int)(8192 / sizeof(T)),
                                 (int)(16384 / sizeof(T)), (int)(24576 / sizeof(T))};
  cfg.wmax[0] = 0;
  for (int b = 1; b <= 4; ++b) cfg.wmax[b] = per_warp_elems[b];
  cfg.wmax[5] = (int)(SMEM_BUDGET / sizeof(T));
  cfg.wmax[6] = INT_MAX;
  static const bool hash_on = [] { const char* e = std::getenv("NTB_HASH_BIN"); return !(e && e[0] == '0'); }();
  cfg.wmax[7] = hash_on ? 1 : 0;             // (a switch, not a width: bin 7 is chosen by product count, see k_bounds)

  // ---- tile path first: it needs only the useful-product count, not the per-column windows
  if constexpr (!scalar_traits<T>::is_complex) {
    if (tile_path_enabled() && Xl.nnz > 0 && Yl.nnz > 0) {
      DevBuf<unsigned long long> fl(1);
      fl.zero();
      NTB_LAUNCH((k_useful_products<T>), min(div_up(Xl.nnz, 256 * 8), kNumSMs * 16), 256, 0, X, Y, Xl.nnz, fl.get());
      unsigned long long h_fl = 0;
      d2h(&h_fl, fl.get(), 1);
      // worth it only when columns are long enough to fill tiles
      if ((double)h_fl >= 16.0 * (double)ncols && spgemm_tile(Xl, Yl, alpha, thr, rules, Z, (double)h_fl, shift, want)) {
        if (stats) {
          stats->shift_applied = shift && shift->sigma != 0.0;
          stats->flops = 2.0 * (double)h_fl;
          stats->tmp_entries = 0;
        }
        account_bytes(Z.nnz);
        return;
      }
    }
  }

  DevBuf<int> lo(ncols), wid(ncols), cap(ncols), binid(ncols), cnt(ncols);
  DevBuf<int> lists((size_t)NBINS * ncols);
  DevBuf<int> bin_count(NBINS);
  DevBuf<unsigned long long> flops(1);
  DevBuf<long long> tmp_off((size_t)ncols + 1);
  bin_count.zero();
  flops.zero();
  cnt.zero();

  {
    int blocks = min(div_up((long long)ncols * 32, 256), kNumSMs * 16);
    NTB_LAUNCH((k_bounds<T>), blocks, 256, 0, X, Y, lo.get(), wid.get(), cap.get(), binid.get(), cfg, flops.get());
    NTB_LAUNCH(k_binlists, div_up(ncols, 256), 256, 0, binid.get(), ncols, lists.get(), bin_count.get());
  }
  exclusive_scan(cap.get(), tmp_off.get(), ncols);

  // one small read-back: bin populations, staging size, flop count
  int h_bins[NBINS];
  long long h_tmp_total = 0;
  unsigned long long h_flops = 0;
  readback_async(h_bins, bin_count.get(), sizeof(h_bins));
  readback_async(&h_tmp_total, tmp_off.get() + ncols, sizeof(long long));
  readback_async(&h_flops, flops.get(), sizeof(h_flops));
  stream_sync();

  auto account = [&](long long nnz_out) {
    account_bytes(nnz_out);
    if (stats) {
      stats->flops = 2.0 * (double)h_flops * (scalar_traits<T>::is_complex ? 4.0 : 1.0);
      stats->tmp_entries = h_tmp_total;
      for (int b2 = 0; b2 < NBINS; ++b2) stats->bins[b2] = h_bins[b2];
    }
  };

  DevBuf<int> tmp_idx((size_t)h_tmp_total);
  DevBuf<T> tmp_val((size_t)h_tmp_total);

  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (rt().profile) {
    CUDA_CHECK(cudaEventCreate(&ev0));
    CUDA_CHECK(cudaEventCreate(&ev1));
    CUDA_CHECK(cudaEventRecord(ev0, rt().stream));
  }

  for (int b = 1; b <= 4; ++b) {
    if (h_bins[b] == 0) continue;
    const size_t smem = (size_t)WARPS * cfg.wmax[b] * sizeof(T);
    CUDA_CHECK(cudaFuncSetAttribute(k_numeric_warp<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUDGET));
    int per_sm = (int)max((size_t)1, min((size_t)8, (size_t)(220 * 1024) / (smem + 1024)));
    int blocks = min(div_up(h_bins[b], WARPS), kNumSMs * per_sm);
    NTB_LAUNCH((k_numeric_warp<T>), blocks, WARPS * 32, smem, X, Y, lists.get() + (size_t)b * ncols,
               bin_count.get() + b, lo.get(), wid.get(), tmp_off.get(), alpha, thr, rules, tmp_idx.get(),
               tmp_val.get(), cnt.get(), cfg.wmax[b]);
  }
  if (h_bins[5] > 0) {
    CUDA_CHECK(cudaFuncSetAttribute((k_numeric_cta<T, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUDGET));
    int blocks = min(h_bins[5], kNumSMs * 4);
    NTB_LAUNCH((k_numeric_cta<T, false>), blocks, CTA_T, SMEM_BUDGET, X, Y, lists.get() + (size_t)5 * ncols,
               bin_count.get() + 5, lo.get(), wid.get(), tmp_off.get(), alpha, thr, rules, tmp_idx.get(),
               tmp_val.get(), cnt.get(), (T*)nullptr);
  }
  if (h_bins[7] > 0) {
    const size_t smem = sizeof(HashSmem<T>) * HASH_WARPS;
    CUDA_CHECK(cudaFuncSetAttribute((k_numeric_hash<T>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = min(div_up(h_bins[7], HASH_WARPS), kNumSMs * 8);
    NTB_LAUNCH((k_numeric_hash<T>), blocks, HASH_WARPS * 32, smem, X, Y, lists.get() + (size_t)7 * ncols, bin_count.get() + 7,
               tmp_off.get(), alpha, thr, rules, tmp_idx.get(), tmp_val.get(), cnt.get());
    rt().hash_columns += (unsigned long long)h_bins[7];
  }
  DevBuf<T> slab;
  if (h_bins[6] > 0) {
    int blocks = min(h_bins[6], kNumSMs * 2);
    slab.alloc((size_t)blocks * nrows);
    slab.zero();
    NTB_LAUNCH((k_numeric_cta<T, true>), blocks, CTA_T, 0, X, Y, lists.get() + (size_t)6 * ncols,
               bin_count.get() + 6, lo.get(), wid.get(), tmp_off.get(), alpha, thr, rules, tmp_idx.get(),
               tmp_val.get(), cnt.get(), slab.get());
  }

  if (rt().profile) {
    CUDA_CHECK(cudaEventRecord(ev1, rt().stream));
    rt().prof_events.emplace_back(ev0, ev1);
  }
  exclusive_scan(cnt.get(), Z.outer.get(), ncols);
  int h_nnz = 0;
  d2h(&h_nnz, Z.outer.get() + ncols, 1);
  Z.alloc_entries(h_nnz);
  if (h_nnz > 0) {
    int blocks = min(div_up((long long)ncols * 32, 256), kNumSMs * 16);
    NTB_LAUNCH((k_compact<T>), blocks, 256, 0, ncols, tmp_off.get(), cnt.get(), Z.outer.get(), tmp_idx.get(),
               tmp_val.get(), Z.inner.get(), Z.val.get());
  }
  account(h_nnz);
}

template void spgemm<double>(const LocalCsc<double>&, const LocalCsc<double>&, double, double, const RuleView&,
                             LocalCsc<double>&, GemmStats*, const DiagShift*, unsigned);
template void spgemm<cplx>(const LocalCsc<cplx>&, const LocalCsc<cplx>&, double, double, const RuleView&,
                           LocalCsc<cplx>&, GemmStats*, const DiagShift*, unsigned);

}  // namespace ntb
