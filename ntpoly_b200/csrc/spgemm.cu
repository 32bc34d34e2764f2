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
//                bin 6  : one CTA per column, window in a per-CTA global slab
//                         (bins 5/6 by default through k_numeric_cta_atomic: 32 k per warp, floating-point atomics into
//                         the window, barrier-free ordered sweep; NTB_ATOMIC_BINS=0: the serial-k kernel k_numeric_cta)
//                bin 7  : one warp per column, SHARED-MEMORY HASH accumulator: scattered columns whose row window
//                         is wide but whose product count is small (graph-like patterns, permuted / load-balanced
//                         matrices) - work proportional to the products, not to the window
//                each: accumulate in k order, then sweep the window in row order applying
//                the drop rule (|alpha*v|>thr or dense-branch |v|>thr) -> sorted, filtered,
//                alpha-scaled entries written to a staging area at a bound-derived offset
//   scan + k_compact   exact CSC of the kept entries
// The window sweep makes sort and filter free. Bins 1-4 and 7 (and the serial-k variant of 5/6) use no atomics and sum
// every element in the reference's order (k ascending); complex operands that are locally dense take the FP64
// tensor-core tile path through their real embeddings instead (spgemm_complex_tiles below).
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

// |v| for the drop test of the heavy-column kernel: sqrt(x^2 + y^2) while the squares are comfortably inside the
// double range, hypot() otherwise (the library routine is ~60 instructions and was most of the 8e9 instructions of
// a c5 sweep launch, profiles/r02g_c5_atomic_slab.keys.txt); differs from hypot() by an ulp at most, i.e. only for
// entries that sit on the threshold itself
__device__ __forceinline__ double mag_fast(double v) { return fabs(v); }
__device__ __forceinline__ double mag_fast(cplx v) {
  const double m2 = v.x * v.x + v.y * v.y;
  return (m2 > 1e-280 && m2 < 1e280) ? sqrt(m2) : hypot(v.x, v.y);
}
__device__ __forceinline__ bool is_zero(double v) { return v == 0.0; }
__device__ __forceinline__ bool is_zero(cplx v) { return v.x == 0.0 && v.y == 0.0; }

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
// bins 5/6, HEAVY columns (many products into a wide window: the iterates of graph-like / permuted matrices, c5): the
// k loop of k_numeric_cta is serial with a block-wide barrier per k, which leaves the CTA idle when the Y columns
// are short (25 entries under 256 threads) or when there are thousands of k. Here every warp takes 32 consecutive k at
// a time (coalesced loads of the X entries and of the Y column bounds) and sends its products to the accumulator window
// with floating-point atomic adds (RED.ADD.F64 in L2 for the global slab, shared-memory atomics for bin 5). The sum of
// an entry is then formed in arrival order instead of k-ascending order: still within rounding of the reference (the
// parity bound is 1e-10 relative), but no longer bit-reproducible from run to run for these columns -
// NTB_ATOMIC_BINS=0 selects the serial kernel. The ordered sweep (threshold rule, alpha, sorted emit) is unchanged.
// (global slab: RED.E.ADD.F64 - a reduction without a return value, so the warp does not wait for the round trip to L2;
// atomicAdd() with an unused result still compiled to ATOMG here)
template <bool GLOBAL> __device__ __forceinline__ void acc_add1(double* a, double v) {
  if (GLOBAL) asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory");
  else atomicAdd(a, v);
}
template <bool GLOBAL> __device__ __forceinline__ void acc_add(double* a, double v) { acc_add1<GLOBAL>(a, v); }
template <bool GLOBAL> __device__ __forceinline__ void acc_add(cplx* a, cplx v) { acc_add1<GLOBAL>(&a->x, v.x); acc_add1<GLOBAL>(&a->y, v.y); }
__device__ __forceinline__ double ld_acc(const double* a, bool global) { return global ? __ldcg(a) : *a; }
__device__ __forceinline__ cplx ld_acc(const cplx* a, bool global) {
  if (!global) return *a;
  const double2 t = __ldcg(reinterpret_cast<const double2*>(a));
  return cplx{t.x, t.y};
}
template <typename T, bool GLOBAL_SLAB>
__global__ void __launch_bounds__(CTA_T)
k_numeric_cta_atomic(CscView<T> X, CscView<T> Y, const int* __restrict__ list, const int* __restrict__ nlist_p,
                     const int* __restrict__ lo, const int* __restrict__ wid, const long long* __restrict__ tmp_off,
                     double alpha, double thr, RuleView rules, int* __restrict__ tmp_idx, T* __restrict__ tmp_val,
                     int* __restrict__ cnt, T* __restrict__ slab) {
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
    // entries s0 + start, + stride, ... of one Y column, four at a time: the loads of a group are issued before its
    // first reduction (a single load per iteration left the warp waiting one L2 latency per product)
    auto walk = [&](int s0, int e0, T b, int start, int stride) {
      int q = s0 + start;
      for (; q + 3 * stride < e0; q += 4 * stride) {
        const int i0 = Y.inner[q], i1 = Y.inner[q + stride], i2 = Y.inner[q + 2 * stride], i3 = Y.inner[q + 3 * stride];
        const T v0 = Y.val[q], v1 = Y.val[q + stride], v2 = Y.val[q + 2 * stride], v3 = Y.val[q + 3 * stride];
        acc_add<GLOBAL_SLAB>(&acc[i0 - base], s_mul(v0, b));
        acc_add<GLOBAL_SLAB>(&acc[i1 - base], s_mul(v1, b));
        acc_add<GLOBAL_SLAB>(&acc[i2 - base], s_mul(v2, b));
        acc_add<GLOBAL_SLAB>(&acc[i3 - base], s_mul(v3, b));
      }
      for (; q < e0; q += stride) acc_add<GLOBAL_SLAB>(&acc[Y.inner[q] - base], s_mul(Y.val[q], b));
    };
    // a SHORT X column (c5: 26 entries of G against 650-entry columns of the iterate) would occupy one warp only:
    // then all eight warps take the same 32 k and share every Y column, 256 entries per sweep
    const bool shared_k = (xe - xs) <= 64;
    for (int p0 = xs + (shared_k ? 0 : warp * 32); p0 < xe; p0 += (shared_k ? 32 : CTA_T)) {
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
      // short Y columns (<= 16 entries on average): one lane per X entry walks its own column (no idle lanes);
      // otherwise the warp (or the CTA) strides over one column at a time
      const int total = __reduce_add_sync(0xffffffffu, ye - ys);
      if (total <= 16 * nv) {
        if (!shared_k || warp == 0) walk(ys, ye, xv, 0, 1);
      } else {
        for (int t = 0; t < nv; ++t) {
          const int s0 = shfl(ys, t), e0 = shfl(ye, t);
          const T b = shfl(xv, t);
          if (shared_k) walk(s0, e0, b, (int)threadIdx.x, CTA_T);
          else walk(s0, e0, b, lane, 32);
        }
      }
    }
    if (GLOBAL_SLAB) __threadfence();
    __syncthreads();
    // ordered sweep in two passes without block-wide barriers in the loops: warp w owns the contiguous rows
    // [w*seg, (w+1)*seg) of the window, counts its kept entries, the eight counts are prefixed once, and the second
    // pass re-reads the segment (L2 / shared memory), emits in row order at the warp's offset and clears the slab.
    // Four 32-row groups per iteration, loads first. (One barrier per 256 rows held the sweep of a 32768-row window
    // at ~200 us per column - more than the products of a column with a few thousand of them.)
    const long long off = tmp_off[j];
    const int seg = ((w + CTA_T - 1) / CTA_T) * 32;
    const int r0 = min(w, warp * seg), r1 = min(w, r0 + seg);
    auto kept = [&](T v, int t, T& sv) {
      sv = s_scale(alpha, v);
      if (is_zero(v)) return false;            // an untouched row of the window (|0| > thr is false for every thr >= 0)
      return keep_entry(mag_fast(sv), mag_fast(v), thr, rule_for(rules, base + t, j));
    };
    int mine = 0;
    for (int t0 = r0; t0 < r1; t0 += 128) {
      T v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int t = t0 + 32 * u + lane; v[u] = (t < r1) ? ld_acc(&acc[t], GLOBAL_SLAB) : zero_of<T>(); }
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int t = t0 + 32 * u + lane; T sv; if (t < r1 && kept(v[u], t, sv)) ++mine; }
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0) s_warp_cnt[warp] = mine;
    __syncthreads();
    int before = 0, total = 0;
    for (int ww = 0; ww < CTA_T / 32; ++ww) { const int c = s_warp_cnt[ww]; total += c; if (ww < warp) before += c; }
    for (int t0 = r0; t0 < r1; t0 += 128) {
      T v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int t = t0 + 32 * u + lane; v[u] = (t < r1) ? ld_acc(&acc[t], GLOBAL_SLAB) : zero_of<T>(); }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = t0 + 32 * u + lane;
        bool keep = false;
        T sv = zero_of<T>();
        if (t < r1) {
          keep = kept(v[u], t, sv);
          if (GLOBAL_SLAB) acc[t] = zero_of<T>();
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const long long pos = off + before + __popc(m & ((1u << lane) - 1));
          tmp_idx[pos] = base + t;
          tmp_val[pos] = sv;
        }
        before += __popc(m);
      }
    }
    if (threadIdx.x == 0) s_running = total;
    if (threadIdx.x == 0) cnt[j] = s_running;
    if (GLOBAL_SLAB) __threadfence();
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
// COMPLEX128 ON THE FP64 TENSOR CORES (the reference's ZGEMM dense branch, DMatrixModule.F90:517-593, at tile
// granularity). A complex product C = A*B is ONE real product of twice the size with exactly the 8 real flops per
// complex multiply-add that the arithmetic needs - no 2x redundancy as with the full real embedding of both operands:
//     B^ = rows (2k, 2k+1) <- (Re, Im) of row k of B                           (2n x m, "stacked" form)
//     A^ = [[Re A, -Im A], [Im A, Re A]] with rows AND columns interleaved     (2n x 2n, embedded form)
//     C^ = A^ * B^ = stacked form of C:  C^(2i,j) = Re C(i,j),  C^(2i+1,j) = Im C(i,j)
// The interleaving keeps a complex 4x2 block inside one real 8x4 tile, so locally dense complex blocks (filled-in
// exponentials / inverses, banded or block-sparse Hermitian matrices) fill the tiles of k_tile_numeric9 as well as real
// ones do. The real product runs with threshold 0 (it keeps every non-zero); the reference's rule on the COMPLEX
// magnitude (|alpha*v| > thr, or |v| > thr under the dense-block rule) is applied when the pairs are zipped back.
__global__ void __launch_bounds__(256) k_embed_right(CscView<cplx> X, long long nnz, int* __restrict__ outer_h,
                                                     int* __restrict__ inner_h, double* __restrict__ val_h) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += stride) {
    const int k = X.inner[p];
    const cplx v = X.val[p];
    *reinterpret_cast<int2*>(inner_h + 2 * p) = make_int2(2 * k, 2 * k + 1);
    *reinterpret_cast<double2*>(val_h + 2 * p) = make_double2(v.x, v.y);
  }
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j <= X.cols; j += stride) outer_h[j] = 2 * X.outer[j];
}
// one warp per column k of Y: real columns 2k = (Re, Im) and 2k+1 = (-Im, Re)
__global__ void __launch_bounds__(256) k_embed_left(CscView<cplx> Y, int* __restrict__ outer_h, int* __restrict__ inner_h,
                                                    double* __restrict__ val_h) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int k = gw; k < Y.cols; k += nw) {
    const int s = Y.outer[k], e = Y.outer[k + 1];
    const long long b0 = 4ll * s, b1 = 2ll * s + 2ll * e;
    if (lane == 0) { outer_h[2 * k] = (int)b0; outer_h[2 * k + 1] = (int)b1; if (k == Y.cols - 1) outer_h[2 * k + 2] = 4 * e; }
    for (int p = s + lane; p < e; p += 32) {
      const int i = Y.inner[p];
      const cplx a = Y.val[p];
      const long long q = 2ll * (p - s);
      *reinterpret_cast<int2*>(inner_h + b0 + q) = make_int2(2 * i, 2 * i + 1);
      *reinterpret_cast<double2*>(val_h + b0 + q) = make_double2(a.x, a.y);
      *reinterpret_cast<int2*>(inner_h + b1 + q) = make_int2(2 * i, 2 * i + 1);
      *reinterpret_cast<double2*>(val_h + b1 + q) = make_double2(-a.y, a.x);
    }
  }
}
// stacked real result -> complex CSC with the threshold rule on the complex magnitude. One warp per column; an entry of
// the real column opens a complex entry when its predecessor belongs to another complex row. FILL=false counts.
template <bool FILL>
__global__ void __launch_bounds__(256) k_zip_complex(CscView<double> Zh, double alpha, double thr, RuleView rules,
                                                     int* __restrict__ cnt, const int* __restrict__ outer,
                                                     int* __restrict__ inner, cplx* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int j = gw; j < Zh.cols; j += nw) {
    const int s = Zh.outer[j], e = Zh.outer[j + 1];
    int count = 0;
    const int dst = FILL ? outer[j] : 0;
    for (int p0 = s; p0 < e; p0 += 32) {
      const int p = p0 + lane;
      bool keep = false;
      int row = 0;
      cplx v = cplx{0.0, 0.0};
      if (p < e) {
        const int r = Zh.inner[p];
        const bool head = (p == s) || ((Zh.inner[p - 1] >> 1) != (r >> 1));
        if (head) {
          row = r >> 1;
          if (r & 1) v.y = Zh.val[p];
          else {
            v.x = Zh.val[p];
            if (p + 1 < e && Zh.inner[p + 1] == r + 1) v.y = Zh.val[p + 1];
          }
          const cplx sv = s_scale(alpha, v);
          keep = keep_entry(s_abs(sv), s_abs(v), thr, rule_for(rules, row, j));
          v = sv;
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (FILL && keep) {
        const int pos = dst + count + __popc(m & ((1u << lane) - 1));
        inner[pos] = row;
        val[pos] = v;
      }
      count += __popc(m);
    }
    if (!FILL && lane == 0) cnt[j] = count;
  }
}

static bool complex_tile_enabled() {
  static const bool on = [] { const char* e = std::getenv("NTB_COMPLEX_TILE"); return !(e && e[0] == '0'); }();
  return on;
}
// C = A*B for complex blocks through the real tile product of the embeddings; false = declined (scattered pattern, too
// large for 32-bit entry counts, tile path off) - the caller then runs the scalar window / hash kernels
static bool spgemm_complex_tiles(const LocalCsc<cplx>& Xl, const LocalCsc<cplx>& Yl, double alpha, double thr,
                                 const RuleView& rules, LocalCsc<cplx>& Z, double useful_products) {
  const long long nx = Xl.nnz, ny = Yl.nnz;
  if (nx <= 0 || ny <= 0 || 4 * ny >= (1ll << 31) || 2 * nx >= (1ll << 31) || Yl.rows >= (1 << 30) || Yl.cols >= (1 << 30)) return false;
  const CscView<cplx> X = Xl.view(), Y = Yl.view();
  LocalCsc<double> Xh, Yh, Zh;
  Xh.rows = 2 * Xl.rows; Xh.cols = Xl.cols;
  Xh.outer.alloc((size_t)Xl.cols + 1);
  Xh.alloc_entries(2 * nx);
  NTB_LAUNCH(k_embed_right, min(div_up(nx, 256 * 4), kNumSMs * 16), 256, 0, X, nx, Xh.outer.get(), Xh.inner.get(), Xh.val.get());
  Yh.rows = 2 * Yl.rows; Yh.cols = 2 * Yl.cols;
  Yh.outer.alloc((size_t)2 * Yl.cols + 1);
  Yh.alloc_entries(4 * ny);
  NTB_LAUNCH(k_embed_left, max(1, min(div_up((long long)Yl.cols * 32, 256), kNumSMs * 16)), 256, 0, Y, Yh.outer.get(), Yh.inner.get(),
             Yh.val.get());
  // 4 real products per complex one; threshold 0 keeps every non-zero of the real product
  if (!spgemm_tile(Xh, Yh, 1.0, 0.0, RuleView{}, Zh, 4.0 * useful_products, nullptr, WANT_CSC)) return false;
  const int ncols = Xl.cols;
  const CscView<double> Zv = Zh.view();
  Z.rows = Yl.rows;
  Z.cols = ncols;
  Z.outer.alloc((size_t)ncols + 1);
  DevBuf<int> cnt((size_t)ncols);
  const int blocks = max(1, min(div_up((long long)ncols * 32, 256), kNumSMs * 16));
  NTB_LAUNCH((k_zip_complex<false>), blocks, 256, 0, Zv, alpha, thr, rules, cnt.get(), (const int*)nullptr, (int*)nullptr, (cplx*)nullptr);
  exclusive_scan(cnt.get(), Z.outer.get(), ncols);
  int h_nnz = 0;
  d2h(&h_nnz, Z.outer.get() + ncols, 1);
  Z.alloc_entries(h_nnz);
  if (h_nnz > 0)
    NTB_LAUNCH((k_zip_complex<true>), blocks, 256, 0, Zv, alpha, thr, rules, (int*)nullptr, Z.outer.get(), Z.inner.get(), Z.val.get());
  rt().complex_tile_products++;
  return true;
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

  // ---- complex operands: the same tile path through the real embeddings when the blocks are locally dense
  if constexpr (scalar_traits<T>::is_complex) {
    if (tile_path_enabled() && complex_tile_enabled() && Xl.nnz > 0 && Yl.nnz > 0 && !(shift && shift->sigma != 0.0)) {
      DevBuf<unsigned long long> fl(1);
      fl.zero();
      NTB_LAUNCH((k_useful_products<T>), min(div_up(Xl.nnz, 256 * 8), kNumSMs * 16), 256, 0, X, Y, Xl.nnz, fl.get());
      unsigned long long h_fl = 0;
      d2h(&h_fl, fl.get(), 1);
      // worth it only when columns are long enough to fill tiles
      if ((double)h_fl >= 16.0 * (double)ncols && spgemm_complex_tiles(Xl, Yl, alpha, thr, rules, Z, (double)h_fl)) {
        if (stats) { stats->flops = 8.0 * (double)h_fl; stats->tmp_entries = 0; }
        account_bytes(Z.nnz);
        return;
      }
      Z.rows = nrows;
      Z.cols = ncols;
      Z.outer.alloc((size_t)ncols + 1);
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
  static const bool atomic_bins = [] { const char* e = std::getenv("NTB_ATOMIC_BINS"); return !(e && e[0] == '0'); }();
  if (h_bins[5] > 0) {
    int blocks = min(h_bins[5], kNumSMs * 4);
    if (atomic_bins) {
      CUDA_CHECK(cudaFuncSetAttribute((k_numeric_cta_atomic<T, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUDGET));
      NTB_LAUNCH((k_numeric_cta_atomic<T, false>), blocks, CTA_T, SMEM_BUDGET, X, Y, lists.get() + (size_t)5 * ncols,
                 bin_count.get() + 5, lo.get(), wid.get(), tmp_off.get(), alpha, thr, rules, tmp_idx.get(),
                 tmp_val.get(), cnt.get(), (T*)nullptr);
    } else {
      CUDA_CHECK(cudaFuncSetAttribute((k_numeric_cta<T, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUDGET));
      NTB_LAUNCH((k_numeric_cta<T, false>), blocks, CTA_T, SMEM_BUDGET, X, Y, lists.get() + (size_t)5 * ncols,
                 bin_count.get() + 5, lo.get(), wid.get(), tmp_off.get(), alpha, thr, rules, tmp_idx.get(),
                 tmp_val.get(), cnt.get(), (T*)nullptr);
    }
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
    // How many accumulator windows (one per CTA) are in flight. Measured on c5 (N=32768 complex, 512 KB per window;
    // gpurun_out/r2c35*): 48 MB of windows 971 ms per step, 96 MB 890, 192 MB 745, 288 / 512 MB the same - more
    // columns in flight beat keeping every window inside the 126 MB L2, up to ~2.6 CTAs per SM
    static const long long slab_mb = [] { const char* e = std::getenv("NTB_SLAB_L2_MB"); return e ? std::max(1ll, std::atoll(e)) : 192ll; }();
    const long long l2_fit = max(1ll, (slab_mb << 20) / ((long long)nrows * (long long)sizeof(T)));
    int blocks = (int)min((long long)min(h_bins[6], kNumSMs * 4), atomic_bins ? max((long long)kNumSMs, l2_fit) : (long long)kNumSMs * 2);
    slab.alloc((size_t)blocks * nrows);
    slab.zero();
    if (atomic_bins)
      NTB_LAUNCH((k_numeric_cta_atomic<T, true>), blocks, CTA_T, 0, X, Y, lists.get() + (size_t)6 * ncols,
                 bin_count.get() + 6, lo.get(), wid.get(), tmp_off.get(), alpha, thr, rules, tmp_idx.get(),
                 tmp_val.get(), cnt.get(), slab.get());
    else
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
