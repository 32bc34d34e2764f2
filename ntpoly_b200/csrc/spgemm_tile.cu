// Tile SpGEMM on the FP64 tensor cores (DMMA.8x8x4) for locally dense operands
// (banded, block-sparse, filled-in purification iterates) — the B200 answer to
// NTPoly's dense-switch idea (reference sparse_includes/GemmMatrix.f90:59-61,
// DenseBranch.f90) applied at 8x4 / 4x8 tile granularity instead of whole blocks.
//
//   C(:,J) = sum_K  A(:,K) * B(K,J)        J: 8 output columns, K: 4 inner indices
//
//   A (the Y operand)  -> tile-CSC of 8x4 tiles, stored in DMMA A-fragment order
//   B (the X operand)  -> tile-CSC of 4x8 tiles, stored in DMMA B-fragment order
//   every (row tile I, K, J) triple with both tiles present = ONE mma.sync.m8n8k4.f64
//   (256 FMAs for 2 coalesced 256-byte loads, accumulators in registers).
//
// Pipeline: k_tile_count / k_tile_fill (CSC -> tiles, bitmap ranked), k_tile_bounds
// (row-tile window per J), k_tile_numeric (DMMA), k_tile_kept + scan + k_tile_emit
// (threshold rule, alpha, ordered compaction into CSC). Exact zeros introduced by tile
// padding never survive the strict |v| > thr test, so results equal the scalar path up
// to summation order.
#include "csc.cuh"

namespace ntb {

constexpr int BM_WORDS = 64;                 // bitmap words per warp: 2048 tiles of reach per tile column
constexpr int BM_BITS = BM_WORDS * 32;
constexpr int TW = 8;                        // warps per CTA

struct TileCsc {
  int tr = 0, tc = 0;                        // tile rows x cols (8x4 for A, 4x8 for B)
  int ntc = 0;                               // number of tile columns
  long long ntiles = 0;
  DevBuf<int> tptr;                          // [ntc+1]
  DevBuf<int> tid;                           // [ntiles] row-tile ids, ascending per tile column
  DevBuf<int> first, last;                   // [ntc] first / last tile id (last < first when empty)
  DevBuf<double> tval;                       // [ntiles*32] fragment-ordered values
};

template <int TR, int TC> __device__ __forceinline__ int frag_pos(int r, int c) {
  // A fragment (8x4): lane = r*4 + c ; B fragment (4x8): lane = c*4 + r
  return (TR == 8) ? (r * 4 + c) : (c * 4 + r);
}

// pass 1 (FILL=false): tiles per tile column; pass 2 (FILL=true): ids + values
template <int TR, int TC, bool FILL>
__global__ void __launch_bounds__(TW * 32)
k_tile_build(CscView<double> M, int ntc, int* __restrict__ tcount, int* __restrict__ first, int* __restrict__ last,
             int* __restrict__ overflow, const int* __restrict__ tptr, int* __restrict__ tid, double* __restrict__ tval) {
  __shared__ unsigned bm[TW][BM_WORDS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned* b = bm[warp];
  for (int q = blockIdx.x * TW + warp; q < ntc; q += gridDim.x * TW) {
    const int c0 = q * TC, c1 = min(M.cols, c0 + TC);
    int rmin = INT_MAX, rmax = -1;
    if (lane < c1 - c0) {
      const int s = M.outer[c0 + lane], e = M.outer[c0 + lane + 1];
      if (e > s) { rmin = M.inner[s]; rmax = M.inner[e - 1]; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, d));
      rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, d));
    }
    if (rmax < 0) {
      if (!FILL && lane == 0) { tcount[q] = 0; first[q] = 0; last[q] = -1; }
      continue;
    }
    const int tmin = rmin / TR, tmax = rmax / TR;
    if (tmax - tmin + 1 > BM_BITS) {
      if (!FILL && lane == 0) { tcount[q] = 0; first[q] = 0; last[q] = -1; atomicExch(overflow, 1); }
      continue;
    }
    b[lane] = 0; b[lane + 32] = 0;
    __syncwarp();
    for (int c = c0; c < c1; ++c)
      for (int p = M.outer[c] + lane; p < M.outer[c + 1]; p += 32) {
        const int t = M.inner[p] / TR - tmin;
        atomicOr(&b[t >> 5], 1u << (t & 31));
      }
    __syncwarp();
    // exclusive prefix of popcounts over the 64 words (2 per lane)
    const unsigned w0 = b[2 * lane], w1 = b[2 * lane + 1];
    const int mine = __popc(w0) + __popc(w1);
    int inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    if (!FILL) {
      if (lane == 0) { tcount[q] = total; first[q] = tmin; last[q] = tmax; }
      __syncwarp();
      continue;
    }
    const int base = tptr[q];
    // tile ids
    int rank = base + inc - mine;
    unsigned w = w0;
    while (w) { const int bit = __ffs(w) - 1; w &= w - 1; tid[rank++] = tmin + 2 * lane * 32 + bit; }
    w = w1;
    while (w) { const int bit = __ffs(w) - 1; w &= w - 1; tid[rank++] = tmin + (2 * lane + 1) * 32 + bit; }
    // word-exclusive prefix back into shared memory (reuse: store prefix in a second array via shuffles)
    __syncwarp();
    __shared__ int pre[TW][BM_WORDS];
    pre[warp][2 * lane] = inc - mine;
    pre[warp][2 * lane + 1] = inc - mine + __popc(w0);
    __syncwarp();
    for (int c = c0; c < c1; ++c)
      for (int p = M.outer[c] + lane; p < M.outer[c + 1]; p += 32) {
        const int r = M.inner[p];
        const int t = r / TR - tmin;
        const int rk = pre[warp][t >> 5] + __popc(b[t >> 5] & ((1u << (t & 31)) - 1));
        tval[((size_t)(base + rk)) * 32 + frag_pos<TR, TC>(r - (r / TR) * TR, c - c0)] = M.val[p];
      }
    __syncwarp();
  }
}

template <int TR, int TC>
static bool build_tiles(const CscView<double>& M, TileCsc& T) {
  T.tr = TR; T.tc = TC;
  T.ntc = div_up(M.cols, TC);
  const int ntc = T.ntc;
  DevBuf<int> tcount((size_t)ntc), overflow(1);
  T.first.alloc((size_t)ntc); T.last.alloc((size_t)ntc); T.tptr.alloc((size_t)ntc + 1);
  overflow.zero();
  const int grid = max(1, min(div_up(ntc, TW), kNumSMs * 8));
  NTB_LAUNCH((k_tile_build<TR, TC, false>), grid, TW * 32, 0, M, ntc, tcount.get(), T.first.get(), T.last.get(),
             overflow.get(), (const int*)nullptr, (int*)nullptr, (double*)nullptr);
  exclusive_scan(tcount.get(), T.tptr.get(), ntc);
  int h[2] = {0, 0};
  CUDA_CHECK(cudaMemcpyAsync(&h[0], T.tptr.get() + ntc, sizeof(int), cudaMemcpyDeviceToHost, rt().stream));
  CUDA_CHECK(cudaMemcpyAsync(&h[1], overflow.get(), sizeof(int), cudaMemcpyDeviceToHost, rt().stream));
  stream_sync();
  if (h[1]) return false;
  T.ntiles = h[0];
  T.tid.alloc((size_t)T.ntiles);
  T.tval.alloc((size_t)T.ntiles * 32);
  T.tval.zero();
  NTB_LAUNCH((k_tile_build<TR, TC, true>), grid, TW * 32, 0, M, ntc, (int*)nullptr, (int*)nullptr, (int*)nullptr,
             (int*)nullptr, T.tptr.get(), T.tid.get(), T.tval.get());
  return true;
}

struct TileView {
  const int* tptr; const int* tid; const int4* meta; const double* tval; int ntc;
};

// per tile column: {offset of its first tile, tile count, first tile id, last tile id}
__global__ void __launch_bounds__(256) k_tile_meta(int ntc, const int* __restrict__ tptr, const int* __restrict__ first,
                                                   const int* __restrict__ last, int4* __restrict__ meta) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < ntc) meta[q] = make_int4(tptr[q], tptr[q + 1] - tptr[q], first[q], last[q]);
}

// per output tile column J: row-tile window aligned to blocks of 8 row tiles, and the DMMA count
__global__ void __launch_bounds__(256) k_tile_bounds(TileView A, TileView B, int* __restrict__ imin8, int* __restrict__ nI8,
                                                     unsigned long long* __restrict__ ndmma) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  unsigned long long mine = 0;
  for (int J = gw; J < B.ntc; J += nw) {
    int mn = INT_MAX, mx = -1;
    for (int t = B.tptr[J] + lane; t < B.tptr[J + 1]; t += 32) {
      const int4 m = A.meta[B.tid[t]];
      if (m.y > 0) { mn = min(mn, m.z); mx = max(mx, m.w); mine += (unsigned long long)m.y; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if (lane == 0) {
      if (mx >= 0) { imin8[J] = (mn >> 3) << 3; nI8[J] = (((mx >> 3) + 1) << 3) - ((mn >> 3) << 3); }
      else { imin8[J] = 0; nI8[J] = 0; }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
  if (lane == 0 && mine) atomicAdd(ndmma, mine);
}

// groups of 8 tile columns (64 output columns): union of their 64-row block ranges
__global__ void __launch_bounds__(256) k_group_bounds(int nJ, int nG, const int* __restrict__ imin8, const int* __restrict__ nI8,
                                                      int* __restrict__ gbmin, int* __restrict__ gnb) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nG) return;
  int mn = INT_MAX, mx = -1;
  for (int J = g * 8; J < min(nJ, g * 8 + 8); ++J)
    if (nI8[J] > 0) { mn = min(mn, imin8[J] >> 3); mx = max(mx, ((imin8[J] + nI8[J]) >> 3) - 1); }
  gbmin[g] = (mx >= 0) ? mn : 0;
  gnb[g] = (mx >= 0) ? (mx - mn + 1) : 0;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// CTA task = (group of 8 tile columns, block of 8 row tiles) = a 64x64 output block;
// warp w owns tile column 8g+w: a 64x8 strip with its 8 accumulator tiles in registers.
// Loop order: B tiles of the column outermost (each loaded once), the <=8 A tiles that meet
// the strip innermost (consecutive in memory); the 8 warps walk the same A tiles, so they
// are served by L1 after the first touch.
__global__ void __launch_bounds__(TW * 32)
k_tile_numeric(TileView A, TileView B, const int* __restrict__ imin8, const int* __restrict__ nI8,
               const long long* __restrict__ stg_off, const int* __restrict__ gbmin, const int* __restrict__ gtask_off,
               int nG, double* __restrict__ stg) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ntasks = gtask_off[nG];
  for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
    int lo_g = 0, hi_g = nG;               // upper_bound(gtask_off, task) - 1
    while (lo_g < hi_g) { const int mid = (lo_g + hi_g) >> 1; if (gtask_off[mid + 1] <= task) lo_g = mid + 1; else hi_g = mid; }
    const int g = lo_g;
    const int J = g * 8 + warp;
    if (J >= B.ntc) continue;
    const int I0 = (gbmin[g] + (task - gtask_off[g])) << 3;
    const int iw0 = imin8[J], iw1 = iw0 + nI8[J];
    if (I0 < iw0 || I0 >= iw1) continue;
    const int xb = B.tptr[J], nK = B.tptr[J + 1] - xb;
    double acc[8][2];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) { acc[ii][0] = 0.0; acc[ii][1] = 0.0; }
    int4 m_next = (nK > 0) ? A.meta[B.tid[xb]] : make_int4(0, 0, 0, -1);
    for (int t = 0; t < nK; ++t) {
      const int4 m = m_next;
      if (t + 1 < nK) m_next = A.meta[B.tid[xb + t + 1]];   // prefetch: hides the tid -> meta dependency
      const int lo = max(m.z, I0), hi = min(m.w, I0 + 7);
      if (m.y == 0 || lo > hi) continue;
      unsigned mask;
      int base;
      if (m.w - m.z + 1 == m.y) {          // one contiguous run of row tiles: tile ii sits at a fixed offset
        const double bv = B.tval[(size_t)(xb + t) * 32 + lane];
        const double* ap = A.tval + ((long long)m.x + (I0 - m.z)) * 32 + lane;   // tile I0 (may precede the run)
        const int l0 = lo - I0, h0 = hi - I0;
        double av[8];
        if (l0 == 0 && h0 == 7) {            // the run covers the whole 64-row strip: no predicates at all
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) av[ii] = ap[ii * 32];
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) dmma884(acc[ii][0], acc[ii][1], av[ii], bv);
          continue;
        }
#pragma unroll
        for (int ii = 0; ii < 8; ++ii)
          if (ii >= l0 && ii <= h0) av[ii] = ap[ii * 32];
#pragma unroll
        for (int ii = 0; ii < 8; ++ii)
          if (ii >= l0 && ii <= h0) dmma884(acc[ii][0], acc[ii][1], av[ii], bv);
        continue;
      } else {                             // several runs: locate the block inside the sorted id list
        int l2 = 0, h2 = m.y;
        while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (A.tid[m.x + mid] < I0) l2 = mid + 1; else h2 = mid; }
        const int id = (lane < 8 && l2 + lane < m.y) ? A.tid[m.x + l2 + lane] : INT_MAX;
        mask = __reduce_or_sync(0xffffffffu, (id < I0 + 8) ? (1u << (id - I0)) : 0u);
        if (mask == 0u) continue;
        base = m.x + l2;
      }
      const double bv = B.tval[(size_t)(xb + t) * 32 + lane];
      const double* ap = A.tval + (size_t)base * 32 + lane;
#pragma unroll
      for (int ii = 0; ii < 8; ++ii)
        if ((mask >> ii) & 1u) {
          const double av = ap[(size_t)__popc(mask & ((1u << ii) - 1u)) * 32];
          dmma884(acc[ii][0], acc[ii][1], av, bv);
        }
    }
    // C fragment: row = lane/4, cols = 2*(lane%4), +1 ; staging is column-major per tile column
    const int wlen = nI8[J] * 8;
    const int r = lane >> 2, cc = (lane & 3) * 2;
    double* o = stg + stg_off[J] * 64 + (size_t)cc * wlen + (size_t)(I0 - iw0) * 8 + r;
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) { o[ii * 8] = acc[ii][0]; o[wlen + ii * 8] = acc[ii][1]; }
  }
}

__device__ __forceinline__ bool tile_rule(const RuleView& r, int inner_idx, int outer_idx) {
  if (r.tbl == nullptr) return false;
  return r.tbl[(inner_idx / r.rb) * r.nJ + (outer_idx / r.cb)] != 0;
}

// kept entries per output column (EMIT=false) / ordered emit into CSC (EMIT=true)
template <bool EMIT>
__global__ void __launch_bounds__(256)
k_tile_emit(int ncols, int nrows, const int* __restrict__ imin8, const int* __restrict__ nI8,
            const long long* __restrict__ stg_off, const double* __restrict__ stg, double alpha, double thr,
            RuleView rules, int* __restrict__ cnt, const int* __restrict__ outer, int* __restrict__ inner,
            double* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int j = gw; j < ncols; j += nw) {
    const int J = j >> 3, jj = j & 7;
    const int wlen = nI8[J] * 8;
    const int base = imin8[J] * 8;
    const double* src = stg + stg_off[J] * 64 + (size_t)jj * wlen;
    int count = 0;
    const int dst = EMIT ? outer[j] : 0;
    for (int t0 = 0; t0 < wlen; t0 += 32) {
      const int t = t0 + lane;
      bool keep = false;
      double sv = 0.0;
      if (t < wlen && base + t < nrows) {
        const double v = src[t];
        sv = alpha * v;
        keep = (tile_rule(rules, base + t, j) ? fabs(v) : fabs(sv)) > thr;
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (EMIT && keep) {
        const int pos = dst + count + __popc(m & ((1u << lane) - 1));
        inner[pos] = base + t;
        val[pos] = sv;
      }
      count += __popc(m);
    }
    if (!EMIT && lane == 0) cnt[j] = count;
  }
}

// returns false when the operands are not locally dense enough (caller falls back to the
// scalar window kernels). useful_products = sum over B entries of the A column lengths.
bool spgemm_tile(const CscView<double>& X, const CscView<double>& Y, double alpha, double thr, const RuleView& rules,
                 LocalCsc<double>& Z, double useful_products, long long nnzX, long long nnzY) {
  const int ncols = X.cols, nrows = Y.rows;
  if (ncols == 0 || nnzX == 0 || nnzY == 0) return false;
  TileCsc A, B;
  if (!build_tiles<8, 4>(Y, A)) return false;
  if ((double)nnzY < 0.20 * 32.0 * (double)A.ntiles) return false;   // tiles mostly padding
  if (!build_tiles<4, 8>(X, B)) return false;
  if ((double)nnzX < 0.20 * 32.0 * (double)B.ntiles) return false;
  const int nJ = B.ntc, nG = div_up(nJ, 8);
  DevBuf<int4> metaA((size_t)A.ntc);
  NTB_LAUNCH(k_tile_meta, div_up(A.ntc, 256), 256, 0, A.ntc, A.tptr.get(), A.first.get(), A.last.get(), metaA.get());
  const TileView Av{A.tptr.get(), A.tid.get(), metaA.get(), A.tval.get(), A.ntc};
  const TileView Bv{B.tptr.get(), B.tid.get(), nullptr, B.tval.get(), B.ntc};
  DevBuf<int> imin8((size_t)nJ), nI8((size_t)nJ), gbmin((size_t)nG), gnb((size_t)nG), gtask_off((size_t)nG + 1);
  DevBuf<long long> stg_off((size_t)nJ + 1);
  DevBuf<unsigned long long> ndmma(1);
  ndmma.zero();
  NTB_LAUNCH(k_tile_bounds, max(1, min(div_up((long long)nJ * 32, 256), kNumSMs * 16)), 256, 0, Av, Bv, imin8.get(),
             nI8.get(), ndmma.get());
  NTB_LAUNCH(k_group_bounds, div_up(nG, 256), 256, 0, nJ, nG, imin8.get(), nI8.get(), gbmin.get(), gnb.get());
  exclusive_scan(nI8.get(), stg_off.get(), nJ);        // staging in units of 64 doubles (8 cols x 8 rows per row tile)
  exclusive_scan(gnb.get(), gtask_off.get(), nG);
  long long h_stg = 0;
  unsigned long long h_ndmma = 0;
  int h_tasks = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h_stg, stg_off.get() + nJ, sizeof(long long), cudaMemcpyDeviceToHost, rt().stream));
  CUDA_CHECK(cudaMemcpyAsync(&h_ndmma, ndmma.get(), sizeof(h_ndmma), cudaMemcpyDeviceToHost, rt().stream));
  CUDA_CHECK(cudaMemcpyAsync(&h_tasks, gtask_off.get() + nG, sizeof(int), cudaMemcpyDeviceToHost, rt().stream));
  stream_sync();
  // tensor-core work must not dwarf the useful work (256 FMAs per DMMA)
  if ((double)h_ndmma * 256.0 > 12.0 * useful_products) return false;

  DevBuf<double> stg((size_t)h_stg * 64);
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (rt().profile) {
    CUDA_CHECK(cudaEventCreate(&ev0));
    CUDA_CHECK(cudaEventCreate(&ev1));
    CUDA_CHECK(cudaEventRecord(ev0, rt().stream));
  }
  if (h_tasks > 0)
    NTB_LAUNCH(k_tile_numeric, min(h_tasks, kNumSMs * 8), TW * 32, 0, Av, Bv, imin8.get(), nI8.get(), stg_off.get(),
               gbmin.get(), gtask_off.get(), nG, stg.get());
  if (rt().profile) {
    CUDA_CHECK(cudaEventRecord(ev1, rt().stream));
    rt().prof_events.emplace_back(ev0, ev1);
  }
  DevBuf<int> cnt((size_t)ncols);
  const int egrid = max(1, min(div_up((long long)ncols * 32, 256), kNumSMs * 16));
  NTB_LAUNCH((k_tile_emit<false>), egrid, 256, 0, ncols, nrows, imin8.get(), nI8.get(), stg_off.get(), stg.get(), alpha,
             thr, rules, cnt.get(), (const int*)nullptr, (int*)nullptr, (double*)nullptr);
  Z.rows = nrows; Z.cols = ncols;
  Z.outer.alloc((size_t)ncols + 1);
  exclusive_scan(cnt.get(), Z.outer.get(), ncols);
  int h_nnz = 0;
  d2h(&h_nnz, Z.outer.get() + ncols, 1);
  Z.alloc_entries(h_nnz);
  if (h_nnz > 0)
    NTB_LAUNCH((k_tile_emit<true>), egrid, 256, 0, ncols, nrows, imin8.get(), nI8.get(), stg_off.get(), stg.get(),
               alpha, thr, rules, (int*)nullptr, Z.outer.get(), Z.inner.get(), Z.val.get());
  rt().tile_products++;
  rt().dmma_issued += (double)h_ndmma;
  return true;
}

}  // namespace ntb
