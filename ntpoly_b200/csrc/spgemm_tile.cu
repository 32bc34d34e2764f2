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

__device__ __forceinline__ bool tile_rule(const RuleView& r, int inner_idx, int outer_idx) {
  if (r.tbl == nullptr) return false;
  return r.tbl[(inner_idx / r.rb) * r.nJ + (outer_idx / r.cb)] != 0;
}

// groups of 8 tile columns (64 output columns): union of their 64-row block ranges, and the
// range of inner tile indices K their B tiles span
__global__ void __launch_bounds__(256) k_group_bounds(int nJ, int nG, const int* __restrict__ imin8, const int* __restrict__ nI8,
                                                      const int4* __restrict__ metaB, int* __restrict__ gbmin,
                                                      int* __restrict__ gnb, int2* __restrict__ gk) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nG) return;
  int mn = INT_MAX, mx = -1, kmn = INT_MAX, kmx = -1;
  for (int J = g * 8; J < min(nJ, g * 8 + 8); ++J)
    if (nI8[J] > 0) {
      mn = min(mn, imin8[J] >> 3); mx = max(mx, ((imin8[J] + nI8[J]) >> 3) - 1);
      const int4 m = metaB[J];
      kmn = min(kmn, m.z); kmx = max(kmx, m.w);
    }
  gbmin[g] = (mx >= 0) ? mn : 0;
  gnb[g] = (mx >= 0) ? (mx - mn + 1) : 0;
  gk[g] = (kmx >= 0) ? make_int2(kmn, kmx - kmn + 1) : make_int2(0, 0);
}

// task t = (group g, 64-row block): {g, first row tile I0, first live inner tile, live inner tile count | window mask << 24}
// "live" = the range of inner tiles K whose A tile column meets the rows of the block (dead chunks at both ends of
// the group's K range are never visited); window mask bit jj = block lies inside the row window of tile column 8g+jj.
// One warp per group, lanes over its row blocks; the A.meta reads are warp-uniform.
__global__ void __launch_bounds__(256) k_task_table(int nG, int nJ, const int* __restrict__ gbmin, const int* __restrict__ gtask_off,
                                                    const int2* __restrict__ gk, const int4* __restrict__ metaA,
                                                    const int* __restrict__ imin8, const int* __restrict__ nI8,
                                                    int4* __restrict__ tasks) {
  const int lane = threadIdx.x & 31;
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= nG) return;
  const int t0 = gtask_off[g], n = gtask_off[g + 1] - t0;
  if (n == 0) return;
  const int2 k2 = gk[g];
  const int b0 = gbmin[g];
  for (int bb = 0; bb < n; bb += 32) {
    const int b = bb + lane;
    const int I0 = (b0 + b) << 3;
    int klo = INT_MAX, khi = -1;
    for (int k = 0; k < k2.y; ++k) {
      const int4 m = metaA[k2.x + k];
      if (m.y > 0 && m.z <= I0 + 7 && m.w >= I0) { klo = min(klo, k); khi = k; }
    }
    unsigned win = 0;
    for (int jj = 0; jj < 8; ++jj) {
      const int J = g * 8 + jj;
      if (J < nJ) { const int iw0 = imin8[J]; if (I0 >= iw0 && I0 < iw0 + nI8[J]) win |= 1u << jj; }
    }
    if (b < n)
      tasks[t0 + b] = (khi >= 0) ? make_int4(g, I0, k2.x + klo, (khi - klo + 1) | (int)(win << 24))
                                 : make_int4(g, I0, k2.x, 0 | (int)(win << 24));
  }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- mbarrier + bulk-copy (TMA, 1-D) primitives; SASS: SYNCS.*, UBLKCP.S.G
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// pipeline shape: KC inner tiles (of 4 indices) per stage, NSTAGE stages; a stage holds the A slab
// [KC][8 row tiles] and the B slab [8 tile columns][KC] (KC * 4 KB), 96 KB per CTA, two CTAs per SM
constexpr int META_BYTES = 32;                               // maskA[8] maskB[8] flags g I0 pad
template <int KC, int NSTAGE> struct PipeCfg {
  static constexpr int STAGE_DOUBLES = 2 * KC * 8 * 32;
  static constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
  static constexpr int SMEM = NSTAGE * STAGE_BYTES + NSTAGE * META_BYTES + 2 * NSTAGE * 8;
};
constexpr int CW = 16;                                       // DMMA warps: 8 tile columns x 2 halves of the row block
constexpr int NUMERIC_THREADS = (CW + 1) * 32;               // + 1 copy warp

// tiles whose ids are the set bits of `mask` lie consecutively in memory from tile `src`; copy those also set in
// `want` to their natural slots (slot = bit) with as few bulk copies as the id runs allow. Returns nothing: the
// byte count was announced to the barrier beforehand (256 B per bit of mask & want).
__device__ __forceinline__ void copy_runs(unsigned mask, unsigned want, long long src, const double* __restrict__ tval,
                                          unsigned dst_slot0, unsigned bar) {
  while (mask) {
    const int b = __ffs(mask) - 1;
    const int len = __ffs(~(mask >> b)) - 1;             // run of consecutive ids
    const unsigned runbits = ((1u << len) - 1u) << b;
    unsigned w = want & runbits;
    while (w) {
      const int wb = __ffs(w) - 1;
      const int wl = __ffs(~(w >> wb)) - 1;
      bulk_g2s(dst_slot0 + wb * 256, tval + (src + (wb - b)) * 32, wl * 256, bar);
      w &= ~(((1u << wl) - 1u) << wb);
    }
    src += len;
    mask &= ~runbits;
  }
}

// Persistent CTAs; task = 64x64 output block (8 tile columns x 8 row tiles), handed out by an atomic counter.
// Warp 8 (copy warp): for every chunk of KC inner tiles it works out which A tiles (rows of the block) and which
// B tiles (columns of the group) exist, publishes the two 8x8 presence masks, and moves the tiles with 1-D bulk
// copies (TMA) into a 3-stage shared-memory ring guarded by full/empty mbarriers. Warps 0-7: warp w owns tile
// column 8g+w, i.e. a 64x8 strip with its 8 accumulator tiles in registers, and issues one DMMA.8x8x4 per
// (present A tile, present B tile) pair from conflict-free 256-byte shared-memory fragments. The strip is
// written to the dense staging window and the kept-entry counts of its 8 columns are accumulated on the fly
// (threshold rule fused), so the emit pass is a single sweep.
template <int KC, int NSTAGE>
__global__ void __launch_bounds__(NUMERIC_THREADS, 2)
k_tile_numeric(TileView A, TileView B, const int* __restrict__ imin8, const int* __restrict__ nI8,
               const long long* __restrict__ stg_off, const int4* __restrict__ tasks,
               int ntasks, int* __restrict__ task_counter, double* __restrict__ stg, int* __restrict__ cnt,
               int nrows, int ncols, double alpha, double thr, RuleView rules) {
  constexpr int STAGE_DOUBLES = PipeCfg<KC, NSTAGE>::STAGE_DOUBLES;
  constexpr int STAGE_BYTES = PipeCfg<KC, NSTAGE>::STAGE_BYTES;
  extern __shared__ __align__(128) unsigned char smem[];
  double* slab = reinterpret_cast<double*>(smem);
  unsigned char* meta = smem + NSTAGE * STAGE_BYTES;
  const unsigned bar0 = smem_u32(smem + NSTAGE * STAGE_BYTES + NSTAGE * META_BYTES);   // full[s] at +8s, empty[s] at +8(NSTAGE+s)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 8 * (NSTAGE + s), CW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  int st = 0;
  unsigned ph = 0;
  if (warp == CW) {
    // ------------------------------------------------------------------ copy warp
    const int4 none = make_int4(0, 0, 0, -1);
    int task = 0;
    if (lane == 0) task = atomicAdd(task_counter, 1);
    task = __shfl_sync(0xffffffffu, task, 0);
    int4 tk = (task < ntasks) ? tasks[task] : make_int4(0, 0, 0, 0);
    for (;;) {
      const bool done = task >= ntasks;
      const int g = tk.x, I0 = tk.y, kmin = tk.z, nk = tk.w & 0xffffff;
      const unsigned win = (unsigned)tk.w >> 24;
      // claim the next task now; its table entry is read after this task's chunks are under way
      int task_n = 0;
      if (!done && lane == 0) task_n = atomicAdd(task_counter, 1);
      // B role (lanes 8..15): tile column J = 8g + lane - 8, skipped when the block is outside its row window
      int4 mb4 = none;
      if (!done && lane >= 8 && lane < 16 && ((win >> (lane - 8)) & 1u)) mb4 = B.meta[g * 8 + lane - 8];
      int4 ma_next = (!done && lane < KC && lane < nk) ? A.meta[kmin + lane] : none;
      long long pB = mb4.x;
      const long long endB = (long long)mb4.x + mb4.y;
      const bool contigB = (mb4.w - mb4.z + 1 == mb4.y);
      if (!contigB) {                               // skip the tiles below the live range
        while (pB < endB && B.tid[pB] < kmin) ++pB;
      }
      const int nch = done ? 1 : max(1, (nk + KC - 1) / KC);
      int4 tk_n = make_int4(0, 0, 0, 0);
      for (int c = 0; c < nch; ++c) {
        const int K0 = kmin + c * KC, Kend = min(K0 + KC, kmin + nk);
        const bool last = (c == nch - 1);
        const int4 m = ma_next;
        ma_next = (lane < KC && K0 + KC + lane < kmin + nk) ? A.meta[K0 + KC + lane] : none;
        if (c == 0 && !done) {
          task_n = __shfl_sync(0xffffffffu, task_n, 0);
          tk_n = (task_n < ntasks) ? tasks[task_n] : make_int4(0, 0, 0, 0);
        }
        unsigned mask = 0;
        long long src = 0;
        if (lane < 8) {
          if (m.y > 0 && m.z <= I0 + 7 && m.w >= I0) {
            if (m.w - m.z + 1 == m.y) {
              const int lo = max(m.z, I0), hi = min(m.w, I0 + 7);
              mask = ((1u << (hi - lo + 1)) - 1u) << (lo - I0);
              src = (long long)m.x + (lo - m.z);
            } else {
              int l2 = 0, h2 = m.y;
              while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (A.tid[m.x + mid] < I0) l2 = mid + 1; else h2 = mid; }
              src = (long long)m.x + l2;
              for (int p = l2; p < m.y; ++p) {
                const int id = A.tid[m.x + p];
                if (id >= I0 + 8) break;
                mask |= 1u << (id - I0);
              }
            }
          }
        } else if (lane < 16 && mb4.y > 0) {
          if (contigB) {
            const int lo = max(mb4.z, K0), hi = min(mb4.w, Kend - 1);
            if (lo <= hi) { mask = ((1u << (hi - lo + 1)) - 1u) << (lo - K0); src = (long long)mb4.x + (lo - mb4.z); }
          } else {
            src = pB;
            while (pB < endB) {
              const int id = B.tid[pB];
              if (id >= Kend) break;
              mask |= 1u << (id - K0);
              ++pB;
            }
          }
        }
        // inner tiles that have both an A tile in the block and a B tile in the group
        const unsigned kA = __ballot_sync(0xffffffffu, lane < 8 && mask != 0u) & 0xffu;
        const unsigned kB = __reduce_or_sync(0xffffffffu, (lane >= 8 && lane < 16) ? mask : 0u);
        const unsigned live = kA & kB;
        if (live == 0u && !last) continue;
        unsigned want = 0;
        if (lane < 8) want = ((live >> lane) & 1u) ? mask : 0u;
        else if (lane < 16) want = mask & live;
        const unsigned bytes = __reduce_add_sync(0xffffffffu, (unsigned)__popc(want) * 256u);
        mbar_wait(bar0 + 8 * (NSTAGE + st), ph ^ 1u);        // consumers have released this slot
        unsigned char* mt = meta + st * META_BYTES;
        if (lane < 16) mt[lane] = (unsigned char)want;
        if (lane == 16) {
          int* mi = reinterpret_cast<int*>(mt + 16);
          mi[0] = (last ? 1 : 0) | (done ? 2 : 0);
          mi[1] = g;
          mi[2] = I0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(bar0 + 8 * st, bytes);
        __syncwarp();
        if (want) {
          const unsigned slab_s = smem_u32(slab + (size_t)st * STAGE_DOUBLES);
          if (lane < 8) copy_runs(mask, want, src, A.tval, slab_s + lane * (8 * 256), bar0 + 8 * st);
          else copy_runs(mask, want, src, B.tval, slab_s + KC * 8 * 256 + (lane - 8) * (KC * 256), bar0 + 8 * st);
        }
        if (++st == NSTAGE) { st = 0; ph ^= 1u; }
      }
      if (done) break;
      task = task_n;
      tk = tk_n;
    }
    return;
  }

  // -------------------------------------------------------------------- DMMA warps
  const int wj = warp & 7, half = warp >> 3;        // tile column of the group, upper/lower 4 row tiles of the block
  for (;;) {
    double acc[4][2];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) { acc[ii][0] = 0.0; acc[ii][1] = 0.0; }
    int g = 0, I0 = 0;
    unsigned fl = 0;
    do {
      mbar_wait(bar0 + 8 * st, ph);
      const unsigned char* mt = meta + st * META_BYTES;
      const uint4 mm = *reinterpret_cast<const uint4*>(mt);               // maskA[0..7], maskB[0..7]
      const int4 mi = *reinterpret_cast<const int4*>(mt + 16);
      fl = (unsigned)mi.x; g = mi.y; I0 = mi.z;
      const unsigned mb = ((wj < 4 ? mm.z >> (8 * wj) : mm.w >> (8 * (wj - 4)))) & 0xffu;
      // this warp's 4 row tiles: nibble `half` of every maskA byte
      const unsigned mlo = (mm.x >> (4 * half)) & 0x0f0f0f0fu, mhi = (mm.y >> (4 * half)) & 0x0f0f0f0fu;
      if (mb && (mlo | mhi)) {
        const double* As = slab + (size_t)st * STAGE_DOUBLES + half * (4 * 32) + lane;
        const double* Bs = slab + (size_t)st * STAGE_DOUBLES + KC * 8 * 32 + wj * (KC * 32) + lane;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
          const unsigned ma = ((kk < 4 ? mlo >> (8 * kk) : mhi >> (8 * (kk - 4)))) & 0xfu;
          if (((mb >> kk) & 1u) == 0u || ma == 0u) continue;
          const double bv = Bs[kk * 32];
          const double* ap = As + kk * (8 * 32);
          if (ma == 0xfu) {
            double av[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) av[ii] = ap[ii * 32];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) dmma884(acc[ii][0], acc[ii][1], av[ii], bv);
          } else {
            // a predicated-off DMMA still occupies the FP64 tensor pipe for its full 16 cycles (measured,
            // scripts/micro/dmma_shapes.cu), so absent tiles are skipped with real branches: enter the
            // unrolled sequence at the first tile of each id run, leave it after the last.
            double av[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
              if ((ma >> ii) & 1u) av[ii] = ap[ii * 32];
            unsigned m = ma;
#define NTB_RUN_STEP(i) dmma884(acc[i][0], acc[i][1], av[i], bv); if (h0 == i) break;
            do {
              const int l0 = __ffs(m) - 1;
              const int len = __ffs(~(m >> l0)) - 1;
              const int h0 = l0 + len - 1;
              m &= ~(((1u << len) - 1u) << l0);
              switch (l0) {
                case 0: NTB_RUN_STEP(0)
                case 1: NTB_RUN_STEP(1)
                case 2: NTB_RUN_STEP(2)
                default: dmma884(acc[3][0], acc[3][1], av[3], bv);
              }
            } while (m);
#undef NTB_RUN_STEP
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8 * (NSTAGE + st));
      if (++st == NSTAGE) { st = 0; ph ^= 1u; }
    } while ((fl & 1u) == 0u);
    if (fl & 2u) break;
    const int J = g * 8 + wj;
    if (J >= B.ntc) continue;
    const int iw0 = imin8[J], nI = nI8[J];
    if (I0 < iw0 || I0 >= iw0 + nI) continue;
    // C fragment: row = lane/4, cols = 2*(lane%4), +1 ; staging is column-major per tile column
    const int wlen = nI * 8;
    const int r = lane >> 2, cc = (lane & 3) * 2;
    const int Ih = I0 + 4 * half;
    double* o = stg + stg_off[J] * 64 + (size_t)cc * wlen + (size_t)(Ih - iw0) * 8 + r;
    const int j0 = J * 8 + cc;
    int c0 = 0, c1 = 0;
    if (rules.tbl == nullptr) {                 // sparse rule everywhere: |alpha*v| > thr
      const bool in0 = j0 < ncols, in1 = j0 + 1 < ncols;
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        const double v0 = acc[ii][0], v1 = acc[ii][1];
        o[ii * 8] = v0;
        o[wlen + ii * 8] = v1;
        const bool rin = (Ih + ii) * 8 + r < nrows;
        c0 += (rin && in0 && fabs(alpha * v0) > thr) ? 1 : 0;
        c1 += (rin && in1 && fabs(alpha * v1) > thr) ? 1 : 0;
      }
    } else {
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        const double v0 = acc[ii][0], v1 = acc[ii][1];
        o[ii * 8] = v0;
        o[wlen + ii * 8] = v1;
        const int row = (Ih + ii) * 8 + r;
        if (row < nrows) {
          if (j0 < ncols) c0 += ((tile_rule(rules, row, j0) ? fabs(v0) : fabs(alpha * v0)) > thr) ? 1 : 0;
          if (j0 + 1 < ncols) c1 += ((tile_rule(rules, row, j0 + 1) ? fabs(v1) : fabs(alpha * v1)) > thr) ? 1 : 0;
        }
      }
    }
#pragma unroll
    for (int d = 4; d < 32; d <<= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, d);
      c1 += __shfl_xor_sync(0xffffffffu, c1, d);
    }
    if (lane < 4) {
      if (c0) atomicAdd(&cnt[j0], c0);
      if (c1) atomicAdd(&cnt[j0 + 1], c1);
    }
  }
}

// ordered emit of the kept entries of every output column into CSC (counts came from the numeric kernel)
__global__ void __launch_bounds__(256)
k_tile_emit(int ncols, int nrows, const int* __restrict__ imin8, const int* __restrict__ nI8,
            const long long* __restrict__ stg_off, const double* __restrict__ stg, double alpha, double thr,
            RuleView rules, const int* __restrict__ outer, int* __restrict__ inner, double* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int j = gw; j < ncols; j += nw) {
    const int J = j >> 3, jj = j & 7;
    const int wlen = nI8[J] * 8;
    const int base = imin8[J] * 8;
    const double* src = stg + stg_off[J] * 64 + (size_t)jj * wlen;
    int count = 0;
    const int dst = outer[j];
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
      if (keep) {
        const int pos = dst + count + __popc(m & ((1u << lane) - 1));
        inner[pos] = base + t;
        val[pos] = sv;
      }
      count += __popc(m);
    }
  }
}

// returns false when the operands are not locally dense enough (caller falls back to the
// scalar window kernels). useful_products = sum over B entries of the A column lengths.
bool spgemm_tile(const CscView<double>& X, const CscView<double>& Y, double alpha, double thr, const RuleView& rules,
                 LocalCsc<double>& Z, double useful_products, long long nnzX, long long nnzY) {
  const int ncols = X.cols, nrows = Y.rows;
  if (ncols == 0 || nnzX == 0 || nnzY == 0 || !(thr >= 0.0)) return false;
  TileCsc A, B;
  if (!build_tiles<8, 4>(Y, A)) return false;
  if ((double)nnzY < 0.20 * 32.0 * (double)A.ntiles) return false;   // tiles mostly padding
  if (!build_tiles<4, 8>(X, B)) return false;
  if ((double)nnzX < 0.20 * 32.0 * (double)B.ntiles) return false;
  const int nJ = B.ntc, nG = div_up(nJ, 8);
  DevBuf<int4> metaA((size_t)A.ntc), metaB((size_t)nJ);
  NTB_LAUNCH(k_tile_meta, div_up(A.ntc, 256), 256, 0, A.ntc, A.tptr.get(), A.first.get(), A.last.get(), metaA.get());
  NTB_LAUNCH(k_tile_meta, div_up(nJ, 256), 256, 0, nJ, B.tptr.get(), B.first.get(), B.last.get(), metaB.get());
  const TileView Av{A.tptr.get(), A.tid.get(), metaA.get(), A.tval.get(), A.ntc};
  const TileView Bv{B.tptr.get(), B.tid.get(), metaB.get(), B.tval.get(), B.ntc};
  DevBuf<int> imin8((size_t)nJ), nI8((size_t)nJ), gbmin((size_t)nG), gnb((size_t)nG), gtask_off((size_t)nG + 1);
  DevBuf<int2> gk((size_t)nG);
  DevBuf<long long> stg_off((size_t)nJ + 1);
  DevBuf<unsigned long long> ndmma(1);
  ndmma.zero();
  NTB_LAUNCH(k_tile_bounds, max(1, min(div_up((long long)nJ * 32, 256), kNumSMs * 16)), 256, 0, Av, Bv, imin8.get(),
             nI8.get(), ndmma.get());
  NTB_LAUNCH(k_group_bounds, div_up(nG, 256), 256, 0, nJ, nG, imin8.get(), nI8.get(), metaB.get(), gbmin.get(),
             gnb.get(), gk.get());
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
  DevBuf<int4> tasks((size_t)max(h_tasks, 1));
  DevBuf<int> cnt((size_t)nJ * 8), task_counter(1);
  cnt.zero();
  task_counter.zero();
  if (h_tasks > 0)
    NTB_LAUNCH(k_task_table, div_up((long long)nG * 32, 256), 256, 0, nG, nJ, gbmin.get(), gtask_off.get(), gk.get(), metaA.get(),
               imin8.get(), nI8.get(), tasks.get());
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (rt().profile) {
    CUDA_CHECK(cudaEventCreate(&ev0));
    CUDA_CHECK(cudaEventCreate(&ev1));
    CUDA_CHECK(cudaEventRecord(ev0, rt().stream));
  }
  if (h_tasks > 0) {
    static int shape = -1;
    if (shape < 0) {
      const char* e = std::getenv("NTB_TILE_PIPE");      // developer knob: 0 = 8x3, 1 = 4x6, 2 = 2x12
      shape = e ? std::atoi(e) : 0;
      CUDA_CHECK(cudaFuncSetAttribute((k_tile_numeric<8, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, PipeCfg<8, 3>::SMEM));
      CUDA_CHECK(cudaFuncSetAttribute((k_tile_numeric<4, 6>), cudaFuncAttributeMaxDynamicSharedMemorySize, PipeCfg<4, 6>::SMEM));
      CUDA_CHECK(cudaFuncSetAttribute((k_tile_numeric<2, 12>), cudaFuncAttributeMaxDynamicSharedMemorySize, PipeCfg<2, 12>::SMEM));
    }
    const int grid = min(h_tasks, kNumSMs * 2);
#define NTB_NUMERIC_ARGS Av, Bv, imin8.get(), nI8.get(), stg_off.get(), tasks.get(), h_tasks, task_counter.get(), stg.get(), \
                         cnt.get(), nrows, ncols, alpha, thr, rules
    if (shape == 0) NTB_LAUNCH((k_tile_numeric<8, 3>), grid, NUMERIC_THREADS, (PipeCfg<8, 3>::SMEM), NTB_NUMERIC_ARGS);
    else if (shape == 2) NTB_LAUNCH((k_tile_numeric<2, 12>), grid, NUMERIC_THREADS, (PipeCfg<2, 12>::SMEM), NTB_NUMERIC_ARGS);
    else NTB_LAUNCH((k_tile_numeric<4, 6>), grid, NUMERIC_THREADS, (PipeCfg<4, 6>::SMEM), NTB_NUMERIC_ARGS);
#undef NTB_NUMERIC_ARGS
  }
  if (rt().profile) {
    CUDA_CHECK(cudaEventRecord(ev1, rt().stream));
    rt().prof_events.emplace_back(ev0, ev1);
  }
  const int egrid = max(1, min(div_up((long long)ncols * 32, 256), kNumSMs * 16));
  Z.rows = nrows; Z.cols = ncols;
  Z.outer.alloc((size_t)ncols + 1);
  exclusive_scan(cnt.get(), Z.outer.get(), ncols);
  int h_nnz = 0;
  d2h(&h_nnz, Z.outer.get() + ncols, 1);
  Z.alloc_entries(h_nnz);
  if (h_nnz > 0)
    NTB_LAUNCH(k_tile_emit, egrid, 256, 0, ncols, nrows, imin8.get(), nI8.get(), stg_off.get(), stg.get(),
               alpha, thr, rules, Z.outer.get(), Z.inner.get(), Z.val.get());
  rt().tile_products++;
  rt().dmma_issued += (double)h_ndmma;
  return true;
}

}  // namespace ntb
