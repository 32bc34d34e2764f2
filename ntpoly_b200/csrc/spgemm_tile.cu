// Tile SpGEMM on the FP64 tensor cores (DMMA.8x8x4) for locally dense operands
// (banded, block-sparse, filled-in purification iterates) — the B200 answer to
// NTPoly's dense-switch idea (reference sparse_includes/GemmMatrix.f90:59-61,
// DenseBranch.f90) applied at 8x4 / 4x8 tile granularity instead of whole blocks.
//
//   C(:,J) = sum_K  A(:,K) * B(K,J)        J: 8 output columns, K: 4 inner indices
//
// Operand layout ("chunked tiles", built per product from CSC):
//   A (the Y operand): 8x4 tiles in DMMA A-fragment order, grouped into SUPER-TILES of
//       64 rows x 32 columns (8 row tiles x 8 inner tiles). The present tiles of a super-tile are
//       contiguous in memory in (inner tile kk, row tile ii) order and described by one
//       64-bit mask (bit kk*8+ii). Super-tiles are listed per chunk column (32 matrix
//       columns), row block ascending.
//   B (the X operand): 4x8 tiles in DMMA B-fragment order, super-tiles of 32 rows x 64
//       columns (8 inner tiles x 8 tile columns), (tile column jj, inner tile kk) order, mask bit
//       jj*8+kk, listed per group of 64 output columns, inner chunk ascending.
//   => the operands of one pipeline stage of the numeric kernel (64x64 output block, 32
//      inner indices) are exactly TWO contiguous byte ranges: one 1-D bulk copy (TMA) each.
//   every (row tile I, K, J) triple with both tiles present = ONE mma.sync.m8n8k4.f64
//   (256 FMAs), accumulators in registers.
//
// Pipeline: k_ct_build (CSC -> chunked tiles, bitmap ranked, 2 passes), k_tile_bounds
// (row-tile window per J), k_tile_numeric (TMA-fed DMMA, threshold counts fused), scan,
// k_tile_emit (threshold rule, alpha, ordered compaction into CSC). Exact zeros introduced
// by tile padding never survive the strict |v| > thr test, so results equal the scalar
// path up to summation order.
#include "csc.cuh"
#include <chrono>

namespace ntb {

constexpr int TW = 8;                        // warps per CTA in the build kernels
constexpr int MAXR = 256;                    // super-tiles of reach per chunk column (bitmap words per warp)
constexpr int IPL = MAXR / 32;               // bitmap words owned by a lane

struct CtView {
  const int4* colmeta; const int4* ent; const double* tval; const int4* kmeta; int ncc;
};

// A: super-tile id = row/64, bit = ((col%32)/4)*8 + (row/8)%8, fragment lane = (row%8)*4 + col%4
// B: super-tile id = row/32, bit = ((col%64)/8)*8 + (row/4)%8, fragment lane = (col%8)*4 + row%4
template <bool ISA> struct CtGeom {
  static constexpr int CW = ISA ? 32 : 64;   // matrix columns per chunk column
  static constexpr int RB = ISA ? 64 : 32;   // matrix rows per super-tile
  __device__ static __forceinline__ int bit(int r, int c) {
    return ISA ? (((c & 31) >> 2) * 8 + ((r >> 3) & 7)) : (((c & 63) >> 3) * 8 + ((r >> 2) & 7));
  }
  __device__ static __forceinline__ int frag(int r, int c) { return ISA ? ((r & 7) * 4 + (c & 3)) : ((c & 7) * 4 + (r & 3)); }
};

__device__ __forceinline__ int popc64(unsigned long long v) { return __popcll(v); }

// pass 1 (FILL=false): super-tile and tile counts per chunk column (+ per-K meta for A);
// pass 2 (FILL=true): entries + values. One warp per chunk column.
template <bool ISA, bool FILL>
__global__ void __launch_bounds__(TW * 32)
k_ct_build(CscView<double> M, int ncc, int* __restrict__ scount, int* __restrict__ tcount, int4* __restrict__ colmeta,
           int4* __restrict__ kmeta, int nk, int* __restrict__ overflow, const int* __restrict__ sptr,
           const int* __restrict__ tptr, int4* __restrict__ ent, double* __restrict__ tval) {
  using G = CtGeom<ISA>;
  __shared__ unsigned long long bm[TW][MAXR];
  __shared__ int pre[TW][MAXR];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long* b = bm[warp];
  for (int q = blockIdx.x * TW + warp; q < ncc; q += gridDim.x * TW) {
    const int c0 = q * G::CW, c1 = min(M.cols, c0 + G::CW);
    int rmin = INT_MAX, rmax = -1;
    for (int c = c0 + lane; c < c1; c += 32) {
      const int s = M.outer[c], e = M.outer[c + 1];
      if (e > s) { rmin = min(rmin, M.inner[s]); rmax = max(rmax, M.inner[e - 1]); }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, d));
      rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, d));
    }
    const bool empty = rmax < 0;
    const int idmin = empty ? 0 : rmin / G::RB, idmax = empty ? -1 : rmax / G::RB;
    const bool over = (idmax - idmin + 1 > MAXR);
    if (empty || over) {
      if (!FILL) {
        if (lane == 0) { scount[q] = 0; tcount[q] = 0; colmeta[q] = make_int4(0, 0, 0, -1); if (over) atomicExch(overflow, 1); }
        if (ISA && lane < 8 && q * 8 + lane < nk) kmeta[q * 8 + lane] = make_int4(0, 0, 0, -1);
      }
      continue;
    }
#pragma unroll
    for (int w = 0; w < IPL; ++w) b[lane * IPL + w] = 0ull;
    __syncwarp();
    for (int c = c0; c < c1; ++c)
      for (int p = M.outer[c] + lane; p < M.outer[c + 1]; p += 32) {
        const int r = M.inner[p];
        const int bit = G::bit(r, c);           // native 32-bit shared-memory atomics on the two halves of the word
        atomicOr(reinterpret_cast<unsigned*>(&b[r / G::RB - idmin]) + (bit >> 5), 1u << (bit & 31));
      }
    __syncwarp();
    // lane owns IPL consecutive ids: prefix of non-empty super-tiles and of tiles
    int ns = 0, nt = 0;
#pragma unroll
    for (int w = 0; w < IPL; ++w) { const unsigned long long m = b[lane * IPL + w]; ns += (m != 0ull); nt += popc64(m); }
    int is = ns, it = nt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int os = __shfl_up_sync(0xffffffffu, is, d), ot = __shfl_up_sync(0xffffffffu, it, d);
      if (lane >= d) { is += os; it += ot; }
    }
    const int tot_s = __shfl_sync(0xffffffffu, is, 31), tot_t = __shfl_sync(0xffffffffu, it, 31);
    if (!FILL) {
      if (lane == 0) { scount[q] = tot_s; tcount[q] = tot_t; }
      if (ISA) {
        // per inner tile kk of this chunk column: tile count, first and last row tile
        for (int kk = 0; kk < 8; ++kk) {
          int cntk = 0, fk = INT_MAX, lk = -1;
#pragma unroll
          for (int w = 0; w < IPL; ++w) {
            const unsigned byte = (unsigned)(b[lane * IPL + w] >> (kk * 8)) & 0xffu;
            if (byte) {
              const int base = (idmin + lane * IPL + w) * 8;
              cntk += __popc(byte);
              fk = min(fk, base + __ffs(byte) - 1);
              lk = max(lk, base + 31 - __clz(byte));
            }
          }
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            cntk += __shfl_xor_sync(0xffffffffu, cntk, d);
            fk = min(fk, __shfl_xor_sync(0xffffffffu, fk, d));
            lk = max(lk, __shfl_xor_sync(0xffffffffu, lk, d));
          }
          if (lane == 0 && q * 8 + kk < nk) kmeta[q * 8 + kk] = (cntk > 0) ? make_int4(0, cntk, fk, lk) : make_int4(0, 0, 0, -1);
        }
      }
      __syncwarp();
      continue;
    }
    const int base_s = sptr[q], base_t = tptr[q];
    if (lane == 0) colmeta[q] = make_int4(base_s, tot_s, idmin, idmax);
    int rs = base_s + is - ns, rt_ = it - nt;
#pragma unroll
    for (int w = 0; w < IPL; ++w) {
      const unsigned long long m = b[lane * IPL + w];
      pre[warp][lane * IPL + w] = rt_;
      if (m != 0ull) {
        ent[rs++] = make_int4(idmin + lane * IPL + w, base_t + rt_, (int)(unsigned)(m & 0xffffffffull), (int)(unsigned)(m >> 32));
        rt_ += popc64(m);
      }
    }
    __syncwarp();
    for (int c = c0; c < c1; ++c)
      for (int p = M.outer[c] + lane; p < M.outer[c + 1]; p += 32) {
        const int r = M.inner[p];
        const int id = r / G::RB - idmin;
        const int bit = G::bit(r, c);
        const int rk = pre[warp][id] + popc64(b[id] & ((1ull << bit) - 1ull));
        tval[((size_t)(base_t + rk)) * 32 + G::frag(r, c)] = M.val[p];
      }
    __syncwarp();
  }
}

template <bool ISA>
static bool build_chunk_tiles(const CscView<double>& M, ChunkTiles& T) {
  using G = CtGeom<ISA>;
  T.ncc = div_up(M.cols, G::CW);
  const int ncc = T.ncc;
  const int nk = ISA ? div_up(M.cols, 4) : 0;
  DevBuf<int> scount((size_t)ncc), tcount((size_t)ncc), sptr((size_t)ncc + 1), tptr((size_t)ncc + 1), overflow(1);
  T.colmeta.alloc((size_t)ncc);
  if (ISA) T.kmeta.alloc((size_t)nk);
  overflow.zero();
  const int grid = max(1, min(div_up(ncc, TW), kNumSMs * 8));
  NTB_LAUNCH((k_ct_build<ISA, false>), grid, TW * 32, 0, M, ncc, scount.get(), tcount.get(), T.colmeta.get(),
             ISA ? T.kmeta.get() : (int4*)nullptr, nk, overflow.get(), (const int*)nullptr, (const int*)nullptr,
             (int4*)nullptr, (double*)nullptr);
  exclusive_scan(scount.get(), sptr.get(), ncc);
  exclusive_scan(tcount.get(), tptr.get(), ncc);
  int h[3] = {0, 0, 0};
  CUDA_CHECK(cudaMemcpyAsync(&h[0], sptr.get() + ncc, sizeof(int), cudaMemcpyDeviceToHost, rt().stream));
  CUDA_CHECK(cudaMemcpyAsync(&h[1], tptr.get() + ncc, sizeof(int), cudaMemcpyDeviceToHost, rt().stream));
  CUDA_CHECK(cudaMemcpyAsync(&h[2], overflow.get(), sizeof(int), cudaMemcpyDeviceToHost, rt().stream));
  stream_sync();
  if (h[2]) return false;
  T.nsuper = h[0];
  T.ntiles = h[1];
  T.ent.alloc((size_t)max(T.nsuper, 1));
  T.tval.alloc((size_t)T.ntiles * 32);
  T.tval.zero();
  NTB_LAUNCH((k_ct_build<ISA, true>), grid, TW * 32, 0, M, ncc, (int*)nullptr, (int*)nullptr, T.colmeta.get(),
             (int4*)nullptr, nk, (int*)nullptr, sptr.get(), tptr.get(), T.ent.get(), T.tval.get());
  if (ISA) T.coltile = std::move(tptr);
  return true;
}

// per output tile column J: row-tile window aligned to blocks of 8 row tiles, and the DMMA count
__global__ void __launch_bounds__(256) k_tile_bounds(CtView A, CtView B, int nJ, int* __restrict__ imin8, int* __restrict__ nI8,
                                                     unsigned long long* __restrict__ ndmma, int diag_on, int dd,
                                                     int ncols_diag, int nrows) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  unsigned long long mine = 0;
  for (int J = gw; J < nJ; J += nw) {
    const int4 cm = B.colmeta[J >> 3];
    const int jj = J & 7;
    int mn = INT_MAX, mx = -1;
    for (int e = lane; e < cm.y; e += 32) {
      const int4 en = B.ent[cm.x + e];
      const unsigned long long m = ((unsigned long long)(unsigned)en.w << 32) | (unsigned)en.z;
      unsigned byte = (unsigned)(m >> (jj * 8)) & 0xffu;
      while (byte) {
        const int kk = __ffs(byte) - 1;
        byte &= byte - 1;
        const int4 km = A.kmeta[en.x * 8 + kk];
        if (km.y > 0) { mn = min(mn, km.z); mx = max(mx, km.w); mine += (unsigned long long)km.y; }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if (diag_on && J * 8 < ncols_diag) {      // the window must hold the rows of the shifted diagonal entries
      const int r0 = J * 8 + dd, r1 = min(J * 8 + 7, ncols_diag - 1) + dd;
      if (r1 >= 0 && r0 < nrows) { mn = min(mn, max(r0, 0) >> 3); mx = max(mx, min(r1, nrows - 1) >> 3); }
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

// task t = (group g, 64-row block Ib)
__global__ void __launch_bounds__(256) k_task_table(int nG, const int* __restrict__ gbmin, const int* __restrict__ gtask_off,
                                                    int2* __restrict__ tasks) {
  const int lane = threadIdx.x & 31;
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= nG) return;
  const int t0 = gtask_off[g], n = gtask_off[g + 1] - t0, b0 = gbmin[g];
  for (int b = lane; b < n; b += 32) tasks[t0 + b] = make_int2(g, b0 + b);
}

__device__ __forceinline__ bool tile_rule(const RuleView& r, int inner_idx, int outer_idx) {
  if (r.tbl == nullptr) return false;
  return r.tbl[(inner_idx / r.rb) * r.nJ + (outer_idx / r.cb)] != 0;
}

// how an accumulated product entry becomes an output entry: threshold rule, alpha, optional diagonal shift
struct EmitSpec {
  double alpha, thr, sigma;
  int dd, ncols_diag;
  RuleView rules;
};
template <bool RULES>
__device__ __forceinline__ double final_value(const EmitSpec& e, double v, int row, int col, bool& keep) {
  double sv = e.alpha * v;
  keep = ((RULES && tile_rule(e.rules, row, col)) ? fabs(v) : fabs(sv)) > e.thr;
  if (!keep) sv = 0.0;
  if (e.sigma != 0.0 && row == col + e.dd && col < e.ncols_diag) { sv += e.sigma; keep = (sv != 0.0); }
  return sv;
}

// out-of-line: keeps the (rare) rule-table / shifted-diagonal test out of the numeric kernel's hot code
__device__ __noinline__ bool keep_general(const EmitSpec& e, double v, int row, int col) {
  bool keep;
  (void)final_value<true>(e, v, row, col, keep);
  return keep;
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

// HBM -> L2 prefetch of a byte range that a later bulk copy will read (no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// pipeline shape: a stage = one A super-tile (<= 64 tiles, 16 KB) + one B super-tile (16 KB); 3 stages = 96 KB per
// CTA, two CTAs per SM
constexpr int NSTAGE_DEFAULT = 3;
constexpr int SLAB_DOUBLES = 64 * 32;                        // one super-tile, all tiles present
constexpr int STAGE_DOUBLES = 2 * SLAB_DOUBLES;
constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;               // 32 KB
constexpr int META_BYTES = 32;                               // maskA (8 B), maskB (8 B), flags, g, Ib, pad
constexpr int numeric_smem(int nstage) { return nstage * STAGE_BYTES + nstage * META_BYTES + 2 * nstage * 8; }
constexpr int CW = 8;                                        // DMMA warps: one per tile column of the group
constexpr int NUMERIC_THREADS = (CW + 1) * 32;               // + 1 copy warp

__device__ __forceinline__ unsigned long long mask64(const int4& e) {
  return ((unsigned long long)(unsigned)e.w << 32) | (unsigned)e.z;
}
// bit kk set iff byte kk of m is non-zero
__device__ __forceinline__ unsigned nonzero_bytes(unsigned long long m) {
  unsigned r = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) r |= (((unsigned)(m >> (8 * k)) & 0xffu) ? 1u : 0u) << k;
  return r;
}
__device__ __forceinline__ unsigned or_bytes(unsigned long long m) {
  unsigned long long t = m | (m >> 32);
  t |= t >> 16;
  t |= t >> 8;
  return (unsigned)t & 0xffu;
}

// entry index of super-tile `id` in a chunk column (binary search unless the ids are one contiguous run), or -1
__device__ __forceinline__ int ct_find(const CtView& A, const int4& ca, int id) {
  if (ca.y <= 0 || id < ca.z || id > ca.w) return -1;
  if (ca.w - ca.z + 1 == ca.y) return ca.x + (id - ca.z);
  int l2 = 0, h2 = ca.y;
  while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (A.ent[ca.x + mid].x < id) l2 = mid + 1; else h2 = mid; }
  return (l2 < ca.y) ? ca.x + l2 : -1;
}
// the (A super-tile, B super-tile) pair of one inner chunk: masks and tile offsets when they share an inner tile
__device__ __forceinline__ void ct_pair(const int4& ea, const int4& eb, int Ib, bool found, unsigned long long& mA,
                                        unsigned long long& mB, int& offA, int& offB) {
  mA = 0ull; mB = 0ull; offA = 0; offB = 0;
  if (found && ea.x == Ib) {
    const unsigned long long ma = mask64(ea), mb = mask64(eb);
    if (nonzero_bytes(ma) & or_bytes(mb)) { mA = ma; mB = mb; offA = ea.y; offB = eb.y; }
  }
}

// Persistent CTAs; task = 64x64 output block (8 tile columns x 8 row tiles), handed out by an atomic counter.
// Warp 8 (copy warp): at the start of a task each lane looks up one inner chunk: the B super-tile (chunk, group)
// and the A super-tile (row block, chunk). For every chunk where both exist and share an inner tile, one lane
// publishes the two 64-bit presence masks and issues TWO 1-D bulk copies (TMA) into a 3-stage shared-memory ring
// guarded by full/empty mbarriers. Warps 0-7: warp w owns tile column 8g + w, i.e. a 64x8 strip of the block with
// its 8 accumulator tiles in registers, and issues one DMMA.8x8x4 per (present A tile, present B tile) pair from
// conflict-free 256-byte shared-memory fragments; tiles are located by popcount rank in the masks. Absent tiles
// are skipped with real branches (a predicated-off DMMA occupies the pipe for its full 16 cycles). The strip is
// written to the dense staging window and the kept-entry counts of its 8 columns are accumulated on the fly
// (threshold rule fused), so the emit pass is a single sweep.
template <int NSTAGE, int MINB>
__global__ void __launch_bounds__(NUMERIC_THREADS, MINB)
k_tile_numeric(CtView A, CtView B, int nJ, const int* __restrict__ imin8, const int* __restrict__ nI8,
               const long long* __restrict__ stg_off, const int2* __restrict__ tasks, int ntasks,
               int* __restrict__ task_counter, double* __restrict__ stg, int* __restrict__ cnt,
               unsigned char* __restrict__ fmA, unsigned char* __restrict__ fmB, int nrows, int ncols, EmitSpec es) {
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
    // The look-ups of a task are a chain of five dependent global loads (task -> B column meta -> B super-tile
    // -> A column meta -> A super-tile). They are software-pipelined: while the stages of task n are pushed,
    // one link of the chain of task n+1 is advanced per pushed stage, so the latency hides behind the ring.
    const int4 none = make_int4(0, 0, 0, -1);
    int task_raw = 0;                               // lane 0: task id returned by the atomic counter
    int pstep = 0;
    bool valid_n = false, have_n = false, found_n = false;
    int tid_n = 0;
    int2 tk_n = make_int2(0, 0);
    int4 cmB_n = none, eb_n = none, ca_n = none, ea_n = none;
    unsigned long long mA_n = 0ull, mB_n = 0ull;
    int offA_n = 0, offB_n = 0;
    auto advance = [&]() {
      switch (pstep) {
        case 0: {
          const int t = __shfl_sync(0xffffffffu, task_raw, 0);
          tid_n = t;
          valid_n = t < ntasks;
          tk_n = valid_n ? tasks[t] : make_int2(0, 0);
          break;
        }
        case 1: cmB_n = valid_n ? B.colmeta[tk_n.x] : none; break;
        case 2: have_n = valid_n && lane < cmB_n.y; eb_n = have_n ? B.ent[cmB_n.x + lane] : none; break;
        case 3: ca_n = have_n ? A.colmeta[eb_n.x] : none; break;
        case 4: {
          const int idx = have_n ? ct_find(A, ca_n, tk_n.y) : -1;
          found_n = idx >= 0;
          ea_n = found_n ? A.ent[idx] : none;
          break;
        }
        case 5: ct_pair(ea_n, eb_n, tk_n.y, found_n, mA_n, mB_n, offA_n, offB_n); break;
        default: break;
      }
      ++pstep;
    };
    if (lane == 0) task_raw = atomicAdd(task_counter, 1);
    while (pstep < 6) advance();
    for (;;) {
      const bool done = !valid_n;
      const int g = tk_n.x, Ib = tk_n.y, task = tid_n;
      const int4 cmB = cmB_n;
      unsigned long long mA = mA_n, mB = mB_n;
      int offA = offA_n, offB = offB_n;
      if (!done) {
        if (lane == 0) task_raw = atomicAdd(task_counter, 1);
        pstep = 0;
      }
      const int nb = done ? 1 : max(1, (cmB.y + 31) / 32);
      for (int bb = 0; bb < nb; ++bb) {
        if (bb > 0) {                               // more than 32 inner chunks in this group: synchronous look-up
          const int e = bb * 32 + lane;
          const bool have = e < cmB.y;
          const int4 eb = have ? B.ent[cmB.x + e] : none;
          const int4 ca = have ? A.colmeta[eb.x] : none;
          const int idx = have ? ct_find(A, ca, Ib) : -1;
          const int4 ea = (idx >= 0) ? A.ent[idx] : none;
          ct_pair(ea, eb, Ib, idx >= 0, mA, mB, offA, offB);
        }
        unsigned todo = __ballot_sync(0xffffffffu, mA != 0ull);
        const bool final_batch = (bb == nb - 1);
        if (todo == 0u && final_batch) todo = 1u;             // a task always ends with a (possibly empty) last stage
        while (todo) {
          const int l = __ffs(todo) - 1;
          todo &= todo - 1;
          const bool last = final_batch && todo == 0u;
          const unsigned long long sA = __shfl_sync(0xffffffffu, mA, l), sB = __shfl_sync(0xffffffffu, mB, l);
          const int oA = __shfl_sync(0xffffffffu, offA, l), oB = __shfl_sync(0xffffffffu, offB, l);
          mbar_wait(bar0 + 8 * (NSTAGE + st), ph ^ 1u);        // consumers have released this slot
          if (lane == 0) {
            unsigned char* mt = meta + st * META_BYTES;
            *reinterpret_cast<unsigned long long*>(mt) = sA;
            *reinterpret_cast<unsigned long long*>(mt + 8) = sB;
            *reinterpret_cast<int4*>(mt + 16) = make_int4((last ? 1 : 0) | (done ? 2 : 0), g, Ib, task);
            const unsigned bA = (unsigned)popc64(sA) * 256u, bB = (unsigned)popc64(sB) * 256u;
            mbar_arrive_expect_tx(bar0 + 8 * st, bA + bB);
            const unsigned slab_s = smem_u32(slab + (size_t)st * STAGE_DOUBLES);
            if (bA) bulk_g2s(slab_s, A.tval + (size_t)oA * 32, bA, bar0 + 8 * st);
            if (bB) bulk_g2s(slab_s + SLAB_DOUBLES * 8, B.tval + (size_t)oB * 32, bB, bar0 + 8 * st);
          }
          __syncwarp();
          if (++st == NSTAGE) { st = 0; ph ^= 1u; }
          if (!done) advance();
        }
      }
      if (done) break;
      while (pstep < 6) advance();
    }
    return;
  }

  // -------------------------------------------------------------------- DMMA warps
  const int wj = warp;                              // tile column of the group
  for (;;) {
    double acc[8][2];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) { acc[ii][0] = 0.0; acc[ii][1] = 0.0; }
    int g = 0, Ib = 0, task = 0;
    unsigned fl = 0;
    do {
      mbar_wait(bar0 + 8 * st, ph);
      const unsigned char* mt = meta + st * META_BYTES;
      const ulonglong2 mm = *reinterpret_cast<const ulonglong2*>(mt);     // maskA, maskB
      const int4 mi = *reinterpret_cast<const int4*>(mt + 16);
      fl = (unsigned)mi.x; g = mi.y; Ib = mi.z; task = mi.w;
      const unsigned mb = (unsigned)(mm.y >> (8 * wj)) & 0xffu;           // my tile column: bits over kk
      if (mb != 0u && mm.x != 0ull) {
        // tiles are rank-packed: byte kk of `excl` = number of A tiles stored before inner tile kk
        unsigned long long x = mm.x - ((mm.x >> 1) & 0x5555555555555555ull);
        x = (x & 0x3333333333333333ull) + ((x >> 2) & 0x3333333333333333ull);
        x = (x + (x >> 4)) & 0x0f0f0f0f0f0f0f0full;
        const unsigned long long excl = (x * 0x0101010101010101ull) << 8;
        const double* As = slab + (size_t)st * STAGE_DOUBLES + lane;
        const double* Bs = As + SLAB_DOUBLES + popc64(mm.y & ((1ull << (8 * wj)) - 1ull)) * 32;
        // one copy of the loop body (not unrolled over kk): the four code paths below times eight would not
        // fit the instruction cache (measured: stall_no_instruction dominated)
        unsigned live = mb & nonzero_bytes(mm.x);
#pragma unroll 1
        while (live) {
          const int kk = __ffs(live) - 1;
          live &= live - 1u;
          const unsigned ma = (unsigned)(mm.x >> (8 * kk)) & 0xffu;
          const double bv = Bs[__popc(mb & ((1u << kk) - 1u)) * 32];
          const double* ap = As + ((unsigned)(excl >> (8 * kk)) & 0xffu) * 32;
          double av[8];
          if (ma == 0xffu) {
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) av[ii] = ap[ii * 32];
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) dmma884(acc[ii][0], acc[ii][1], av[ii], bv);
          } else if ((ma & (ma + 1u)) == 0u) {
            // Absent tiles are skipped with branches, not predicates: a predicated-off DMMA still holds the FP64
            // tensor pipe for 16 cycles (scripts/micro/dmma_shapes.cu).
            // prefix run (row tiles 0..h0, the band edge leaving the block): one jump into a descending sequence
            const int h0 = __popc(ma) - 1;
#pragma unroll
            for (int ii = 0; ii < 7; ++ii)
              if (ii <= h0) av[ii] = ap[ii * 32];
#define NTB_D(i) dmma884(acc[i][0], acc[i][1], av[i], bv);
            switch (h0) {
              case 6: NTB_D(6)
              case 5: NTB_D(5)
              case 4: NTB_D(4)
              case 3: NTB_D(3)
              case 2: NTB_D(2)
              case 1: NTB_D(1)
              default: NTB_D(0)
            }
          } else if (((ma | (ma - 1u)) & 0xffu) == 0xffu) {
            // suffix run (row tiles l0..7, the band edge entering the block): one jump into an ascending sequence
            const int l0 = __ffs(ma) - 1;
            const double* aq = ap - l0 * 32;
#pragma unroll
            for (int ii = 1; ii < 8; ++ii)
              if (ii >= l0) av[ii] = aq[ii * 32];
            switch (l0) {
              case 1: NTB_D(1)
              case 2: NTB_D(2)
              case 3: NTB_D(3)
              case 4: NTB_D(4)
              case 5: NTB_D(5)
              case 6: NTB_D(6)
              default: NTB_D(7)
            }
#undef NTB_D
          } else {
            // general pattern: load the present tiles (rank-packed), then enter the unrolled DMMA sequence at the
            // first tile of each id run and leave it after the last
#pragma unroll
            for (int ii = 0; ii < 8; ++ii)
              if ((ma >> ii) & 1u) av[ii] = ap[__popc(ma & ((1u << ii) - 1u)) * 32];
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
                case 3: NTB_RUN_STEP(3)
                case 4: NTB_RUN_STEP(4)
                case 5: NTB_RUN_STEP(5)
                case 6: NTB_RUN_STEP(6)
                default: dmma884(acc[7][0], acc[7][1], av[7], bv);
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
    if (J >= nJ) continue;
    const int I0 = Ib << 3;
    const int iw0 = imin8[J], nI = nI8[J];
    if (I0 < iw0 || I0 >= iw0 + nI) continue;
    // C fragment: row = lane/4, cols = 2*(lane%4), +1 ; staging is column-major per tile column.
    // Alongside the raw strip: kept-entry counts of the 8 columns and the presence bytes of the strip's tiles in
    // the two tile forms of the RESULT (the next product reads them without ever rebuilding tiles from CSC).
    const int wlen = nI * 8;
    const int r = lane >> 2, cc = (lane & 3) * 2;
    double* o = stg + stg_off[J] * 64 + (size_t)cc * wlen + (size_t)(I0 - iw0) * 8 + r;
    const int j0 = J * 8 + cc;
    const bool in0 = j0 < ncols, in1 = j0 + 1 < ncols;
    // the shifted diagonal touches only the strips that cross it: everything else takes the plain test
    const bool on_diag = es.sigma != 0.0 && (I0 * 8 <= J * 8 + 7 + es.dd) && (I0 * 8 + 63 >= J * 8 + es.dd);
    const bool plain = es.rules.tbl == nullptr && !on_diag;
    int c0 = 0, c1 = 0;
    unsigned bR0 = 0, bR1 = 0, bL0 = 0, bL1 = 0;   // right-form bytes (rows 0-31 / 32-63), left-form bytes (cols 0-3 / 4-7)
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      const double v0 = acc[ii][0], v1 = acc[ii][1];
      o[ii * 8] = v0;
      o[wlen + ii * 8] = v1;
      const int row = (I0 + ii) * 8 + r;
      bool k0 = false, k1 = false;
      if (row < nrows) {
        if (plain) {                            // sparse rule everywhere, no shift: |alpha*v| > thr
          k0 = in0 && fabs(es.alpha * v0) > es.thr;
          k1 = in1 && fabs(es.alpha * v1) > es.thr;
        } else {
          if (in0) k0 = keep_general(es, v0, row, j0);
          if (in1) k1 = keep_general(es, v1, row, j0 + 1);
        }
      }
      c0 += k0 ? 1 : 0;
      c1 += k1 ? 1 : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, k0 || k1);
      const unsigned lo = (bal & 0x0000ffffu) ? 1u : 0u, hi = (bal & 0xffff0000u) ? 1u : 0u;
      if (ii < 4) bR0 |= (lo << (2 * ii)) | (hi << (2 * ii + 1));
      else bR1 |= (lo << (2 * (ii - 4))) | (hi << (2 * (ii - 4) + 1));
      bL0 |= ((bal & 0x33333333u) ? 1u : 0u) << ii;
      bL1 |= ((bal & 0xccccccccu) ? 1u : 0u) << ii;
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
    if (lane == 0) {
      fmB[((size_t)task * 2 + 0) * 8 + wj] = (unsigned char)bR0;
      fmB[((size_t)task * 2 + 1) * 8 + wj] = (unsigned char)bR1;
      unsigned char* fa = fmA + ((size_t)task * 2 + (wj >> 2)) * 8 + 2 * (wj & 3);
      fa[0] = (unsigned char)bL0;
      fa[1] = (unsigned char)bL1;
    }
  }
}

// ---- v9 of the numeric kernel: same pipeline, same results, leaner DMMA warps ---------------------------------------
// ncu's per-instruction samples of the kernel above (profiles/r01b_numeric_source_top.txt) show the DMMA warps spending
// 18 % of their time in per-stage mask arithmetic and barrier turn-around, 18 % in the per-inner-tile address chain
// (BREV/FLO/POPC run on the quarter-rate XU pipe and sit on the critical path of every DMMA burst) and 22 % in the
// epilogue (two dependent global loads before the first store). Here
//   * the COPY WARP digests the masks once per stage (it idles on the empty barrier 95 % of the time anyway): per
//     inner tile kk of the A super-tile a descriptor {presence byte, first tile, first/last row tile, kind}, per
//     (tile column, kk) the index of the B tile in the slab; the DMMA warps only extract bit fields;
//   * the inner loop counts kk up instead of bit-scanning, and prefetches the next descriptor;
//   * the window constants of the epilogue are loaded when a task STARTS, so their latency hides behind the DMMAs;
//   * the presence bytes of the result's tile forms come from two warp-wide OR reductions instead of 8 ballots.
constexpr int META9 = 128;
constexpr int numeric_smem9(int nstage) { return nstage * STAGE_BYTES + nstage * META9 + 2 * nstage * 8; }
//  per stage: +0   int4 {flags | nzA << 8, g, Ib, task}   flags: 1 = last stage of the task, 2 = no more tasks
//             +16  u32 adesc[8]  ma | first tile << 8 | l0 << 16 | h0 << 19 | kind << 22   (kind: 0 full, 1 prefix
//                                run 0..h0, 2 suffix run l0..7, 3 anything else)
//             +48  u8 mbyte[8]   presence bits over kk of tile column jj of the B super-tile
//             +64  u8 boff[64]   [jj*8+kk] index of B tile (jj,kk) in the slab
template <int NSTAGE, int MINB>
__global__ void __launch_bounds__(NUMERIC_THREADS, MINB)
k_tile_numeric9(CtView A, CtView B, int nJ, const int* __restrict__ imin8, const int* __restrict__ nI8,
                const long long* __restrict__ stg_off, const int2* __restrict__ tasks, int ntasks,
                int* __restrict__ task_counter, double* __restrict__ stg, int* __restrict__ cnt,
                unsigned char* __restrict__ fmA, unsigned char* __restrict__ fmB, int nrows, int ncols, EmitSpec es) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* slab = reinterpret_cast<double*>(smem);
  unsigned char* meta = smem + NSTAGE * STAGE_BYTES;
  const unsigned bar0 = smem_u32(smem + NSTAGE * STAGE_BYTES + NSTAGE * META9);   // full[s] at +8s, empty[s] at +8(NSTAGE+s)
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
    // ------------------------------------------------------------------ copy warp (look-up chain as in v8)
    const int4 none = make_int4(0, 0, 0, -1);
    int task_raw = 0;
    int pstep = 0;
    bool valid_n = false, have_n = false, found_n = false;
    int tid_n = 0;
    int2 tk_n = make_int2(0, 0);
    int4 cmB_n = none, eb_n = none, ca_n = none, ea_n = none;
    unsigned long long mA_n = 0ull, mB_n = 0ull;
    int offA_n = 0, offB_n = 0;
    auto advance = [&]() {
      switch (pstep) {
        case 0: {
          const int t = __shfl_sync(0xffffffffu, task_raw, 0);
          tid_n = t;
          valid_n = t < ntasks;
          tk_n = valid_n ? tasks[t] : make_int2(0, 0);
          break;
        }
        case 1: cmB_n = valid_n ? B.colmeta[tk_n.x] : none; break;
        case 2: have_n = valid_n && lane < cmB_n.y; eb_n = have_n ? B.ent[cmB_n.x + lane] : none; break;
        case 3: ca_n = have_n ? A.colmeta[eb_n.x] : none; break;
        case 4: {
          const int idx = have_n ? ct_find(A, ca_n, tk_n.y) : -1;
          found_n = idx >= 0;
          ea_n = found_n ? A.ent[idx] : none;
          break;
        }
        case 5:
          ct_pair(ea_n, eb_n, tk_n.y, found_n, mA_n, mB_n, offA_n, offB_n);
          // the stages of the NEXT task are now known: pull their tiles into L2 while this task is still being
          // consumed, so that their bulk copies find them there (the DMMA warps were waiting 12 % of their time for
          // a full barrier, i.e. for HBM latency at the short band-edge stages)
          if (mA_n != 0ull) {
            bulk_prefetch_l2(A.tval + (size_t)offA_n * 32, (unsigned)popc64(mA_n) * 256u);
            bulk_prefetch_l2(B.tval + (size_t)offB_n * 32, (unsigned)popc64(mB_n) * 256u);
          }
          break;
        default: break;
      }
      ++pstep;
    };
    if (lane == 0) task_raw = atomicAdd(task_counter, 1);
    while (pstep < 6) advance();
    for (;;) {
      const bool done = !valid_n;
      const int g = tk_n.x, Ib = tk_n.y, task = tid_n;
      const int4 cmB = cmB_n;
      unsigned long long mA = mA_n, mB = mB_n;
      int offA = offA_n, offB = offB_n;
      if (!done) {
        if (lane == 0) task_raw = atomicAdd(task_counter, 1);
        pstep = 0;
      }
      const int nb = done ? 1 : max(1, (cmB.y + 31) / 32);
      for (int bb = 0; bb < nb; ++bb) {
        if (bb > 0) {
          const int e = bb * 32 + lane;
          const bool have = e < cmB.y;
          const int4 eb = have ? B.ent[cmB.x + e] : none;
          const int4 ca = have ? A.colmeta[eb.x] : none;
          const int idx = have ? ct_find(A, ca, Ib) : -1;
          const int4 ea = (idx >= 0) ? A.ent[idx] : none;
          ct_pair(ea, eb, Ib, idx >= 0, mA, mB, offA, offB);
        }
        unsigned todo = __ballot_sync(0xffffffffu, mA != 0ull);
        const bool final_batch = (bb == nb - 1);
        if (todo == 0u && final_batch) todo = 1u;
        while (todo) {
          const int l = __ffs(todo) - 1;
          todo &= todo - 1;
          const bool last = final_batch && todo == 0u;
          const unsigned long long sA = __shfl_sync(0xffffffffu, mA, l), sB = __shfl_sync(0xffffffffu, mB, l);
          const int oA = __shfl_sync(0xffffffffu, offA, l), oB = __shfl_sync(0xffffffffu, offB, l);
          // digest the masks while the slot may still be busy
          unsigned desc = 0;
          if (lane < 8) {
            const unsigned ma = (unsigned)(sA >> (8 * lane)) & 0xffu;
            if (ma) {
              const unsigned aoff = (unsigned)popc64(sA & ((1ull << (8 * lane)) - 1ull));
              const unsigned l0 = (unsigned)__ffs(ma) - 1u, h0 = 31u - (unsigned)__clz(ma), pc = (unsigned)__popc(ma);
              unsigned kind = 3u;
              if (ma == 0xffu) kind = 0u;
              else if (l0 == 0u && pc == h0 + 1u) kind = 1u;
              else if (h0 == 7u && pc == 8u - l0) kind = 2u;
              desc = ma | (aoff << 8) | (l0 << 16) | (h0 << 19) | (kind << 22);
            }
          }
          const unsigned b0 = (unsigned)popc64(sB & ((1ull << (2 * lane)) - 1ull));
          const unsigned b1 = b0 + (unsigned)((sB >> (2 * lane)) & 1ull);
          const unsigned nzA = nonzero_bytes(sA);
          mbar_wait(bar0 + 8 * (NSTAGE + st), ph ^ 1u);        // consumers have released this slot
          unsigned char* mt = meta + st * META9;
          if (lane < 8) reinterpret_cast<unsigned*>(mt + 16)[lane] = desc;
          reinterpret_cast<unsigned short*>(mt + 64)[lane] = (unsigned short)(b0 | (b1 << 8));
          if (lane == 0) {
            *reinterpret_cast<unsigned long long*>(mt + 48) = sB;
            *reinterpret_cast<int4*>(mt) = make_int4((int)((last ? 1u : 0u) | (done ? 2u : 0u) | (nzA << 8)), g, Ib, task);
          }
          __syncwarp();                                        // the other lanes' meta stores happen before the arrive
          if (lane == 0) {
            const unsigned bA = (unsigned)popc64(sA) * 256u, bB = (unsigned)popc64(sB) * 256u;
            mbar_arrive_expect_tx(bar0 + 8 * st, bA + bB);
            const unsigned slab_s = smem_u32(slab + (size_t)st * STAGE_DOUBLES);
            if (bA) bulk_g2s(slab_s, A.tval + (size_t)oA * 32, bA, bar0 + 8 * st);
            if (bB) bulk_g2s(slab_s + SLAB_DOUBLES * 8, B.tval + (size_t)oB * 32, bB, bar0 + 8 * st);
          }
          __syncwarp();
          if (++st == NSTAGE) { st = 0; ph ^= 1u; }
          if (!done) advance();
        }
      }
      if (done) break;
      while (pstep < 6) advance();
    }
    return;
  }

  // -------------------------------------------------------------------- DMMA warps
  const int wj = warp;                              // tile column of the group
  for (;;) {
    double acc[8][2];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) { acc[ii][0] = 0.0; acc[ii][1] = 0.0; }
    int g = 0, Ib = 0, task = 0;
    unsigned fl = 0;
    bool fresh = true;
    int iw0 = 0, nI = 0;
    long long so = 0;
    do {
      mbar_wait(bar0 + 8 * st, ph);
      const unsigned char* mt = meta + st * META9;
      const int4 mi = *reinterpret_cast<const int4*>(mt);
      fl = (unsigned)mi.x & 0xffu; g = mi.y; Ib = mi.z; task = mi.w;
      if (fresh) {                                   // a task starts: issue the loads its epilogue will need
        fresh = false;
        const int J = g * 8 + wj;
        if (!(fl & 2u) && J < nJ) { iw0 = imin8[J]; nI = nI8[J]; so = stg_off[J]; }
      }
      const unsigned mb = mt[48 + wj];
      const unsigned live = mb & ((unsigned)mi.x >> 8) & 0xffu;
      if (live != 0u) {
        const unsigned long long bo = *reinterpret_cast<const unsigned long long*>(mt + 64 + 8 * wj);
        const unsigned* adesc = reinterpret_cast<const unsigned*>(mt + 16);
        const double* As = slab + (size_t)st * STAGE_DOUBLES + lane;
        const double* Bs = As + SLAB_DOUBLES;
        unsigned dn = adesc[0];
#pragma unroll 1
        for (int kk = 0; (live >> kk) != 0u; ++kk) {
          const unsigned d = dn;
          dn = adesc[(kk + 1) & 7];
          if (((live >> kk) & 1u) == 0u) continue;
          const unsigned ma = d & 0xffu;
          const double bv = Bs[((unsigned)(bo >> (8 * kk)) & 0xffu) * 32];
          const double* ap = As + ((d >> 8) & 0xffu) * 32;
          const unsigned kind = (d >> 22) & 3u;
          double av[8];
#define NTB_D(i) dmma884(acc[i][0], acc[i][1], av[i], bv);
          if (kind == 0u) {
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) av[ii] = ap[ii * 32];
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) dmma884(acc[ii][0], acc[ii][1], av[ii], bv);
          } else if (kind == 1u) {
            // prefix run (row tiles 0..h0): one jump into a descending sequence (real branches: a predicated-off DMMA
            // still holds the FP64 tensor pipe for 16 cycles)
            const int h0 = (int)((d >> 19) & 7u);
#pragma unroll
            for (int ii = 0; ii < 7; ++ii)
              if (ii <= h0) av[ii] = ap[ii * 32];
            switch (h0) {
              case 6: NTB_D(6)
              case 5: NTB_D(5)
              case 4: NTB_D(4)
              case 3: NTB_D(3)
              case 2: NTB_D(2)
              case 1: NTB_D(1)
              default: NTB_D(0)
            }
          } else if (kind == 2u) {
            // suffix run (row tiles l0..7): one jump into an ascending sequence
            const int l0 = (int)((d >> 16) & 7u);
            const double* aq = ap - l0 * 32;
#pragma unroll
            for (int ii = 1; ii < 8; ++ii)
              if (ii >= l0) av[ii] = aq[ii * 32];
            switch (l0) {
              case 1: NTB_D(1)
              case 2: NTB_D(2)
              case 3: NTB_D(3)
              case 4: NTB_D(4)
              case 5: NTB_D(5)
              case 6: NTB_D(6)
              default: NTB_D(7)
            }
          } else {
#pragma unroll
            for (int ii = 0; ii < 8; ++ii)
              if ((ma >> ii) & 1u) av[ii] = ap[__popc(ma & ((1u << ii) - 1u)) * 32];
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
                case 3: NTB_RUN_STEP(3)
                case 4: NTB_RUN_STEP(4)
                case 5: NTB_RUN_STEP(5)
                case 6: NTB_RUN_STEP(6)
                default: dmma884(acc[7][0], acc[7][1], av[7], bv);
              }
            } while (m);
#undef NTB_RUN_STEP
          }
#undef NTB_D
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8 * (NSTAGE + st));
      if (++st == NSTAGE) { st = 0; ph ^= 1u; }
    } while ((fl & 1u) == 0u);
    if (fl & 2u) break;
    const int J = g * 8 + wj;
    if (J >= nJ) continue;
    const int I0 = Ib << 3;
    if (I0 < iw0 || I0 >= iw0 + nI) continue;
    // C fragment: row = lane/4, cols = 2*(lane%4), +1 ; staging is column-major per tile column.
    const int wlen = nI * 8;
    const int r = lane >> 2, cc = (lane & 3) * 2;
    double* o = stg + so * 64 + (size_t)cc * wlen + (size_t)(I0 - iw0) * 8 + r;
    const int j0 = J * 8 + cc;
    const bool in0 = j0 < ncols, in1 = j0 + 1 < ncols;
    const bool on_diag = es.sigma != 0.0 && (I0 * 8 <= J * 8 + 7 + es.dd) && (I0 * 8 + 63 >= J * 8 + es.dd);
    const bool norules = es.rules.tbl == nullptr;
    int c0 = 0, c1 = 0;
    unsigned km = 0;                               // bit ii: this lane keeps an entry of row tile ii
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      const double v0 = acc[ii][0], v1 = acc[ii][1];
      o[ii * 8] = v0;
      o[wlen + ii * 8] = v1;
      const int row = (I0 + ii) * 8 + r;
      bool k0 = false, k1 = false;
      if (row < nrows) {
        if (norules) {                          // sparse rule everywhere: |alpha*v| > thr
          const double s0 = es.alpha * v0, s1 = es.alpha * v1;
          k0 = in0 && fabs(s0) > es.thr;
          k1 = in1 && fabs(s1) > es.thr;
          if (on_diag) {                        // a shifted diagonal entry is kept iff it is non-zero (final_value)
            if (in0 && row == j0 + es.dd && j0 < es.ncols_diag) k0 = ((k0 ? s0 : 0.0) + es.sigma) != 0.0;
            if (in1 && row == j0 + 1 + es.dd && j0 + 1 < es.ncols_diag) k1 = ((k1 ? s1 : 0.0) + es.sigma) != 0.0;
          }
        } else {
          if (in0) k0 = keep_general(es, v0, row, j0);
          if (in1) k1 = keep_general(es, v1, row, j0 + 1);
        }
      }
      c0 += k0 ? 1 : 0;
      c1 += k1 ? 1 : 0;
      km |= ((k0 || k1) ? 1u : 0u) << ii;
    }
    // presence bytes of the strip's tiles in the two forms of the RESULT: right form = (rows 0-31 / 32-63) x inner
    // tiles of 4 rows (lanes 0-15 hold rows 0-3 of a row tile, lanes 16-31 rows 4-7); left form = (columns 0-3 /
    // 4-7) x row tiles (lanes with lane%4 < 2 hold columns 0-3)
    unsigned x = km;
    x = (x | (x << 4)) & 0x0f0fu;
    x = (x | (x << 2)) & 0x3333u;
    x = (x | (x << 1)) & 0x5555u;                  // bit ii -> bit 2*ii
    const unsigned bR = __reduce_or_sync(0xffffffffu, (lane & 16) ? (x << 1) : x);
    const unsigned bL = __reduce_or_sync(0xffffffffu, (lane & 2) ? (km << 8) : km);
#pragma unroll
    for (int d = 4; d < 32; d <<= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, d);
      c1 += __shfl_xor_sync(0xffffffffu, c1, d);
    }
    if (lane < 4) {
      if (c0) atomicAdd(&cnt[j0], c0);
      if (c1) atomicAdd(&cnt[j0 + 1], c1);
    }
    if (lane == 0) {
      fmB[((size_t)task * 2 + 0) * 8 + wj] = (unsigned char)(bR & 0xffu);
      fmB[((size_t)task * 2 + 1) * 8 + wj] = (unsigned char)((bR >> 8) & 0xffu);
      unsigned char* fa = fmA + ((size_t)task * 2 + (wj >> 2)) * 8 + 2 * (wj & 3);
      fa[0] = (unsigned char)(bL & 0xffu);
      fa[1] = (unsigned char)((bL >> 8) & 0xffu);
    }
  }
}

// ordered emit of the kept entries of every output column into CSC (counts came from the numeric kernel)
__global__ void __launch_bounds__(256)
k_tile_emit(int ncols, int nrows, const int* __restrict__ imin8, const int* __restrict__ nI8,
            const long long* __restrict__ stg_off, const double* __restrict__ stg, EmitSpec es,
            const int* __restrict__ outer, int* __restrict__ inner, double* __restrict__ val) {
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
      if (t < wlen && base + t < nrows) sv = final_value<true>(es, src[t], base + t, j, keep);
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

// ---- tile forms of the result, straight from the staging strips --------------------------------------------
// Every task (group g, row block Ib) owns two super-tiles of each form: left form (Ib, chunk column 2g+c), right
// form (inner chunk 2Ib+h, group g); their presence masks were written by the numeric kernel (fmA / fmB).
// Slots are enumerated in chunk-column order: right form slot = 2*task + h; left form slot = 2*t0(g) + c*n(g) + b.
constexpr int FI_T = 1024, FI_ITEMS = 8;

// slot s of a form -> its task-owned mask, super-tile id and the slot range [sa, sb) of its chunk column
__device__ __forceinline__ void form_slot(bool left, int s, const int2* __restrict__ tasks, const int* __restrict__ gtask_off,
                                          const unsigned long long* __restrict__ fmA, const unsigned long long* __restrict__ fmB,
                                          unsigned long long& mask, int& id, int& q, int& sa, int& sb) {
  const int t = s >> 1;
  const int2 tk = tasks[t];
  const int t0 = gtask_off[tk.x], nn = gtask_off[tk.x + 1] - t0;
  if (!left) { mask = fmB[s]; id = 2 * tk.y + (s & 1); q = tk.x; sa = 2 * t0; sb = sa + 2 * nn; return; }
  const int local = s - 2 * t0, c = local / nn, b = local - c * nn;
  mask = fmA[(size_t)(t0 + b) * 2 + c];
  id = tasks[t0 + b].y;
  q = 2 * tk.x + c; sa = 2 * t0 + c * nn; sb = sa + nn;
}

// (1) masks in slot order, both forms (blockIdx.y: 0 = left, 1 = right)
__global__ void __launch_bounds__(256)
k_forms_slots(int ntasks, const int2* __restrict__ tasks, const int* __restrict__ gtask_off,
              const unsigned long long* __restrict__ fmA, const unsigned long long* __restrict__ fmB,
              unsigned long long* __restrict__ smaskL, unsigned long long* __restrict__ smaskR) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= 2 * ntasks) return;
  const bool left = blockIdx.y == 0;
  unsigned long long m; int id, q, sa, sb;
  form_slot(left, s, tasks, gtask_off, fmA, fmB, m, id, q, sa, sb);
  (left ? smaskL : smaskR)[s] = m;
}

// (2) one CTA per form: exclusive scans of (non-empty, tile count) over the slots. Every thread owns one contiguous
// chunk of slots: chunk totals -> one block-wide scan -> outputs (two sweeps over L2-resident data, three barriers)
__global__ void __launch_bounds__(FI_T)
k_forms_scan(int ntasks, const unsigned long long* __restrict__ smaskL, const unsigned long long* __restrict__ smaskR,
             int* __restrict__ seidxL, int* __restrict__ seidxR, int* __restrict__ stoffL, int* __restrict__ stoffR,
             int* __restrict__ totals) {
  const bool left = blockIdx.x == 0;
  const unsigned long long* smask = left ? smaskL : smaskR;
  int* seidx = left ? seidxL : seidxR;
  int* stoff = left ? stoffL : stoffR;
  const int n = 2 * ntasks;
  __shared__ int wsum[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (n + FI_T - 1) / FI_T;
  const int s0 = min(n, (int)threadIdx.x * per), s1 = min(n, s0 + per);
  int f = 0, pc = 0;
#pragma unroll 4
  for (int i = s0; i < s1; ++i) { const unsigned long long m = smask[i]; f += (m != 0ull); pc += popc64(m); }
  int fi = f, pi = pc;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, fi, d), b = __shfl_up_sync(0xffffffffu, pi, d);
    if (lane >= d) { fi += a; pi += b; }
  }
  if (lane == 31) { wsum[0][warp] = fi; wsum[1][warp] = pi; }
  __syncthreads();
  if (warp == 0) {
    const int a = wsum[0][lane], b = wsum[1][lane];
    int ai = a, bi = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, ai, d), y = __shfl_up_sync(0xffffffffu, bi, d);
      if (lane >= d) { ai += x; bi += y; }
    }
    wsum[0][lane] = ai - a; wsum[1][lane] = bi - b;      // exclusive warp offsets
  }
  __syncthreads();
  int e = wsum[0][warp] + fi - f, t = wsum[1][warp] + pi - pc;
#pragma unroll 4
  for (int i = s0; i < s1; ++i) {
    const unsigned long long m = smask[i];
    seidx[i] = e; stoff[i] = t;
    if (m != 0ull) { ++e; t += popc64(m); }
  }
  if (threadIdx.x == FI_T - 1) { seidx[n] = e; totals[left ? 0 : 2] = e; totals[left ? 1 : 3] = t; }
}

// (3) entries and chunk-column meta (blockIdx.y: 0 = left, 1 = right); threads beyond the slots mark empty columns
__global__ void __launch_bounds__(256)
k_forms_write(int ntasks, int nG, const int2* __restrict__ tasks, const int* __restrict__ gtask_off,
              const unsigned long long* __restrict__ fmA, const unsigned long long* __restrict__ fmB,
              const int* __restrict__ seidxL, const int* __restrict__ seidxR, const int* __restrict__ stoffL,
              const int* __restrict__ stoffR, int4* __restrict__ entL, int4* __restrict__ entR,
              int4* __restrict__ colmetaL, int4* __restrict__ colmetaR, int* __restrict__ coltileL,
              const int* __restrict__ totals) {
  const bool left = blockIdx.y == 0;
  const int* seidx = left ? seidxL : seidxR;
  const int* stoff = left ? stoffL : stoffR;
  int4* ent = left ? entL : entR;
  int* cm = reinterpret_cast<int*>(left ? colmetaL : colmetaR);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = 2 * ntasks;
  if (i < n) {
    unsigned long long m; int id, q, sa, sb;
    form_slot(left, i, tasks, gtask_off, fmA, fmB, m, id, q, sa, sb);
    if (m != 0ull) {
      const int e = seidx[i], e0 = seidx[sa], e1 = seidx[sb];
      ent[e] = make_int4(id, stoff[i], (int)(unsigned)(m & 0xffffffffull), (int)(unsigned)(m >> 32));
      if (e == e0) { cm[4 * q + 0] = e0; cm[4 * q + 1] = e1 - e0; cm[4 * q + 2] = id; }
      if (e + 1 == e1) cm[4 * q + 3] = id;
    }
    return;
  }
  const int q = i - n;                       // one extra thread per chunk column: empty columns
  const int ncc = left ? 2 * nG : nG;
  if (q >= ncc) return;
  const int g = left ? (q >> 1) : q;
  const int t0 = gtask_off[g], nn = gtask_off[g + 1] - t0;
  const int sa = left ? 2 * t0 + (q & 1) * nn : 2 * t0, sb = sa + (left ? nn : 2 * nn);
  if (seidx[sa] == seidx[sb]) reinterpret_cast<int4*>(cm)[q] = make_int4(0, 0, 0, -1);
  if (left) {
    coltileL[q] = (sa < n) ? stoff[sa] : totals[1];
    if (q == ncc - 1) coltileL[ncc] = totals[1];
  }
}

// left form: per inner tile K of the result (4 columns): tile count, first and last row tile
__global__ void __launch_bounds__(256)
k_forms_kmeta(int nk, int nG, const int2* __restrict__ tasks, const int* __restrict__ gtask_off,
              const unsigned long long* __restrict__ fmA, int4* __restrict__ kmeta) {
  const int K = blockIdx.x * blockDim.x + threadIdx.x;
  if (K >= nk) return;
  const int g = K >> 4, c = (K >> 3) & 1, kk = K & 7;
  int cnt = 0, fk = INT_MAX, lk = -1;
  if (g < nG) {
    const int t0 = gtask_off[g], nn = gtask_off[g + 1] - t0;
    for (int b = 0; b < nn; ++b) {
      const unsigned byte = (unsigned)(fmA[(size_t)(t0 + b) * 2 + c] >> (8 * kk)) & 0xffu;
      if (byte) {
        const int base = tasks[t0 + b].y * 8;
        cnt += __popc(byte);
        fk = min(fk, base + __ffs(byte) - 1);
        lk = max(lk, base + 31 - __clz(byte));
      }
    }
  }
  kmeta[K] = (cnt > 0) ? make_int4(0, cnt, fk, lk) : make_int4(0, 0, 0, -1);
}

// one CTA per task, warp w <-> tile column 8g+w: thresholded, scaled (and shifted) values into both tile forms
__global__ void __launch_bounds__(256)
k_forms_fill(int nJ, int nrows, int ncols, const int2* __restrict__ tasks, const int* __restrict__ gtask_off,
             const int* __restrict__ imin8, const int* __restrict__ nI8, const long long* __restrict__ stg_off,
             const double* __restrict__ stg, EmitSpec es, const unsigned long long* __restrict__ fmA,
             const unsigned long long* __restrict__ fmB, const int* __restrict__ stoffL, const int* __restrict__ stoffR,
             double* __restrict__ tvalL, double* __restrict__ tvalR, unsigned want) {
  const int task = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int2 tk = tasks[task];
  const int g = tk.x, I0 = tk.y << 3, J = g * 8 + w;
  if (J >= nJ) return;
  const int iw0 = imin8[J], nI = nI8[J];
  if (I0 < iw0 || I0 >= iw0 + nI) return;
  const int wlen = nI * 8;
  const double* src = stg + stg_off[J] * 64 + (size_t)(I0 - iw0) * 8;
  // right form: tiles (h, jj = w, kk): rows 32h + 4kk + lane%4, column lane/4
  if (want & WANT_RIGHT) {
    const int c = lane >> 2, r = lane & 3, col = J * 8 + c;
    for (int h = 0; h < 2; ++h) {
      const unsigned long long m = fmB[(size_t)task * 2 + h];
      unsigned byte = (unsigned)(m >> (8 * w)) & 0xffu;
      long long t = (long long)stoffR[task * 2 + h] + popc64(m & ((1ull << (8 * w)) - 1ull));
      while (byte) {
        const int kk = __ffs(byte) - 1;
        byte &= byte - 1;
        const int lr = 32 * h + 4 * kk + r, row = I0 * 8 + lr;
        bool keep = false;
        double v = 0.0;
        if (row < nrows && col < ncols) v = final_value<true>(es, src[(size_t)c * wlen + lr], row, col, keep);
        tvalR[t * 32 + lane] = keep ? v : 0.0;
        ++t;
      }
    }
  }
  // left form: chunk column c = w/4, inner tiles kk = 2(w%4) + ch: rows 8ii + lane/4, column 4ch + lane%4
  if (want & WANT_LEFT) {
    const int cch = w >> 2, r = lane >> 2, c4 = lane & 3;
    const int t0 = gtask_off[g], nn = gtask_off[g + 1] - t0;
    const unsigned long long m = fmA[(size_t)task * 2 + cch];
    const int slot = 2 * t0 + cch * nn + (task - t0);
    for (int ch = 0; ch < 2; ++ch) {
      const int kk = 2 * (w & 3) + ch;
      unsigned byte = (unsigned)(m >> (8 * kk)) & 0xffu;
      long long t = (long long)stoffL[slot] + popc64(m & ((1ull << (8 * kk)) - 1ull));
      const int lc = 4 * ch + c4, col = J * 8 + lc;
      while (byte) {
        const int ii = __ffs(byte) - 1;
        byte &= byte - 1;
        const int lr = 8 * ii + r, row = I0 * 8 + lr;
        bool keep = false;
        double v = 0.0;
        if (row < nrows && col < ncols) v = final_value<true>(es, src[(size_t)lc * wlen + lr], row, col, keep);
        tvalL[t * 32 + lane] = keep ? v : 0.0;
        ++t;
      }
    }
  }
}

// Deferred CSC entries of a product, from its right form (the kept entries of a product are exactly its non-zero
// values: |alpha*v| > thr >= 0, or a non-zero shifted diagonal entry). One warp per column; a lane is one row of a
// super-tile (inner tile lane/4, row lane%4), so entries come out in ascending row order.
__global__ void __launch_bounds__(256)
k_right_to_csc(int ncols, CtView R, const int* __restrict__ outer, int* __restrict__ inner, double* __restrict__ val,
               int* __restrict__ mismatch) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int j = gw; j < ncols; j += nw) {
    const int4 cm = R.colmeta[j >> 6];
    const int bit = ((j >> 3) & 7) * 8 + (lane >> 2);
    const int fr = (j & 7) * 4 + (lane & 3);
    int pos = outer[j];
    const int end = outer[j + 1];
    for (int e = 0; e < cm.y; ++e) {
      const int4 en = R.ent[cm.x + e];
      const unsigned long long m = mask64(en);
      double v = 0.0;
      if ((m >> bit) & 1ull) v = R.tval[((size_t)en.y + popc64(m & ((1ull << bit) - 1ull))) * 32 + fr];
      const bool keep = v != 0.0;
      const unsigned b = __ballot_sync(0xffffffffu, keep);
      const int p = pos + __popc(b & ((1u << lane) - 1u));
      if (keep && p < end) { inner[p] = en.x * 32 + lane; val[p] = v; }
      pos += __popc(b);
    }
    if (lane == 0 && pos != end) atomicExch(mismatch, 1);
  }
}
void tile_materialize_entries(const LocalCsc<double>& M) {
  NTB_CHECK(M.forms && M.forms->has_right == 1, "deferred entries without a right tile form");
  const ChunkTiles& R = M.forms->right;
  M.inner.alloc((size_t)M.nnz);
  M.val.alloc((size_t)M.nnz);
  rt().deferred_materialized++;
  if (M.nnz == 0 || M.cols == 0) return;
  const CtView Rv{R.colmeta.get(), R.ent.get(), R.tval.get(), nullptr, R.ncc};
  DevBuf<int> bad(1);
  bad.zero();
  NTB_LAUNCH(k_right_to_csc, max(1, min(div_up((long long)M.cols * 32, 256), kNumSMs * 16)), 256, 0, M.cols, Rv,
             M.outer.get(), M.inner.get(), M.val.get(), bad.get());
  int h = 0;
  d2h(&h, bad.get(), 1);
  NTB_CHECK(h == 0, "deferred entries: the right form does not match the column counts");
}

// Column sums of |alpha*A + B| from the RIGHT tile forms of two blocks of equal shape: the convergence norm of the
// Newton-Schulz style drivers on iterates that live as tile forms (their CSC entries may be deferred). One warp per
// tile column (8 matrix columns): it walks the two id-sorted super-tile lists of its chunk column like a merge; a
// lane is one fragment position (column lane/4, row lane%4 of a 4x8 tile). Absent tiles and dropped entries are
// zeros, so the value equals the CSC kernel's (ops.cu: k_diff_col_abs) up to summation order.
__global__ void __launch_bounds__(256)
k_form_diff_col_abs(CtView A, CtView B, int nJ, int ncols, double alpha, double* __restrict__ colsum) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const int4 none = make_int4(INT_MAX, 0, 0, 0);
  for (int J = gw; J < nJ; J += nw) {
    const int q = J >> 3, sh = (J & 7) * 8;
    const int4 ca = A.colmeta[q], cb = B.colmeta[q];
    int ia = 0, ib = 0;
    int4 ea = (ca.y > 0) ? A.ent[ca.x] : none, eb = (cb.y > 0) ? B.ent[cb.x] : none;
    double s = 0.0;
    while (ea.x != INT_MAX || eb.x != INT_MAX) {
      const int id = min(ea.x, eb.x);
      const bool ta = ea.x == id, tb = eb.x == id;
      const unsigned long long wa = ta ? mask64(ea) : 0ull, wb = tb ? mask64(eb) : 0ull;
      const unsigned ma = (unsigned)(wa >> sh) & 0xffu, mb = (unsigned)(wb >> sh) & 0xffu;
      const size_t ba = (size_t)ea.y + popc64(wa & ((1ull << sh) - 1ull)), bb = (size_t)eb.y + popc64(wb & ((1ull << sh) - 1ull));
      unsigned m = ma | mb;
      while (m) {
        const int kk = __ffs(m) - 1;
        m &= m - 1u;
        const unsigned below = (1u << kk) - 1u;
        double a = 0.0, b = 0.0;
        if ((ma >> kk) & 1u) a = A.tval[(ba + __popc(ma & below)) * 32 + lane];
        if ((mb >> kk) & 1u) b = B.tval[(bb + __popc(mb & below)) * 32 + lane];
        s += fabs(alpha * a + b);
      }
      if (ta) { ++ia; ea = (ia < ca.y) ? A.ent[ca.x + ia] : none; }
      if (tb) { ++ib; eb = (ib < cb.y) ? B.ent[cb.x + ib] : none; }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    const int col = J * 8 + (lane >> 2);
    if ((lane & 3) == 0 && col < ncols) colsum[col] = s;
  }
}
bool tile_diff_col_abs_sums(const LocalCsc<double>& A, const LocalCsc<double>& B, double alpha, double* d_colsum) {
  if (!A.forms || !B.forms || A.forms->has_right != 1 || B.forms->has_right != 1) return false;
  const ChunkTiles& Ra = A.forms->right;
  const ChunkTiles& Rb = B.forms->right;
  if (A.cols != B.cols || A.rows != B.rows || Ra.ncc != Rb.ncc || A.cols == 0) return false;
  const CtView Av{Ra.colmeta.get(), Ra.ent.get(), Ra.tval.get(), nullptr, Ra.ncc};
  const CtView Bv{Rb.colmeta.get(), Rb.ent.get(), Rb.tval.get(), nullptr, Rb.ncc};
  const int nJ = div_up(A.cols, 8);
  NTB_LAUNCH(k_form_diff_col_abs, max(1, min(div_up((long long)nJ * 32, 256), kNumSMs * 16)), 256, 0, Av, Bv, nJ, A.cols,
             alpha, d_colsum);
  return true;
}

// the cached (or freshly built) tile form of an operand; nullptr when its pattern cannot be tiled
const ChunkTiles* tile_operand_form(const LocalCsc<double>& M, bool left) {
  if (!M.forms) M.forms = std::make_shared<TileForms>();
  TileForms& f = *M.forms;
  int& has = left ? f.has_left : f.has_right;
  ChunkTiles& T = left ? f.left : f.right;
  if (has == 0) {
    const bool ok = left ? build_chunk_tiles<true>(M.view(), T) : build_chunk_tiles<false>(M.view(), T);
    has = ok ? 1 : -1;
    rt().tile_builds++;
  }
  return has == 1 ? &T : nullptr;
}

// ---- gathered left form (halo exchange along a process row): rebase one rank's piece ---------------------------
__global__ void __launch_bounds__(256) k_fixup_left(LeftPiece pc, int rank, int4* __restrict__ colmeta, int4* __restrict__ ent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < pc.ncols_chunk) {                     // chunk columns of this rank, global index rank*ncols_chunk + i
    int4 cm = colmeta[(size_t)rank * pc.ncols_chunk + i];
    if (i >= pc.a && i < pc.b && cm.y > 0) cm.x += pc.ent_base;
    else cm = make_int4(0, 0, 0, -1);           // not received: must never be referenced
    colmeta[(size_t)rank * pc.ncols_chunk + i] = cm;
  }
  if (i < pc.nent) {
    int4 e = ent[(size_t)pc.ent_base + i];
    e.y = e.y - pc.tile_lo + pc.recv_base;
    ent[(size_t)pc.ent_base + i] = e;
  }
}
void tile_fixup_gathered_left(ChunkTiles& G, const LeftPiece* pieces, int npieces) {
  for (int p = 0; p < npieces; ++p) {
    const int n = max(pieces[p].ncols_chunk, pieces[p].nent);
    if (n > 0) NTB_LAUNCH(k_fixup_left, div_up(n, 256), 256, 0, pieces[p], p, G.colmeta.get(), G.ent.get());
  }
}

// returns false when the operands are not locally dense enough (caller falls back to the
// scalar window kernels). useful_products = sum over B entries of the A column lengths.
bool spgemm_tile(const LocalCsc<double>& Xl, const LocalCsc<double>& Yl, double alpha, double thr, const RuleView& rules,
                 LocalCsc<double>& Z, double useful_products, const DiagShift* shift, unsigned want) {
  if (Xl.cols == 0 || Xl.nnz == 0 || Yl.nnz == 0 || !(thr >= 0.0)) return false;
  const ChunkTiles* A = tile_operand_form(Yl, true);
  if (!A || (double)Yl.nnz < 0.20 * 32.0 * (double)A->ntiles) return false;   // tiles mostly padding
  const ChunkTiles* B = tile_operand_form(Xl, false);
  if (!B || (double)Xl.nnz < 0.20 * 32.0 * (double)B->ntiles) return false;
  return spgemm_tile_core(*A, *B, Xl.cols, Yl.rows, alpha, thr, rules, Z, useful_products, shift, false, want);
}

bool spgemm_tile_core(const ChunkTiles& Aform, const ChunkTiles& Bform, int ncols, int nrows, double alpha, double thr,
                      const RuleView& rules, LocalCsc<double>& Z, double useful_products, const DiagShift* shift,
                      bool force, unsigned want) {
  // entries can be deferred only when the kept entries are exactly the non-zero values (no dense-rule blocks)
  if (!(want & WANT_CSC)) { if (rules.tbl != nullptr) want = WANT_ALL; else want |= WANT_RIGHT; }
  const ChunkTiles* A = &Aform;
  const ChunkTiles* B = &Bform;
  static const bool timing = std::getenv("NTB_TILE_TIMING") != nullptr;      // developer probe: wall time per phase
  auto now = [&]() { if (timing) stream_sync(); return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = now();
  const int nJ = div_up(ncols, 8), nG = B->ncc;
  const CtView Av{A->colmeta.get(), A->ent.get(), A->tval.get(), A->kmeta.get(), A->ncc};
  const CtView Bv{B->colmeta.get(), B->ent.get(), B->tval.get(), nullptr, B->ncc};
  EmitSpec es;
  es.alpha = alpha; es.thr = thr; es.rules = rules;
  es.sigma = shift ? shift->sigma : 0.0;
  es.dd = shift ? shift->dd : 0;
  es.ncols_diag = shift ? shift->ncols_diag : 0;
  DevBuf<int> imin8((size_t)nJ), nI8((size_t)nJ), gbmin((size_t)nG), gnb((size_t)nG), gtask_off((size_t)nG + 1);
  DevBuf<long long> stg_off((size_t)nJ + 1);
  DevBuf<unsigned long long> ndmma(1);
  ndmma.zero();
  NTB_LAUNCH(k_tile_bounds, max(1, min(div_up((long long)nJ * 32, 256), kNumSMs * 16)), 256, 0, Av, Bv, nJ, imin8.get(),
             nI8.get(), ndmma.get(), es.sigma != 0.0 ? 1 : 0, es.dd, es.ncols_diag, nrows);
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
  if (!force && useful_products >= 0.0 && (double)h_ndmma * 256.0 > 12.0 * useful_products) return false;

  const auto t1 = now();
  DevBuf<double> stg((size_t)h_stg * 64);
  DevBuf<int2> tasks((size_t)max(h_tasks, 1));
  DevBuf<int> cnt((size_t)nJ * 8), task_counter(1);
  DevBuf<unsigned long long> fmA((size_t)max(h_tasks, 1) * 2), fmB((size_t)max(h_tasks, 1) * 2);
  cnt.zero();
  task_counter.zero();
  fmA.zero();
  fmB.zero();
  if (h_tasks > 0)
    NTB_LAUNCH(k_task_table, div_up((long long)nG * 32, 256), 256, 0, nG, gbmin.get(), gtask_off.get(), tasks.get());
  const auto t2 = now();
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (rt().profile) {
    CUDA_CHECK(cudaEventCreate(&ev0));
    CUDA_CHECK(cudaEventCreate(&ev1));
    CUDA_CHECK(cudaEventRecord(ev0, rt().stream));
  }
  if (h_tasks > 0) {
    // pipeline shape: 3 stages x 2 CTAs per SM (default) or 2 stages x 3 CTAs per SM (NTB_NUMERIC_SHAPE=23)
    static const int shape = [] { const char* e = std::getenv("NTB_NUMERIC_SHAPE"); return e ? std::atoi(e) : 32; }();
    auto launch = [&](auto kern, int nstage, int per_sm) {
      static bool attr_set = false;
      if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, numeric_smem(nstage)));
        attr_set = true;
      }
      NTB_LAUNCH(kern, min(h_tasks, kNumSMs * per_sm), NUMERIC_THREADS, numeric_smem(nstage), Av, Bv, nJ, imin8.get(),
                 nI8.get(), stg_off.get(), tasks.get(), h_tasks, task_counter.get(), stg.get(), cnt.get(),
                 reinterpret_cast<unsigned char*>(fmA.get()), reinterpret_cast<unsigned char*>(fmB.get()), nrows, ncols, es);
    };
    static const int ver = [] { const char* e = std::getenv("NTB_NUMERIC_VER"); return e ? std::atoi(e) : 9; }();
    auto launch9 = [&](auto kern, int nstage, int per_sm) {
      static bool attr_set = false;
      if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, numeric_smem9(nstage)));
        attr_set = true;
      }
      NTB_LAUNCH(kern, min(h_tasks, kNumSMs * per_sm), NUMERIC_THREADS, numeric_smem9(nstage), Av, Bv, nJ, imin8.get(),
                 nI8.get(), stg_off.get(), tasks.get(), h_tasks, task_counter.get(), stg.get(), cnt.get(),
                 reinterpret_cast<unsigned char*>(fmA.get()), reinterpret_cast<unsigned char*>(fmB.get()), nrows, ncols, es);
    };
    if (ver == 9 && shape == 23) launch9(k_tile_numeric9<2, 3>, 2, 3);
    else if (ver == 9) launch9(k_tile_numeric9<NSTAGE_DEFAULT, 2>, NSTAGE_DEFAULT, 2);
    else if (shape == 23) launch(k_tile_numeric<2, 3>, 2, 3);
    else launch(k_tile_numeric<NSTAGE_DEFAULT, 2>, NSTAGE_DEFAULT, 2);
  }
  if (rt().profile) {
    CUDA_CHECK(cudaEventRecord(ev1, rt().stream));
    rt().prof_events.emplace_back(ev0, ev1);
  }
  const auto t3 = now();
  // ---- CSC of the result
  const int egrid = max(1, min(div_up((long long)ncols * 32, 256), kNumSMs * 16));
  Z.rows = nrows; Z.cols = ncols;
  Z.outer.alloc((size_t)ncols + 1);
  exclusive_scan(cnt.get(), Z.outer.get(), ncols);
  // ---- index of the result's tile forms (same read-back as the entry count)
  auto forms = std::make_shared<TileForms>();
  ChunkTiles& L = forms->left;
  ChunkTiles& R = forms->right;
  DevBuf<int> seidxL, seidxR, stoffL, stoffR, totals(4);
  const int nk = div_up(ncols, 4);
  if (h_tasks > 0) {
    const size_t ns = (size_t)h_tasks * 2;
    L.ent.alloc(ns); R.ent.alloc(ns);
    L.colmeta.alloc((size_t)nG * 2); R.colmeta.alloc((size_t)nG);
    L.kmeta.alloc((size_t)nk);
    L.coltile.alloc((size_t)nG * 2 + 1);
    seidxL.alloc(ns + 1); seidxR.alloc(ns + 1); stoffL.alloc(ns); stoffR.alloc(ns);
    DevBuf<unsigned long long> smaskL(ns), smaskR(ns);
    NTB_LAUNCH(k_forms_slots, dim3(div_up((long long)ns, 256), 2), 256, 0, h_tasks, tasks.get(), gtask_off.get(), fmA.get(),
               fmB.get(), smaskL.get(), smaskR.get());
    NTB_LAUNCH(k_forms_scan, 2, FI_T, 0, h_tasks, smaskL.get(), smaskR.get(), seidxL.get(), seidxR.get(), stoffL.get(),
               stoffR.get(), totals.get());
    NTB_LAUNCH(k_forms_write, dim3(div_up((long long)ns + 2 * nG, 256), 2), 256, 0, h_tasks, nG, tasks.get(), gtask_off.get(),
               fmA.get(), fmB.get(), seidxL.get(), seidxR.get(), stoffL.get(), stoffR.get(), L.ent.get(), R.ent.get(),
               L.colmeta.get(), R.colmeta.get(), L.coltile.get(), totals.get());
    NTB_LAUNCH(k_forms_kmeta, div_up(nk, 256), 256, 0, nk, nG, tasks.get(), gtask_off.get(), fmA.get(), L.kmeta.get());
  } else {
    totals.zero();
  }
  int h_nnz = 0, h_tot[4] = {0, 0, 0, 0};
  CUDA_CHECK(cudaMemcpyAsync(&h_nnz, Z.outer.get() + ncols, sizeof(int), cudaMemcpyDeviceToHost, rt().stream));
  CUDA_CHECK(cudaMemcpyAsync(h_tot, totals.get(), sizeof(h_tot), cudaMemcpyDeviceToHost, rt().stream));
  stream_sync();
  const auto t4 = now();
  const bool with_forms = h_tasks > 0 && h_nnz > 0;
  const bool defer = with_forms && !(want & WANT_CSC);
  if (defer) {
    Z.alloc_entries(0);
    Z.nnz = h_nnz;                             // inner/val are filled from the right form on first use
  } else {
    Z.alloc_entries(h_nnz);
    if (h_nnz > 0)
      NTB_LAUNCH(k_tile_emit, egrid, 256, 0, ncols, nrows, imin8.get(), nI8.get(), stg_off.get(), stg.get(), es,
                 Z.outer.get(), Z.inner.get(), Z.val.get());
  }
  const auto t5 = now();
  if (with_forms) {
    L.ncc = div_up(ncols, 32); L.nsuper = h_tot[0]; L.ntiles = h_tot[1];
    R.ncc = nG; R.nsuper = h_tot[2]; R.ntiles = h_tot[3];
    const bool wl = (want & WANT_LEFT) != 0, wr = (want & WANT_RIGHT) != 0;
    if (wl) L.tval.alloc((size_t)L.ntiles * 32);
    if (wr) R.tval.alloc((size_t)R.ntiles * 32);
    if (wl || wr)
      NTB_LAUNCH(k_forms_fill, h_tasks, 256, 0, nJ, nrows, ncols, tasks.get(), gtask_off.get(), imin8.get(), nI8.get(),
                 stg_off.get(), stg.get(), es, fmA.get(), fmB.get(), stoffL.get(), stoffR.get(), L.tval.get(), R.tval.get(),
                 want);
    forms->has_left = wl ? 1 : 0;              // a form that was not asked for is rebuilt from CSC if it is ever needed
    forms->has_right = wr ? 1 : 0;
    Z.forms = forms;
    Z.deferred = defer;
  }
  const auto t6 = now();
  if (timing)
    std::fprintf(stderr, "[tile] bounds %.3f  alloc+tasks %.3f  numeric %.3f  index %.3f  emit %.3f  fill %.3f ms  (tasks %d, stg %.0f MB, nnz %d)\n",
                 ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, t5), ms(t5, t6), h_tasks, (double)h_stg * 512.0 / 1e6, h_nnz);
  rt().tile_products++;
  if (Z.deferred) rt().deferred_products++;
  rt().dmma_issued += (double)h_ndmma;
  return true;
}

}  // namespace ntb
