#include "device.cuh"
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <unordered_map>

namespace ntb {

static Runtime g_rt;
Runtime& rt() { return g_rt; }

void ensure_init() {
  if (g_rt.inited) return;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    // No silent CPU fallback: the product path is CUDA only.
    NTB_FATAL("no CUDA device visible: ntpoly_b200 has no CPU fallback");
  }
  int dev = 0;
  if (const char* lr = std::getenv("LOCAL_RANK")) dev = std::atoi(lr) % ndev;
  CUDA_CHECK(cudaSetDevice(dev));
  g_rt.device = dev;
  if (!g_rt.stream) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&g_rt.stream, cudaStreamNonBlocking));
    g_rt.owns_stream = true;
  }
  g_rt.inited = true;
}

PhaseScope::PhaseScope(int tag) {
  if (!g_rt.profile) return;
  Runtime::PhaseEv ev{tag, nullptr, nullptr};
  CUDA_CHECK(cudaEventCreate(&ev.e0));
  CUDA_CHECK(cudaEventCreate(&ev.e1));
  readback_flush();
  CUDA_CHECK(cudaEventRecord(ev.e0, g_rt.stream));
  idx = (int)g_rt.phase_events.size();
  g_rt.phase_events.push_back(ev);
}
PhaseScope::~PhaseScope() {
  if (idx < 0 || idx >= (int)g_rt.phase_events.size()) return;
  readback_flush();
  cudaEventRecord(g_rt.phase_events[(size_t)idx].e1, g_rt.stream);
}

void set_stream(cudaStream_t s) {
  ensure_init();
  if (g_rt.stream) CUDA_CHECK(cudaStreamSynchronize(g_rt.stream));
  CUDA_CHECK(cudaDeviceSynchronize());   // cached arena blocks may still be in use on the old stream
  if (g_rt.owns_stream && g_rt.stream) CUDA_CHECK(cudaStreamDestroy(g_rt.stream));
  g_rt.stream = s;
  g_rt.owns_stream = false;
}

// ---------------------------------------------------------------------------
// Device arena: NTPoly's memory pool idea (scratch that survives across multiplies,
// reference MatrixMemoryPoolModule.F90 / PMatrixMemoryPoolModule.F90) applied to every
// temporary of the hot path. All work is ordered on ONE stream, so a block handed back
// by dfree() can be reused by the next dmalloc() without any synchronisation; after the
// first iterations of a solver no call reaches cudaMalloc any more.
// ---------------------------------------------------------------------------
namespace {
struct Arena {
  std::multimap<size_t, void*> free_blocks;          // capacity -> block
  std::unordered_map<void*, size_t> capacity;        // every live or cached block
  size_t bytes_reserved = 0;
  void release_cached() {
    for (auto& kv : free_blocks) { cudaFree(kv.second); capacity.erase(kv.second); bytes_reserved -= kv.first; }
    free_blocks.clear();
  }
} g_arena;

size_t round_size(size_t bytes) {
  if (bytes < 512) return 512;
  if (bytes < (1u << 20)) return (bytes + 511) & ~size_t(511);
  // large blocks: 1/8-octave bins (at least 2 MiB granules) so that the slightly different sizes of successive
  // iterates map to the same capacity and hit the cache instead of cudaMalloc
  size_t gran = size_t(2) << 20;
  int lg = 63 - __builtin_clzll((unsigned long long)bytes);
  if (lg - 3 > 21) gran = size_t(1) << (lg - 3);
  return (bytes + gran - 1) & ~(gran - 1);
}
}  // namespace


// ---------------------------------------------------------------------------
// Peer-visible slab (multi-GPU, see peer.cu): ONE big allocation per process that every other rank of the box has
// mapped through CUDA IPC. Buffers that peers read in place over NVLink (the left tile forms of distributed products)
// are carved out of it, so a peer addresses them as (its mapping of this slab) + offset. Allocation is host-side
// first fit with coalescing. A freed block is QUARANTINED until the next peer barrier has been enqueued
// (shared_slab_epoch): a peer may still be reading it, and every rank frees at the same point of the (SPMD) program,
// so one barrier later nobody can be.
// ---------------------------------------------------------------------------
namespace {
struct SharedSlab {
  unsigned char* base = nullptr;
  size_t bytes = 0;
  std::map<size_t, size_t> avail;                    // offset -> size, coalesced, reusable now
  std::unordered_map<size_t, size_t> live;           // offset -> size
  struct Q { size_t off, size; unsigned long long epoch; };
  std::deque<Q> quarantine;
  unsigned long long epoch = 0;
  size_t in_use = 0, peak = 0;
  void insert_free(size_t off, size_t size) {
    auto nx = avail.lower_bound(off);
    if (nx != avail.begin()) {
      auto pv = std::prev(nx);
      if (pv->first + pv->second == off) { off = pv->first; size += pv->second; avail.erase(pv); }
    }
    if (nx != avail.end() && off + size == nx->first) { size += nx->second; avail.erase(nx); }
    avail[off] = size;
  }
} g_slab;
}  // namespace

void shared_slab_attach(void* base, size_t bytes, size_t reserved_prefix) {
  g_slab = SharedSlab();
  g_slab.base = static_cast<unsigned char*>(base);
  g_slab.bytes = bytes;
  if (base && bytes > reserved_prefix) g_slab.avail[reserved_prefix] = bytes - reserved_prefix;
}
void shared_slab_detach() { g_slab = SharedSlab(); }
void shared_slab_epoch() { ++g_slab.epoch; }
bool is_shared_ptr(const void* p) {
  const unsigned char* q = static_cast<const unsigned char*>(p);
  return g_slab.base && q >= g_slab.base && q < g_slab.base + g_slab.bytes;
}
long long shared_offset(const void* p) {
  return is_shared_ptr(p) ? (long long)(static_cast<const unsigned char*>(p) - g_slab.base) : -1;
}
size_t shared_slab_peak() { return g_slab.peak; }
void* dmalloc_shared(size_t bytes) {
  if (!g_slab.base) return dmalloc(bytes);
  while (!g_slab.quarantine.empty() && g_slab.quarantine.front().epoch < g_slab.epoch) {
    g_slab.insert_free(g_slab.quarantine.front().off, g_slab.quarantine.front().size);
    g_slab.quarantine.pop_front();
  }
  const size_t need = ((bytes ? bytes : 1) + 255) & ~size_t(255);
  auto best = g_slab.avail.end();
  for (auto it = g_slab.avail.begin(); it != g_slab.avail.end(); ++it)
    if (it->second >= need && (best == g_slab.avail.end() || it->second < best->second)) best = it;
  if (best == g_slab.avail.end()) return dmalloc(bytes);         // slab full: an ordinary (not peer-visible) block
  const size_t off = best->first, size = best->second;
  g_slab.avail.erase(best);
  if (size > need) g_slab.avail[off + need] = size - need;
  g_slab.live[off] = need;
  g_slab.in_use += need;
  if (g_slab.in_use > g_slab.peak) g_slab.peak = g_slab.in_use;
  return g_slab.base + off;
}

void* dmalloc(size_t bytes) {
  ensure_init();
  const size_t need = round_size(bytes);
  auto it = g_arena.free_blocks.lower_bound(need);
  // accept a cached block unless it would waste more than half of itself (keeps the big
  // staging buffers from being burnt on small requests)
  if (it != g_arena.free_blocks.end() && (it->first <= 2 * need || it->first - need <= (size_t(4) << 20))) {
    void* p = it->second;
    g_arena.free_blocks.erase(it);
    return p;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, need);
  if (e != cudaSuccess) {
    cudaGetLastError();
    CUDA_CHECK(cudaStreamSynchronize(g_rt.stream));
    g_arena.release_cached();
    CUDA_CHECK(cudaMalloc(&p, need));
  }
  g_arena.capacity[p] = need;
  g_arena.bytes_reserved += need;
  return p;
}
void dfree(void* p) {
  if (!p) return;
  if (is_shared_ptr(p)) {
    const size_t off = (size_t)(static_cast<unsigned char*>(p) - g_slab.base);
    auto it = g_slab.live.find(off);
    NTB_CHECK(it != g_slab.live.end(), "dfree of a slab pointer that is not live");
    g_slab.quarantine.push_back({off, it->second, g_slab.epoch});
    g_slab.in_use -= it->second;
    g_slab.live.erase(it);
    return;
  }
  auto it = g_arena.capacity.find(p);
  NTB_CHECK(it != g_arena.capacity.end(), "dfree of a pointer the arena does not own");
  g_arena.free_blocks.emplace(it->second, p);
}
size_t arena_bytes_reserved() { return g_arena.bytes_reserved; }

// ---------------------------------------------------------------------------
// Small read-backs (task counts, nnz, norms, flags: about ten per solver step) do not use a copy engine: a copy
// engine serves its queue in order, so a 4-byte cudaMemcpyAsync of the library stream waits behind a whole bulk
// device-to-host transfer that another stream has queued (the asynchronous egress of the previous result) - measured:
// the step of a pipelined host loop could not start before the previous result had left completely. Instead one
// thread block stores the bytes straight into mapped pinned host memory; stream_sync() hands them to the caller.
// ---------------------------------------------------------------------------
namespace {
constexpr size_t RB_SCRATCH = 64 << 10, RB_MAX = 4096;
struct PendingReadback { void* host; size_t off, bytes; };
unsigned char* g_rb_host = nullptr;                  // mapped pinned scratch (host address)
unsigned char* g_rb_dev = nullptr;                   // the same memory as seen from the device
size_t g_rb_used = 0;
std::vector<PendingReadback> g_rb_pending;

// consecutive read-backs (no kernel of the library launched in between) travel in ONE launch: up to RB_SEG segments
constexpr int RB_SEG = 8;
struct ReadbackBatch { const unsigned char* src[RB_SEG]; unsigned char* dst[RB_SEG]; unsigned bytes[RB_SEG]; int n; };
__global__ void __launch_bounds__(128) k_readback(ReadbackBatch b) {
  for (int s = 0; s < b.n; ++s) {
    const unsigned char* src = b.src[s];
    unsigned char* dst = b.dst[s];
    const unsigned bytes = b.bytes[s];
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | bytes) & 3u) == 0) {
      const unsigned n = bytes >> 2;
      for (unsigned i = threadIdx.x; i < n; i += blockDim.x)
        reinterpret_cast<unsigned*>(dst)[i] = reinterpret_cast<const unsigned*>(src)[i];
    } else {
      for (unsigned i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
    }
  }
  __threadfence_system();
}
ReadbackBatch g_rb_batch{};
unsigned long long g_rb_batch_launches = 0;       // rt().launches when the open batch got its last segment
}  // namespace
void readback_flush() {
  if (g_rb_batch.n == 0) return;
  k_readback<<<1, 128, 0, g_rt.stream>>>(g_rb_batch);
  CUDA_CHECK(cudaGetLastError());
  g_rt.launches++;
  g_rb_batch.n = 0;
}

void readback_async(void* host, const void* dev, size_t bytes) {
  if (bytes == 0) return;
  const size_t padded = (bytes + 15) & ~size_t(15);
  if (bytes > RB_MAX || g_rb_used + padded > RB_SCRATCH) {          // bulk, or scratch exhausted: the copy engine
    readback_flush();
    CUDA_CHECK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, g_rt.stream));
    return;
  }
  if (!g_rb_host) {
    CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&g_rb_host), RB_SCRATCH, cudaHostAllocMapped));
    CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_rb_dev), g_rb_host, 0));
  }
  // a segment may join the open batch only if nothing was launched since (the batch's copy runs where its LAST
  // segment was enqueued; a kernel in between could have changed an earlier segment's source)
  if (g_rb_batch.n == RB_SEG || (g_rb_batch.n > 0 && g_rb_batch_launches != g_rt.launches)) readback_flush();
  g_rb_batch.src[g_rb_batch.n] = static_cast<const unsigned char*>(dev);
  g_rb_batch.dst[g_rb_batch.n] = g_rb_dev + g_rb_used;
  g_rb_batch.bytes[g_rb_batch.n] = (unsigned)bytes;
  g_rb_batch.n++;
  g_rb_batch_launches = g_rt.launches;
  g_rb_pending.push_back(PendingReadback{host, g_rb_used, bytes});
  g_rb_used += padded;
}
// A kernel of the caller stores `bytes` at the returned device address (mapped pinned host memory); they are handed to
// `host` by the next stream_sync(), like readback_async().
void* readback_reserve(void* host, size_t bytes) {
  const size_t padded = (bytes + 15) & ~size_t(15);
  NTB_CHECK(padded <= RB_SCRATCH, "readback_reserve: request exceeds the scratch");
  if (g_rb_used + padded > RB_SCRATCH) stream_sync();
  if (!g_rb_host) {
    CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&g_rb_host), RB_SCRATCH, cudaHostAllocMapped));
    CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_rb_dev), g_rb_host, 0));
  }
  void* d = g_rb_dev + g_rb_used;
  g_rb_pending.push_back(PendingReadback{host, g_rb_used, bytes});
  g_rb_used += padded;
  return d;
}
namespace {
std::vector<std::function<void()>> g_sync_hooks;
}
void on_next_sync(std::function<void()> fn) { g_sync_hooks.push_back(std::move(fn)); }
void stream_sync() {
  readback_flush();
  CUDA_CHECK(cudaStreamSynchronize(g_rt.stream));
  if (!g_rb_pending.empty()) {
    for (const PendingReadback& r : g_rb_pending) std::memcpy(r.host, g_rb_host + r.off, r.bytes);
    g_rb_pending.clear();
  }
  g_rb_used = 0;
  g_rt.syncs++;
  if (!g_sync_hooks.empty()) {                   // the read-backs they wait for have landed
    std::vector<std::function<void()>> hooks;
    hooks.swap(g_sync_hooks);
    for (auto& h : hooks) h();
  }
}

// ---------------------------------------------------------------------------
// device-wide exclusive scan: block partials -> single-block top scan -> final
// ---------------------------------------------------------------------------
constexpr int SCAN_T = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_T * SCAN_ITEMS;

template <typename Out>
__device__ __forceinline__ Out block_exclusive_scan(Out v, Out* smem_warp, Out& block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Out inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Out o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) smem_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    Out w = (lane < (SCAN_T / 32)) ? smem_warp[lane] : Out(0);
    Out winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      Out o = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += o;
    }
    if (lane < (SCAN_T / 32)) smem_warp[lane] = winc - w;  // exclusive warp offsets
    if (lane == (SCAN_T / 32) - 1) smem_warp[SCAN_T / 32] = winc;
  }
  __syncthreads();
  block_total = smem_warp[SCAN_T / 32];
  Out r = smem_warp[warp] + inc - v;
  __syncthreads();
  return r;
}

template <typename Out>
__global__ void __launch_bounds__(SCAN_T) k_scan_partials(const int* __restrict__ in, int n, Out* __restrict__ partial) {
  __shared__ Out sw[SCAN_T / 32 + 1];
  const long long base = (long long)blockIdx.x * SCAN_TILE;
  Out s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    long long i = base + (long long)k * SCAN_T + threadIdx.x;
    if (i < n) s += (Out)in[i];
  }
  Out total;
  (void)block_exclusive_scan<Out>(s, sw, total);
  if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

template <typename Out>
__global__ void __launch_bounds__(SCAN_T) k_scan_top(Out* __restrict__ partial, int nblocks) {
  __shared__ Out sw[SCAN_T / 32 + 1];
  Out carry = 0;
  for (int base = 0; base < nblocks; base += SCAN_T) {
    int i = base + threadIdx.x;
    Out v = (i < nblocks) ? partial[i] : Out(0);
    Out total;
    Out ex = block_exclusive_scan<Out>(v, sw, total);
    if (i < nblocks) partial[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) partial[nblocks] = carry;
}

template <typename Out>
__global__ void __launch_bounds__(SCAN_T) k_scan_final(const int* __restrict__ in, Out* __restrict__ out, int n,
                                                       const Out* __restrict__ partial, int nblocks) {
  __shared__ Out sw[SCAN_T / 32 + 1];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  Out v[SCAN_ITEMS];
  Out s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    long long i = base + k;
    v[k] = (i < n) ? (Out)in[i] : Out(0);
    s += v[k];
  }
  Out total;
  Out ex = block_exclusive_scan<Out>(s, sw, total) + partial[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    long long i = base + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = partial[nblocks];
}

// small inputs: one block does the whole scan (one launch instead of three; the scans of the hot path are over a few
// thousand groups or a rank's columns, and at that size the launches cost more than the work)
constexpr int SCAN1_T = 1024;
constexpr int SCAN1_MAX = 1 << 13;     // beyond that the per-thread chunks are long and their loads uncoalesced: three launches win
template <typename Out>
__global__ void __launch_bounds__(SCAN1_T) k_scan_single(const int* __restrict__ in, Out* __restrict__ out, int n) {
  __shared__ Out sw[33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (n + SCAN1_T - 1) / SCAN1_T;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  Out s = 0;
  for (int i = lo; i < hi; ++i) s += (Out)in[i];
  Out inc = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const Out o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) sw[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const Out w = sw[lane];
    Out winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const Out o = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += o;
    }
    sw[lane] = winc - w;
    if (lane == 31) sw[32] = winc;
  }
  __syncthreads();
  Out ex = sw[warp] + inc - s;
  for (int i = lo; i < hi; ++i) { out[i] = ex; ex += (Out)in[i]; }
  if (threadIdx.x == 0) out[n] = sw[32];
}

template <typename Out> static void scan_impl(const int* in, Out* out, int n) {
  if (n <= 0) {
    readback_flush();
    CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(Out), rt().stream));
    return;
  }
  if (n <= SCAN1_MAX && static_cast<const void*>(in) != static_cast<const void*>(out)) {
    NTB_LAUNCH((k_scan_single<Out>), 1, SCAN1_T, 0, in, out, n);
    return;
  }
  int nblocks = div_up(n, SCAN_TILE);
  DevBuf<Out> partial((size_t)nblocks + 1);
  NTB_LAUNCH((k_scan_partials<Out>), nblocks, SCAN_T, 0, in, n, partial.get());
  NTB_LAUNCH((k_scan_top<Out>), 1, SCAN_T, 0, partial.get(), nblocks);
  NTB_LAUNCH((k_scan_final<Out>), nblocks, SCAN_T, 0, in, out, n, partial.get(), nblocks);
}

void exclusive_scan(const int* in, int* out, int n) { scan_impl<int>(in, out, n); }
void exclusive_scan(const int* in, long long* out, int n) { scan_impl<long long>(in, out, n); }

}  // namespace ntb
