// Runtime plumbing: the library stream, stream-ordered allocation, launch
// accounting and device-wide scans used by every kernel family.
#pragma once
#include "common.cuh"
#include <functional>
#include <memory>
#include <utility>
#include <vector>

namespace ntb {

constexpr int PEER_MAX = 16;             // ranks of one box whose memory a kernel can address (peer.h)

struct Runtime {
  int device = -1;
  cudaStream_t stream = nullptr;   // all kernels of the hot path are issued here
  bool owns_stream = false;
  bool inited = false;
  unsigned long long launches = 0; // kernels launched by this library (bench: gpu_launches)
  unsigned long long syncs = 0;    // host waits for the library stream (stream_sync)
  double flops_useful = 0.0;       // 2*sum_{(i,k) in A} nnz(B(k,:)) accumulated over multiplies
  unsigned long long multiplies = 0;
  unsigned long long dense_rule_blocks = 0;
  unsigned long long tile_products = 0;  // local products that ran on the DMMA tile path
  unsigned long long complex_tile_products = 0;  // complex local products that ran on the DMMA tile path (real embedding)
  unsigned long long fused_norms = 0;    // difference norms that came out of a product's epilogue (no separate pass)
  unsigned long long hash_columns = 0;   // output columns served by the shared-memory hash accumulator (scattered patterns)
  unsigned long long tile_combines = 0;  // tile-space linear combinations (fused driver steps: no CSC round trip)
  unsigned long long tile_builds = 0;    // CSC -> tile-form conversions (0 per product once operands carry their forms)
  unsigned long long halo_products = 0;  // distributed products that fetched the left operand as a tile halo
  unsigned long long peer_products = 0;  // ... of which read the halo tiles in place from the peers' memory (no copy, no NCCL)
  double halo_bytes = 0.0;               // tile bytes received from the peers for those halos
  double dmma_issued = 0.0;              // DMMA.8x8x4 instructions (x256 FMAs) issued by the tile path
  unsigned long long deferred_products = 0;      // tile products emitted without CSC entries (outer + right form only)
  unsigned long long deferred_materialized = 0; // ... whose entries had to be produced later after all
  unsigned long long sorted_ingests = 0;        // triplet lists taken as they were (own block, column-major): no sort, no gather
  bool count_flops = false;        // instrumentation: count the useful products of every multiply (one extra sweep + read-back)
  double alg_bytes = 0.0;          // compulsory bytes of the local products: bytes(A)+bytes(B)+bytes(C_kept)
  // optional device timing of the dominant (numeric SpGEMM) kernels, for bench.py's roofline
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  // ... and of the phases around them (bench.py: where the rest of a step goes): tag, begin, end
  struct PhaseEv { int tag; cudaEvent_t e0, e1; };
  std::vector<PhaseEv> phase_events;
};
// device time of a phase of the hot path, only while rt().profile is on: PhaseScope p(tag) brackets what is enqueued
// during its lifetime with two events. Tags: 0 symbolic phase of a tile product (bounds ... task count), 2 its tail
// (outer scan, per-K meta, publication), 3 convergence norm / scalars, 4 tile-space combine
struct PhaseScope {
  int idx = -1;
  explicit PhaseScope(int tag);
  ~PhaseScope();
};
Runtime& rt();
void ensure_init();
void set_stream(cudaStream_t s);

void* dmalloc(size_t bytes);
void dfree(void* p);
// waits for the library stream and completes the read-backs enqueued with readback_async()
void stream_sync();
// device -> host copy on the library stream whose result is valid after the next stream_sync(); small sizes bypass the
// copy engines (see device.cu)
void readback_async(void* host, const void* dev, size_t bytes);
// Consecutive readback_async() calls travel in one launch; the batch is launched before ANY later work of the library
// stream (NTB_LAUNCH, memsets, copies, collectives call this), so a read-back still sees what it saw when it was a
// launch of its own.
void readback_flush();
// fn runs inside the next stream_sync(), after the read-backs enqueued so far have landed in their host addresses:
// how a product finishes its bookkeeping (entry count, published descriptors, byte accounting) WITHOUT a wait of its
// own - the values arrive with whatever wait comes next (the following product's task count, a norm)
void on_next_sync(std::function<void()> fn);
void* readback_reserve(void* host, size_t bytes);
size_t arena_bytes_reserved();
// peer-visible slab (device.cu / peer.cu): falls back to dmalloc when there is no slab or it is full
void shared_slab_attach(void* base, size_t bytes, size_t reserved_prefix);
void shared_slab_detach();
void shared_slab_epoch();
bool is_shared_ptr(const void* p);
long long shared_offset(const void* p);      // byte offset inside the slab, -1 for any other pointer
size_t shared_slab_peak();
void* dmalloc_shared(size_t bytes);

// RAII device array on the library stream
template <typename T> struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    p = static_cast<T*>(dmalloc((count ? count : 1) * sizeof(T)));
  }
  // prefers the peer-visible slab (multi-GPU runs); an ordinary arena block otherwise
  void alloc_shared(size_t count) {
    release();
    n = count;
    p = static_cast<T*>(dmalloc_shared((count ? count : 1) * sizeof(T)));
  }
  void release() { if (p) { dfree(p); p = nullptr; n = 0; } }
  void zero() { readback_flush(); CUDA_CHECK(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), rt().stream)); }
  T* get() const { return p; }
};

// An entry count that may still be on its way from the device: a tile product whose entries are deferred enqueues the
// read-back of its count and returns; the value lands with the next stream_sync() of anybody. Reading the count
// (implicit conversion) waits only if it has not landed yet. Copies share the landing cell.
struct PendingCount { int raw = 0; bool arrived = false; };
struct LazyCount {
  mutable long long v = 0;
  mutable std::shared_ptr<PendingCount> pend;
  LazyCount() = default;
  LazyCount(long long x) : v(x) {}
  LazyCount& operator=(long long x) { v = x; pend.reset(); return *this; }
  void settle() const {
    if (!pend) return;
    if (!pend->arrived) stream_sync();
    v = pend->raw;
    pend.reset();
  }
  operator long long() const { settle(); return v; }
  bool pending() const { return pend && !pend->arrived; }
  // false only when the count is known to be zero
  bool maybe_nonzero() const { return pending() ? true : (long long)(*this) > 0; }
  // a reader for later (inside an on_next_sync hook, when the value has landed)
  std::function<long long()> later() const {
    if (!pend) { const long long x = v; return [x] { return x; }; }
    std::shared_ptr<PendingCount> p = pend;
    return [p] { return (long long)p->raw; };
  }
};

template <typename T> inline void d2h(T* host, const T* dev, size_t count) {
  readback_async(host, dev, count * sizeof(T));
  stream_sync();
}
template <typename T> inline void h2d(T* dev, const T* host, size_t count) {
  readback_flush();
  CUDA_CHECK(cudaMemcpyAsync(dev, host, count * sizeof(T), cudaMemcpyHostToDevice, rt().stream));
}
template <typename T> inline void d2d(T* dst, const T* src, size_t count) {
  readback_flush();
  if (count) CUDA_CHECK(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToDevice, rt().stream));
}

#define NTB_LAUNCH(kernel, grid, block, smem, ...)                              \
  do {                                                                          \
    ::ntb::readback_flush();                                                    \
    kernel<<<(grid), (block), (smem), ::ntb::rt().stream>>>(__VA_ARGS__);       \
    ::ntb::rt().launches++;                                                     \
    CUDA_CHECK(cudaGetLastError());                                             \
  } while (0)

// out[0..n] = exclusive prefix sums of in[0..n-1] (out has n+1 entries; out[n] = total).
void exclusive_scan(const int* in, int* out, int n);
void exclusive_scan(const int* in, long long* out, int n);

}  // namespace ntb
