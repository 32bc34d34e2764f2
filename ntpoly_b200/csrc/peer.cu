#include "peer.h"
#include <cstring>
#include <vector>

namespace ntb {

namespace {
PeerSpace g_peer;
void* g_slab = nullptr;                          // this process's slab (lives until process exit)
size_t g_slab_bytes = 0;
std::vector<void*> g_opened;                     // IPC mappings of the other ranks' slabs

// control block at the start of every slab
constexpr size_t FLAG_OFF = 0;                   // u64 flags[PEER_MAX]: flags[p] = last epoch rank p has reached
constexpr size_t INBOX_OFF = 1024;               // PeerPayload inbox[RING][PEER_MAX]
constexpr int RING = 4;
constexpr size_t CTRL_BYTES = 64 << 10;
static_assert(INBOX_OFF + sizeof(PeerPayload) * RING * PEER_MAX <= CTRL_BYTES, "control block");

struct PeerCtl { unsigned char* base[PEER_MAX]; int n, me; };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// One block. (1) every rank stores its payload into slot (epoch % RING) of every rank's inbox - peer stores over
// NVLink -, (2) fences and raises its flag at every rank to `epoch`, (3) waits until all flags of its own slab have
// reached `epoch`, (4) hands the received payloads to the host through mapped memory. A rank can be at most one
// exchange ahead of any other, so a ring of 4 slots is never overwritten while it is still being read.
__global__ void __launch_bounds__(128) k_peer_exchange(PeerCtl c, PeerPayload pl, const unsigned long long* dev8, int n8,
                                                       const int* dev_i32_w7, unsigned long long epoch, int slot,
                                                       unsigned long long* host_out) {
  __shared__ PeerPayload sp;
  const int t = threadIdx.x;
  if (t == 0) {
    sp = pl;
    for (int i = 0; i < n8; ++i) sp.w[i] = dev8[i];
    if (dev_i32_w7) sp.w[7] = (unsigned long long)(long long)*dev_i32_w7;
  }
  __syncthreads();
  if (t < c.n * 8) {
    const int p = t >> 3, w = t & 7;
    unsigned long long* dst =
        reinterpret_cast<unsigned long long*>(c.base[p] + INBOX_OFF + ((size_t)slot * PEER_MAX + c.me) * sizeof(PeerPayload)) + w;
    *reinterpret_cast<volatile unsigned long long*>(dst) = sp.w[w];
  }
  __threadfence_system();
  __syncthreads();
  if (t < c.n) {
    st_release_sys(reinterpret_cast<unsigned long long*>(c.base[t] + FLAG_OFF) + c.me, epoch);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(c.base[c.me] + FLAG_OFF) + t;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(mine) < epoch) {
      if (global_ns() - t0 > 120ull * 1000000000ull) {        // a peer died or left the SPMD sequence: do not hang the box
        printf("[ntpoly_b200] peer_exchange: rank %d gave up waiting for rank %d at epoch %llu\n", c.me, t, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  if (host_out && t < c.n * 8) {
    const unsigned long long* src =
        reinterpret_cast<const unsigned long long*>(c.base[c.me] + INBOX_OFF + ((size_t)slot * PEER_MAX + (t >> 3)) * sizeof(PeerPayload)) + (t & 7);
    host_out[t] = *reinterpret_cast<const volatile unsigned long long*>(src);
  }
  __threadfence_system();
}
}  // namespace

PeerSpace& peer() { return g_peer; }

void peer_teardown() {
  if (g_peer.ok) shared_slab_detach();
  g_peer = PeerSpace();
}

bool peer_setup(CommHandle* wc) {
  const int n = comm_size(wc), me = comm_rank(wc);
  if (g_peer.ok && g_peer.n == n && g_peer.me == me) return true;
  g_peer = PeerSpace();
  if (n <= 1 || n > PEER_MAX) return false;
  if (const char* e = std::getenv("NTB_P2P")) if (e[0] == '0') return false;   // same environment on every rank
  ensure_init();
  // ---- own slab (kept for the life of the process) and its IPC handle
  double ok_local = 1.0;
  if (!g_slab) {
    size_t free_b = 0, total_b = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    size_t want = std::min<size_t>(size_t(32) << 30, total_b / 5);
    if (const char* e = std::getenv("NTB_PEER_SLAB_MB")) want = (size_t)std::atoll(e) << 20;
    want = std::min(want, free_b / 2) & ~size_t((2 << 20) - 1);
    if (want < (CTRL_BYTES << 4) || cudaMalloc(&g_slab, want) != cudaSuccess) { cudaGetLastError(); g_slab = nullptr; ok_local = 0.0; }
    else g_slab_bytes = want;
  }
  struct Hello { cudaIpcMemHandle_t handle; unsigned long long bytes; int device; int ok; };
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  Hello mine{};
  mine.device = rt().device;
  mine.bytes = g_slab_bytes;
  if (g_slab) {
    readback_flush();
    CUDA_CHECK(cudaMemsetAsync(g_slab, 0, CTRL_BYTES, rt().stream));            // flags and inbox start at zero
    if (cudaIpcGetMemHandle(&mine.handle, g_slab) != cudaSuccess) { cudaGetLastError(); ok_local = 0.0; }
  }
  mine.ok = ok_local != 0.0 ? 1 : 0;
  std::vector<Hello> all((size_t)n);
  {
    DevBuf<Hello> d_mine(1), d_all((size_t)n);
    h2d(d_mine.get(), &mine, 1);
    comm_allgather_bytes(wc, d_mine.get(), d_all.get(), sizeof(Hello));
    CUDA_CHECK(cudaMemcpyAsync(all.data(), d_all.get(), sizeof(Hello) * n, cudaMemcpyDeviceToHost, rt().stream));
    stream_sync();
  }
  // ---- map the other slabs
  unsigned char* base[PEER_MAX] = {};
  for (int p = 0; p < n && ok_local != 0.0; ++p) {
    if (!all[p].ok) { ok_local = 0.0; break; }
    if (p == me) { base[p] = static_cast<unsigned char*>(g_slab); continue; }
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, rt().device, all[p].device) != cudaSuccess || !can) { cudaGetLastError(); ok_local = 0.0; break; }
    void* q = nullptr;
    if (cudaIpcOpenMemHandle(&q, all[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok_local = 0.0; break; }
    g_opened.push_back(q);
    base[p] = static_cast<unsigned char*>(q);
  }
  // ---- everybody or nobody (this all-reduce is also the barrier behind the memsets of the control blocks)
  {
    DevBuf<double> d(1);
    h2d(d.get(), &ok_local, 1);
    comm_allreduce_f64(wc, d.get(), 1, RedOp::Min);
    CUDA_CHECK(cudaMemcpyAsync(&ok_local, d.get(), sizeof(double), cudaMemcpyDeviceToHost, rt().stream));
    stream_sync();
  }
  if (ok_local == 0.0) {
    for (void* q : g_opened) cudaIpcCloseMemHandle(q);
    g_opened.clear();
    return false;
  }
  g_peer.ok = true; g_peer.n = n; g_peer.me = me; g_peer.slab_bytes = g_slab_bytes;
  for (int p = 0; p < n; ++p) g_peer.base[p] = base[p];
  shared_slab_attach(g_slab, g_slab_bytes, CTRL_BYTES);
  return true;
}

void peer_exchange(const PeerPayload& mine, PeerPayload* out, const void* dev8, int n8, const int* dev_i32_w7) {
  NTB_CHECK(g_peer.ok, "peer_exchange without a peer space");
  PeerCtl c{};
  for (int p = 0; p < g_peer.n; ++p) c.base[p] = g_peer.base[p];
  c.n = g_peer.n; c.me = g_peer.me;
  const unsigned long long epoch = ++g_peer.epoch;
  unsigned long long* host_out =
      out ? static_cast<unsigned long long*>(readback_reserve(out, sizeof(PeerPayload) * g_peer.n)) : nullptr;
  NTB_CHECK(n8 >= 0 && n8 <= 8, "peer_exchange: at most 8 words");
  readback_flush();
  k_peer_exchange<<<1, 128, 0, rt().stream>>>(c, mine, static_cast<const unsigned long long*>(dev8), dev8 ? n8 : 0,
                                              dev_i32_w7, epoch, (int)(epoch % RING), host_out);
  CUDA_CHECK(cudaGetLastError());
  rt().launches++;
  g_peer.exchanges++;
  shared_slab_epoch();                           // blocks freed before this point are reusable by work enqueued after it
}

}  // namespace ntb
