#include "device.cuh"

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
  // keep freed blocks cached in the stream-ordered pool: the per-iteration
  // temporaries of the solvers then never hit cudaMalloc again.
  cudaMemPool_t pool;
  CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
  uint64_t thresh = UINT64_MAX;
  CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  g_rt.inited = true;
}

void set_stream(cudaStream_t s) {
  ensure_init();
  if (g_rt.stream) CUDA_CHECK(cudaStreamSynchronize(g_rt.stream));
  if (g_rt.owns_stream && g_rt.stream) CUDA_CHECK(cudaStreamDestroy(g_rt.stream));
  g_rt.stream = s;
  g_rt.owns_stream = false;
}

void* dmalloc(size_t bytes) {
  ensure_init();
  void* p = nullptr;
  CUDA_CHECK(cudaMallocAsync(&p, bytes ? bytes : 16, g_rt.stream));
  return p;
}
void dfree(void* p) {
  if (p) CUDA_CHECK(cudaFreeAsync(p, g_rt.stream));
}
void stream_sync() { CUDA_CHECK(cudaStreamSynchronize(g_rt.stream)); }

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

template <typename Out> static void scan_impl(const int* in, Out* out, int n) {
  if (n <= 0) {
    CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(Out), rt().stream));
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
