// Do DMMA.8x8x4 (FP64 tensor pipe) and DFMA (FP64 pipe) share issue/execution resources on B200?
// mode 0: all 8 warps per CTA run DMMA chains; mode 1: all run DFMA chains; mode 2: warps 0-3 DMMA, 4-7 DFMA
// (same per-warp work as in modes 0/1). If the pipes are independent, t(mode 2) ~ max(t0, t1)/2 ... see printout.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) k(double* out, int iters, int mode) {
  const int warp = threadIdx.x >> 5;
  const bool dmma = mode == 0 || (mode == 2 && warp < 4);
  const bool dfma = mode == 1 || (mode == 2 && warp >= 4);
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  if (dmma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mma884(c[i][0], c[i][1], a, b);
    }
  }
  if (dfma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {          // 8 x 16 = 128 DFMA per iteration per thread
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i][0] = fma(c[i][0], a, b); c[i][1] = fma(c[i][1], b, a); }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 4 * 256 * 8);
  const int iters = 4000;
  for (int bps : {1, 2}) {
    const int blocks = 148 * bps;
    float t0 = timeit([&] { k<<<blocks, 256>>>(out, iters, 0); });
    float t1 = timeit([&] { k<<<blocks, 256>>>(out, iters, 1); });
    float t2 = timeit([&] { k<<<blocks, 256>>>(out, iters, 2); });
    // mode 0: 8 warps x 8 DMMA x 512 flop; mode 1: 8 warps x 32 lanes x 128 DFMA x 2 flop
    const double f0 = 8.0 * 8 * 512 * iters * blocks, f1 = 8.0 * 32 * 128 * 2 * iters * blocks;
    printf("blocks/SM=%d  DMMA only: %.3f ms %.2f TF/s | DFMA only: %.3f ms %.2f TF/s | half/half: %.3f ms %.2f TF/s "
           "(independent pipes would give %.3f ms, a shared pipe %.3f ms)\n", bps, t0, f0 / t0 / 1e9, t1, f1 / t1 / 1e9, t2,
           (f0 + f1) / 2 / t2 / 1e9, (t0 > t1 ? t0 : t1) / 2, (t0 + t1) / 2);
  }
  return 0;
}
