// Microbenchmark: FP64 DMMA.8x8x4 and DFMA issue throughput on B200 (roofline denominators for
// the tile path). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  double c[CHAINS][2];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  double c[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8 * 8);
  const int iters = 20000;
  for (int bps : {1, 2, 4, 8}) {
    int blocks = 148 * bps;
    float ms = timeit([&] { k_dmma<8><<<blocks, 256>>>(out, iters); });
    double fl = 2.0 * 256 * 8 * (double)iters * blocks * 8;  // 8 warps/block, 8 chains
    printf("DMMA.8x8x4 chains=8 blocks/SM=%d: %.2f ms  %.2f TFLOP/s\n", bps, ms, fl / ms / 1e9);
    ms = timeit([&] { k_dmma<2><<<blocks, 256>>>(out, iters); });
    fl = 2.0 * 256 * 2 * (double)iters * blocks * 8;
    printf("DMMA.8x8x4 chains=2 blocks/SM=%d: %.2f ms  %.2f TFLOP/s\n", bps, ms, fl / ms / 1e9);
    ms = timeit([&] { k_dfma<8><<<blocks, 256>>>(out, iters); });
    fl = 2.0 * 8 * (double)iters * blocks * 256;
    printf("DFMA chains=8 blocks/SM=%d: %.2f ms  %.2f TFLOP/s\n", bps, ms, fl / ms / 1e9);
  }
  return 0;
}
