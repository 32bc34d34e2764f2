// In-run measurement of the FP64 tensor-pipe peak (bench.py's roofline denominator for the DMMA tile path): the issue
// rate of DMMA.8x8x4 with eight independent accumulator chains per warp, the only FP64 tensor-core instruction of sm_100a
// (the larger mma.sync f64 shapes compile to sequences of it; profiles/r01_micro_dmma_shapes.txt).
#include "device.cuh"

namespace ntb {
namespace {
__global__ void __launch_bounds__(256) k_dmma_peak(double* out, int iters) {
  const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

// TFLOP/s (2 flops per FMA, 256 FMAs per instruction), best of `repeats` launches timed with CUDA events on the library stream
double measure_dmma_peak_tflops(int repeats) {
  ensure_init();
  const int blocks = kNumSMs * 2, iters = 4000;
  DevBuf<double> out((size_t)blocks * 256);
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  NTB_LAUNCH(k_dmma_peak, blocks, 256, 0, out.get(), iters);
  stream_sync();
  double best = 0.0;
  for (int r = 0; r < repeats; ++r) {
    CUDA_CHECK(cudaEventRecord(e0, rt().stream));
    NTB_LAUNCH(k_dmma_peak, blocks, 256, 0, out.get(), iters);
    CUDA_CHECK(cudaEventRecord(e1, rt().stream));
    stream_sync();
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 256.0 * 8.0 * iters * (double)blocks * 8.0 / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}
}  // namespace ntb
