#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma884(double* c, const double* a, const double* b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
}
__device__ __forceinline__ void mma1684(double* c, const double* a, const double* b){
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b){
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* c, const double* a, const double* b){
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};" : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
template <int SHAPE, int CHAINS>
__global__ void __launch_bounds__(256) k(double* out, int iters) {
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-6 * i;
  double c[CHAINS][4];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (SHAPE == 0) mma884(c[i], a, b);
      if (SHAPE == 1) mma1684(c[i], a, b);
      if (SHAPE == 2) mma1688(c[i], a, b);
      if (SHAPE == 3) mma16816(c[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 8 live DMMAs + 8 DMMAs under a warp-uniform predicate that is false at run time: do they cost pipe time?
__global__ void __launch_bounds__(256) kpred(double* out, int iters, int flag) {
  double a[1] = {threadIdx.x * 1e-3}, b[1] = {1.0 + threadIdx.x * 1e-6};
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mma884(c[i], a, b);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %4, 0;\n\t@p mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n\t}"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]), "r"(flag));
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
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8 * 8);
  const int iters = 4000;
  const double fma[4] = {256, 512, 1024, 2048};
  const char* nm[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
  for (int bps : {1, 2, 4}) {
    int blocks = 148 * bps;
    float ms[4];
    ms[0] = timeit([&] { k<0, 8><<<blocks, 256>>>(out, iters); });
    ms[1] = timeit([&] { k<1, 8><<<blocks, 256>>>(out, iters); });
    ms[2] = timeit([&] { k<2, 8><<<blocks, 256>>>(out, iters); });
    ms[3] = timeit([&] { k<3, 8><<<blocks, 256>>>(out, iters); });
    for (int s = 0; s < 4; ++s)
      printf("%s chains=8 blocks/SM=%d: %.3f ms %.2f TFLOP/s  (%.1f clk/inst/SMSP @1.965GHz)\n", nm[s], bps, ms[s],
             2.0 * fma[s] * 8 * iters * blocks * 8 / ms[s] / 1e9, ms[s] * 1e-3 * 1.965e9 / (8.0 * iters * bps * 2));
  }
  {
    float t0 = timeit([&] { k<0, 8><<<148, 256>>>(out, iters); });
    float t1 = timeit([&] { kpred<<<148, 256>>>(out, iters, 0); });
    float t2 = timeit([&] { kpred<<<148, 256>>>(out, iters, 1); });
    printf("8 DMMA: %.3f ms; 8 DMMA + 8 predicated-off: %.3f ms; 8 + 8 predicated-on: %.3f ms\n", t0, t1, t2);
  }
  float m1 = timeit([&] { k<0, 1><<<148, 128>>>(out, iters); });
  printf("m8n8k4 1 chain 1 warp/SMSP: dependent latency %.1f clk\n", m1 * 1e-3 * 1.965e9 / iters);
  m1 = timeit([&] { k<3, 1><<<148, 128>>>(out, iters); });
  printf("m16n8k16 1 chain 1 warp/SMSP: dependent latency %.1f clk\n", m1 * 1e-3 * 1.965e9 / iters);
  return 0;
}
