// ntpoly_b200 — common device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace ntb {

// NTPoly has no error returns on its C ABI: failures print and MPI_Abort
// (reference Source/Fortran/ErrorModule.F90:193-205). Same contract here.
[[noreturn]] inline void fatal(const char* what, const char* file, int line) {
  std::fprintf(stderr, "[ntpoly_b200] fatal: %s (%s:%d)\n", what, file, line);
  std::fflush(stderr);
  std::abort();
}
#define NTB_FATAL(msg) ::ntb::fatal((msg), __FILE__, __LINE__)
#define NTB_CHECK(cond, msg) \
  do { if (!(cond)) ::ntb::fatal((msg), __FILE__, __LINE__); } while (0)
#define CUDA_CHECK(expr)                                                        \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      std::fprintf(stderr, "[ntpoly_b200] CUDA error %s: %s\n", #expr,          \
                   cudaGetErrorString(_e));                                     \
      ::ntb::fatal("CUDA call failed", __FILE__, __LINE__);                     \
    }                                                                           \
  } while (0)

// ---- scalar types: NTREAL = double, NTCOMPLEX = complex128 (DataTypesModule.F90:10-18)
struct cplx {
  double x, y;
};
static_assert(sizeof(cplx) == 16, "complex128 layout");

template <typename T> struct scalar_traits;
template <> struct scalar_traits<double> { static constexpr bool is_complex = false; };
template <> struct scalar_traits<cplx>   { static constexpr bool is_complex = true; };

__host__ __device__ __forceinline__ double s_zero(double) { return 0.0; }
__host__ __device__ __forceinline__ cplx s_zero(cplx) { return cplx{0.0, 0.0}; }
template <typename T> __host__ __device__ __forceinline__ T zero_of() { return s_zero(T{}); }

__host__ __device__ __forceinline__ double s_add(double a, double b) { return a + b; }
__host__ __device__ __forceinline__ cplx s_add(cplx a, cplx b) { return cplx{a.x + b.x, a.y + b.y}; }
__host__ __device__ __forceinline__ double s_mul(double a, double b) { return a * b; }
__host__ __device__ __forceinline__ cplx s_mul(cplx a, cplx b) {
  return cplx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
// acc + a*b
__device__ __forceinline__ double s_fma(double a, double b, double acc) { return fma(a, b, acc); }
__device__ __forceinline__ cplx s_fma(cplx a, cplx b, cplx acc) {
  cplx r;
  r.x = fma(a.x, b.x, acc.x);
  r.x = fma(-a.y, b.y, r.x);
  r.y = fma(a.x, b.y, acc.y);
  r.y = fma(a.y, b.x, r.y);
  return r;
}
__host__ __device__ __forceinline__ double s_scale(double alpha, double v) { return alpha * v; }
__host__ __device__ __forceinline__ cplx s_scale(double alpha, cplx v) { return cplx{alpha * v.x, alpha * v.y}; }
__host__ __device__ __forceinline__ double s_abs(double v) { return fabs(v); }
__host__ __device__ __forceinline__ double s_abs(cplx v) { return hypot(v.x, v.y); }
__host__ __device__ __forceinline__ double s_conj(double v) { return v; }
__host__ __device__ __forceinline__ cplx s_conj(cplx v) { return cplx{v.x, -v.y}; }
__host__ __device__ __forceinline__ double s_real(double v) { return v; }
__host__ __device__ __forceinline__ double s_real(cplx v) { return v.x; }

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ cplx shfl(cplx v, int src) {
  return cplx{__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)};
}

// ---- local sparse block: the reference's Matrix_lsr/lsc (SMatrixModule.F90:15-30):
// CSC, outer[cols+1], inner = row ids ascending within a column. 0-based on device.
template <typename T> struct CscView {
  int rows;            // inner dimension
  int cols;            // outer dimension
  const int* outer;    // [cols+1]
  const int* inner;    // [nnz]
  const T* val;        // [nnz]
};

constexpr int kNumSMs = 148;  // B200

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace ntb
