/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C + OpenMP) of the local kernels on NTPoly's
 * MatrixMultiply hot path. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * shipped CUDA path never links or calls it.
 *
 * Parity status: the reference (Fortran 2003 + MPI) cannot be compiled in this
 * image (no gfortran / MPI), so this restatement is pinned against the only
 * golden vector the reference ships for the path
 * (Examples/PremadeMatrix/Density-Reference.mtx, see tests/test_oracle_golden.py)
 * and against SciPy the same way the reference's own unit tests are
 * (UnitTests/test_psmatrixalgebra.py, test_matrix.py: Frobenius error <= 1e-4).
 * Threshold>0 / alpha,beta != default / slices>1 behaviour is "parity unpinned"
 * by any reference test: it follows the cited source lines only.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

/* real instantiation */
#define SCALAR double
#define ABSF(x) fabs(x)
#define FN(name) CAT(name, _r)
#include "local_kernels.inc.h"
#undef SCALAR
#undef ABSF
#undef FN

/* complex instantiation (Fortran COMPLEX(NTCOMPLEX) == C double _Complex) */
#define SCALAR double _Complex
#define ABSF(x) cabs(x)
#define FN(name) CAT(name, _c)
#include "local_kernels.inc.h"
#undef SCALAR
#undef ABSF
#undef FN

void orc_free(void *p) { free(p); }

int orc_max_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
