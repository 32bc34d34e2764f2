/* TEST INFRASTRUCTURE: the entry points NTPoly's C++ classes reference but libntpoly_b200.so deliberately does not
 * export (MPI-IO binary format, eigensolver-based Dense* drivers). An executable needs them resolved;
 * calling one is an error. */
#include <stdio.h>
#include <stdlib.h>
#define OUT_OF_SCOPE(name) void name(void) { fprintf(stderr, #name " is outside the path libntpoly_b200 implements\n"); abort(); }
OUT_OF_SCOPE(ConstructMatrixFromBinary_ps_wrp)
OUT_OF_SCOPE(ConstructMatrixFromBinaryPG_ps_wrp)
OUT_OF_SCOPE(WriteMatrixToBinary_ps_wrp)
OUT_OF_SCOPE(DenseDensity_wrp)
OUT_OF_SCOPE(DenseSignFunction_wrp)
OUT_OF_SCOPE(DenseInvert_wrp)
OUT_OF_SCOPE(DenseSquareRoot_wrp)
OUT_OF_SCOPE(DenseInverseSquareRoot_wrp)
OUT_OF_SCOPE(ComputeDenseExponential_wrp)
OUT_OF_SCOPE(ComputeDenseLogarithm_wrp)
OUT_OF_SCOPE(ComputeExponentialPade_wrp)
OUT_OF_SCOPE(ComputeLogarithm_wrp)
