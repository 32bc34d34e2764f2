!> ISO_C_BINDING interface to libntpoly_b200.so for Fortran hosts.
!>
!> NOT compiled in this repository's image (no Fortran compiler). It shows the thin shim the
!> north star asks for: NTPoly's PSMatrixAlgebraModule keeps its generic names and argument
!> lists (PSMatrixAlgebraModule.F90:108-123, 414-422, 463-467, 535-539, 362-366, 387-393) and
!> forwards to the device-resident implementation through opaque handles, instead of running
!> the Fortran kernels. A Matrix_ps on the Fortran side then only carries the handle.
MODULE NTPolyB200Shim
  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE
  INTEGER, PARAMETER :: SIZE_wrp = 12
  !> Drop-in stand-in for TYPE(Matrix_ps): the data lives in GPU memory.
  TYPE, PUBLIC :: Matrix_ps
     INTEGER(C_INT) :: ih(SIZE_wrp) = 0
  END TYPE Matrix_ps
  TYPE, PUBLIC :: MatrixMemoryPool_p
     INTEGER(C_INT) :: ih(SIZE_wrp) = 0
  END TYPE MatrixMemoryPool_p
  !> Stand-ins for the local layer (SMatrixModule.F90:15-30, MatrixMemoryPoolModule.F90:13-53): one device CSC block.
  TYPE, PUBLIC :: Matrix_lsr
     INTEGER(C_INT) :: ih(SIZE_wrp) = 0
  END TYPE Matrix_lsr
  TYPE, PUBLIC :: MatrixMemoryPool_lr
     INTEGER(C_INT) :: ih(SIZE_wrp) = 0
  END TYPE MatrixMemoryPool_lr

  INTERFACE
     SUBROUTINE MatrixMultiply_ps_wrp(ih_matA, ih_matB, ih_matC, alpha_in, beta_in, &
          & threshold_in, ih_memory_pool_in) BIND(C, NAME="MatrixMultiply_ps_wrp")
       IMPORT :: C_INT, C_DOUBLE
       INTEGER(C_INT), INTENT(IN) :: ih_matA(*), ih_matB(*)
       INTEGER(C_INT), INTENT(INOUT) :: ih_matC(*), ih_memory_pool_in(*)
       REAL(C_DOUBLE), INTENT(IN) :: alpha_in, beta_in, threshold_in
     END SUBROUTINE MatrixMultiply_ps_wrp
     SUBROUTINE IncrementMatrix_ps_wrp(ih_matA, ih_matB, alpha_in, threshold_in) &
          & BIND(C, NAME="IncrementMatrix_ps_wrp")
       IMPORT :: C_INT, C_DOUBLE
       INTEGER(C_INT), INTENT(IN) :: ih_matA(*)
       INTEGER(C_INT), INTENT(INOUT) :: ih_matB(*)
       REAL(C_DOUBLE), INTENT(IN) :: alpha_in, threshold_in
     END SUBROUTINE IncrementMatrix_ps_wrp
     SUBROUTINE ScaleMatrix_ps_wrp(ih_this, constant) BIND(C, NAME="ScaleMatrix_ps_wrp")
       IMPORT :: C_INT, C_DOUBLE
       INTEGER(C_INT), INTENT(INOUT) :: ih_this(*)
       REAL(C_DOUBLE), INTENT(IN) :: constant
     END SUBROUTINE ScaleMatrix_ps_wrp
     SUBROUTINE MatrixTrace_ps_wrp(ih_this, trace_val) BIND(C, NAME="MatrixTrace_ps_wrp")
       IMPORT :: C_INT, C_DOUBLE
       INTEGER(C_INT), INTENT(IN) :: ih_this(*)
       REAL(C_DOUBLE), INTENT(OUT) :: trace_val
     END SUBROUTINE MatrixTrace_ps_wrp
     FUNCTION MatrixNorm_ps_wrp(ih_this) RESULT(n) BIND(C, NAME="MatrixNorm_ps_wrp")
       IMPORT :: C_INT, C_DOUBLE
       INTEGER(C_INT), INTENT(IN) :: ih_this(*)
       REAL(C_DOUBLE) :: n
     END FUNCTION MatrixNorm_ps_wrp
     SUBROUTINE DotMatrix_psr_wrp(ih_matA, ih_matB, product) BIND(C, NAME="DotMatrix_psr_wrp")
       IMPORT :: C_INT, C_DOUBLE
       INTEGER(C_INT), INTENT(IN) :: ih_matA(*), ih_matB(*)
       REAL(C_DOUBLE), INTENT(OUT) :: product
     END SUBROUTINE DotMatrix_psr_wrp
     SUBROUTINE ConstructMatrixMemoryPool_p_wrp(ih_this, ih_matrix) &
          & BIND(C, NAME="ConstructMatrixMemoryPool_p_wrp")
       IMPORT :: C_INT
       INTEGER(C_INT), INTENT(INOUT) :: ih_this(*)
       INTEGER(C_INT), INTENT(IN) :: ih_matrix(*)
     END SUBROUTINE ConstructMatrixMemoryPool_p_wrp
     SUBROUTINE DestructMatrixMemoryPool_p_wrp(ih_this) BIND(C, NAME="DestructMatrixMemoryPool_p_wrp")
       IMPORT :: C_INT
       INTEGER(C_INT), INTENT(INOUT) :: ih_this(*)
     END SUBROUTINE DestructMatrixMemoryPool_p_wrp
     !> the local kernel (Source/C/SMatrix_c.h:24-28, SMatrixAlgebraModule.F90:221-254)
     SUBROUTINE MatrixMultiply_lsr_wrp(ih_matA, ih_matB, ih_matC, IsATransposed, IsBTransposed, alpha, &
          & beta, threshold, ih_matrix_memory_pool) BIND(C, NAME="MatrixMultiply_lsr_wrp")
       IMPORT :: C_INT, C_DOUBLE, C_BOOL
       INTEGER(C_INT), INTENT(IN) :: ih_matA(*), ih_matB(*)
       INTEGER(C_INT), INTENT(INOUT) :: ih_matC(*), ih_matrix_memory_pool(*)
       LOGICAL(C_BOOL), INTENT(IN) :: IsATransposed, IsBTransposed
       REAL(C_DOUBLE), INTENT(IN) :: alpha, beta, threshold
     END SUBROUTINE MatrixMultiply_lsr_wrp
     SUBROUTINE ConstructMatrixMemoryPool_lr_wrp(ih_this, columns, rows) &
          & BIND(C, NAME="ConstructMatrixMemoryPool_lr_wrp")
       IMPORT :: C_INT
       INTEGER(C_INT), INTENT(INOUT) :: ih_this(*)
       INTEGER(C_INT), INTENT(IN) :: columns, rows
     END SUBROUTINE ConstructMatrixMemoryPool_lr_wrp
     SUBROUTINE DestructMatrixMemoryPool_lr_wrp(ih_this) BIND(C, NAME="DestructMatrixMemoryPool_lr_wrp")
       IMPORT :: C_INT
       INTEGER(C_INT), INTENT(INOUT) :: ih_this(*)
     END SUBROUTINE DestructMatrixMemoryPool_lr_wrp
  END INTERFACE

  INTERFACE MatrixMultiply
     MODULE PROCEDURE MatrixMultiply_b200
     MODULE PROCEDURE GemmMatrix_lsr_b200
  END INTERFACE MatrixMultiply
  INTERFACE IncrementMatrix
     MODULE PROCEDURE IncrementMatrix_b200
  END INTERFACE IncrementMatrix
CONTAINS
  !> Same optional-argument surface as PSMatrixAlgebraModule::MatrixMultiply_ps.
  SUBROUTINE MatrixMultiply_b200(matA, matB, matC, alpha_in, beta_in, threshold_in, memory_pool_in)
    TYPE(Matrix_ps), INTENT(IN) :: matA, matB
    TYPE(Matrix_ps), INTENT(INOUT) :: matC
    REAL(C_DOUBLE), OPTIONAL, INTENT(IN) :: alpha_in, beta_in, threshold_in
    TYPE(MatrixMemoryPool_p), OPTIONAL, INTENT(INOUT) :: memory_pool_in
    REAL(C_DOUBLE) :: alpha, beta, threshold
    TYPE(MatrixMemoryPool_p) :: pool
    alpha = 1.0_C_DOUBLE; beta = 0.0_C_DOUBLE; threshold = 0.0_C_DOUBLE
    IF (PRESENT(alpha_in)) alpha = alpha_in
    IF (PRESENT(beta_in)) beta = beta_in
    IF (PRESENT(threshold_in)) threshold = threshold_in
    IF (PRESENT(memory_pool_in)) THEN
       CALL MatrixMultiply_ps_wrp(matA%ih, matB%ih, matC%ih, alpha, beta, threshold, memory_pool_in%ih)
    ELSE
       CALL ConstructMatrixMemoryPool_p_wrp(pool%ih, matA%ih)
       CALL MatrixMultiply_ps_wrp(matA%ih, matB%ih, matC%ih, alpha, beta, threshold, pool%ih)
       CALL DestructMatrixMemoryPool_p_wrp(pool%ih)
    END IF
  END SUBROUTINE MatrixMultiply_b200
  !> Same optional-argument surface as SMatrixAlgebraModule::GemmMatrix_lsr (:221-254).
  SUBROUTINE GemmMatrix_lsr_b200(matA, matB, matC, IsATransposed_in, IsBTransposed_in, alpha_in, beta_in, &
       & threshold_in, blocked_memory_pool_in)
    TYPE(Matrix_lsr), INTENT(IN) :: matA, matB
    TYPE(Matrix_lsr), INTENT(INOUT) :: matC
    LOGICAL, OPTIONAL, INTENT(IN) :: IsATransposed_in, IsBTransposed_in
    REAL(C_DOUBLE), OPTIONAL, INTENT(IN) :: alpha_in, beta_in, threshold_in
    TYPE(MatrixMemoryPool_lr), OPTIONAL, INTENT(INOUT) :: blocked_memory_pool_in
    LOGICAL(C_BOOL) :: ta, tb
    REAL(C_DOUBLE) :: alpha, beta, threshold
    TYPE(MatrixMemoryPool_lr) :: pool
    ta = .FALSE._C_BOOL; tb = .FALSE._C_BOOL
    alpha = 1.0_C_DOUBLE; beta = 0.0_C_DOUBLE; threshold = 0.0_C_DOUBLE
    IF (PRESENT(IsATransposed_in)) ta = LOGICAL(IsATransposed_in, KIND=C_BOOL)
    IF (PRESENT(IsBTransposed_in)) tb = LOGICAL(IsBTransposed_in, KIND=C_BOOL)
    IF (PRESENT(alpha_in)) alpha = alpha_in
    IF (PRESENT(beta_in)) beta = beta_in
    IF (PRESENT(threshold_in)) threshold = threshold_in
    IF (PRESENT(blocked_memory_pool_in)) THEN
       CALL MatrixMultiply_lsr_wrp(matA%ih, matB%ih, matC%ih, ta, tb, alpha, beta, threshold, &
            & blocked_memory_pool_in%ih)
    ELSE
       CALL ConstructMatrixMemoryPool_lr_wrp(pool%ih, 1_C_INT, 1_C_INT)
       CALL MatrixMultiply_lsr_wrp(matA%ih, matB%ih, matC%ih, ta, tb, alpha, beta, threshold, pool%ih)
       CALL DestructMatrixMemoryPool_lr_wrp(pool%ih)
    END IF
  END SUBROUTINE GemmMatrix_lsr_b200
  SUBROUTINE IncrementMatrix_b200(matA, matB, alpha_in, threshold_in)
    TYPE(Matrix_ps), INTENT(IN) :: matA
    TYPE(Matrix_ps), INTENT(INOUT) :: matB
    REAL(C_DOUBLE), OPTIONAL, INTENT(IN) :: alpha_in, threshold_in
    REAL(C_DOUBLE) :: alpha, threshold
    alpha = 1.0_C_DOUBLE; threshold = 0.0_C_DOUBLE
    IF (PRESENT(alpha_in)) alpha = alpha_in
    IF (PRESENT(threshold_in)) threshold = threshold_in
    CALL IncrementMatrix_ps_wrp(matA%ih, matB%ih, alpha, threshold)
  END SUBROUTINE IncrementMatrix_b200
END MODULE NTPolyB200Shim
