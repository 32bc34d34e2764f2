/* TEST INFRASTRUCTURE: stand-in for <mpi.h> used only to compile NTPoly's own C++ front end and its PremadeMatrix
 * example (from where they lie in the reference checkout) against libntpoly_b200.so, where ranks come from the
 * environment and NCCL replaces MPI. The front end touches MPI only through these names. */
#pragma once
typedef int MPI_Comm;
typedef int MPI_Fint;
#define MPI_COMM_WORLD 0
#define MPI_THREAD_SERIALIZED 2
static inline MPI_Fint MPI_Comm_c2f(MPI_Comm c) { return c; }
static inline int MPI_Init_thread(int *argc, char ***argv, int required, int *provided) {
  (void)argc; (void)argv;
  *provided = required;
  return 0;
}
static inline int MPI_Finalize(void) { return 0; }
/* one process per GPU: rank and size come from the launcher's environment (torchrun, srun --export, mpirun -x) */
#include <stdlib.h>
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { const char *e = getenv("RANK"); (void)c; *rank = e ? atoi(e) : 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *size) { const char *e = getenv("WORLD_SIZE"); (void)c; *size = e ? atoi(e) : 1; return 0; }
