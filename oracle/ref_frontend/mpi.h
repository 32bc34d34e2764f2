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
