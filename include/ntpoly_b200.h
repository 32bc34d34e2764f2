/* ntpoly_b200 — C ABI of the B200-native NTPoly hot path (libntpoly_b200.so).
 *
 * Drop-in boundary: the symbols in sections 1-7 are exactly the `*_wrp` entry
 * points NTPoly's own C ABI exposes for this path (reference headers under
 * /root/reference/Source/C/, Fortran shims under Source/Wrapper/): same names,
 * same argument order, everything by pointer, opaque `int ih[SIZE_wrp]` handles
 * (SIZE_wrp = 12, Source/C/Wrapper.h:4). NTPoly's C++ classes (Source/CPlusPlus)
 * and its SWIG Python module bind these symbols unchanged.
 *
 * Behavioural contract kept from the reference: no error returns (a failure
 * prints and aborts, Source/Fortran/ErrorModule.F90:193-205); every `_ps_` call
 * is collective over the matrix's process grid; handles own heap objects that
 * Construct* allocates and Destruct* frees.
 *
 * What differs: matrices live in GPU memory (one CSC block per GPU) between
 * calls; `world_comm` arguments (Fortran MPI communicator integers) are accepted
 * and ignored — the rank layout comes from ntb_world_init() or the RANK /
 * WORLD_SIZE / LOCAL_RANK environment (one process per GPU, NCCL underneath).
 *
 * Section 8 lists the `ntb_` extensions (bootstrap, bulk triplet transfer,
 * counters) that have no counterpart in the reference.
 */
#ifndef NTPOLY_B200_H
#define NTPOLY_B200_H

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTB_SIZE_wrp 12

/* ---- 1. process grid (Source/C/ProcessGrid_c.h:4-39; ProcessGridModule_wrp.F90) */
void ConstructGlobalProcessGrid_wrp(const int *world_comm, const int *process_rows,
                                    const int *process_columns, const int *process_slices);
void ConstructGlobalProcessGrid_onlyslice_wrp(const int *world_comm, const int *process_slices);
void ConstructGlobalProcessGrid_default_wrp(const int *world_comm);
void CopyProcessGrid_wrp(const int *ih_old_grid, int *ih_new_grid);
int GetGlobalMySlice_wrp(void);
int GetGlobalMyColumn_wrp(void);
int GetGlobalMyRow_wrp(void);
bool GetGlobalIsRoot_wrp(void);
int GetGlobalNumSlices_wrp(void);
int GetGlobalNumColumns_wrp(void);
int GetGlobalNumRows_wrp(void);
void WriteGlobalProcessGridInfo_wrp(void);
void DestructGlobalProcessGrid_wrp(void);
void ConstructProcessGrid_wrp(int *ih_grid, const int *world_comm, const int *process_rows,
                              const int *process_columns, const int *process_slices);
void ConstructProcessGrid_onlyslice_wrp(int *ih_grid, const int *world_comm, const int *process_slices);
void ConstructProcessGrid_default_wrp(int *ih_grid, const int *world_comm);
int GetMySlice_wrp(const int *ih_grid);
int GetMyColumn_wrp(const int *ih_grid);
int GetMyRow_wrp(const int *ih_grid);
int GetNumSlices_wrp(const int *ih_grid);
int GetNumColumns_wrp(const int *ih_grid);
int GetNumRows_wrp(const int *ih_grid);
void WriteProcessGridInfo_wrp(const int *ih_grid);
void DestructProcessGrid_wrp(int *ih_grid);

/* ---- 2. triplet lists (Source/C/TripletList_c.h:4-38) */
void ConstructTripletList_r_wrp(int *ih_this, const int *size);
void ResizeTripletList_r_wrp(int *ih_this, const int *size);
void AppendToTripletList_r_wrp(int *ih_this, const int *index_column, const int *index_row,
                               const double *point_value);
void SetTripletAt_r_wrp(int *ih_this, const int *index, const int *index_column, const int *index_row,
                        const double *point_value);
void GetTripletAt_r_wrp(const int *ih_this, const int *index, int *index_column, int *index_row,
                        double *point_value);
void DestructTripletList_r_wrp(int *ih_this);
int GetTripletListSize_r_wrp(const int *ih_this);
void ConstructTripletList_c_wrp(int *ih_this, const int *size);
void ResizeTripletList_c_wrp(int *ih_this, const int *size);
void AppendToTripletList_c_wrp(int *ih_this, const int *index_column, const int *index_row,
                               const double *point_value_real, const double *point_value_imag);
void SetTripletAt_c_wrp(int *ih_this, const int *index, const int *index_column, const int *index_row,
                        const double *point_value_real, const double *point_value_imag);
void GetTripletAt_c_wrp(const int *ih_this, const int *index, int *index_column, int *index_row,
                        double *point_value_real, double *point_value_imag);
void DestructTripletList_c_wrp(int *ih_this);
int GetTripletListSize_c_wrp(const int *ih_this);
/* sorted copy (by column, then row) in a freshly allocated list whose handle is written to h_sorted. Three arguments
 * like the reference's C header and C++ caller (TripletList_c.h:15-16, TripletList.cc:88-96); the Fortran shim behind
 * them takes four (TripletListModule_wrp.F90:135-151), a mismatch inside the reference. */
void SortTripletList_r_wrp(const int *ih_this, const int *matrix_size, int *h_sorted);
void SortTripletList_c_wrp(const int *ih_this, const int *matrix_size, int *h_sorted);

/* ---- 3. distributed matrix container (Source/C/PSMatrix_c.h:4-49; PSMatrixModule_wrp.F90) */
void ConstructEmptyMatrix_ps_wrp(int *ih_this, const int *matrix_dim);
void ConstructEmptyMatrixPG_ps_wrp(int *ih_this, const int *matrix_dim, const int *ih_grid);
void CopyMatrix_ps_wrp(const int *ih_matA, int *ih_matB);
void DestructMatrix_ps_wrp(int *ih_this);
void ConstructMatrixFromMatrixMarket_ps_wrp(int *ih_this, const char *file_name, const int *name_size);
void ConstructMatrixFromMatrixMarketPG_ps_wrp(int *ih_this, const char *file_name, const int *name_size,
                                              const int *ih_grid);
void WriteMatrixToMatrixMarket_ps_wrp(const int *ih_this, const char *file_name, const int *name_size);
void FillMatrixFromTripletList_psr_wrp(const int *ih_this, const int *ih_triplet_list);
void FillMatrixFromTripletList_psc_wrp(const int *ih_this, const int *ih_triplet_list);
void FillMatrixPermutation_ps_wrp(int *ih_this, const int *ih_permutation, const bool *permuterows);
void FillMatrixIdentity_ps_wrp(int *ih_this);
void GetMatrixActualDimension_ps_wrp(const int *ih_this, int *size);
void GetMatrixLogicalDimension_ps_wrp(const int *ih_this, int *size);
void GetMatrixSize_ps_wrp(const int *ih_this, long int *size);
void GetMatrixTripletList_psr_wrp(const int *ih_this, int *ih_triplet_list);
void GetMatrixTripletList_psc_wrp(const int *ih_this, int *ih_triplet_list);
void TransposeMatrix_ps_wrp(const int *ih_matA, int *ih_transmat);
void ConjugateMatrix_ps_wrp(int *ih_matA);
void GetMatrixProcessGrid_ps_wrp(const int *ih_this, int *ih_grid);
int IsIdentity_ps_wrp(const int *ih_this);
/* container utilities either side of the path (host-side logic over the ingest / egress primitives):
 * PSMatrixModule.F90:958-990, 1153-1225, 1718-1741, distributed_includes/{FillMatrixDense,GetMatrixBlock,SliceMatrix,
 * ResizeMatrix}.f90; Source/C/MatrixConversion_c.h:4 (MatrixConversionModule.F90:21-61) */
void FillMatrixDense_ps_wrp(int *ih_this);
void GetMatrixBlock_psr_wrp(const int *ih_this, int *ih_triplet_list, int *start_row, int *end_row,
                            int *start_column, int *end_column);
void GetMatrixBlock_psc_wrp(const int *ih_this, int *ih_triplet_list, int *start_row, int *end_row,
                            int *start_column, int *end_column);
void GetMatrixSlice_wrp(const int *ih_this, int *ih_submatrix, int *start_row, int *end_row, int *start_column,
                        int *end_column);
void ResizeMatrix_ps_wrp(int *ih_this, const int *new_size);
void SnapMatrixToSparsityPattern_wrp(int *ih_matA, const int *ih_matB);

/* ---- 4. THE HOT PATH: distributed algebra (Source/C/PSMatrix_c.h:51-68;
 *         PSMatrixAlgebraModule_wrp.F90:29-198 -> PSMatrixAlgebraModule.F90) */
void MatrixMultiply_ps_wrp(const int *ih_matA, const int *ih_matB, int *ih_matC, const double *alpha_in,
                           const double *beta_in, const double *threshold_in, int *ih_memory_pool_in);
void IncrementMatrix_ps_wrp(const int *ih_matA, int *ih_matB, const double *alpha_in,
                            const double *threshold_in);
void ScaleMatrix_ps_wrp(int *ih_this, const double *constant);
void MatrixTrace_ps_wrp(const int *ih_this, double *trace_val);
double MatrixNorm_ps_wrp(const int *ih_this);
void DotMatrix_psr_wrp(const int *ih_matA, const int *ih_matB, double *product);
void DotMatrix_psc_wrp(const int *ih_matA, const int *ih_matB, double *product_real, double *product_imag);
void MatrixPairwiseMultiply_ps_wrp(const int *ih_matA, const int *ih_matB, int *ih_matC);
double MeasureAsymmetry_ps_wrp(const int *ih_this);
void SymmetrizeMatrix_ps_wrp(int *ih_this);
/* column j scaled by every listed (index_column = j, point_value); PSMatrixAlgebraModule.F90:507-532 */
void MatrixDiagonalScale_psr_wrp(int *ih_mat, const int *ih_tlist);
void MatrixDiagonalScale_psc_wrp(int *ih_mat, const int *ih_tlist);

/* ---- 4b. THE LOCAL KERNEL: Matrix_lsr / Matrix_lsc and SMatrixAlgebraModule (Source/C/SMatrix_c.h:3-83;
 *          SMatrixModule_wrp.F90, SMatrixAlgebraModule_wrp.F90 -> SMatrixAlgebraModule.F90:116-287,
 *          sparse_includes/{GemmMatrix,SparseBranch,DenseBranch,MultiplyBlock,PruneList}.f90).
 *          One device-resident CSC block per handle; MatrixMultiply_ls*_wrp is the local product the
 *          distributed multiply runs per block pair: C = alpha*op(A)*op(B) + beta*C, op(X) = X^T when the
 *          flag is set, entries kept by the dense rule (|v| > threshold before alpha) when both operands are
 *          more than 10 % full and by the sparse rule (|alpha*v| > threshold) otherwise. */
void ConstructMatrixFromFile_lsr_wrp(int *ih_this, const char *file_name, const int *name_size);
void ConstructMatrixFromTripletList_lsr_wrp(int *ih_this, const int *ih_triplet_list, const int *rows,
                                            const int *columns);
void ConstructZeroMatrix_lsr_wrp(int *ih_this, const int *rows, const int *columns);
void DestructMatrix_lsr_wrp(int *ih_this);
void CopyMatrix_lsr_wrp(const int *ih_matA, int *ih_matB);
void GetMatrixRows_lsr_wrp(const int *ih_this, int *rows);
void GetMatrixColumns_lsr_wrp(const int *ih_this, int *columns);
void ExtractMatrixRow_lsr_wrp(const int *ih_this, int *row_number, int *ih_row_out);
void ExtractMatrixColumn_lsr_wrp(const int *ih_this, int *column_number, int *ih_column_out);
void ScaleMatrix_lsr_wrp(int *ih_this, const double *constant);
void IncrementMatrix_lsr_wrp(const int *ih_matA, int *ih_matB, const double *alpha_in, const double *threshold_in);
void DotMatrix_lsr_wrp(const int *ih_matA, const int *ih_matB, double *product);
void PairwiseMultiplyMatrix_lsr_wrp(const int *ih_matA, const int *ih_matB, int *ih_matC);
void MatrixMultiply_lsr_wrp(const int *ih_matA, const int *ih_matB, int *ih_matC, const bool *IsATransposed,
                            const bool *IsBTransposed, const double *alpha, const double *beta,
                            const double *threshold, int *ih_matrix_memory_pool);
void TransposeMatrix_lsr_wrp(const int *ih_matA, int *ih_matAT);
void PrintMatrix_lsr_wrp(const int *ih_this);
void PrintMatrixF_lsr_wrp(const int *ih_this, const char *file_name, const int *name_size);
void MatrixToTripletList_lsr_wrp(const int *ih_this, int *ih_triplet_list);
void MatrixDiagonalScale_lsr_wrp(int *ih_mat, const int *ih_tlist);
void ConstructMatrixFromFile_lsc_wrp(int *ih_this, const char *file_name, const int *name_size);
void ConstructMatrixFromTripletList_lsc_wrp(int *ih_this, const int *ih_triplet_list, const int *rows,
                                            const int *columns);
void ConstructZeroMatrix_lsc_wrp(int *ih_this, const int *rows, const int *columns);
void DestructMatrix_lsc_wrp(int *ih_this);
void CopyMatrix_lsc_wrp(const int *ih_matA, int *ih_matB);
void GetMatrixRows_lsc_wrp(const int *ih_this, int *rows);
void GetMatrixColumns_lsc_wrp(const int *ih_this, int *columns);
void ExtractMatrixRow_lsc_wrp(const int *ih_this, int *row_number, int *ih_row_out);
void ExtractMatrixColumn_lsc_wrp(const int *ih_this, int *column_number, int *ih_column_out);
void ScaleMatrix_lsc_wrp(int *ih_this, const double *constant);
void IncrementMatrix_lsc_wrp(const int *ih_matA, int *ih_matB, const double *alpha_in, const double *threshold_in);
void DotMatrix_lsc_wrp(const int *ih_matA, const int *ih_matB, double *product_real, double *product_complex);
void PairwiseMultiplyMatrix_lsc_wrp(const int *ih_matA, const int *ih_matB, int *ih_matC);
void MatrixMultiply_lsc_wrp(const int *ih_matA, const int *ih_matB, int *ih_matC, const bool *IsATransposed,
                            const bool *IsBTransposed, const double *alpha, const double *beta,
                            const double *threshold, int *ih_matrix_memory_pool);
void TransposeMatrix_lsc_wrp(const int *ih_matA, int *ih_matAT);
void ConjugateMatrix_lsc_wrp(int *ih_matA);
void PrintMatrix_lsc_wrp(const int *ih_this);
void PrintMatrixF_lsc_wrp(const int *ih_this, const char *file_name, const int *name_size);
void MatrixToTripletList_lsc_wrp(const int *ih_this, int *ih_triplet_list);
void MatrixDiagonalScale_lsc_wrp(int *ih_mat, const int *ih_tlist);

/* ---- 5. memory pools (Source/C/PMatrixMemoryPool_c.h:4-5, MatrixMemoryPool_c.h:3-8). The reference's pools
 *         hold 36 B per element of the dense local block (MatrixMemoryPoolModule.F90:13-53); here the scratch of a
 *         product comes from the device arena and the handles only keep the API and the shape. */
void ConstructMatrixMemoryPool_p_wrp(int *ih_this, const int *ih_matrix);
void DestructMatrixMemoryPool_p_wrp(int *ih_this);
void ConstructMatrixMemoryPool_lr_wrp(int *ih_this, const int *columns, const int *rows);
void DestructMatrixMemoryPool_lr_wrp(int *ih_this);
void ConstructMatrixMemoryPool_lc_wrp(int *ih_this, const int *columns, const int *rows);
void DestructMatrixMemoryPool_lc_wrp(int *ih_this);

/* ---- 6. solver parameters / permutation / load balancer
 *         (Source/C/SolverParameters_c.h, Permutation_c.h, LoadBalancer_c.h) */
void ConstructSolverParameters_wrp(int *ih_this);
void SetParametersConvergeDiff_wrp(int *ih_this, const double *new_value);
void SetParametersMaxIterations_wrp(int *ih_this, const int *new_value);
void SetParametersBeVerbose_wrp(int *ih_this, const bool *new_value);
void SetParametersThreshold_wrp(int *ih_this, const double *new_value);
void SetParametersLoadBalance_wrp(int *ih_this, const int *ih_permutation);
void SetParametersStepThreshold_wrp(int *ih_this, const double *new_value);
void SetParametersMonitorConvergence_wrp(int *ih_this, const bool *new_value);
void DestructSolverParameters_wrp(int *ih_this);
void ConstructDefaultPermutation_wrp(int *ih_this, const int *matrix_dimension);
void ConstructReversePermutation_wrp(int *ih_this, const int *matrix_dimension);
void ConstructRandomPermutation_wrp(int *ih_this, const int *matrix_dimension);
void DestructPermutation_wrp(int *ih_this);
void PermuteMatrix_wrp(const int *ih_mat_in, int *ih_mat_out, const int *ih_permutation, int *ih_memorypool);
void UndoPermuteMatrix_wrp(const int *ih_mat_in, int *ih_mat_out, const int *ih_permutation,
                           int *ih_memorypool);

/* ---- 7. drivers that iterate on the hot path (signatures frozen)
 *         Source/C/DensityMatrixSolvers_c.h, SignSolvers_c.h, InverseSolvers_c.h,
 *         SquareRootSolvers_c.h, ExponentialSolvers_c.h, EigenBounds_c.h.
 *         (the reference headers mark energy/chemical potential `const double*`
 *          although the Fortran side writes them; they are outputs.) */
void TRS2_wrp(const int *ih_Hamiltonian, const int *ih_InverseSquareRoot, const double *trace,
              int *ih_Density, double *energy_value_out, double *chemical_potential_out,
              const int *ih_solver_parameters);
void TRS4_wrp(const int *ih_Hamiltonian, const int *ih_InverseSquareRoot, const double *trace,
              int *ih_Density, double *energy_value_out, double *chemical_potential_out,
              const int *ih_solver_parameters);
void PM_wrp(const int *ih_Hamiltonian, const int *ih_InverseSquareRoot, const double *trace,
            int *ih_Density, double *energy_value_out, double *chemical_potential_out,
            const int *ih_solver_parameters);
void HPCP_wrp(const int *ih_Hamiltonian, const int *ih_InverseSquareRoot, const double *trace,
              int *ih_Density, double *energy_value_out, double *chemical_potential_out,
              const int *ih_solver_parameters);
void ScaleAndFold_wrp(const int *ih_Hamiltonian, const int *ih_InverseSquareRoot, const double *trace,
                      int *ih_Density, const double *homo, const double *lumo, double *energy_value_out,
                      const int *ih_solver_parameters);
void EnergyDensityMatrix_wrp(const int *ih_Hamiltonian, const int *ih_Density, int *ih_EnergyDensity,
                             const double *threshold);
void McWeenyStep_wrp(const int *ih_D, int *ih_DOut, const double *threshold);
void McWeenyStepS_wrp(const int *ih_D, int *ih_DOut, const int *ih_S, const double *threshold);
void SignFunction_wrp(const int *ih_mat1, int *ih_signmat, const int *ih_solver_parameters);
void PolarDecomposition_wrp(const int *ih_mat1, int *ih_umat, int *ih_hmat, const int *ih_solver_parameters);
void Invert_wrp(const int *ih_Hamiltonian, int *ih_Inverse, const int *ih_solver_parameters);
void PseudoInverse_wrp(const int *ih_Hamiltonian, int *ih_Inverse, const int *ih_solver_parameters);
void SquareRoot_wrp(const int *ih_Input, int *ih_Output, const int *ih_solver_parameters);
void InverseSquareRoot_wrp(const int *ih_Input, int *ih_Output, const int *ih_solver_parameters);
void ComputeExponential_wrp(const int *ih_Input, int *ih_Output, const int *ih_solver_parameters);
void GershgorinBounds_wrp(const int *ih_Hamiltonian, double *max_value, double *min_value);
void PowerBounds_wrp(const int *ih_Hamiltonian, double *max_value, const int *ih_solver_parameters);

/* ---- 7b. logging (Source/C/Logging_c.h:4-7): accepted so that front ends which activate NTPoly's YAML logger link and
 *          run; the logger itself (host text output) is outside the path */
void ActivateLogger_wrp(const bool *start_document);
void ActivateLoggerFile_wrp(const bool *start_document, const char *file_name, const int *name_size);
void DeactivateLogger_wrp(void);

/* ---- 8. ntb_ extensions (no counterpart in the reference) ------------------- */
/* bootstrap: rank/size of this process and, for size>1, the 128-byte ncclUniqueId that
 * rank 0 obtained from ntb_nccl_unique_id() and the host broadcast (e.g. torch.distributed). */
void ntb_nccl_unique_id(void *out128);
void ntb_world_init(int rank, int size, const void *nccl_unique_id_128);
int ntb_world_rank(void);
int ntb_world_size(void);
/* run the hot path on a caller-owned CUDA stream (cudaStream_t passed as void*) */
void ntb_set_stream(void *cuda_stream);
void ntb_synchronize(void);
/* bulk triplet transfer: 1-based global indices, n entries */
void ntb_TripletList_r_set(int *ih_this, long long n, const int *rows, const int *cols, const double *vals);
void ntb_TripletList_r_get(const int *ih_this, int *rows, int *cols, double *vals);
void ntb_TripletList_c_set(int *ih_this, long long n, const int *rows, const int *cols,
                           const double *vals_interleaved_re_im);
void ntb_TripletList_c_get(const int *ih_this, int *rows, int *cols, double *vals_interleaved_re_im);
/* host arrays straight into / out of a matrix (skips the triplet-list object) */
void ntb_FillMatrixFromArrays_ps(int *ih_this, long long n, const int *rows, const int *cols,
                                 const double *vals, int is_complex_interleaved);
long long ntb_GetMatrixLocalSize_ps(const int *ih_this);
void ntb_GetMatrixArrays_ps(const int *ih_this, int *rows, int *cols, double *vals);
/* The same for a real matrix without waiting for the device-to-host copies (they run on a second stream and overlap
 * whatever is enqueued next, e.g. the ingest of the next matrix); returns the local entry count. The host arrays
 * (pinned memory, or the copies serialise) are complete after ntb_EgressWait(). capacity = entries the three host
 * arrays can hold: when the block has more, NOTHING is copied and minus the required count is returned. */
long long ntb_GetMatrixArraysAsync_ps(const int *ih_this, long long capacity, int *rows, int *cols, double *vals);
void ntb_EgressWait(void);
/* Ingest in two halves, so that the host-to-device copies of the NEXT matrix overlap the work on the current one:
 * ntb_StageArrays only enqueues the copies of a real 1-based global list (pinned host arrays, valid until the fill) on a
 * copy stream and returns; ntb_FillMatrixFromStaged_ps makes the library stream wait for them, builds the matrix like
 * ntb_FillMatrixFromArrays_ps and releases the stage (handle zeroed). */
void ntb_StageArrays(int *ih_stage, long long n, const int *rows, const int *cols, const double *vals);
void ntb_FillMatrixFromStaged_ps(int *ih_this, int *ih_stage);
/* triplet lists since the last reset that were taken as they came (every rank's list its own block in column-major
 * order without duplicates): no sort, no gather */
double ntb_sorted_ingests(void);
void ntb_ConstructEmptyMatrixComplex_ps(int *ih_this, const int *matrix_dim, const int *is_complex);
int ntb_MatrixIsComplex_ps(const int *ih_this);
void ntb_FilterMatrix_ps(int *ih_this, const double *threshold);
void ntb_ScaleMatrixComplex_ps(int *ih_this, const double *re, const double *im);
/* solver variants the reference exposes only through Fortran optional arguments */
void ntb_InverseSquareRootOrder_wrp(const int *ih_Input, int *ih_Output, const int *ih_solver_parameters,
                                    const int *order);
void ntb_SquareRootOrder_wrp(const int *ih_Input, int *ih_Output, const int *ih_solver_parameters,
                             const int *order);
void ntb_ConstructRandomPermutationSeeded(int *ih_this, const int *matrix_dimension, const long long *seed);
void ntb_SetPermutation(int *ih_this, const int *matrix_dimension, const int *index_lookup_1based);
/* counters: [0] kernels launched by this library, [1] multiplies, [2] useful flops,
 *           [3] block pairs that used the dense-branch rule */
void ntb_get_counters(double *out4);
void ntb_reset_counters(void);
/* compulsory bytes bytes(A)+bytes(B)+bytes(C_kept) accumulated over the local products since the
 * last reset (A counted once when A and B are the same matrix) */
double ntb_algorithmic_bytes(void);
/* out2 = {local products that ran on the FP64 tensor-core tile path, DMMA.8x8x4 instructions issued} */
void ntb_get_tile_counters(double *out2);
/* 1 (default): locally dense real products run on the FP64 tensor-core tile path; 0: scalar kernels only */
void ntb_set_tile_path(int on);
/* 1 (default): MatrixMultiplyShift fuses the identity shift into the product's emit pass; 0: two reference calls */
void ntb_set_fused_shift(int on);
/* 1 (default): the sign / polar iteration takes ||X_new - X|| out of the epilogue of the product that computes X_new
 * (tile path, one rank or a column-split grid) instead of a separate pass over both iterates; 0 (NTB_FUSED_NORM=0):
 * always the separate norm kernel. ntb_fused_norms: norms obtained that way since the reset. */
void ntb_set_fused_norm(int on);
double ntb_fused_norms(void);
/* CSC -> tile-form conversions since the last reset (0 per product once operands carry their tile forms) */
double ntb_tile_builds(void);
/* distributed products whose left operand was fetched as a tile halo (1 x C x 1 grids): {count, tile bytes received
 * from the peers (the rank's own tiles are used in place)} */
void ntb_get_halo_counters(double *out2);
/* peer memory over NVLink (column-split grids): {1 when every rank has mapped every other rank's slab, products that
 * read the neighbours' operand tiles in place (no copy, no NCCL), barrier/exchange kernels enqueued, peak bytes of the
 * peer-visible slab in use} */
void ntb_get_peer_counters(double *out4);
/* measured issue peak of the FP64 tensor-core instruction DMMA.8x8x4 on this GPU, TFLOP/s (best of `repeats` launches of
 * a register-only micro-kernel): the denominator of the FP64-tensor roofline fraction, measured in the same run */
double ntb_measure_dmma_peak_tflops(int repeats);
/* host waits for the library stream since the last ntb_reset_counters (a sign iteration in tile space needs 3: the two
 * products' task counts and the convergence norm) */
double ntb_get_sync_count(void);
/* device milliseconds per phase of the hot path accumulated while ntb_profile_enable(1) was on, then cleared:
 * out8[0] symbolic phases of the tile products, [2] their tails (outer index, per-K meta, publication), [3] convergence
 * norms / scalars, [4] tile-space combines (the numeric kernels themselves: ntb_profile_read) */
void ntb_profile_read_phases(double *out8);
/* 1 (default): column-split grids use the tile halo exchange; 0: always the reference-style CSC panel gather */
void ntb_set_halo_path(int on);
/* 0 (default): PermuteMatrix / UndoPermuteMatrix relabel the indices on the device; 1: the reference's two products by
 * permutation matrices (LoadBalancerModule.F90:38-47, 77-86). Same result bit for bit. */
void ntb_set_permute_gemm(int on);
/* 1 (default): TRS2 / TRS4 evaluate 2X - X^2, Fx + sigma*Gx, Tr(X2 Fx), Tr(X2 Gx), Tr(XH), Tr(X) straight from the tile
 * forms of iterates that came out of tile products; 0: the reference's call sequence on CSC entries.
 * Same results up to the summation order of the scalars. ntb_tile_combines: tile-space combinations since the reset. */
void ntb_set_fused_steps(int on);
double ntb_tile_combines(void);
/* output columns of local products served by the shared-memory hash accumulator (scattered patterns: wide row window,
 * few products) since the reset; NTB_HASH_BIN=0 in the environment sends them back to the window kernels */
double ntb_hash_columns(void);
/* complex local products that ran on the FP64 tensor-core tile path since the reset: a complex product C = A*B is one
 * real tile product of the embeddings A^ = [[Re A, -Im A], [Im A, Re A]] (rows and columns interleaved) and
 * B^ = (Re B; Im B) (rows interleaved) - the reference's ZGEMM dense branch (DMatrixModule.F90:517-593) at tile
 * granularity; NTB_COMPLEX_TILE=0 in the environment keeps complex products on the scalar window / hash kernels */
double ntb_complex_tile_products(void);
/* C = alpha*A*B (thresholded), then IncrementMatrix(Identity, C, sigma) with threshold 0 — the call pair of
 * SignSolversModule.F90:226-229 / SquareRootSolversModule.F90 as one entry point. */
void ntb_MatrixMultiplyShift_ps(const int *ih_matA, const int *ih_matB, int *ih_matC, const double *alpha,
                                const double *threshold, const double *sigma, const int *ih_identity,
                                int *ih_memory_pool);
/* One pass of the loop body of SignFunction (SignSolversModule.F90:213-234), exactly as the SignFunction_wrp driver
 * runs it: X advances in place, T1/T2 are work matrices (their contents afterwards are unspecified scratch: T2 holds
 * the previous iterate), the return value is ||X_new - X_old||. */
double ntb_SignIteration(int *ih_X, const int *ih_identity, int *ih_T1, int *ih_T2, const double *alpha_k,
                         const double *threshold, int *ih_memory_pool);
/* The same loop body out of place: X_next receives the next iterate and X is left untouched; the return value is
 * ||X_next - X||. SignFunction_wrp runs this and then exchanges the contents of X and X_next. */
double ntb_SignStep(const int *ih_X, const int *ih_identity, int *ih_T1, int *ih_Xnext, const double *alpha_k,
                    const double *threshold, int *ih_memory_pool);
/* Tile-space helpers of the fused TRS2 / TRS4 steps (DensityMatrixSolversModule.F90:394-400, 591-625), exposed for their
 * parity tests. Operands must be real matrices that live as tile forms (results of tile products); return 1 when the
 * helper ran, 0 when the caller has to issue the reference's call sequence instead.
 *   ntb_TileCombine_ps  mode 0: Out = alpha*P + beta*Q, a matched entry kept iff |v| > threshold (ScaleMatrix(Q, beta);
 *                       IncrementMatrix(P, Q, alpha, threshold));  mode 1 (P = X^2, Q = X): Out = Fx + sigma*Gx with
 *                       Fx = 4X - 3X^2, Gx = I - 2X + X^2
 *   ntb_TileScalars_ps  mode 0: out2[0] = DotMatrix(A, B);  mode 1 (A = X^2, B = X): out2 = {DotMatrix(X2, Fx),
 *                       DotMatrix(X2, Gx)};  mode 2: out2[0] = MatrixTrace(A) (B = NULL) */
int ntb_TileCombine_ps(const int *ih_P, const int *ih_Q, int mode, double alpha, double beta, double threshold, double sigma,
                       int *ih_Out);
int ntb_TileScalars_ps(int mode, const int *ih_A, const int *ih_B, double *out2);
/* Instrumentation, 0 by default: when 1 every multiply also counts its useful products
 * F = sum over the entries (k,j) of B of nnz(A(:,k)) (counters [2], ntb_last_solve [4]); one extra sweep over B and
 * one read-back per product. When 0 the counts are only taken where the path choice needs them. */
void ntb_set_flop_counting(int on);
/* out2 = {tile products emitted as outer index + right tile form only (their CSC entries deferred),
 *         deferred products whose entries had to be materialized later} */
void ntb_get_deferred_counters(double *out2);
/* device timing of the numeric SpGEMM kernels (CUDA events on the library stream):
 * enable, run, then read out2 = {total ms, number of timed products}; reading clears the record */
void ntb_profile_enable(int on);
void ntb_profile_read(double *out2);
/* record of the last solver call: [0] loop counter at exit, [1] last monitored value,
 *                                 [2] energy, [3] multiplies, [4] useful flops */
void ntb_last_solve(double *out5);
/* bytes of this rank's local block as counted for the roofline:
 * nnz*(sizeof(value)+4) + (cols+1)*4    */
long long ntb_MatrixAlgorithmicBytes_ps(const int *ih_this);
/* pure host arithmetic, usable without a GPU: the block rank `rank` owns on a rows x cols x slices grid.
 * out12 = {my_slice, my_row, my_col, logical_dim, local_rows, local_cols, start_row, start_col (0-based),
 *          row blocks, column blocks, row-communicator colour, column-communicator colour} */
void ntb_grid_layout(int rank, int size, int rows, int cols, int slices, int matrix_dim, int *out12);
/* NTPoly's automatic grid for `size` processes: out3 = {rows, cols, slices} (ProcessGridModule.F90:576-638) */
void ntb_default_grid(int size, int *out3);
const char *ntb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NTPOLY_B200_H */
