// Host drivers that iterate on the device hot path. Signatures and iteration logic
// mirror the reference solver modules; every matrix stays on the GPU for the whole solve.
#pragma once
#include "psmatrix.h"
#include <vector>

namespace ntb {

// PermutationModule.F90:13-20 (1-based lookups over the logical dimension)
struct Permutation {
  std::vector<int> index_lookup;
  std::vector<int> reverse_index_lookup;
};
void permutation_default(Permutation& p, int n);
void permutation_reverse(Permutation& p, int n);
void permutation_random(Permutation& p, int n, unsigned long long seed);

// ConvergenceMonitorModule.F90:12-28
struct Monitor {
  std::vector<double> win_short, win_long;
  double loose_cutoff = 1e-2, tight_cutoff = 1e-8;
  bool automatic = true;
  int nval = 0;
  void construct(bool automatic_in, double tight);
  void append(double v);
  bool converged(bool be_verbose) const;
};

// SolverParametersModule.F90:14-33
struct SolverParameters {
  double converge_diff = 1e-6;
  int max_iterations = 1000;
  double threshold = 0.0;
  bool be_verbose = false;
  bool do_load_balancing = false;
  Permutation balance_permutation;
  double step_thresh = 1e-2;
  bool monitor_convergence = true;
};

// per-solve record (extension: lets tests check "identical iteration counts")
struct SolveRecord {
  int loop_counter = 0;        // value of the Fortran loop variable II at exit
  double last_value = 0.0;     // last value fed to the monitor
  double energy = 0.0;
  unsigned long long multiplies = 0;
  double flops = 0.0;
};
SolveRecord& last_solve();

void set_permute_gemm(int on);
void set_fused_steps(int on);   // 1: PermuteMatrix as two products with permutation matrices (reference form), 0 (default): index relabelling
void permute_matrix(const Matrix& in, Matrix& out, const Permutation& p, MemoryPool* pool);
void undo_permute_matrix(const Matrix& in, Matrix& out, const Permutation& p, MemoryPool* pool);

void solve_trs2(const Matrix& H, const Matrix& ISQ, double trace, Matrix& K, double* energy_out, double* chempot_out,
                const SolverParameters& params);
void solve_trs4(const Matrix& H, const Matrix& ISQ, double trace, Matrix& K, double* energy_out, double* chempot_out,
                const SolverParameters& params);
void solve_pm(const Matrix& H, const Matrix& ISQ, double trace, Matrix& K, double* energy_out, double* chempot_out,
              const SolverParameters& params);
void solve_hpcp(const Matrix& H, const Matrix& ISQ, double trace, Matrix& K, double* energy_out, double* chempot_out,
                const SolverParameters& params);
void solve_scale_and_fold(const Matrix& H, const Matrix& ISQ, double trace, Matrix& K, double homo, double lumo,
                          double* energy_out, const SolverParameters& params);
void solve_sign(const Matrix& In, Matrix& Out, const SolverParameters& params);
double sign_step(const Matrix& X, const Matrix& Identity, Matrix& T1, Matrix& Xn, Matrix& OutT, double alpha_k,
                 double threshold, bool needs_transpose, MemoryPool* pool);
double sign_iteration(Matrix& X, const Matrix& Identity, Matrix& T1, Matrix& T2, Matrix& OutT, double alpha_k,
                      double threshold, bool needs_transpose, MemoryPool* pool);
void solve_polar(const Matrix& In, Matrix& U, Matrix* Hmat, const SolverParameters& params);
void solve_invert(const Matrix& In, Matrix& Out, const SolverParameters& params);
void solve_sqrt(const Matrix& In, Matrix& Out, const SolverParameters& params, bool inverse, int order);
void solve_power_bounds(const Matrix& M, double* max_value, const SolverParameters& params, bool default_params);
void solve_exponential(const Matrix& In, Matrix& Out, const SolverParameters& params);
void mcweeny_step(const Matrix& D, Matrix& Dout, const Matrix* S, double threshold);
void energy_density_matrix(const Matrix& H, const Matrix& D, Matrix& ED, double threshold);

}  // namespace ntb
