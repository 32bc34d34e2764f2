// Solver drivers over the device-resident hot path. Each routine follows the
// reference driver line by line (citations per function) so that iteration
// counts and convergence decisions are the same; all matrix arithmetic is the
// CUDA path in psmatrix.cu / spgemm.cu / ops.cu.
#include "solvers.h"
#include <cmath>
#include <cstdio>

namespace ntb {

static SolveRecord g_last;
SolveRecord& last_solve() { return g_last; }

// ---------------------------------------------------------------------------
void permutation_default(Permutation& p, int n) {       // PermutationModule.F90:30-47
  p.index_lookup.resize(n); p.reverse_index_lookup.resize(n);
  for (int i = 0; i < n; ++i) { p.index_lookup[i] = i + 1; p.reverse_index_lookup[i] = i + 1; }
}
void permutation_reverse(Permutation& p, int n) {       // PermutationModule.F90:50-68
  p.index_lookup.resize(n); p.reverse_index_lookup.resize(n);
  for (int i = 0; i < n; ++i) { p.index_lookup[i] = n - i; p.reverse_index_lookup[i] = i + 1; }
}
void permutation_random(Permutation& p, int n, unsigned long long seed) {   // PermutationModule.F90:71-115
  permutation_default(p, n);
  // same shuffle structure as the reference (swap the LAST entry with a random one, n times);
  // the reference draws from Fortran RANDOM_NUMBER, here a fixed-seed splitmix64 that every
  // rank evaluates identically (stands in for the MPI_Bcast of the lookup).
  unsigned long long s = seed ? seed : 0x9E3779B97F4A7C15ull;
  for (int ii = n; ii >= 1; --ii) {
    s += 0x9E3779B97F4A7C15ull;
    unsigned long long z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    const int ri = (int)std::floor(n * u);  // 0-based
    std::swap(p.index_lookup[n - 1], p.index_lookup[ri]);
  }
  for (int i = 0; i < n; ++i) p.reverse_index_lookup[p.index_lookup[i] - 1] = i + 1;
}

// ---------------------------------------------------------------------------
void Monitor::construct(bool automatic_in, double tight) {   // ConvergenceMonitorModule.F90:35-89
  win_short.assign(3, 0.0);
  win_long.assign(6, 0.0);
  loose_cutoff = 1e-2;
  tight_cutoff = tight;
  automatic = automatic_in;
  nval = 0;
}
void Monitor::append(double v) {                              // :99-119
  for (size_t i = 0; i + 1 < win_short.size(); ++i) win_short[i] = win_short[i + 1];
  for (size_t i = 0; i + 1 < win_long.size(); ++i) win_long[i] = win_long[i + 1];
  win_short.back() = v;
  win_long.back() = v;
  nval++;
}
bool Monitor::converged(bool be_verbose) const {              // :122-191
  const double last = win_short[win_short.size() - 1];
  const double last2 = win_short[win_short.size() - 2];
  if (be_verbose && world().rank == 0) std::printf("  - Convergence: %.15e\n", last);
  bool conv = !(std::fabs(last) > tight_cutoff);
  if (!automatic || conv) return conv;
  conv = true;
  if (nval < (int)win_long.size()) conv = false;
  double avs = 0, avl = 0;
  for (double v : win_short) avs += v;
  for (double v : win_long) avl += v;
  avs /= win_short.size();
  avl /= win_long.size();
  if (!(10 * avs > avl && avs / 10 < avl)) conv = false;
  if (!(10 * last > avl && last / 10 < avl)) conv = false;
  if (last < 0) conv = false;
  if (std::fabs(last) < std::fabs(last2)) conv = false;
  if (avl > loose_cutoff) conv = false;
  return conv;
}

// ---------------------------------------------------------------------------
// LoadBalancerModule.F90:16-92. The reference forms the row and column permutation matrices and multiplies twice
// (out = PR*in*PC resp. PC*in*PR); the same matrix is obtained here by relabelling the indices on the device
// (psmatrix.cu: mat_relabel): in(r,c) lands at (rev[r], rev[c]) for PermuteMatrix and at (fwd[r], fwd[c]) for
// UndoPermuteMatrix. NTB_PERMUTE_GEMM=1 / ntb_set_permute_gemm(1) selects the reference's two products instead.
// 1 (default): TRS2 / TRS4 evaluate their per-iteration helpers in tile space (csc.cuh: tile_combine,
// tile_form_scalars); 0 (NTB_FUSED_STEPS=0): always the reference's call sequence on CSC entries
static int g_fused_steps = -1;
void set_fused_steps(int on) { g_fused_steps = on ? 1 : 0; }
static bool fused_steps_enabled() {
  if (g_fused_steps < 0) { const char* e = std::getenv("NTB_FUSED_STEPS"); g_fused_steps = (e && e[0] == '0') ? 0 : 1; }
  return g_fused_steps == 1;
}
static int g_permute_gemm = -1;
void set_permute_gemm(int on) { g_permute_gemm = on ? 1 : 0; }
static bool permute_gemm() {
  if (g_permute_gemm < 0) { const char* e = std::getenv("NTB_PERMUTE_GEMM"); g_permute_gemm = (e && e[0] == '1') ? 1 : 0; }
  return g_permute_gemm == 1;
}
// only index_lookup is used, like the reference's FillMatrixPermutation (ConstructReversePermutation leaves an identity
// in reverse_index_lookup, PermutationModule.F90:64-67, so the inverse is formed here)
static void relabel_with(const Matrix& in, Matrix& out, const Permutation& p, bool inverse) {
  NTB_CHECK((int)p.index_lookup.size() >= in.logical_dim, "permutation shorter than the logical matrix dimension");
  std::vector<int> map0((size_t)in.logical_dim, 0);
  for (int i = 0; i < in.logical_dim; ++i) {
    const int l = p.index_lookup[(size_t)i] - 1;
    NTB_CHECK(l >= 0 && l < in.logical_dim, "permutation entry outside the logical matrix dimension");
    if (inverse) map0[(size_t)l] = i; else map0[(size_t)i] = l;
  }
  mat_relabel(in, out, map0.data());
}
void permute_matrix(const Matrix& in, Matrix& out, const Permutation& p, MemoryPool* pool) {   // :16-52
  if (!permute_gemm()) { relabel_with(in, out, p, true); return; }
  Matrix PR, PC, T;
  mat_construct_like(PR, in);
  mat_construct_like(PC, in);
  mat_fill_permutation(PR, p.index_lookup.data(), true);
  mat_fill_permutation(PC, p.index_lookup.data(), false);
  mat_multiply(PR, in, T, 1.0, 0.0, 0.0, pool);
  mat_multiply(T, PC, out, 1.0, 0.0, 0.0, pool);
}
void undo_permute_matrix(const Matrix& in, Matrix& out, const Permutation& p, MemoryPool* pool) {  // :55-92
  if (!permute_gemm()) { relabel_with(in, out, p, false); return; }
  Matrix PR, PC, T;
  mat_construct_like(PR, in);
  mat_construct_like(PC, in);
  mat_fill_permutation(PR, p.index_lookup.data(), true);
  mat_fill_permutation(PC, p.index_lookup.data(), false);
  mat_multiply(PC, in, T, 1.0, 0.0, 0.0, pool);
  mat_multiply(T, PR, out, 1.0, 0.0, 0.0, pool);
}

namespace {
struct SolveScope {   // per-solve accounting for SolveRecord
  unsigned long long m0; double f0;
  SolveScope() : m0(rt().multiplies), f0(rt().flops_useful) { g_last = SolveRecord(); }
  ~SolveScope() { g_last.multiplies = rt().multiplies - m0; g_last.flops = rt().flops_useful - f0; }
};

double dot_real(const Matrix& A, const Matrix& B) { double re, im; mat_dot(A, B, &re, &im); return re; }

struct DensitySetup { Matrix WH, IMat, ISQT; MemoryPool pool; };

// common prologue of PM/TRS2/TRS4/HPCP (e.g. DensityMatrixSolversModule.F90:342-362)
void density_setup(const Matrix& H, const Matrix& ISQ, const SolverParameters& p, DensitySetup& s) {
  mat_construct_like(s.IMat, H);
  mat_fill_identity(s.IMat);
  mat_transpose(ISQ, s.ISQT);
  mat_similarity_transform(H, ISQ, s.ISQT, s.WH, &s.pool, p.threshold);
  if (p.do_load_balancing) {
    permute_matrix(s.WH, s.WH, p.balance_permutation, &s.pool);
    permute_matrix(s.IMat, s.IMat, p.balance_permutation, &s.pool);
  }
}
// common epilogue (e.g. :421-433)
void density_finish(Matrix& X, const Matrix& ISQ, DensitySetup& s, const SolverParameters& p, Matrix& K) {
  if (p.do_load_balancing) undo_permute_matrix(X, X, p.balance_permutation, &s.pool);
  mat_similarity_transform(X, s.ISQT, ISQ, K, &s.pool, p.threshold);
}
}  // namespace

// ---------------------------------------------------------------------------
// TRS2  (DensityMatrixSolversModule.F90:285-481)
void solve_trs2(const Matrix& H, const Matrix& ISQ, double trace, Matrix& K, double* energy_out, double* chempot_out,
                const SolverParameters& p) {
  SolveScope scope;
  Monitor mon;
  mon.construct(p.monitor_convergence, p.converge_diff);
  std::vector<double> sigma((size_t)p.max_iterations + 1, 0.0);
  DensitySetup s;
  density_setup(H, ISQ, p, s);
  double e_min, e_max;
  mat_gershgorin(s.WH, &e_min, &e_max);
  Matrix X, X2;
  mat_copy(s.WH, X);
  mat_scale(X, -1.0);
  mat_increment(s.IMat, X, e_max, 0.0);
  mat_scale(X, 1.0 / (e_max - e_min));
  double energy = 0.0;
  int II = 1;
  for (II = 1; II <= p.max_iterations; ++II) {
    const double tv = mat_trace(X);
    sigma[II] = (trace - tv < 0.0) ? -1.0 : 1.0;
    // X^2 stays in tile space (both forms, entries deferred) when the product runs on the tile path
    mat_multiply(X, X, X2, 1.0, 0.0, p.threshold, &s.pool, WANT_LEFT | WANT_RIGHT);
    if (sigma[II] > 0.0) {
      // 2X - X^2 (ScaleMatrix(X, 2); IncrementMatrix(X2, X, -1, threshold)): straight from the tile forms when both
      // iterates live there, the reference's two calls otherwise
      Matrix T;
      if (fused_steps_enabled() && mat_tile_combine(X2, X, 0, -1.0, 2.0, p.threshold, 0.0, T, WANT_LEFT | WANT_RIGHT)) {
        std::swap(X, T);
      } else {
        mat_scale(X, 2.0);
        mat_increment(X2, X, -1.0, p.threshold);
      }
    } else {
      mat_copy(X2, X);
    }
    const double old = energy;
    energy = dot_real(X, s.WH);
    mon.append(energy - old);
    g_last.last_value = energy - old;
    if (mon.converged(p.be_verbose)) break;
  }
  const int total_iterations = II - 1;
  g_last.loop_counter = II;
  g_last.energy = energy;
  if (energy_out) *energy_out = energy;
  density_finish(X, ISQ, s, p, K);
  if (chempot_out) {
    double a = 0.0, b = 1.0, mid = 0.0;
    for (int it = 1; it <= p.max_iterations; ++it) {
      mid = (b - a) / 2.0 + a;
      double z = mid;
      for (int jj = 1; jj <= total_iterations; ++jj) z = (sigma[jj] < 0.0) ? z * z : 2.0 * z - z * z;
      if (z < 0.5) a = mid; else b = mid;
      if (std::fabs(z - 0.5) < p.converge_diff) break;
    }
    *chempot_out = e_max + (e_min - e_max) * mid;
  }
}

// ---------------------------------------------------------------------------
// Scale and Fold  (DensityMatrixSolversModule.F90:953-1117)
void solve_scale_and_fold(const Matrix& H, const Matrix& ISQ, double trace, Matrix& K, double homo, double lumo,
                          double* energy_out, const SolverParameters& p) {
  SolveScope scope;
  Monitor mon;
  mon.construct(p.monitor_convergence, p.converge_diff);
  DensitySetup s;
  density_setup(H, ISQ, p, s);
  double e_min, e_max;
  mat_gershgorin(s.WH, &e_min, &e_max);
  Matrix X, X2;
  mat_copy(s.WH, X);
  mat_scale(X, -1.0);
  mat_increment(s.IMat, X, e_max, 0.0);
  mat_scale(X, 1.0 / (e_max - e_min));
  double Beta = (e_max - lumo) / (e_max - e_min);
  double BetaBar = (e_max - homo) / (e_max - e_min);
  double energy = 0.0;
  int II = 1;
  for (II = 1; II <= p.max_iterations; ++II) {
    const double tv = mat_trace(X);
    if (tv > trace) {
      const double alpha = 2.0 / (2.0 - Beta);
      mat_scale(X, alpha);
      mat_increment(s.IMat, X, 1.0 - alpha, 0.0);
      mat_multiply(X, X, X2, 1.0, 0.0, p.threshold, &s.pool);
      mat_copy(X2, X);
      Beta = (alpha * Beta + 1 - alpha) * (alpha * Beta + 1 - alpha);
      BetaBar = (alpha * BetaBar + 1 - alpha) * (alpha * BetaBar + 1 - alpha);
    } else {
      const double alpha = 2.0 / (1.0 + BetaBar);
      mat_multiply(X, X, X2, 1.0, 0.0, p.threshold, &s.pool);
      mat_scale(X, 2 * alpha);
      mat_increment(X2, X, -1.0 * alpha * alpha, 0.0);
      Beta = 2.0 * alpha * Beta - alpha * alpha * Beta * Beta;
      BetaBar = 2.0 * alpha * BetaBar - alpha * alpha * BetaBar * BetaBar;
    }
    const double old = energy;
    energy = 2.0 * dot_real(X, s.WH);
    mon.append(energy - old);
    g_last.last_value = energy - old;
    if (mon.converged(p.be_verbose)) break;
  }
  g_last.loop_counter = II;
  g_last.energy = energy;
  if (energy_out) *energy_out = energy;
  density_finish(X, ISQ, s, p, K);
}

// ---------------------------------------------------------------------------
// TRS4  (DensityMatrixSolversModule.F90:485-716)
void solve_trs4(const Matrix& H, const Matrix& ISQ, double trace, Matrix& K, double* energy_out, double* chempot_out,
                const SolverParameters& p) {
  SolveScope scope;
  const double sigma_min = 0.0, sigma_max = 6.0;
  Monitor mon;
  mon.construct(p.monitor_convergence, p.converge_diff);
  std::vector<double> sigma((size_t)p.max_iterations + 1, 0.0);
  DensitySetup s;
  density_setup(H, ISQ, p, s);
  double e_min, e_max;
  mat_gershgorin(s.WH, &e_min, &e_max);
  Matrix X, X2, Fx, Gx, T;
  mat_copy(s.WH, X);
  mat_scale(X, -1.0);
  mat_increment(s.IMat, X, e_max, 0.0);
  mat_scale(X, 1.0 / (e_max - e_min));
  double energy = 0.0;
  int II = 1;
  for (II = 1; II <= p.max_iterations; ++II) {
    // X^2 stays in tile space (both forms, entries deferred) when the product runs on the tile path
    mat_multiply(X, X, X2, 1.0, 0.0, p.threshold, &s.pool, WANT_LEFT | WANT_RIGHT);
    // Fused step: Tr(X2 Fx), Tr(X2 Gx) and Fx + sigma Gx are evaluated entry by entry from the tile
    // forms of X2 and X with the reference's rounding sequence - Fx, Gx are never formed, nothing leaves tile space.
    // Otherwise (complex, scattered, general grids): the reference's call sequence below.
    double tfx = 0.0, tgx = 0.0;
    double both[2];
    const bool fused = fused_steps_enabled() && mat_tile_scalars(1, X2, &X, both);
    if (fused) {
      tfx = both[0]; tgx = both[1];
    } else {
      mat_copy(X2, Fx);
      mat_scale(Fx, -3.0);
      mat_increment(X, Fx, 4.0, 0.0);
      mat_copy(s.IMat, Gx);
      mat_increment(X, Gx, -2.0, 0.0);
      mat_increment(X2, Gx, 1.0, 0.0);
      tfx = dot_real(X2, Fx);
      tgx = dot_real(X2, Gx);
    }
    sigma[II] = (std::fabs(tgx) < 1.0e-14) ? 0.5 * (sigma_max - sigma_min) : (trace - tfx) / tgx;
    if (sigma[II] > sigma_max) {
      if (!(fused && mat_tile_combine(X2, X, 0, -1.0, 2.0, 0.0, 0.0, T, WANT_LEFT | WANT_RIGHT))) {
        mat_copy(X, T);
        mat_scale(T, 2.0);
        mat_increment(X2, T, -1.0, 0.0);
      }
    } else if (sigma[II] < sigma_min) {
      mat_copy(X2, T);
    } else {
      bool done = false;
      if (fused) {
        Matrix FG;                               // Fx + sigma*Gx, needed as a right operand only
        if (mat_tile_combine(X2, X, 1, 0.0, 0.0, 0.0, sigma[II], FG, WANT_RIGHT)) {
          mat_multiply(X2, FG, T, 1.0, 0.0, p.threshold, &s.pool, WANT_LEFT | WANT_RIGHT);
          done = true;
        }
      }
      if (!done) {
        if (fused) {                             // (the combine declined: form Fx, Gx after all)
          mat_copy(X2, Fx);
          mat_scale(Fx, -3.0);
          mat_increment(X, Fx, 4.0, 0.0);
          mat_copy(s.IMat, Gx);
          mat_increment(X, Gx, -2.0, 0.0);
          mat_increment(X2, Gx, 1.0, 0.0);
        }
        mat_scale(Gx, sigma[II]);
        mat_increment(Fx, Gx, 1.0, 0.0);
        mat_multiply(X2, Gx, T, 1.0, 0.0, p.threshold, &s.pool);
      }
    }
    // reference :624-625 first forms X_k - TempMat and then overwrites it with TempMat;
    // the discarded difference has no effect on any result and is not computed here.
    std::swap(X, T);                             // (the reference copies TempMat into X; T is scratch)
    const double old = energy;
    energy = dot_real(X, s.WH);
    mon.append(energy - old);
    g_last.last_value = energy - old;
    if (mon.converged(p.be_verbose)) break;
  }
  const int total_iterations = II - 1;
  g_last.loop_counter = II;
  g_last.energy = energy;
  if (energy_out) *energy_out = energy;
  density_finish(X, ISQ, s, p, K);
  if (chempot_out) {
    double a = 0.0, b = 1.0, mid = 0.0;
    for (int it = 1; it <= p.max_iterations; ++it) {
      mid = (b - a) / 2.0 + a;
      double z = mid;
      for (int jj = 1; jj <= total_iterations; ++jj) {
        if (sigma[jj] > sigma_max) z = 2.0 * z - z * z;
        else if (sigma[jj] < sigma_min) z = z * z;
        else {
          const double fx = (z * z) * (4.0 * z - 3.0 * z * z);
          const double gx = (z * z) * (1.0 - z) * (1.0 - z);
          z = fx + sigma[jj] * gx;
        }
      }
      if (z < 0.5) a = mid; else b = mid;
      if (std::fabs(z - 0.5) < p.converge_diff) break;
    }
    *chempot_out = e_max + (e_min - e_max) * mid;
  }
}

// ---------------------------------------------------------------------------
// PM  (DensityMatrixSolversModule.F90:37-281)
void solve_pm(const Matrix& H, const Matrix& ISQ, double trace, Matrix& K, double* energy_out, double* chempot_out,
              const SolverParameters& p) {
  SolveScope scope;
  Monitor mon;
  mon.construct(p.monitor_convergence, p.converge_diff);
  std::vector<double> sigma((size_t)p.max_iterations + 1, 0.0);
  DensitySetup s;
  density_setup(H, ISQ, p, s);
  double e_min, e_max;
  mat_gershgorin(s.WH, &e_min, &e_max);
  const double n = (double)H.actual_dim;
  Matrix X, X2, X3, T;
  mat_copy(s.WH, X);
  const double lambda = mat_trace(X) / n;
  const double alpha = std::min(trace / (e_max - lambda), (n - trace) / (lambda - e_min));
  mat_scale(X, -alpha / n);
  mat_increment(s.IMat, X, (alpha * lambda + trace) / n, 0.0);
  double energy = 0.0;
  int II = 1;
  for (II = 1; II <= p.max_iterations; ++II) {
    mat_multiply(X, X, X2, 1.0, 0.0, p.threshold, &s.pool);
    mat_multiply(X, X2, X3, 1.0, 0.0, p.threshold, &s.pool);
    mat_copy(X, T);
    mat_increment(X2, T, -1.0, p.threshold);
    const double tv = mat_trace(T);
    const double tv2 = dot_real(T, X);
    sigma[II] = (tv <= 2.2250738585072014e-308) ? 1.0 : tv2 / tv;
    double a1, a2, a3;
    if (sigma[II] > 0.5) {
      a1 = 0.0; a2 = 1.0 + 1.0 / sigma[II]; a3 = -1.0 / sigma[II];
    } else {
      a1 = (1.0 - 2.0 * sigma[II]) / (1.0 - sigma[II]);
      a2 = (1.0 + sigma[II]) / (1.0 - sigma[II]);
      a3 = -1.0 / (1.0 - sigma[II]);
    }
    mat_scale(X, a1);
    mat_increment(X2, X, a2, p.threshold);
    mat_increment(X3, X, a3, p.threshold);
    const double old = energy;
    energy = dot_real(X, s.WH);
    mon.append(energy - old);
    g_last.last_value = energy - old;
    if (mon.converged(p.be_verbose)) break;
  }
  const int total_iterations = II - 1;
  g_last.loop_counter = II;
  g_last.energy = energy;
  if (energy_out) *energy_out = energy;
  density_finish(X, ISQ, s, p, K);
  if (chempot_out) {
    double a = 0.0, b = 1.0, mid = 0.0;
    for (int it = 1; it <= p.max_iterations; ++it) {
      mid = (b - a) / 2.0 + a;
      double z = mid;
      for (int jj = 1; jj <= total_iterations; ++jj) {
        if (sigma[jj] > 0.5) {
          z = ((1.0 + sigma[jj]) * z * z) - (z * z * z);
          z = z / sigma[jj];
        } else {
          z = ((1.0 - 2.0 * sigma[jj]) * z) + ((1.0 + sigma[jj]) * z * z) - (z * z * z);
          z = z / (1.0 - sigma[jj]);
        }
      }
      if (z < 0.5) a = mid; else b = mid;
      if (std::fabs(z - 0.5) < p.converge_diff) break;
    }
    *chempot_out = lambda - (n * mid - trace) / alpha;
  }
}

// ---------------------------------------------------------------------------
// HPCP  (DensityMatrixSolversModule.F90:720-946)
void solve_hpcp(const Matrix& H, const Matrix& ISQ, double trace, Matrix& K, double* energy_out, double* chempot_out,
                const SolverParameters& p) {
  SolveScope scope;
  Monitor mon;
  mon.construct(p.monitor_convergence, p.converge_diff);
  std::vector<double> sigma_array((size_t)p.max_iterations + 1, 0.0);
  DensitySetup s;
  density_setup(H, ISQ, p, s);
  const double n = (double)H.actual_dim;
  double e_min, e_max;
  mat_gershgorin(s.WH, &e_min, &e_max);
  const double mu = mat_trace(s.WH) / n;
  const double sigma_bar = (n - trace) / n;
  const double sigma = 1.0 - sigma_bar;
  const double beta = sigma / (e_max - mu);
  const double beta_bar = sigma_bar / (mu - e_min);
  const double beta_1 = sigma;
  const double beta_2 = std::min(beta, beta_bar);
  Matrix D1, DH, DDH, D2DH, T;
  mat_copy(s.IMat, D1);
  mat_scale(D1, beta_1);
  mat_copy(s.IMat, T);
  mat_scale(T, mu);
  mat_increment(s.WH, T, -1.0, 0.0);
  mat_scale(T, beta_2);
  mat_increment(T, D1, 1.0, 0.0);
  double energy = 0.0;
  int II = 1;
  for (II = 1; II <= p.max_iterations; ++II) {
    mat_copy(D1, DH);
    mat_increment(s.IMat, DH, -1.0, 0.0);
    mat_scale(DH, -1.0);
    mat_multiply(D1, DH, DDH, 1.0, 0.0, p.threshold, &s.pool);
    const double tv = mat_trace(DDH);
    mat_multiply(D1, DDH, D2DH, 1.0, 0.0, p.threshold, &s.pool);
    sigma_array[II] = mat_trace(D2DH) / tv;
    mat_increment(D2DH, D1, 2.0, 0.0);
    mat_increment(DDH, D1, -1.0 * 2.0 * sigma_array[II], 0.0);
    const double old = energy;
    energy = dot_real(D1, s.WH);
    mon.append(energy - old);
    g_last.last_value = energy - old;
    if (mon.converged(p.be_verbose)) break;
  }
  const int total_iterations = II - 1;
  g_last.loop_counter = II;
  g_last.energy = energy;
  if (energy_out) *energy_out = energy;
  density_finish(D1, ISQ, s, p, K);
  if (chempot_out) {
    double a = 0.0, b = 1.0, mid = 0.0;
    for (int it = 1; it <= p.max_iterations; ++it) {
      mid = (b - a) / 2.0 + a;
      double z = mid;
      for (int jj = 1; jj <= total_iterations; ++jj)
        z = z + 2.0 * ((z * z) * (1.0 - z) - sigma_array[jj] * z * (1.0 - z));
      if (z < 0.5) a = mid; else b = mid;
      if (std::fabs(z - 0.5) < p.converge_diff) break;
    }
    *chempot_out = mu + (beta_1 - mid) / beta_2;
  }
}

// ---------------------------------------------------------------------------
// EnergyDensityMatrix / McWeenyStep (DensityMatrixSolversModule.F90:1163-1231)
void energy_density_matrix(const Matrix& H, const Matrix& D, Matrix& ED, double threshold) {
  MemoryPool pool;
  mat_similarity_transform(H, D, D, ED, &pool, threshold);
}
void mcweeny_step(const Matrix& D, Matrix& Dout, const Matrix* S, double threshold) {
  MemoryPool pool;
  Matrix DS, DSD;
  if (S) mat_multiply(D, *S, DS, 1.0, 0.0, threshold, &pool); else mat_copy(D, DS);
  mat_multiply(DS, D, DSD, 1.0, 0.0, threshold, &pool);
  Matrix out;
  mat_multiply(DS, DSD, out, -2.0, 0.0, threshold, &pool);
  mat_increment(DSD, out, 3.0, 0.0);
  Dout = std::move(out);
}

// ---------------------------------------------------------------------------
// sign function / polar decomposition core (SignSolversModule.F90:150-258)
// One pass of the loop body (:213-234): X <- 0.5*a*X*(3I - a^2 X^T X); returns ||X_new - X_old||.
// "Gemm then IncrementMatrix(Identity, T1, 3)" is issued as one fused product (mat_multiply_shift).
// One pass of the loop body of SignFunction / PolarDecomposition (SignSolversModule.F90:207-240), out of place:
// Xn = 1/2 a X (3I - a^2 X^T X), X untouched; returns ||Xn - X||. T1 = 3I - a^2 X^T X only ever feeds the second
// product as its right operand, so it is emitted as a right tile form with deferred entries (csc.cuh).
double sign_step(const Matrix& X, const Matrix& Identity, Matrix& T1, Matrix& Xn, Matrix& OutT, double alpha_k,
                 double threshold, bool needs_transpose, MemoryPool* pool) {
  if (needs_transpose) {
    mat_transpose(X, OutT);
    if (OutT.is_complex) mat_conjugate(OutT);
    mat_multiply_shift(OutT, X, T1, -1.0 * alpha_k * alpha_k, threshold, 3.0, Identity, pool, WANT_RIGHT);
  } else {
    mat_multiply_shift(X, X, T1, -1.0 * alpha_k * alpha_k, threshold, 3.0, Identity, pool, WANT_RIGHT);
  }
  // the next iterate is read by the norm below and by the products of the next pass, all of which work on tile
  // forms: its CSC entries stay deferred until somebody asks for them (the caller reading the result)
  // reference: IncrementMatrix(T2, X, -1); norm = MatrixNorm(X); CopyMatrix(T2, X) — X - T2 is never needed itself,
  // and on the tile path not even a second pass over the two iterates: ||Xn - X|| comes out of the product's epilogue
  double norm_value = 0.0;
  if (mat_multiply_diffnorm(X, T1, Xn, 0.5 * alpha_k, threshold, pool, WANT_LEFT | WANT_RIGHT, &norm_value)) return norm_value;
  return mat_diff_norm(Xn, X, -1.0);
}
// ... and in place, as the drivers use it: X becomes the next iterate by exchanging it with the work matrix T2
// (the reference copies Temp2 into X; T2 is scratch either way and holds the previous iterate afterwards).
double sign_iteration(Matrix& X, const Matrix& Identity, Matrix& T1, Matrix& T2, Matrix& OutT, double alpha_k,
                      double threshold, bool needs_transpose, MemoryPool* pool) {
  const double norm_value = sign_step(X, Identity, T1, T2, OutT, alpha_k, threshold, needs_transpose, pool);
  std::swap(X, T2);
  return norm_value;
}

static void sign_core(const Matrix& In, Matrix& Out, const SolverParameters& p, bool needs_transpose) {
  const double alpha = 1.69770248526;
  Monitor mon;
  mon.construct(p.monitor_convergence, p.converge_diff);
  MemoryPool pool;
  Matrix Identity, T1, T2, OutT, X;
  mat_construct_like(Identity, In);
  mat_fill_identity(Identity);
  if (p.do_load_balancing) {
    permute_matrix(Identity, Identity, p.balance_permutation, &pool);
    permute_matrix(In, X, p.balance_permutation, &pool);
  } else {
    mat_copy(In, X);
  }
  double e_min, e_max;
  mat_gershgorin(In, &e_min, &e_max);
  double xk = std::fabs(e_min / e_max);
  mat_scale(X, 1.0 / std::fabs(e_max));
  int II = 1;
  for (II = 1; II <= p.max_iterations; ++II) {
    const double alpha_k = std::min(std::sqrt(3.0 / (1.0 + xk + xk * xk)), alpha);
    xk = 0.5 * alpha_k * xk * (3.0 - (alpha_k * alpha_k) * xk * xk);
    const double norm_value = sign_iteration(X, Identity, T1, T2, OutT, alpha_k, p.threshold, needs_transpose, &pool);
    mon.append(norm_value);
    g_last.last_value = norm_value;
    if (mon.converged(p.be_verbose)) break;
  }
  g_last.loop_counter = II;
  if (p.do_load_balancing) undo_permute_matrix(X, X, p.balance_permutation, &pool);
  Out = std::move(X);
}
void solve_sign(const Matrix& In, Matrix& Out, const SolverParameters& p) {
  SolveScope scope;
  sign_core(In, Out, p, false);
}
void solve_polar(const Matrix& In, Matrix& U, Matrix* Hmat, const SolverParameters& p) {   // :106-146
  SolveScope scope;
  sign_core(In, U, p, true);
  if (Hmat) {
    Matrix UT;
    mat_transpose(U, UT);
    if (UT.is_complex) mat_conjugate(UT);
    mat_multiply(UT, In, *Hmat, 1.0, 0.0, p.threshold, nullptr);
  }
}

// ---------------------------------------------------------------------------
// Hotelling inverse (InverseSolversModule.F90:29-149)
void solve_invert(const Matrix& In, Matrix& Out, const SolverParameters& p) {
  SolveScope scope;
  Monitor mon;
  mon.construct(p.monitor_convergence, p.converge_diff);
  MemoryPool pool;
  Matrix Identity, Bal, T1, T2, X;
  mat_construct_like(Identity, In);
  mat_fill_identity(Identity);
  if (p.do_load_balancing) {
    permute_matrix(Identity, Identity, p.balance_permutation, &pool);
    permute_matrix(In, Bal, p.balance_permutation, &pool);
  } else {
    mat_copy(In, Bal);
  }
  const double sigma = mat_sigma(Bal);
  mat_copy(Bal, X);
  mat_scale(X, sigma);
  int II = 1;
  for (II = 1; II <= p.max_iterations; ++II) {
    mat_multiply(X, Bal, T1, 1.0, 0.0, p.threshold, &pool);
    const double norm_value = mat_diff_norm(T1, Identity, -1.0);      // ||I - X*A|| (reference: Copy, Increment, Norm)
    mat_multiply(T1, X, T2, -1.0, 0.0, p.threshold, &pool);
    mat_scale(X, 2.0);
    mat_increment(T2, X, 1.0, p.threshold);
    mon.append(norm_value);
    g_last.last_value = norm_value;
    if (mon.converged(p.be_verbose)) break;
  }
  g_last.loop_counter = II;
  if (p.do_load_balancing) undo_permute_matrix(X, X, p.balance_permutation, &pool);
  Out = std::move(X);
}

// ---------------------------------------------------------------------------
// Newton-Schulz (inverse) square root (SquareRootSolversModule.F90:164-531)
static void ns_isr_order2(const Matrix& In, Matrix& Out, const SolverParameters& p, bool inverse) {   // :201-337
  Monitor mon;
  mon.construct(p.monitor_convergence, p.converge_diff);
  MemoryPool pool;
  Matrix Identity, Z, Y, X, T, Tk;
  mat_construct_like(Identity, In);
  mat_fill_identity(Identity);
  mat_construct_like(Z, In);
  mat_fill_identity(Z);
  mat_copy(In, Y);
  if (p.do_load_balancing) {
    permute_matrix(Y, Y, p.balance_permutation, &pool);
    permute_matrix(Identity, Identity, p.balance_permutation, &pool);
    permute_matrix(Z, Z, p.balance_permutation, &pool);
  }
  int II = 1;
  for (II = 1; II <= p.max_iterations; ++II) {
    mat_multiply(Y, Z, X, 1.0, 0.0, p.threshold, &pool);
    double e_min, e_max;
    mat_gershgorin(X, &e_min, &e_max);
    const double lambda = 1.0 / std::max(std::fabs(e_min), std::fabs(e_max));
    mat_scale(X, lambda);
    const double norm_value = mat_diff_norm(X, Identity, -1.0);       // ||I - X|| (reference: Copy, Increment, Norm)
    mat_copy(Identity, Tk);
    mat_scale(Tk, 3.0);
    mat_increment(X, Tk, -1.0, 0.0);
    mat_scale(Tk, 0.5);
    mat_copy(Z, T);
    mat_multiply(T, Tk, Z, 1.0, 0.0, p.threshold, &pool);
    mat_scale(Z, std::sqrt(lambda));
    mat_copy(Y, T);
    mat_multiply(Tk, T, Y, 1.0, 0.0, p.threshold, &pool);
    mat_scale(Y, std::sqrt(lambda));
    mon.append(norm_value);
    g_last.last_value = norm_value;
    if (mon.converged(p.be_verbose)) break;
  }
  g_last.loop_counter = II;
  Matrix res;
  if (inverse) res = std::move(Z); else res = std::move(Y);
  if (p.do_load_balancing) undo_permute_matrix(res, res, p.balance_permutation, &pool);
  Out = std::move(res);
}

static void ns_isr_taylor(const Matrix& In, Matrix& Out, const SolverParameters& p, int order, bool inverse) {  // :340-531
  Monitor mon;
  mon.construct(p.monitor_convergence, p.converge_diff);
  MemoryPool pool;
  Matrix Identity, Z, Y, X, T, T2;
  mat_construct_like(Identity, In);
  mat_fill_identity(Identity);
  double e_min, e_max;
  mat_gershgorin(In, &e_min, &e_max);
  const double lambda = 1.0 / std::max(std::fabs(e_min), std::fabs(e_max));
  mat_construct_like(Z, In);
  mat_fill_identity(Z);
  mat_copy(In, Y);
  mat_scale(Y, lambda);
  if (p.do_load_balancing) {
    permute_matrix(Y, Y, p.balance_permutation, &pool);
    permute_matrix(Identity, Identity, p.balance_permutation, &pool);
    permute_matrix(Z, Z, p.balance_permutation, &pool);
  }
  int II = 1;
  for (II = 1; II <= p.max_iterations; ++II) {
    mat_multiply_shift(Z, Y, X, 1.0, p.threshold, -1.0, Identity, &pool);      // X = Z*Y - I
    const double norm_value = mat_norm(X);
    if (order == 3) {
      mat_multiply(X, X, T, 1.0, 0.0, p.threshold, &pool);
      mat_scale(X, -0.5);
      mat_increment(Identity, X, 1.0, 0.0);
      mat_increment(T, X, 0.375, 0.0);
    } else if (order == 5) {
      const double aa = -40.0 / 35.0, bb = 48.0 / 35.0, cc = -64.0 / 35.0, dd = 128.0 / 35.0;
      const double a = (aa - 1.0) / 2.0;
      const double b = bb * (a + 1.0) - cc - a * (a + 1.0) * (a + 1.0);
      const double c = bb - b - a * (a + 1.0);
      const double d = dd - b * c;
      mat_multiply(X, X, T, 1.0, 0.0, p.threshold, &pool);
      mat_increment(X, T, a, 0.0);
      mat_copy(Identity, T2);
      mat_scale(T2, b);
      mat_increment(X, T2, 1.0, 0.0);
      mat_increment(T, T2, 1.0, 0.0);
      mat_increment(Identity, T, c, 0.0);
      mat_multiply_shift(T2, T, X, 1.0, p.threshold, d, Identity, &pool);
      mat_scale(X, 35.0 / 128.0);
    }
    // any other order falls through the reference's SELECT CASE without a polynomial step
    mat_copy(Z, T);
    mat_multiply(X, T, Z, 1.0, 0.0, p.threshold, &pool);
    mat_copy(Y, T);
    mat_multiply(T, X, Y, 1.0, 0.0, p.threshold, &pool);
    mon.append(norm_value);
    g_last.last_value = norm_value;
    if (mon.converged(p.be_verbose)) break;
  }
  g_last.loop_counter = II;
  Matrix res;
  if (inverse) { mat_scale(Z, std::sqrt(lambda)); res = std::move(Z); }
  else { mat_scale(Y, 1.0 / std::sqrt(lambda)); res = std::move(Y); }
  if (p.do_load_balancing) undo_permute_matrix(res, res, p.balance_permutation, &pool);
  Out = std::move(res);
}

void solve_sqrt(const Matrix& In, Matrix& Out, const SolverParameters& p, bool inverse, int order) {   // :164-198
  SolveScope scope;
  if (order == 2) ns_isr_order2(In, Out, p, inverse);
  else ns_isr_taylor(In, Out, p, order, inverse);
}

// ---------------------------------------------------------------------------
// PowerBounds (EigenBoundsModule.F90:60-189)
void solve_power_bounds(const Matrix& M, double* max_value, const SolverParameters& pin, bool default_params) {
  SolverParameters p = pin;
  if (default_params) p.max_iterations = 10;
  Monitor mon;
  mon.construct(p.monitor_convergence, p.converge_diff);
  MemoryPool pool;
  Matrix vec, vec2;
  mat_construct_like(vec, M);
  {
    std::vector<int> rows, cols;
    std::vector<double> vals;
    if (M.start_row == 0)
      for (int ii = M.start_col; ii < M.start_col + M.local_cols; ++ii) {
        rows.push_back(1); cols.push_back(ii + 1); vals.push_back(1.0 / (double)M.actual_dim);
      }
    mat_fill_from_triplets(vec, rows.data(), cols.data(), vals.data(), nullptr, (long long)rows.size(), true, true);
  }
  double ritz[3] = {0, 0, 0}, aitken[3] = {0, 0, 0};
  double mv = 0.0;
  int II = 1;
  for (II = 1; II <= p.max_iterations; ++II) {
    mat_multiply(M, vec, vec2, 1.0, 0.0, p.threshold, &pool);
    const double sv = dot_real(vec, vec);
    mv = dot_real(vec, vec2) / sv;
    const double scale_value = 1.0 / mat_norm(vec2);
    mat_scale(vec2, scale_value);
    mat_copy(vec2, vec);
    ritz[0] = ritz[1]; ritz[1] = ritz[2]; ritz[2] = mv;
    aitken[0] = aitken[1]; aitken[1] = aitken[2];
    if (II >= 3) {
      const double num = ritz[2] * ritz[0] - ritz[1] * ritz[1];
      const double den = ritz[2] - 2 * ritz[1] + ritz[0];
      aitken[2] = (std::fabs(den) > 1e-14) ? num / den : ritz[2];
    } else {
      aitken[2] = ritz[2];
    }
    mon.append(-(aitken[2] - aitken[1]));
    if (mon.converged(p.be_verbose)) {
      if (std::fabs(aitken[2] - ritz[2]) < mon.loose_cutoff) break;
    }
  }
  *max_value = aitken[2];
}

// ---------------------------------------------------------------------------
// Chebyshev evaluation (ChebyshevSolversModule.F90:83-186)
static void chebyshev_compute(const Matrix& In, Matrix& Out, const std::vector<double>& coef, const SolverParameters& p) {
  MemoryPool pool;
  const int degree = (int)coef.size();
  Matrix Identity, Bal, Tk, Tkm1, Tkm2, Res;
  mat_construct_like(Identity, In);
  mat_fill_identity(Identity);
  mat_copy(In, Bal);
  if (p.do_load_balancing) {
    permute_matrix(Identity, Identity, p.balance_permutation, &pool);
    permute_matrix(Bal, Bal, p.balance_permutation, &pool);
  }
  mat_copy(Identity, Tkm2);
  // One step of the recurrence, T_k = 2*Bal*T_{k-1} - T_{k-2}, Res += c_k*T_k (ChebyshevSolversModule.F90:146-163:
  // MatrixMultiply, IncrementMatrix(Tkm2, Tk, -1), IncrementMatrix(Tk, Res, c_k), both adds with threshold 0). When
  // the product runs on the tile path (real, locally dense iterates, one rank) the two adds are combined straight
  // from the tile forms (k_form_combine through the product's emit_strip) and the iterates never leave tile space:
  // 1*Tk + (-1)*Tkm2 and c*Tk + 1*Res are the same sums as the reference's (-1)*Tkm2 + Tk and c*Tk + Res (the add is
  // commutative, and with threshold 0 the rule "kept iff non-zero, or an untested tail" does not depend on which list
  // is called A), except that a tail entry that is EXACTLY zero does not exist in a tile form.
  auto step = [&](double c) {
    const bool fuse = fused_steps_enabled() && !Bal.is_complex && Bal.grid->size == 1;
    mat_multiply(Bal, Tkm1, Tk, 2.0, 0.0, p.threshold, &pool, fuse ? WANT_RIGHT : WANT_ALL);   // (T_k is only ever a right operand)
    Matrix T;
    if (fuse && mat_tile_combine(Tk, Tkm2, 0, 1.0, -1.0, 0.0, 0.0, T, WANT_RIGHT)) std::swap(Tk, T);
    else mat_increment(Tkm2, Tk, -1.0, 0.0);
    Matrix R;
    if (fuse && mat_tile_combine(Tk, Res, 0, c, 1.0, 0.0, 0.0, R, WANT_RIGHT)) std::swap(Res, R);
    else mat_increment(Tk, Res, c, 0.0);
  };
  if (degree == 1) {
    mat_copy(Tkm2, Res);
    mat_scale(Res, coef[0]);
  } else {
    mat_copy(Bal, Tkm1);
    mat_copy(Tkm2, Res);
    mat_scale(Res, coef[0]);
    mat_increment(Tkm1, Res, coef[1], 0.0);
    if (degree > 2) {
      step(coef[2]);
      for (int ii = 4; ii <= degree; ++ii) {
        mat_copy(Tkm1, Tkm2);
        mat_copy(Tk, Tkm1);
        step(coef[ii - 1]);
      }
    }
  }
  if (p.do_load_balancing) undo_permute_matrix(Res, Res, p.balance_permutation, &pool);
  Out = std::move(Res);
}

// ComputeExponential (ExponentialSolversModule.F90:37-148)
void solve_exponential(const Matrix& In, Matrix& Out, const SolverParameters& p) {
  SolveScope scope;
  SolverParameters sub = p, psub = p;
  psub.max_iterations = 10;
  MemoryPool pool;
  double spectral_radius = 0.0;
  solve_power_bounds(In, &spectral_radius, psub, false);
  double sigma_val = 1.0;
  int sigma_counter = 1;
  while (spectral_radius / sigma_val > 1.0) { sigma_val *= 2; sigma_counter++; }
  Matrix Scaled, Temp, Res;
  mat_copy(In, Scaled);
  mat_scale(Scaled, 1.0 / sigma_val);
  sub.threshold = sub.threshold / sigma_val;
  const std::vector<double> coef = {
      1.266065877752007e+00, 1.130318207984970e+00, 2.714953395340771e-01, 4.433684984866504e-02,
      5.474240442092110e-03, 5.429263119148932e-04, 4.497732295351912e-05, 3.198436462630565e-06,
      1.992124801999838e-07, 1.103677287249654e-08, 5.505891628277851e-10, 2.498021534339559e-11,
      1.038827668772902e-12, 4.032447357431817e-14, 2.127980007794583e-15, -1.629151584468762e-16};
  chebyshev_compute(Scaled, Res, coef, sub);
  if (p.do_load_balancing) permute_matrix(Res, Res, p.balance_permutation, &pool);
  for (int counter = 1; counter <= sigma_counter - 1; ++counter) {
    mat_multiply(Res, Res, Temp, 1.0, 0.0, p.threshold, &pool);
    mat_copy(Temp, Res);
  }
  if (p.do_load_balancing) undo_permute_matrix(Res, Res, p.balance_permutation, &pool);
  g_last.loop_counter = sigma_counter;
  Out = std::move(Res);
}

}  // namespace ntb
