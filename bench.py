#!/usr/bin/env python
"""Benchmark of the NTPoly hot path on B200 (contract: one JSON line on rank 0).

Default workload = BASELINE.json configs[3] (c4), the configuration the metric is quoted on (it fits one GPU):
one Newton-Schulz sign-function iteration (reference SignSolversModule.F90:207-240 - two thresholded distributed
multiplies, two sparse adds, one 1-norm) on the synthetic banded matrix N=262144 (half-bandwidth 82, 165 nnz/row)
shifted to straddle zero, threshold 1e-6, applied to the fixed iterate X_3 so that every step does identical work.
N GPUs share the SAME matrix on NTPoly's process grid (default 1xNx1, column split; --grid RxCxS for the others):
strong scaling.

  python bench.py [--gpus 1] [--steps 20] [--warmup 5] [--config c1|c3|c4|c5]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # CPU restatement of the reference algorithm (oracle/), same config/steps

Other configs (BASELINE.json configs 0, 2, 4; one GPU each unless torchrun is used):
  c1  one MatrixMultiply X*X, banded N=8192, thr 1e-8 (UnitTests/bench.f90 style); L2 flushed between steps
  c3  TRS4 purification of the block-sparse insulator N=65536 (32x32 blocks, 1 % block fill); step = one solve
  c5  complex Hermitian path: Hotelling inverse of a shifted copy + Chebyshev exponential; step = both solves

Every line carries: `roofline` (numeric SpGEMM kernel timed with CUDA events inside the library, against the HBM peak
of MEASURED_PEAKS.json and, for the FP64 tensor path, against a DMMA peak MEASURED IN THIS RUN), `e2e` (host arrays in
and out through the C ABI), `cpu_baseline` (the oracle on a bounded sample), `parity_checked` (the GPU result of a
reduced copy of the workload compared with the oracle's simulation of the benched grid, before anything is timed).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "GFLOP/s"
METRICS = {"c4": "spgemm_useful_gflops_per_sign_iteration", "c1": "spgemm_useful_gflops_single_multiply",
           "c3": "spgemm_useful_gflops_per_trs4_purification", "c5": "spgemm_useful_gflops_complex_inverse_exponential"}
ALPHA_MAX = 1.69770248526
# dram read+write bytes of one numeric launch (mean of the step's two products) from `ncu --set full` of the SHIPPED
# kernel on the c4 step: profiles/r02h_numeric.keys.txt (951 MB + 410 MB and 993 MB + 779 MB). Only quoted for c4 on 1 GPU.
NCU_TRAFFIC_C4 = 0.5 * ((951.27e6 + 410.40e6) + (992.53e6 + 779.18e6))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=["c1", "c3", "c4", "c5"])
    ap.add_argument("--n", type=int, default=0, help="matrix size (0: the config's own)")
    ap.add_argument("--threshold", type=float, default=0.0, help="0: the config's own")
    ap.add_argument("--iterate", type=int, default=3, help="c4: which Newton-Schulz iterate the step is applied to")
    ap.add_argument("--grid", default="", help="RxCxS process grid (default 1xNx1)")
    ap.add_argument("--cpu-n", type=int, default=32768, help="c4: matrix size of the bounded cpu_baseline sample / parity check")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-peaks", action="store_true")
    ap.add_argument("--c5-scale", type=float, default=0.001,
                    help="c5: the exponential's input is G*scale + 0.01*I. The reference example scales its 512-node graph by 0.5 "
                         "and gets a DENSE exponential; at N=32768 the Chebyshev polynomials T_k(x) of the reference's "
                         "evaluation fill in completely unless 560*scale^3 (the x^3 coefficient of T_15) stays below the "
                         "threshold: scale 0.005 -> dense 32768^2 iterates (measured, 8 s per solve), 0.001 -> ~650 per row")
    a = ap.parse_args()
    defaults = {"c1": (8192, 1e-8), "c3": (65536, 1e-6), "c4": (262144, 1e-6), "c5": (32768, 1e-6)}
    if a.n == 0:
        a.n = defaults[a.config][0]
    if a.threshold == 0.0:
        a.threshold = defaults[a.config][1]
    return a


def workload_name(args):
    n, thr = args.n, args.threshold
    if args.config == "c4":
        return (f"newton-schulz sign iteration (SignFunction driver loop body: 2 multiplies, identity shift, convergence "
                f"norm), banded N={n} half-bandwidth 82, thr={thr:g}, iterate X_{args.iterate}")
    if args.config == "c1":
        return f"single MatrixMultiply X*X, banded symmetric N={n} half-bandwidth 82 (165 nnz/row), thr={thr:g}"
    if args.config == "c3":
        return (f"TRS4 purification (whole solve = one step), block-sparse insulator N={n}, 32x32 blocks, 20 blocks per block "
                f"row, trace N/2, thr={thr:g}, converge 1e-5, identity overlap")
    return (f"complex Hermitian path (Guo-transformed directed ER graph, ~25 nnz/row) N={n}: Hotelling inverse of G + (8*||G||_1 + 1)*I "
            f"plus Chebyshev-16 exponential of {args.c5_scale:g}*G, thr={thr:g} (whole pair of solves = one step)")


def config_of(args, world):
    """identical in both arms (the driver compares them)"""
    R, C, S = grid_of(args, world)
    l2 = ("L2 flushed between steps (512 MB written)" if args.config == "c1"
          else "inputs exceed L2 (operands > 500 MB vs 126 MB L2)")
    return {"workload": workload_name(args), "grid": f"{R}x{C}x{S}", "l2": l2}


def grid_of(args, world):
    if args.grid:
        R, C, S = (int(x) for x in args.grid.split("x"))
        assert R * C * S == world, "--grid does not match the number of ranks"
        return R, C, S
    return 1, world, 1


def alpha_sequence(e_min, e_max, count):
    """scaling factors of the reference iteration (SignSolversModule.F90:166,209-211)"""
    xk = abs(e_min / e_max)
    out = []
    for _ in range(count):
        ak = min(math.sqrt(3.0 / (1.0 + xk + xk * xk)), ALPHA_MAX)
        xk = 0.5 * ak * xk * (3.0 - ak * ak * xk * xk)
        out.append(ak)
    return out


# --------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """statistics over the samples that arrived in [t0, t1 + 0.06] (everything when no window is given)"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if t0 is not None and not (t0 <= ts <= t1 + 0.06):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------
# the CPU side (oracle/): reference arm, cpu_baseline leg, parity check. Nothing else touches oracle/.
# --------------------------------------------------------------------------------------
def cpu_sign_iteration_factory(n, threshold, iterate, grid=None):
    from oracle import oracle as O
    from ntpoly_b200.workloads import banded_sign_input
    O.build()
    m = banded_sign_input(n)
    M = O.PSMatrix.from_scipy(m, O.Grid(*grid) if grid else None)
    I = O.identity(M)
    e_min, e_max = O.gershgorin(M)
    alphas = alpha_sequence(e_min, e_max, iterate)
    X = O.scale(M, 1.0 / abs(e_max))

    def step(X, ak, stats=None):
        T1 = O.multiply(X, X, alpha=-ak * ak, thr=threshold, stats=stats)
        T1 = O.increment(I, T1, alpha=3.0)
        T2 = O.multiply(X, T1, alpha=0.5 * ak, thr=threshold, stats=stats)
        D = O.increment(T2, X, alpha=-1.0)
        return T2, O.norm(D)

    for k in range(iterate - 1):
        X, _ = step(X, alphas[k])
    ak = alphas[iterate - 1]
    return O, (lambda stats=None: step(X, ak, stats))


def cpu_workload(args, sample: bool):
    """returns (O, step(stats) -> None, description of the sample). sample=True: the bounded cpu_baseline leg;
    sample=False: the reference arm (the full config wherever the CPU finishes a step in seconds)."""
    from oracle import oracle as O
    import scipy.sparse as sp
    from ntpoly_b200 import workloads as W
    O.build()
    thr = args.threshold
    if args.config == "c4":
        n = args.cpu_n if sample else args.n
        _, stepf = cpu_sign_iteration_factory(n, thr, args.iterate)
        what = (f"the same sign iteration on the banded N={n} matrix" + (f" (1/{args.n // n} of the columns; per-column "
                "work of a banded matrix is size independent)" if n != args.n else " (the full configuration)"))
        return O, (lambda st: stepf(st)), what
    if args.config == "c1":
        X = O.PSMatrix.from_scipy(W.banded(args.n))
        return O, (lambda st: O.multiply(X, X, thr=thr, stats=st)), f"the same product, banded N={args.n} (the full configuration)"
    if args.config == "c3":
        n = 4096
        H = O.PSMatrix.from_scipy(W.block_sparse_hamiltonian(n))
        I = O.identity(H)
        p = O.SolverParameters(converge_diff=1e-5, threshold=thr, max_iterations=2)

        def stepf(st):
            O.TOTAL_STATS = st
            O.trs4(H, I, n // 2, p)
            O.TOTAL_STATS = None
        return O, stepf, (f"the first 2 purification iterations of the same TRS4 solve on the N={n} instance of the "
                          f"generator (1/{args.n // n} of the block rows; the pattern is a block band, work per block row is "
                          "size independent)")
    # c5 on the CPU: the same generator, shift rule and scale on a smaller graph (the restatement's complex scattered
    # products run at ~1 GFLOP/s: N=32768 would take many minutes per solve)
    n = min(args.n, 2048 if sample else 4096)
    g = W.complex_hermitian_graph(n)
    shift = 8.0 * float(np.asarray(abs(g).sum(axis=0)).max()) + 1.0
    A = O.PSMatrix.from_scipy(sp.csc_matrix(g + sp.identity(n) * shift), is_complex=True)
    G = O.PSMatrix.from_scipy(sp.csc_matrix(g * args.c5_scale + sp.identity(n) * 0.01), is_complex=True)
    p = O.SolverParameters(converge_diff=1e-5, threshold=thr)

    def stepf(st):
        O.TOTAL_STATS = st
        O.invert(A, p)
        O.compute_exponential(G, p)
        O.TOTAL_STATS = None
    return O, stepf, f"the same two solves on the N={n} instance of the generator"


def run_cpu(args, steps, warmup, sample):
    O, stepf, what = cpu_workload(args, sample)
    for _ in range(warmup):
        stepf(None)
    st = O.MultiplyStats()
    t0 = time.perf_counter()
    for _ in range(steps):
        stepf(st)
    dt = time.perf_counter() - t0
    return {"gflops": st.flops / dt / 1e9, "ms_per_step": dt / steps * 1e3, "cores": O.lib().orc_max_threads(),
            "flops_per_step": st.flops / steps, "what": what}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    # all host cores (torchrun exports OMP_NUM_THREADS=1 to its workers; the OpenMP runtime has not been loaded yet)
    os.environ["OMP_NUM_THREADS"] = os.environ.get("BENCH_CPU_THREADS", str(os.cpu_count() or 1))
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    r = run_cpu(args, steps, warmup, sample=False)
    sample = (f"{r['what']}, {steps} steps after {warmup} warm-ups; CPU restatement of the reference algorithm (oracle/): "
              f"Gustavson with a windowed dense accumulator, OpenMP over columns, all host threads")
    line = {
        "impl": "reference", "metric": METRICS[args.config], "value": r["gflops"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "c128" if args.config == "c5" else "f64", "data": "synthetic",
        "config": config_of(args, max(world, args.gpus)),
        "cpu_baseline": {"value": r["gflops"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["gflops"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# parity helpers (bench --check): this rank's block against the oracle's simulation of the benched grid
# --------------------------------------------------------------------------------------
def block_parity(got_block, ref_block, thr):
    """relative Frobenius error on the common pattern + entries present on one side only that are NOT within the
    near-threshold band (tests/util.py: compare_sparse, restated without asserts)"""
    import scipy.sparse as sp
    got, ref = sp.csc_matrix(got_block), sp.csc_matrix(ref_block)
    g = sp.csc_matrix((np.ones(got.nnz), got.indices, got.indptr), shape=got.shape)
    r = sp.csc_matrix((np.ones(ref.nnz), ref.indices, ref.indptr), shape=ref.shape)
    common = g.multiply(r)
    bad = 0
    scale = max(float(abs(ref).max()) if ref.nnz else 0.0, 1e-300)
    for only, src in ((g - common, got), (r - common, ref)):
        only = sp.csc_matrix(only)
        only.eliminate_zeros()
        if only.nnz:
            vals = abs(np.asarray(src[only.nonzero()]).ravel())
            band = max(thr * 1e-6, 1e-13 * scale) + thr * 1e-6
            bad += int((np.abs(vals - thr) > band).sum())
    d = got.multiply(common) - ref.multiply(common)
    den = math.sqrt(abs(ref.multiply(common)).power(2).sum())
    num = math.sqrt(abs(d).power(2).sum())
    return (num / den if den > 0 else num), bad


def local_block_of(M):
    import scipy.sparse as sp
    rows, cols, vals = M.get_arrays()
    n = M.GetLogicalDimension()
    return sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(n, n)).tocsc()


def oracle_block_of(OM, rank):
    import scipy.sparse as sp
    rows, cols, vals = OM.local_triplets(rank)
    return sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(OM.N, OM.N)).tocsc()


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
class Env:
    """process world + library handles shared by the workloads"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import ntpoly_b200.api as nt
        self.torch, self.dist, self.nt, self.args = torch, dist, nt, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            # NCCL's log (communicator sizes: the driver checks that N ranks took part) goes to stderr: NCCL writes to
            # file descriptor 1, which now points at stderr; the JSON line goes to the saved original stdout
            if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
                os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "INFO")
                os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        nt.init_world_from_torch()
        assert self.world == args.gpus or self.world == 1, "launch with torchrun --nproc-per-node == --gpus"
        self.grid = grid_of(args, self.world)
        nt.ConstructGlobalProcessGrid(*self.grid)
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        nt.set_stream(self.stream.cuda_stream)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def allsum(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t]

    def fill(self, M, m):
        """every rank generates the same matrix and contributes a disjoint share of the triplets"""
        m = m.tocoo()
        sel = slice(self.rank, None, self.world)
        M.fill_from_arrays(m.row[sel] + 1, m.col[sel] + 1, m.data[sel])

    def pinned(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    def measure_peaks(self):
        """FP64 peaks of THIS GPU in THIS run: the DMMA.8x8x4 issue peak (register-only micro-kernel of the library) and
        a cuBLAS DGEMM (torch.matmul, 8192^3) - SURVEY 8d asks for the latter; the former is the tighter bound"""
        torch, nt = self.torch, self.nt
        out = {"dmma_issue_tflops": nt.measure_dmma_peak_tflops(3)}
        try:
            n = 8192
            a = torch.randn(n, n, dtype=torch.float64, device="cuda")
            b = torch.randn(n, n, dtype=torch.float64, device="cuda")
            torch.matmul(a, b)
            best = 0.0
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(a, b)
                e1.record()
                torch.cuda.synchronize()
                best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
            out["cublas_dgemm_tflops"] = best
            del a, b
        except Exception as ex:          # pragma: no cover
            out["cublas_dgemm_tflops"] = None
            out["cublas_error"] = str(ex)[:120]
        return out


def timed_steps(env, step, steps, flush=None, est_step_ms=None):
    """K steps between two events on the library stream, bracketed by barrier + synchronize; with `flush` (c1) every
    step has its own event pair and the L2 flush between steps is outside the summed time.
    Clocks: nvidia-smi needs ~0.1 s to start and samples every 50 ms, while K steps of the multi-GPU sign iteration
    take 13 ms. When the caller gives est_step_ms (the time of one step, max over ranks, so that every rank runs the
    same number of collective steps) the sampler is therefore started first, IDENTICAL untimed steps run for ~0.45 s
    immediately before the timed region, and the samples used are those taken while that uninterrupted load was
    running (`clocks.window` says so); the timed region itself and the counters read after it are unchanged."""
    torch, nt = env.torch, env.nt
    sampler = ClockSampler(env.local_rank)
    env.barrier()
    pre = 0
    if est_step_ms is not None and est_step_ms * steps < 400.0 and flush is None:
        pre = max(1, int(math.ceil(450.0 / max(est_step_ms, 1e-3))))
    sample = not os.environ.get("BENCH_NO_SAMPLER")
    if sample:
        sampler.start()
    t_load0 = time.time()
    for _ in range(pre):
        step()
    if pre:
        torch.cuda.synchronize()
        env.barrier()
    nt.reset_counters()
    nt.profile_enable(True)
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        env.barrier()
        ms_total = e0.elapsed_time(e1)
    else:
        pairs = []
        for _ in range(steps):
            flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            pairs.append((e0, e1))
        env.barrier()
        ms_total = sum(a.elapsed_time(b) for a, b in pairs)
    prof = nt.profile_read()
    prof["phases"] = nt.profile_read_phases()
    nt.profile_enable(False)
    t_load1 = time.time()
    clocks = sampler.stop(t_load0 + (0.1 if pre else 0.0), t_load1) if sample else {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler off"]}
    clocks["window"] = ("the timed region" if not pre else
                        f"{pre} identical untimed steps + the {steps} timed ones back to back: the timed region alone "
                        f"({ms_total:.1f} ms) is shorter than nvidia-smi's start-up and sampling period")
    return ms_total, clocks, prof


def roofline_of(env, prof, alg_bytes, flops_local, ms_total_local, peaks, kernel, extra=None):
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    nms = prof["numeric_ms"]
    achieved = alg_bytes / (nms * 1e-3) / 1e9 if nms > 0 else 0.0
    tf = flops_local / (nms * 1e-3) / 1e12 if nms > 0 else 0.0
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
         "kernel": kernel, "launches_timed": prof["products"], "peak_source": peak_src,
         "numeric_share_of_step": nms / ms_total_local if ms_total_local > 0 else 0.0,
         "step_ms_by_phase": dict({k: v / max(env.args.steps, 1) for k, v in prof.get("phases", {}).items()},
                                  numeric_ms=nms / max(env.args.steps, 1),
                                  unaccounted_ms=(ms_total_local - nms - sum(prof.get("phases", {}).values())) / max(env.args.steps, 1)),
         "fp64_tflops_useful": tf}
    if peaks:
        r["fp64_tensor_peak_tflops"] = peaks["dmma_issue_tflops"]
        r["fp64_peaks_measured_in_this_run"] = peaks
        r["fp64_frac"] = tf / peaks["dmma_issue_tflops"] if peaks["dmma_issue_tflops"] else None
    if extra:
        r.update(extra)
    return r


# ---------------------------------------------------------------------------------------------------------------
def run_c4(env):
    args, nt, torch = env.args, env.nt, env.torch
    from ntpoly_b200.workloads import banded_sign_input
    n, thr, rank, world = args.n, args.threshold, env.rank, env.world
    R, C, S = env.grid

    def prepare(nn):
        """(X_iterate, I, alpha_k, scratch) of the banded sign input of size nn, advanced to the quoted iterate"""
        m = banded_sign_input(nn)
        M = nt.Matrix_ps(nn)
        env.fill(M, m)
        del m
        I = nt.Matrix_ps(nn)
        I.FillIdentity()
        e_min, e_max = nt.EigenBounds.GershgorinBounds(M)
        alphas = alpha_sequence(e_min, e_max, args.iterate)
        X = nt.Matrix_ps(M)
        X.Scale(1.0 / abs(e_max))
        pool = nt.PMatrixMemoryPool(X)
        T1, T2, W = nt.Matrix_ps(nn), nt.Matrix_ps(nn), nt.Matrix_ps(nn)
        for k in range(args.iterate - 1):         # advance to the iterate the step is quoted on
            nt.sign_iteration(X, I, T1, T2, alphas[k], thr, pool)
        return M, X, I, alphas[args.iterate - 1], pool, T1, W

    # ---- parity of the benched step on the benched grid, reduced size (the oracle finishes it in seconds)
    parity = None
    if not args.no_check:
        nn = args.cpu_n
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
        _, Xc, Ic, akc, poolc, T1c, Wc = prepare(nn)
        nrm = nt.sign_step(Xc, Ic, T1c, Wc, akc, thr, poolc)
        O, cstep = cpu_sign_iteration_factory(nn, thr, args.iterate, grid=(R, C, S))
        ref, ref_norm = cstep()
        err, bad = block_parity(local_block_of(Wc), oracle_block_of(ref, rank), thr)
        worst = env.allmax([err, float(bad), abs(nrm - ref_norm) / max(abs(ref_norm), 1e-300)])
        parity = {"ok": bool(worst[0] <= 1e-10 and worst[1] == 0 and worst[2] <= 1e-8), "rel_fro_max_over_ranks": worst[0],
                  "pattern_differences_outside_threshold_band": int(worst[1]), "norm_rel_diff": worst[2],
                  "what": f"this bench step (sign iteration, iterate X_{args.iterate}, thr {thr:g}) at N={nn} on the {R}x{C}x{S} grid: "
                          f"every rank's block of X_next against the oracle's simulation of that grid",
                  "tolerance": "rel. Frobenius <= 1e-10 on the common pattern; pattern may differ only within 1e-6*thr of thr"}
        del Xc, Ic, T1c, Wc, poolc

    peaks = None if args.no_peaks else env.measure_peaks()
    M, X, I, ak, pool, T1, W = prepare(n)

    def step():
        """The loop body of SignFunction (SignSolversModule.F90:207-240) exactly as the SignFunction_wrp driver of this
        library runs it (csrc/solvers.cu: sign_step; the driver then exchanges X and the work matrix, a pointer swap),
        through the C ABI. W receives X_{k+1}; X_k is left untouched so that every step does identical work."""
        return nt.sign_step(X, I, T1, W, ak, thr, pool)

    # useful flops of one step: counted ONCE on a warm-up step (the timed steps repeat exactly this work); the
    # counting itself is instrumentation (one extra sweep + read-back per product) and is off in the timed region
    step()
    nt.set_flop_counting(True)
    nt.reset_counters()
    step()
    flops_per_step_local = nt.counters()["flops"]
    nt.set_flop_counting(False)
    warm = max(args.warmup, 3)
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    est_step_ms = env.allmax([(time.perf_counter() - t_w) * 1e3 / warm])[0]
    ms_total, clocks, prof = timed_steps(env, step, args.steps, est_step_ms=est_step_ms)
    cnt = nt.counters()
    fused_norms = nt.fused_norms()
    extra = {"tile_form_builds_in_timed_region": nt.tile_builds(), "peer": nt.peer_counters(),
             "fused_norms_in_timed_region": fused_norms,
             "numeric_time_includes": ("the step's convergence norm ||X_new - X||, taken in the epilogue of the second product "
                                       "(about +0.09 ms per step on one GPU instead of a 0.18 ms pass of its own); with "
                                       "NTB_FUSED_NORM=0 the two products alone give fp64_frac 0.49 / frac 0.25 "
                                       "(profiles/r02h_bench_1gpu_c4_separate_norm_kernel.json)") if fused_norms else "the two products only",
             "host_waits_per_step": nt.sync_count() / args.steps,
             "deferred_csc_products_in_timed_region": nt.deferred_counters()["products"],
             "deferred_csc_materialized_in_timed_region": nt.deferred_counters()["materialized"]}
    alg_bytes = nt.algorithmic_bytes()
    ms_local = ms_total
    ms_total = env.allmax([ms_total])[0]
    flops_total = env.allsum([flops_per_step_local * args.steps])[0]
    result = {"ms_per_step": ms_total / args.steps, "value": flops_total / (ms_total * 1e-3) / 1e9, "clocks": clocks,
              "launches": cnt["launches"], "warmup": warm}

    # ---- end to end, (a) the whole driver: host triplets in once, SignFunction_wrp, host triplets out once - the call
    # a user of NTPoly makes; per-iteration figure = solve / iterations. (b) streaming: host triplets in AND out every
    # step (three streams: copy-in, compute, copy-out)
    e2e = None
    if not args.no_e2e:
        rows, cols, vals = M.get_arrays()
        pin_in = [env.pinned(a) for a in (rows, cols, vals)]
        Min, Sout = nt.Matrix_ps(n), nt.Matrix_ps(n)
        sps = nt.SolverParameters()
        sps.SetThreshold(thr)
        sps.SetConvergeDiff(1e-4)
        cap = int(len(rows) * 1.6) + 4096
        pin_out = (torch.empty(cap, dtype=torch.int32).pin_memory().numpy(), torch.empty(cap, dtype=torch.int32).pin_memory().numpy(),
                   torch.empty(cap, dtype=torch.float64).pin_memory().numpy())

        def solve():
            Min.fill_from_arrays(*pin_in)
            nt.SignSolvers.ComputeSign(Min, Sout, sps)
            out = Sout.get_arrays(out=pin_out)
            return sum(a.nbytes for a in out)

        solve()
        nt.set_flop_counting(True)
        nt.reset_counters()
        solve()
        solve_flops_local = nt.counters()["flops"]
        iters = nt.last_solve()["loop_counter"]
        nt.set_flop_counting(False)
        solve()
        nsolve = 3
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(nsolve):
            out_bytes = solve()
        env.barrier()
        dt = env.allmax([time.perf_counter() - t0])[0]
        fl = env.allsum([solve_flops_local * nsolve])[0]
        in_bytes = sum(a.nbytes for a in pin_in)
        e2e = {"value": fl / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(in_bytes / iters),
               "d2h_bytes_per_step": int(out_bytes / iters) + 8, "ms_per_step": dt / nsolve / iters * 1e3,
               "what": "whole SignFunction solve through the C ABI (FillMatrixFromTripletList with pinned host arrays -> "
                       "SignFunction_wrp -> GetMatrixTripletList into pinned host arrays), converge 1e-4; a step = one "
                       "iteration of that solve: its share of the input/result copies plus the norm read-back every iteration",
               "solve": {"iterations": iters, "ms": dt / nsolve * 1e3, "h2d_bytes": int(in_bytes), "d2h_bytes": int(out_bytes),
                         "solves_timed": nsolve, "useful_gflop": fl / nsolve / 1e9}}
        del Min, Sout
        # (b) streaming
        rows, cols, vals = X.get_arrays()
        pin = [env.pinned(a) for a in (rows, cols, vals)]
        Xh = nt.Matrix_ps(n)
        cap = int(len(rows) * 1.5) + 1024
        pout = [(torch.empty(cap, dtype=torch.int32).pin_memory().numpy(),
                 torch.empty(cap, dtype=torch.int32).pin_memory().numpy(),
                 torch.empty(cap, dtype=torch.float64).pin_memory().numpy()) for _ in range(2)]
        e2e_steps = min(max(args.steps, 16), 24)

        def e2e_loop(count):
            d2h = 0
            st = nt.stage_arrays(*pin)                    # H2D of step 0's input (copy stream)
            for i in range(count):
                nxt = nt.stage_arrays(*pin) if i + 1 < count else None
                Xh.fill_from_staged(st)
                nt.sign_step(Xh, I, T1, W, ak, thr, pool)
                nt.egress_wait()                          # result i-1 has landed
                out = W.get_arrays_async(pout[i % 2])     # D2H of the step's result X_{k+1} (second copy stream)
                d2h = sum(a.nbytes for a in out) + 8
                st = nxt
            nt.egress_wait()
            return d2h

        e2e_loop(4)
        env.barrier()
        t0 = time.perf_counter()
        d2h = e2e_loop(e2e_steps)
        env.barrier()
        dt = env.allmax([time.perf_counter() - t0])[0]
        fl = env.allsum([flops_per_step_local * e2e_steps])[0]
        e2e["streaming"] = {"value": fl / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(sum(a.nbytes for a in pin)),
                            "d2h_bytes_per_step": int(d2h), "ms_per_step": dt / e2e_steps * 1e3, "steps": e2e_steps,
                            "what": "host triplets in AND out EVERY step (ntb_StageArrays + ntb_FillMatrixFromStaged_ps, "
                                    "ntb_SignStep, ntb_GetMatrixArraysAsync_ps + ntb_EgressWait), three streams; bound by "
                                    "PCIe: 16-byte triplets each way"}

    roof = None
    if rank == 0:
        extra["traffic_source"] = ("ncu --set full of the shipped kernel, mean of the step's two launches "
                                   "(profiles/r02h_numeric.keys.txt)") if world == 1 and n == 262144 else None
        roof = roofline_of(env, prof, alg_bytes, flops_per_step_local * args.steps, ms_local, peaks,
                           "k_tile_numeric9 (numeric SpGEMM incl. threshold and tile-form output, one launch per product)", extra)
        if world == 1 and n == 262144:
            roof["traffic"] = NCU_TRAFFIC_C4
        roof["note"] = ("arithmetic intensity of this product (~11 flop/B) is above the FP64 machine balance (DMMA peak / "
                        "HBM peak ~ 5.8 flop/B): the kernel is bound by the FP64 tensor pipe; fp64_frac is its share of "
                        "the DMMA peak measured in this run")
    return result, roof, e2e, parity


def run_c1(env):
    args, nt, torch = env.args, env.nt, env.torch
    from ntpoly_b200.workloads import banded
    from oracle import oracle as O
    n, thr, rank, world = args.n, args.threshold, env.rank, env.world
    m = banded(n)
    X, Cm = nt.Matrix_ps(n), nt.Matrix_ps(n)
    env.fill(X, m)
    pool = nt.PMatrixMemoryPool(X)

    def step():
        Cm.Gemm(X, X, pool, threshold=thr)

    parity = None
    if not args.no_check:
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
        step()
        OX = O.PSMatrix.from_scipy(m, O.Grid(*env.grid))
        ref = O.multiply(OX, OX, thr=thr)
        err, bad = block_parity(local_block_of(Cm), oracle_block_of(ref, rank), thr)
        worst = env.allmax([err, float(bad)])
        parity = {"ok": bool(worst[0] <= 1e-10 and worst[1] == 0), "rel_fro_max_over_ranks": worst[0],
                  "pattern_differences_outside_threshold_band": int(worst[1]),
                  "what": f"the benched product itself (full size N={n}) against the oracle on the benched grid"}
    peaks = None if args.no_peaks else env.measure_peaks()
    step()
    nt.set_flop_counting(True)
    nt.reset_counters()
    step()
    flops_local = nt.counters()["flops"]
    nt.set_flop_counting(False)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    ms_total, clocks, prof = timed_steps(env, step, args.steps, flush=lambda: flush_buf.zero_())
    cnt = nt.counters()
    alg_bytes = nt.algorithmic_bytes()
    ms_local = ms_total
    ms_total = env.allmax([ms_total])[0]
    flops_total = env.allsum([flops_local * args.steps])[0]
    result = {"ms_per_step": ms_total / args.steps, "value": flops_total / (ms_total * 1e-3) / 1e9, "clocks": clocks,
              "launches": cnt["launches"], "warmup": warm}
    e2e = None
    if not args.no_e2e:
        rows, cols, vals = X.get_arrays()
        pin_in = [env.pinned(a) for a in (rows, cols, vals)]
        Xh, Ch = nt.Matrix_ps(n), nt.Matrix_ps(n)

        def e2e_step():
            Xh.fill_from_arrays(*pin_in)
            Ch.Gemm(Xh, Xh, pool, threshold=thr)
            return Ch.get_arrays()

        e2e_step()
        env.barrier()
        k = max(args.steps, 10)
        t0 = time.perf_counter()
        for _ in range(k):
            out = e2e_step()
        env.barrier()
        dt = env.allmax([time.perf_counter() - t0])[0]
        e2e = {"value": env.allsum([flops_local * k])[0] / dt / 1e9, "unit": UNIT, "ms_per_step": dt / k * 1e3,
               "h2d_bytes_per_step": int(sum(a.nbytes for a in pin_in)), "d2h_bytes_per_step": int(sum(a.nbytes for a in out)),
               "what": "FillMatrixFromTripletList (pinned host arrays) -> MatrixMultiply_ps_wrp -> GetMatrixTripletList, every step"}
    roof = None
    if rank == 0:
        roof = roofline_of(env, prof, alg_bytes, flops_local * args.steps, ms_local, peaks,
                           "k_tile_numeric9 (numeric SpGEMM incl. threshold and tile-form output)",
                           {"tile_products": nt.tile_counters()["tile_products"]})
    return result, roof, e2e, parity


def run_c3(env):
    args, nt, torch = env.args, env.nt, env.torch
    from ntpoly_b200.workloads import block_sparse_hamiltonian
    from oracle import oracle as O
    n, thr, rank, world = args.n, args.threshold, env.rank, env.world

    def params(maxit=None):
        p = nt.SolverParameters()
        p.SetThreshold(thr)
        p.SetConvergeDiff(1e-5)
        if maxit:
            p.SetMaxIterations(maxit)
        return p

    parity = None
    if not args.no_check:
        nn = 2048
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
        h = block_sparse_hamiltonian(nn)
        Hc, ISQc, Kc = nt.Matrix_ps(nn), nt.Matrix_ps(nn), nt.Matrix_ps(nn)
        env.fill(Hc, h)
        ISQc.FillIdentity()
        e, _ = nt.DensityMatrixSolvers.TRS4(Hc, ISQc, nn // 2, Kc, params())
        its = nt.last_solve()["loop_counter"]
        OH = O.PSMatrix.from_scipy(h, O.Grid(*env.grid))
        ref, info = O.trs4(OH, O.identity(OH), nn // 2, O.SolverParameters(converge_diff=1e-5, threshold=thr))
        err, bad = block_parity(local_block_of(Kc), oracle_block_of(ref, rank), thr)
        worst = env.allmax([err, abs(e - info.energy) / abs(info.energy)])
        parity = {"ok": bool(its == info.iterations and worst[1] <= 1e-8 and worst[0] <= 1e-6),
                  "iterations": [its, info.iterations], "energy_rel_diff": worst[1], "rel_fro_max_over_ranks": worst[0],
                  "pattern_differences_outside_threshold_band": bad,
                  "what": f"the same TRS4 solve on the N={nn} instance of the generator against the oracle: identical "
                          "iteration count, Tr(KH) within 1e-8, density within 1e-6 (entries at the threshold may fall on "
                          "either side after ~40 thresholded products)"}
        del Hc, ISQc, Kc
    peaks = None if args.no_peaks else env.measure_peaks()
    h = block_sparse_hamiltonian(n)
    H, ISQ, K = nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n)
    env.fill(H, h)
    del h
    ISQ.FillIdentity()
    p = params()

    def step():
        return nt.DensityMatrixSolvers.TRS4(H, ISQ, n // 2, K, p)

    step()
    nt.set_flop_counting(True)
    nt.reset_counters()
    step()
    flops_local = nt.counters()["flops"]
    iters = nt.last_solve()["loop_counter"]
    mults = nt.last_solve()["multiplies"]
    nt.set_flop_counting(False)
    warm = max(min(args.warmup, 3), 1)
    for _ in range(warm - 1):
        step()
    ms_total, clocks, prof = timed_steps(env, step, args.steps)
    cnt = nt.counters()
    tc = nt.tile_counters()
    alg_bytes = nt.algorithmic_bytes()
    ms_local = ms_total
    ms_total = env.allmax([ms_total])[0]
    flops_total = env.allsum([flops_local * args.steps])[0]
    result = {"ms_per_step": ms_total / args.steps, "value": flops_total / (ms_total * 1e-3) / 1e9, "clocks": clocks,
              "launches": cnt["launches"], "warmup": warm + 1,
              "config_extra": {"purification_iterations": iters, "multiplies_per_solve": mults,
                               "sec_per_iteration": ms_total / args.steps / iters * 1e-3, "nnz_per_row_density": K.GetSize() / n}}
    e2e = None
    if not args.no_e2e:
        rows, cols, vals = H.get_arrays()
        pin_in = [env.pinned(a) for a in (rows, cols, vals)]
        Hh, Kh = nt.Matrix_ps(n), nt.Matrix_ps(n)

        def e2e_step():
            Hh.fill_from_arrays(*pin_in)
            nt.DensityMatrixSolvers.TRS4(Hh, ISQ, n // 2, Kh, p)
            return Kh.get_arrays()

        out = e2e_step()
        env.barrier()
        k = 2
        t0 = time.perf_counter()
        for _ in range(k):
            out = e2e_step()
        env.barrier()
        dt = env.allmax([time.perf_counter() - t0])[0]
        e2e = {"value": env.allsum([flops_local * k])[0] / dt / 1e9, "unit": UNIT, "ms_per_step": dt / k * 1e3,
               "h2d_bytes_per_step": int(sum(a.nbytes for a in pin_in)), "d2h_bytes_per_step": int(sum(a.nbytes for a in out)),
               "what": "FillMatrixFromTripletList (pinned host arrays) -> TRS4_wrp -> GetMatrixTripletList, every step (= solve)"}
    roof = None
    if rank == 0:
        roof = roofline_of(env, prof, alg_bytes, flops_local * args.steps, ms_local, peaks,
                           "k_tile_numeric9 (numeric SpGEMM incl. threshold and tile-form output, one launch per product)",
                           {"tile_products": tc["tile_products"], "products": cnt["multiplies"],
                            "dmma_useful_fraction": (flops_local * args.steps / 2.0) / (tc["dmma"] * 256.0) if tc["dmma"] else None})
    return result, roof, e2e, parity


def run_c5(env):
    args, nt, torch = env.args, env.nt, env.torch
    import scipy.sparse as sp
    import scipy.linalg as la
    from ntpoly_b200.workloads import complex_hermitian_graph
    n, thr, rank, world = args.n, args.threshold, env.rank, env.world

    def build(nn, scale=None):
        scale = args.c5_scale if scale is None else scale
        g = complex_hermitian_graph(nn)
        # Hotelling input: positive definite shifted copy, G + s*I with s = 8*||G||_1 + 1. (With s = ||G||_1 + 1 the
        # iteration cannot converge at thr = 1e-6: the dropped terms (G/s)^k, k >= 3, have a 1-norm of ~0.1-0.2, the
        # residual ||I - X*A||_1 stagnates there above the monitor's loose cutoff 1e-2 and the solve runs into
        # max_iterations - measured, gpurun_out/r2c16_probe_inv*.log; the reference's monitor behaves the same.)
        shift = 8.0 * float(np.asarray(abs(g).sum(axis=0)).max()) + 1.0
        a = sp.csc_matrix(g + sp.identity(nn) * shift)
        e = sp.csc_matrix(g * scale + sp.identity(nn) * 0.01)   # exponential input (non-zero (1,1), see oracle.compute_exponential)
        return a, e

    def params():
        p = nt.SolverParameters()
        p.SetThreshold(thr)
        p.SetConvergeDiff(1e-5)
        return p

    parity = None
    if not args.no_check:
        # the same pair of solves on the N=1024 instance of the generator (same shift rule, same scale) against the
        # oracle's simulation of the benched grid: identical iteration counts, every rank's block within 1e-8
        from oracle import oracle as O
        nn = 1024
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
        a, e = build(nn, args.c5_scale)
        A, E, Ai, Ee = nt.Matrix_ps(nn, is_complex=True), nt.Matrix_ps(nn, is_complex=True), nt.Matrix_ps(nn), nt.Matrix_ps(nn)
        env.fill(A, a)
        env.fill(E, e)
        nt.InverseSolvers.Invert(A, Ai, params())
        it_inv = nt.last_solve()["loop_counter"]
        nt.ExponentialSolvers.ComputeExponential(E, Ee, params())
        it_exp = nt.last_solve()["loop_counter"]
        op = O.SolverParameters(converge_diff=1e-5, threshold=thr)
        grid = O.Grid(*env.grid)
        ref_inv, info_inv = O.invert(O.PSMatrix.from_scipy(a, grid, is_complex=True), op)
        ref_exp, info_exp = O.compute_exponential(O.PSMatrix.from_scipy(e, grid, is_complex=True), op)
        e1, bad1 = block_parity(local_block_of(Ai), oracle_block_of(ref_inv, rank), thr)
        e2, bad2 = block_parity(local_block_of(Ee), oracle_block_of(ref_exp, rank), thr)
        worst = env.allmax([e1, e2])
        parity = {"ok": bool(it_inv == info_inv.iterations and it_exp == info_exp.iterations and worst[0] <= 1e-8 and worst[1] <= 1e-8),
                  "iterations_inverse": [it_inv, info_inv.iterations], "sigma_counter": [it_exp, info_exp.iterations],
                  "rel_fro_inverse_max_over_ranks": worst[0], "rel_fro_exponential_max_over_ranks": worst[1],
                  "pattern_differences_outside_threshold_band": int(bad1 + bad2),
                  "what": f"the same two solves on the N={nn} instance of the generator against the oracle on the benched grid: "
                          "identical iteration counts, every rank's block of both results within 1e-8 (relative Frobenius, "
                          "common pattern)"}
        del A, E, Ai, Ee
    peaks = None if args.no_peaks else env.measure_peaks()
    a, e = build(n)
    A, E, Ai, Ee = nt.Matrix_ps(n, is_complex=True), nt.Matrix_ps(n, is_complex=True), nt.Matrix_ps(n), nt.Matrix_ps(n)
    env.fill(A, a)
    env.fill(E, e)
    p = params()
    its = {}

    def step():
        nt.InverseSolvers.Invert(A, Ai, p)
        its["inverse"] = nt.last_solve()["loop_counter"]
        nt.ExponentialSolvers.ComputeExponential(E, Ee, p)
        its["sigma_counter"] = nt.last_solve()["loop_counter"]

    step()
    nt.set_flop_counting(True)
    nt.reset_counters()
    step()
    flops_local = nt.counters()["flops"]
    mults = nt.counters()["multiplies"]
    nt.set_flop_counting(False)
    warm = 2
    ms_total, clocks, prof = timed_steps(env, step, args.steps)
    cnt = nt.counters()
    alg_bytes = nt.algorithmic_bytes()
    ms_local = ms_total
    ms_total = env.allmax([ms_total])[0]
    flops_total = env.allsum([flops_local * args.steps])[0]
    result = {"ms_per_step": ms_total / args.steps, "value": flops_total / (ms_total * 1e-3) / 1e9, "clocks": clocks,
              "launches": cnt["launches"], "warmup": warm,
              "config_extra": {"iterations": its, "multiplies_per_step": mults, "sec_per_multiply": ms_total / args.steps / max(mults, 1) * 1e-3,
                               "nnz_per_row_inverse": Ai.GetSize() / n, "nnz_per_row_exponential": Ee.GetSize() / n}}
    e2e = None
    if not args.no_e2e:
        rows, cols, vals = A.get_arrays()
        r2, c2, v2 = E.get_arrays()
        pin_a = [env.pinned(x) for x in (rows, cols, vals)]
        pin_e = [env.pinned(x) for x in (r2, c2, v2)]
        Ah, Eh = nt.Matrix_ps(n, is_complex=True), nt.Matrix_ps(n, is_complex=True)

        def e2e_step():
            Ah.fill_from_arrays(*pin_a)
            Eh.fill_from_arrays(*pin_e)
            nt.InverseSolvers.Invert(Ah, Ai, p)
            nt.ExponentialSolvers.ComputeExponential(Eh, Ee, p)
            return Ai.get_arrays(), Ee.get_arrays()

        env.barrier()
        k = 1
        t0 = time.perf_counter()
        for _ in range(k):
            o1, o2 = e2e_step()
        env.barrier()
        dt = env.allmax([time.perf_counter() - t0])[0]
        e2e = {"value": env.allsum([flops_local * k])[0] / dt / 1e9, "unit": UNIT, "ms_per_step": dt / k * 1e3,
               "h2d_bytes_per_step": int(sum(x.nbytes for x in pin_a + pin_e)),
               "d2h_bytes_per_step": int(sum(x.nbytes for x in o1 + o2)),
               "what": "FillMatrixFromTripletList (complex, pinned host arrays) -> Invert_wrp + ComputeExponential_wrp -> GetMatrixTripletList"}
    roof = None
    if rank == 0:
        roof = roofline_of(env, prof, alg_bytes, flops_local * args.steps, ms_local, None,
                           "k_numeric_warp / k_numeric_cta<cplx> (scalar window kernels: complex128 has no tensor-core path yet)",
                           {"products": cnt["multiplies"]})
    return result, roof, e2e, parity


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    # everything any library writes to file descriptor 1 (NCCL's INFO log) goes to stderr; the JSON line to the real stdout
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    env = Env(args)
    run = {"c4": run_c4, "c1": run_c1, "c3": run_c3, "c5": run_c5}[args.config]
    result, roof, e2e, parity = run(env)
    rank, world = env.rank, env.world

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.environ["OMP_NUM_THREADS"] = os.environ.get("BENCH_CPU_THREADS", str(os.cpu_count() or 1))
        r = run_cpu(args, 2 if args.config in ("c4", "c1") else 1, 1 if args.config in ("c4", "c1") else 0, sample=True)
        cpu_baseline = {"value": r["gflops"], "unit": UNIT, "cores": r["cores"], "kind": "port", "ms_per_step": r["ms_per_step"],
                        "sample": r["what"] + "; CPU restatement of the reference algorithm (oracle/), windowed accumulator, "
                                              "OpenMP over columns"}
    if rank == 0:
        cfg = config_of(args, world)             # identical in both arms
        details = {"sec_per_iteration": result["ms_per_step"] * 1e-3,
                   "flops": "useful flops 2*sum_{(i,k) in A} nnz(B(k,:)) counted once on an identical warm-up step; the "
                            "counting (instrumentation) is off inside the timed region"}
        details.update(result.get("config_extra", {}))
        line = {
            "metric": METRICS[args.config], "value": result["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": result["warmup"], "ms_per_step": result["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c128" if args.config == "c5" else "f64", "data": "synthetic", "config": cfg,
            "clocks": result["clocks"], "e2e": e2e, "gpu_launches": result["launches"], "roofline": roof,
            "cpu_baseline": cpu_baseline, "parity_checked": parity, "details": details,
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
