#!/usr/bin/env python
"""Benchmark of the NTPoly hot path on B200 (contract: one JSON line on rank 0).

Workload (BASELINE.json configs[3], the configuration the metric is quoted on; it fits one GPU):
one Newton-Schulz sign-function iteration (reference SignSolversModule.F90:207-240 — two
thresholded distributed multiplies, two sparse adds, one 1-norm; the result replaces X by exchanging it with the
work matrix) on the synthetic banded matrix
N=262144 (half-bandwidth 82, 165 nnz/row) shifted to straddle zero, threshold 1e-6, applied to
the fixed iterate X_3 so that every step does identical work.  N GPUs share the SAME matrix on
NTPoly's process grid (1x2x1, 1x4x1, 1x8x1: column split): strong scaling.

  python bench.py --gpus 1 --steps 10 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # CPU restatement of the reference algorithm (oracle/)
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

METRIC = "spgemm_useful_gflops_per_sign_iteration"
UNIT = "GFLOP/s"
GRIDS = {1: (1, 1, 1), 2: (1, 2, 1), 4: (1, 4, 1), 8: (1, 8, 1)}   # column split: tile halo exchange, no panel gather
ALPHA_MAX = 1.69770248526
FP64_PEAK_TFLOPS = 37.2          # DMMA.8x8x4 issue peak measured on this pool's B200 (scripts/micro/dmma_shapes.cu)
TRAFFIC_PER_LAUNCH = 1.531e9      # dram read+write bytes per numeric launch, mean of the step's two products, ncu --set full (profiles/r01d_numeric.keys.txt)


def workload_name(n, thr, iterate):
    return (f"newton-schulz sign iteration (SignFunction driver loop body: 2 multiplies, identity shift, convergence "
            f"norm), banded N={n} half-bandwidth 82, thr={thr:g}, iterate X_{iterate}")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=262144)
    ap.add_argument("--threshold", type=float, default=1e-6)
    ap.add_argument("--iterate", type=int, default=3, help="which Newton-Schulz iterate the step is applied to")
    ap.add_argument("--cpu-n", type=int, default=32768, help="matrix size of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


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
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
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
# reference arm / cpu baseline: the oracle (CPU restatement of the reference algorithm)
# --------------------------------------------------------------------------------------
def cpu_sign_iteration_factory(n, threshold, iterate):
    from oracle import oracle as O
    from ntpoly_b200.workloads import banded_sign_input
    O.build()
    m = banded_sign_input(n)
    M = O.PSMatrix.from_scipy(m)
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


def run_cpu(n, threshold, iterate, steps, warmup):
    O, step = cpu_sign_iteration_factory(n, threshold, iterate)
    for _ in range(warmup):
        step()
    st = O.MultiplyStats()
    t0 = time.perf_counter()
    for _ in range(steps):
        step(st)
    dt = time.perf_counter() - t0
    return {"gflops": st.flops / dt / 1e9, "ms_per_step": dt / steps * 1e3, "cores": O.lib().orc_max_threads(),
            "flops_per_step": st.flops / steps}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # all host cores (torchrun exports OMP_NUM_THREADS=1 to its workers; the OpenMP runtime has not been loaded yet)
    os.environ["OMP_NUM_THREADS"] = os.environ.get("BENCH_CPU_THREADS", str(os.cpu_count() or 1))
    n = args.cpu_n
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 3))
    r = run_cpu(n, args.threshold, args.iterate, steps, warmup)
    sample = (f"one sign iteration on the banded N={n} matrix (1/{args.n // n} of the N={args.n} columns; per-column "
              f"work of a banded matrix is size independent), {steps} steps, windowed accumulator, OpenMP over rows")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["gflops"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n, args.threshold, args.iterate),
                   "sampled_n": n},
        "cpu_baseline": {"value": r["gflops"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["gflops"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import ntpoly_b200.api as nt
    from ntpoly_b200.workloads import banded_sign_input

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        if "BENCH_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["BENCH_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)           # NCCL prints its version banner to stdout at VERSION and above
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nt.init_world_from_torch()
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    R, C, S = GRIDS[world]
    nt.ConstructGlobalProcessGrid(R, C, S)
    bench_stream = torch.cuda.Stream()
    torch.cuda.set_stream(bench_stream)
    if not os.environ.get('BENCH_OWN_STREAM'):
        nt.set_stream(bench_stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- input: every rank generates the same matrix and contributes a disjoint share
    n, thr = args.n, args.threshold
    m = banded_sign_input(n).tocoo()
    sel = slice(rank, None, world)
    M = nt.Matrix_ps(n)
    M.fill_from_arrays(m.row[sel] + 1, m.col[sel] + 1, m.data[sel])
    del m
    I = nt.Matrix_ps(n)
    I.FillIdentity()
    e_min, e_max = nt.EigenBounds.GershgorinBounds(M)
    alphas = alpha_sequence(e_min, e_max, args.iterate)
    X = nt.Matrix_ps(M)
    X.Scale(1.0 / abs(e_max))
    pool = nt.PMatrixMemoryPool(X)
    T1, T2 = nt.Matrix_ps(n), nt.Matrix_ps(n)

    W = nt.Matrix_ps(n)

    def step(Xin, ak):
        """The loop body of SignFunction (SignSolversModule.F90:207-240) exactly as the SignFunction_wrp driver of this
        library runs it (csrc/solvers.cu: sign_step; the driver then exchanges X and the work matrix, a pointer swap),
        through the C ABI: two thresholded multiplies (the first with the 3I shift of the following IncrementMatrix
        fused into its emit pass and its result handed to the second as a tile form), the convergence norm
        ||X_{k+1} - X_k||. W receives X_{k+1}; X_k is left untouched so that every step does identical work."""
        return nt.sign_step(Xin, I, T1, W, ak, thr, pool)

    for k in range(args.iterate - 1):         # advance to the iterate the step is quoted on
        nt.sign_iteration(X, I, T1, T2, alphas[k], thr, pool)
    ak = alphas[args.iterate - 1]

    # useful flops of one step: counted ONCE on a warm-up step (the timed steps repeat exactly this work); the
    # counting itself is instrumentation (one extra sweep + read-back per product) and is off in the timed region
    step(X, ak)
    nt.set_flop_counting(True)
    nt.reset_counters()
    step(X, ak)
    flops_per_step_local = nt.counters()["flops"]
    nt.set_flop_counting(False)
    for _ in range(max(args.warmup, 3)):
        step(X, ak)

    # ---- timed region: device events on the launching stream, max over ranks
    sampler = ClockSampler(local_rank)
    barrier()
    nt.reset_counters()
    nt.profile_enable(True)
    if not os.environ.get('BENCH_NO_SAMPLER'):
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(X, ak)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    prof = nt.profile_read()
    nt.profile_enable(False)
    cnt = nt.counters()
    builds_timed = nt.tile_builds()
    alg_bytes = nt.algorithmic_bytes()
    flops_local = flops_per_step_local * args.steps
    dfr = nt.deferred_counters()
    t = torch.tensor([ms_total, flops_local, prof["numeric_ms"], alg_bytes], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total = float(tmax[0])
        flops_total = float(tsum[1])
    else:
        flops_total = flops_local
    ms_per_step = ms_total / args.steps
    value = flops_total / (ms_total * 1e-3) / 1e9

    # ---- end to end: host triplets in, host triplets out, every step (rank-local shares). Every step copies its input
    # from pinned host memory (ntb_StageArrays + ntb_FillMatrixFromStaged_ps) and its result X_{k+1} back to pinned host
    # memory (ntb_GetMatrixArraysAsync_ps). The three stages of consecutive steps are pipelined: while step i computes,
    # the input of step i+1 comes in and the result of step i-1 goes out (PCIe is full duplex); the timed region ends
    # when the last result has landed on the host.
    e2e = None
    if not args.no_e2e:
        rows, cols, vals = X.get_arrays()
        pin = [torch.from_numpy(a).pin_memory().numpy() for a in (rows, cols, vals)]
        Xh = nt.Matrix_ps(n)
        cap = int(len(rows) * 1.5) + 1024                 # pinned landing buffers for the step's result, two sets
        pout = [(torch.empty(cap, dtype=torch.int32).pin_memory().numpy(),
                 torch.empty(cap, dtype=torch.int32).pin_memory().numpy(),
                 torch.empty(cap, dtype=torch.float64).pin_memory().numpy()) for _ in range(2)]
        e2e_steps = min(max(args.steps, 16), 24)         # a pipeline: fill and drain are inside the timed region

        def e2e_loop(count):
            d2h = 0
            st = nt.stage_arrays(*pin)                    # H2D of step 0's input (copy stream)
            for i in range(count):
                # H2D of step i+1's input: enqueued now, it runs while this step computes and while result i-1 goes out
                nxt = nt.stage_arrays(*pin) if i + 1 < count else None
                Xh.fill_from_staged(st)                   # this step's input becomes the matrix (waits for its copies)
                step(Xh, ak)
                nt.egress_wait()                          # result i-1 has landed
                out = W.get_arrays_async(pout[i % 2])     # D2H of the step's result X_{k+1} (second copy stream)
                d2h = sum(a.nbytes for a in out) + 8      # + the norm scalar
                st = nxt
            nt.egress_wait()
            return d2h

        e2e_loop(4)                                       # warm: the arena reaches its steady state (no cudaMalloc)
        barrier()
        nt.reset_counters()
        t0 = time.perf_counter()
        d2h = e2e_loop(e2e_steps)
        barrier()
        dt = time.perf_counter() - t0
        f = torch.tensor([dt, flops_per_step_local * e2e_steps], dtype=torch.float64, device="cuda")
        if world > 1:
            fm = f.clone(); dist.all_reduce(fm, op=dist.ReduceOp.MAX)
            fs = f.clone(); dist.all_reduce(fs, op=dist.ReduceOp.SUM)
            dt, fl = float(fm[0]), float(fs[1])
        else:
            dt, fl = float(f[0]), float(f[1])
        e2e = {"value": fl / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(sum(a.nbytes for a in pin)),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt / e2e_steps * 1e3, "steps": e2e_steps,
               "overlap": "three streams: every step copies its own input in (copy stream 1) and its own result out "
                          "(copy stream 2); the input copy of step i+1 and the result copy of step i-1 run while step i "
                          "computes (PCIe is full duplex)",
               "api": "C ABI with pinned host arrays: ntb_StageArrays + ntb_FillMatrixFromStaged_ps (triplets in), "
                      "ntb_SignStep (the SignFunction driver's loop body), ntb_GetMatrixArraysAsync_ps + ntb_EgressWait "
                      "(triplets out)",
               "sorted_ingests": nt.sorted_ingests()}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (numeric SpGEMM), rank 0's launches
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (prof["numeric_ms"] * 1e-3) / 1e9 if prof["numeric_ms"] > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": TRAFFIC_PER_LAUNCH, "kernel": "k_tile_numeric9 (numeric SpGEMM incl. threshold and tile-form output, one launch per product)",
                "launches_timed": prof["products"], "peak_source": peak_src,
                "numeric_share_of_step": prof["numeric_ms"] / (ms_total if world == 1 else float(t[0])),
                "fp64_tflops_useful": flops_local / (prof["numeric_ms"] * 1e-3) / 1e12 if prof["numeric_ms"] > 0 else 0.0,
                "fp64_tensor_peak_tflops": FP64_PEAK_TFLOPS,
                "fp64_frac": (flops_local / (prof["numeric_ms"] * 1e-3) / 1e12 / FP64_PEAK_TFLOPS) if prof["numeric_ms"] > 0 else 0.0,
                "gustavson_upper_bytes_per_launch": (flops_local / max(prof["products"], 1)) / 2.0 * 12.0,
                "tile_form_builds_in_timed_region": builds_timed,
                "deferred_csc_products_in_timed_region": dfr["products"],
                "deferred_csc_materialized_in_timed_region": dfr["materialized"],
                "note": "arithmetic intensity of this product (~10 flop/B) is above the FP64 machine balance "
                        "(37.2 TF/s DMMA measured, scripts/micro / HBM peak): the kernel is bound by the FP64 tensor "
                        "pipe, fp64_frac is its share of that peak; see DESIGN.md"}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        r = run_cpu(args.cpu_n, thr, args.iterate, 2, 1)
        cpu_baseline = {"value": r["gflops"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                        "ms_per_step": r["ms_per_step"],
                        "sample": f"same sign iteration on the banded N={args.cpu_n} matrix (1/{n // args.cpu_n} of the "
                                  f"columns), 2 steps; CPU restatement of the reference algorithm (oracle/), "
                                  f"windowed accumulator, OpenMP over rows"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, thr, args.iterate),
                   "grid": f"{R}x{C}x{S}", "l2": "inputs exceed L2 (operands > 500 MB vs 126 MB L2)",
                   "sec_per_iteration": ms_per_step * 1e-3,
                   "flops": "useful flops of the step counted once on an identical warm-up step; flop counting "
                            "(instrumentation) is off inside the timed region"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": cnt["launches"], "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
