"""Developer timing probe (not the contract bench): X*X on banded matrices."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import ntpoly_b200.api as nt
from util import banded

nt.ConstructGlobalProcessGrid(1, 1, 1)
nt.set_stream(torch.cuda.current_stream().cuda_stream)
for n in [8192, 65536, 262144]:
    a = banded(n)
    A = nt.Matrix_ps(n); A.fill_from_scipy(a)
    C = nt.Matrix_ps(n)
    for thr in [1e-8]:
        for _ in range(3):
            C.Gemm(A, A, None, threshold=thr)
        nt.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            C.Gemm(A, A, None, threshold=thr)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        c = nt.counters()
        fl = c["flops"] / reps
        by = A.algorithmic_bytes() + C.algorithmic_bytes()
        print(f"n={n} thr={thr} ms={ms:.3f} GFLOP/s={fl/ms/1e6:.1f} algGB/s={by/ms/1e6:.1f} nnzA={A.GetSize()} nnzC={C.GetSize()} launches/step={c['launches']/reps}", flush=True)
    # helper timings
    B = nt.Matrix_ps(A)
    for name, fn in [("increment", lambda: B.Increment(C, -1.0, 1e-8)), ("dot", lambda: A.Dot(C)), ("trace", lambda: A.Trace()), ("norm", lambda: C.Norm()), ("copy", lambda: nt.Matrix_ps(C)), ("scale", lambda: C.Scale(1.0))]:
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize()
        print(f"   {name}: {(time.perf_counter()-t0)/5*1e3:.3f} ms", flush=True)
