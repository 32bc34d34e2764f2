"""developer probe: where the end-to-end step (host triplets in, host triplets out) spends its time"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ntpoly_b200.api as nt
from ntpoly_b200.workloads import banded_sign_input
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
thr = 1e-6
nt.ConstructGlobalProcessGrid(1, 1, 1)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); nt.set_stream(s.cuda_stream)
M = nt.Matrix_ps(n); M.fill_from_scipy(banded_sign_input(n))
I = nt.Matrix_ps(n); I.FillIdentity()
emin, emax = nt.EigenBounds.GershgorinBounds(M)
X = nt.Matrix_ps(M); X.Scale(1.0 / abs(emax))
T1, T2, W = nt.Matrix_ps(n), nt.Matrix_ps(n), nt.Matrix_ps(n)
for k in range(2):
    nt.sign_iteration(X, I, T1, T2, 1.2, thr)
rows, cols, vals = X.get_arrays()
pin = [torch.from_numpy(a).pin_memory().numpy() for a in (rows, cols, vals)]
cap = int(len(rows) * 1.5) + 1024
pout = [(torch.empty(cap, dtype=torch.int32).pin_memory().numpy(), torch.empty(cap, dtype=torch.int32).pin_memory().numpy(),
         torch.empty(cap, dtype=torch.float64).pin_memory().numpy()) for _ in range(2)]
Xh = nt.Matrix_ps(n)
def T(name, fn):
    nt.synchronize(); t0 = time.perf_counter(); r = fn(); nt.synchronize(); nt.egress_wait()
    print(f"  {name:34s} {(time.perf_counter()-t0)*1e3:9.3f} ms", flush=True); return r
print("nnz", len(rows), "bytes in", sum(a.nbytes for a in pin))
# raw copies
d = [torch.empty(len(a), dtype=torch.from_numpy(a).dtype, device="cuda") for a in pin]
for it in range(2):
    T("raw H2D 3 arrays (torch)", lambda: [d[i].copy_(torch.from_numpy(pin[i]), non_blocking=True) for i in range(3)])
    T("raw D2H 3 arrays (torch)", lambda: [torch.from_numpy(pout[0][i][:len(rows)]).copy_(d[i], non_blocking=True) for i in range(3)])
s2 = torch.cuda.Stream()
def both():
    for i in range(3):
        d[i].copy_(torch.from_numpy(pin[i]), non_blocking=True)
    with torch.cuda.stream(s2):
        for i in range(3):
            torch.from_numpy(pout[1][i][:len(rows)]).copy_(d[i], non_blocking=True)
    s2.synchronize()
T("raw H2D + D2H concurrently", both)
T("raw H2D + D2H concurrently", both)
for it in range(3):
    print("serialised step", it)
    T("fill_from_arrays (H2D + ingest)", lambda: Xh.fill_from_arrays(*pin))
    T("sign_step (forms from CSC + 2 products)", lambda: nt.sign_step(Xh, I, T1, W, 1.2, thr))
    T("get_arrays_async + wait", lambda: W.get_arrays_async(pout[it % 2]))
os.environ["NTB_TILE_TIMING"] = "1"
