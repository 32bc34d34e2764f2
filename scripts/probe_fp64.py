import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
for n in [4096, 8192]:
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): c = a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): c = a @ b
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"DGEMM n={n}: {ms:.2f} ms  {2*n**3/ms/1e9:.1f} TFLOP/s", flush=True)
# fp64 elementwise fma rate
x = torch.randn(1 << 26, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.draw", "--format=csv"], capture_output=True, text=True).stdout)
