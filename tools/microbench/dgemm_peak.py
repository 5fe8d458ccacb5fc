"""Measure cuBLAS DGEMM throughput on this GPU (the FP64 roofline denominator)."""
import json, sys, torch
def bench(m, n, k, reps=5):
    a = torch.randn(m, k, dtype=torch.float64, device="cuda")
    b = torch.randn(k, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * m * n * k / best / 1e9
out = {"gpu": torch.cuda.get_device_name(0)}
for shp in [(8192, 8192, 8192), (2048, 65536, 2048), (4096, 32768, 4096)]:
    out["dgemm_%dx%dx%d_tflops" % shp] = bench(*shp)
# sustained
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(60): torch.matmul(a, b)
e1.record(); torch.cuda.synchronize()
out["dgemm_8192_sustained_tflops"] = 60 * 2.0 * 8192**3 / e0.elapsed_time(e1) / 1e9
print(json.dumps(out))
