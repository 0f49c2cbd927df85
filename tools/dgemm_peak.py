import torch, time
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn_like(a)
for _ in range(2): c = a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): c = a @ b
e1.record(); torch.cuda.synchronize()
print("cuBLAS DGEMM 8192^3: %.2f TFLOP/s" % (3 * 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12))
