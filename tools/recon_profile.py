"""Batched reconstruct throughput (developer tool): python tools/recon_profile.py n_c r N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openmeasure_b200 import engine as E
F = 9
n_c, r, N = (int(v) for v in sys.argv[1:4])
n = F * n_c
X = torch.rand(n, 8, dtype=torch.float64, device="cuda")
eng = E.Engine(X, F, group=False)
eng.stats("std", 1)
eng.Ut = torch.rand(eng.ntiles, r, 128, dtype=torch.float64, device="cuda"); eng.r = r
A = torch.rand(N, r, dtype=torch.float64, device="cuda")
out = torch.empty(n, N, dtype=torch.float64, device="cuda")
for _ in range(2): eng.reconstruct(A, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): eng.reconstruct(A, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"n={n} r={r} N={N}: {ms:.3f} ms  {N/ms*1e3:.0f} recon/s  {2.0*n*r*N/ms/1e9:.2f} TFLOP/s  write {8.0*n*N/ms/1e6:.0f} GB/s")
