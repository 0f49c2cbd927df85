"""Run the many-snapshot FP64 contraction kernels once each, timed (developer tool, also the ncu target):
    python tools/big_profile.py [n_c m r N reps]
Gram (+ row means), back-projection and batched reconstruct at a config-3-like shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openmeasure_b200 import synth, engine as E

args = [int(v) for v in sys.argv[1:]]
n_c, m, r, N, reps = (args + [100000, 256, 100, 512, 3][len(args):])[:5]
F = 9 if m != 1024 else 8
Xd = synth.snapshots(F, n_c, m, r)
eng = E.Engine(Xd, F, group=False)
n = F * n_c
ev = lambda: torch.cuda.Event(enable_timing=True)
for rep in range(reps):
    t = [ev() for _ in range(5)]
    t[0].record()
    eng.stats("std", 1, defer_row_means=True)
    t[1].record()
    G = eng.gram()
    t[2].record()
    S, V = eng.eig_pod(G)
    t[3].record()
    eng.backproject(eng.pod_weights[:, :r].contiguous())
    t[4].record()
    torch.cuda.synchronize()
g = [ev(), ev(), ev()]
for rep in range(reps):
    g[0].record()
    eng.gram(centred=False, scaled=False)
    g[1].record()
    eng.cnt.zero_()
    eng.cnt
    g[2].record()
    torch.cuda.synchronize()
print(f"gram uncentred (no row means, no centring)  {g[0].elapsed_time(g[1]):9.3f} ms  {n * m * (m + 1.0) / g[0].elapsed_time(g[1]) / 1e9:8.2f} TFLOP/s algorithmic")
A = torch.rand(N, r, dtype=torch.float64, device="cuda")
rows = min(n, (1 << 30) // (8 * N) // 128 * 128)
out = torch.empty(rows, N, dtype=torch.float64, device="cuda")
u = [ev(), ev()]
for rep in range(reps):
    u[0].record()
    eng.reconstruct(A, row0=0, nrows=rows, out=out)
    u[1].record()
    torch.cuda.synchronize()
ms = [t[i].elapsed_time(t[i + 1]) for i in range(4)] + [u[0].elapsed_time(u[1])]
print(f"F={F} n_c={n_c} n={n} m={m} r={r} N={N}")
print(f"stats        {ms[0]:9.3f} ms  {2 * 8.0 * n * m / ms[0] / 1e6:8.0f} GB/s (2 passes)")
print(f"gram         {ms[1]:9.3f} ms  {n * m * (m + 1.0) / ms[1] / 1e9:8.2f} TFLOP/s algorithmic (n m (m+1))")
print(f"eigh         {ms[2]:9.3f} ms")
print(f"backproject  {ms[3]:9.3f} ms  {2.0 * n * m * r / ms[3] / 1e9:8.2f} TFLOP/s algorithmic (2 n m r)")
print(f"reconstruct  {ms[4]:9.3f} ms  {2.0 * rows * r * N / ms[4] / 1e9:8.2f} TFLOP/s ({rows} rows x {N} vectors, r={r})")
