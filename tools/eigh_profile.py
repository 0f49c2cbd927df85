import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openmeasure_b200 import synth, engine as E, _lib
F, n_c, m, r = 9, 183620, 41, 40
Xd = synth.snapshots(F, n_c, m, r)
eng = E.Engine(Xd, F, group=False); eng.stats("std", 1, defer_row_means=True); G = eng.gram()
w = torch.empty(m, dtype=torch.float64, device="cuda"); V = torch.empty(m, m, dtype=torch.float64, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
for _ in range(3): _lib.call("omb_eigh_jacobi", p(G), m, p(w), p(V), p(info), None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): _lib.call("omb_eigh_jacobi", p(G), m, p(w), p(V), p(info), None)
e1.record(); torch.cuda.synchronize()
print("jacobi ms", e0.elapsed_time(e1) / 10, "sweeps", int(info.item()))
wl = torch.linalg.eigvalsh(G).flip(0)
print("max rel err top 40:", float(((w - wl).abs() / wl)[:40].max()), "w[-1]/w[0]", float(w[-1] / w[0]), float(wl[-1] / wl[0]))
e0.record()
for _ in range(10): torch.linalg.eigh(G)
e1.record(); torch.cuda.synchronize()
print("torch eigh ms", e0.elapsed_time(e1) / 10)
