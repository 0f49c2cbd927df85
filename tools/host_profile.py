"""Where does the host time of one fit + optimal_placement step go?  (developer tool)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openmeasure_b200 import synth, engine as E
from openmeasure_b200.sparse_sensing import SPR

F, n_c, m, r = 9, 183620, 41, 40
Xd = synth.snapshots(F, n_c, m, r)
torch.cuda.synchronize()


def seg(name, fn, sync=True):
    t0 = time.perf_counter()
    out = fn()
    t1 = time.perf_counter()
    if sync:
        torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"  {name:14s} issue {1e3*(t1-t0):7.3f} ms   +sync {1e3*(t2-t0):7.3f} ms")
    return out


for it in range(4):
    print("iter", it)
    eng = seg("Engine()", lambda: E.Engine(Xd, F, group=False))
    seg("stats", lambda: eng.stats("std", 1, defer_row_means=True))
    G = seg("gram", lambda: eng.gram())
    S, V = seg("eigh", lambda: eng.eig_pod(G))
    W = seg("W", lambda: (V[:, :r] / S[:r]).contiguous())
    seg("backproject", lambda: eng.backproject(W))
    seg("qrcp b8", lambda: eng.qrcp(block=8))
    seg("qrcp b1", lambda: eng.qrcp(block=1))
    seg("qrcp b4", lambda: eng.qrcp(block=4))
    seg("qrcp b16", lambda: eng.qrcp(block=16))
    t0 = time.perf_counter()
    spr = SPR.from_device(Xd, F, group=False)
    spr.fit(select_modes="number", n_modes=r)
    t1 = time.perf_counter()
    C = spr.optimal_placement()
    t2 = time.perf_counter()
    print(f"  API fit {1e3*(t1-t0):.3f} ms, optimal_placement {1e3*(t2-t1):.3f} ms")
