"""Time the pivoted-QR placement alone (developer tool): python tools/qr_profile.py [n_c] [m] [r]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openmeasure_b200 import synth, engine as E

F = 9
n_c = int(sys.argv[1]) if len(sys.argv) > 1 else 183620
m = int(sys.argv[2]) if len(sys.argv) > 2 else 41
r = int(sys.argv[3]) if len(sys.argv) > 3 else 40
blocks = [int(b) for b in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 4, 8, 16]
Xd = synth.snapshots(F, n_c, m, r)
eng = E.Engine(Xd, F, group=False)
eng.stats("std", 1, defer_row_means=True)
S, V = eng.eig_pod(eng.gram())
eng.backproject((V[:, :r] / S[:r]).contiguous())
torch.cuda.synchronize()
ref = None
for b in blocks:
    for _ in range(2):
        piv, rd, gap = eng.qrcp(block=b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        piv, rd, gap = eng.qrcp(block=b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    p = piv.cpu()
    if ref is None: ref = p
    print(f"block={b:2d}  {ms:8.3f} ms  pivots_equal_block1={bool((p == ref).all())}  min gap {gap.min().item():.2e}")
