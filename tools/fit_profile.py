"""Run only the fit-side kernels once (developer tool for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openmeasure_b200 import synth, engine as E
F, n_c, m, r = 9, 183620, 41, 40
Xd = synth.snapshots(F, n_c, m, r)
eng = E.Engine(Xd, F, group=False)
for _ in range(2):
    eng.stats("std", 1, defer_row_means=True)
    G = eng.gram()
    S, V = eng.eig_pod(G)
    eng.backproject((V[:, :r] / S[:r]).contiguous())
torch.cuda.synchronize()
print("ok", float(S[0]))
