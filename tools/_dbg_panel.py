import os, sys, ctypes as C
sys.path.insert(0, "/root/repo")
import torch
from openmeasure_b200 import synth, engine as E, _lib
F, n_c, m, r = 9, 183620, 41, 40
Xd = synth.snapshots(F, n_c, m, r)
eng = E.Engine(Xd, F, group=False)
eng.stats("std", 1, defer_row_means=True)
S, V = eng.eig_pod(eng.gram())
eng.backproject((V[:, :r] / S[:r]).contiguous())
for s in (6, 7, 8, 14, 39):
    eng.qrcp(s=s, block=8); torch.cuda.synchronize()
    buf = (C.c_longlong * 16)()
    L = C.CDLL("/root/repo/openmeasure_b200/libomb200.so")
    L.omb_debug_panel_clocks(buf)
    v = list(buf)[:7]
    print("last step", s - 1, "t =", (s - 1) % 8, [v[k + 1] - v[k] for k in range(6)], "total", v[6] - v[0])
