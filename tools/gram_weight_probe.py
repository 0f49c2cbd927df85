import os, sys
sys.path.insert(0, '/root/repo' if os.path.exists('/root/repo/openmeasure_b200') else '.')
import torch
from openmeasure_b200 import synth, engine as E
n_c, m = int(sys.argv[1]), int(sys.argv[2])
F = 9 if m != 1024 else 8
X = synth.snapshots(F, n_c, m, 100)
eng = E.Engine(X, F, group=False)
eng.stats("std", 1)
ev = lambda: torch.cuda.Event(enable_timing=True)
for rep in range(3):
    a, b = ev(), ev()
    a.record(); eng.gram(centred=False, scaled=False); b.record(); torch.cuda.synchronize()
print(os.environ.get("OMB_GB_WDIAG"), f"{a.elapsed_time(b):.3f} ms")
