import sys; sys.path.insert(0, '.')
import torch
from openmeasure_b200 import synth, engine as E
for (n_c, m, r) in [(400000, 256, 100), (400000, 257, 99), (400000, 255, 101)]:
    X = synth.snapshots(9, n_c, m, r)
    eng = E.Engine(X, 9, group=False)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        t = [ev() for _ in range(4)]
        eng.stats("std", 1, defer_row_means=True); t[0].record()
        G = eng.gram(); t[1].record()
        S, V = eng.eig_pod(G); t[2].record()
        eng.backproject(eng.pod_weights[:, :r].contiguous()); t[3].record()
        torch.cuda.synchronize()
    print(n_c, m, r, f"gram {t[0].elapsed_time(t[1]):.2f} ms  backproject {t[2].elapsed_time(t[3]):.2f} ms")
