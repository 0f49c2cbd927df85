"""One process, all visible GPUs, the reference's own constructor (developer tool, run under gpurun --gpus N):
OMB_DEVICES=all python tools/multi_device_check.py [n_c m r]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OMB_DEVICES", "all")
import numpy as np
import torch
from openmeasure_b200 import synth
from openmeasure_b200 import sparse_sensing as sps

n_c, m, r = (int(v) for v in (sys.argv[1:4] + ["200000", "64", "30"][len(sys.argv) - 1:]))
F = 9
X = synth.snapshots(F, n_c, m, r).cpu().numpy()
xyz = np.zeros((n_c, 3))
t0 = time.perf_counter()
spr = sps.SPR(X, F, xyz)
spr.fit(select_modes="number", n_modes=r)
C = spr.optimal_placement()
t1 = time.perf_counter()
sps.ROM.devices = [0]
one = sps.SPR(X, F, xyz)
one.fit(select_modes="number", n_modes=r)
C1 = one.optimal_placement()
t2 = time.perf_counter()
same = bool(np.array_equal(C.pivots, C1.pivots))
ds = float(np.max(np.abs(spr.Sigma_r - one.Sigma_r) / one.Sigma_r))
print(f"devices={getattr(spr, '_md', None) and spr._md.devices} rows={F*n_c} m={m} r={r}: pivots_identical={same} max_rel_dsigma={ds:.2e} "
      f"multi {t1-t0:.2f} s, single {t2-t1:.2f} s -> {'OK' if same and ds < 1e-12 else 'FAIL'}")
sys.exit(0 if same and ds < 1e-12 else 1)
