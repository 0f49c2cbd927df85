import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from openmeasure_b200 import engine as E
from openmeasure_b200.sparse_sensing import SPR
rng = np.random.default_rng(0)
for (F, n_c, m, r, N) in [(2, 333, 128, 64, 34), (1, 2049, 130, 100, 130), (3, 900, 256, 100, 258), (2, 450, 300, 128, 16)]:
    X = rng.random((F * n_c, m)) + 1.0
    spr = SPR(X, F, np.zeros((n_c, 3)))
    spr.fit(select_modes='number', n_modes=r)
    C = spr.optimal_placement()
    spr.train(C)
    A = rng.standard_normal((N, r))
    out = spr.reconstruct(A, chunk_rows=256)
    y = np.zeros((r, 3)); y[:, 0] = X[C.pivots, 0]; y[:, 1] = 0.1; y[:, 2] = C.pivots // n_c
    a, s = spr.predict([y, y])
    print(F, n_c, m, r, N, float(out.sum()) != 0, a.shape)
print("done")
