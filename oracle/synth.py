"""CPU twin of the synthetic snapshot generator (DESIGN.md "Synthetic workload").

TEST INFRASTRUCTURE ONLY.  Not a reference function: the reference ships no data (its
data/ROM/*.npy are git-LFS pointers), so both arms are fed a deterministic, shardable,
bit-reproducible generator.  The CUDA twin lives in openmeasure_b200/csrc/synth.cu.
"""
import math

import numpy as np

from . import clib

SEED = 1234


def tables(m, r, hard=False):
    """(K, amp[K], dec[m], eps).  `hard` = steep spectrum (rho=0.5) that needs the accurate POD path."""
    K = r + 8
    rho = 0.5 if hard else math.pow(10.0, -2.0 / K)
    delta = math.pow(10.0, -3.0 / m)
    amp = np.array([math.pow(rho, k) for k in range(K)], dtype=np.float64)
    dec = np.array([math.pow(delta, j) for j in range(m)], dtype=np.float64)
    return K, amp, dec, 1e-3


def snapshots(F, n_cells, m, r, seed=SEED, cell0=0, ncell_loc=None, hard=False):
    """(F*ncell_loc, m) C-order rows [cell0, cell0+ncell_loc) of every feature block."""
    if ncell_loc is None:
        ncell_loc = n_cells - cell0
    K, amp, dec, eps = tables(m, r, hard)
    out = np.empty((F * ncell_loc, m), dtype=np.float64)
    clib.lib().omo_synth_fill(clib._dp(out), F, n_cells, cell0, ncell_loc, m, K, seed,
                              clib._dp(amp), clib._dp(dec), eps)
    return out
