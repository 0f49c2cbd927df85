"""Synthetic snapshot generator on the GPU (DESIGN.md "Synthetic workload").

Deterministic, shardable by cells, bit-identical to the CPU twin used by the tests' oracle.  The
reference ships no data (its data/ROM/*.npy are git-LFS pointers), so bench.py and the parity
tests feed both arms from this generator.
"""
import ctypes as C
import math

import torch

from . import _lib
from .engine import _p, _stream, require_cuda

SEED = 1234


def tables(m, r, hard=False):
    K = r + 8
    rho = 0.5 if hard else math.pow(10.0, -2.0 / K)
    delta = math.pow(10.0, -3.0 / m)
    amp = [math.pow(rho, k) for k in range(K)]
    dec = [math.pow(delta, j) for j in range(m)]
    return K, amp, dec, 1e-3


def snapshots(F, n_cells, m, r, seed=SEED, cell0=0, ncell_loc=None, hard=False, device=None):
    """(F*ncell_loc, m) float64 CUDA tensor: cells [cell0, cell0+ncell_loc) of every feature."""
    require_cuda()
    if ncell_loc is None:
        ncell_loc = n_cells - cell0
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    K, amp, dec, eps = tables(m, r, hard)
    amp_d = torch.tensor(amp, dtype=torch.float64, device=dev)
    dec_d = torch.tensor(dec, dtype=torch.float64, device=dev)
    X = torch.empty(F * ncell_loc, m, dtype=torch.float64, device=dev)
    ws = torch.empty(int(_lib.load().omb_synth_ws_bytes(F, m, K)), dtype=torch.uint8, device=dev)
    _lib.call("omb_synth_fill", _p(X), F, n_cells, cell0, ncell_loc, m, K, C.c_uint64(seed),
              _p(amp_d), _p(dec_d), eps, _p(ws), _stream())
    return X
