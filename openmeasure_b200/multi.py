"""One process driving several GPUs behind the UNCHANGED constructor: `SPR(X, n_features, xyz)`.

The row-sharded path is written for one rank per GPU (torch.distributed, `SPR.from_host / from_device /
from_npy`).  This module gives the same sharding to a plain `SPR(X_numpy, F, xyz)` call: with
`ROM.devices = 'all'` (or a list of device ids, or OMB_DEVICES=all | 0,1,2,3 in the environment) the
constructor returns an object that keeps one ordinary SPR per device -- each on its own cells of every
feature, driven by its own host thread -- and forwards every public method to all of them.  Results that
are replicated across ranks (Sigma_r, Ar, pivots, Theta, predictions) come from rank 0; row-distributed ones
(X_cnt, X_scl, X0, Ur, reconstructions) are assembled into the reference's global row order.

Cross-device exchanges are the same fixed-order combinations as everywhere else (comm.py), through a
thread communicator whose all-gather copies the peers' small tensors device to device; the pivot exchange of
the placement takes the host-gathered route (one small all-gather per pivot step) -- the kernel-driven NVLink
exchange needs symmetric memory, i.e. torch.distributed.  For throughput runs use one process per GPU.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import comm as _comm


def resolve_devices(setting):
    """[] / [id]: single-device path.  `setting`: None | 'all' | iterable of ids; OMB_DEVICES overrides None."""
    if setting is None:
        setting = os.environ.get("OMB_DEVICES")
    if setting is None or setting == "":
        return []
    if not torch.cuda.is_available():
        return []
    if isinstance(setting, str):
        if setting.strip().lower() == "all":
            return list(range(torch.cuda.device_count()))
        return [int(v) for v in setting.replace(",", " ").split()]
    return [int(v) for v in setting]


class DeviceThreadComm(_comm.ThreadComm):
    """ThreadComm whose ranks live on different devices: the gathered tensors are copied to the caller's device
    (a device-to-device copy is ordered after the producer's stream by the framework)."""

    def allgather(self, t):
        sh = self.shared
        sh.slots[self.rank] = t.contiguous().reshape(-1)
        sh.barrier.wait()
        out = torch.stack([s.to(t.device, non_blocking=True) for s in sh.slots])
        sh.barrier.wait()
        return out


class MultiDevice:
    """G per-device objects of class `cls` (ROM or SPR) + the threads that drive them."""

    def __init__(self, cls, X, n_features, xyz, devices):
        self.devices = list(devices)
        G = self.G = len(self.devices)
        self.F = F = int(n_features)
        n_c = X.shape[0] // F
        if n_c < G:
            raise ValueError("fewer cells per feature than devices")
        base, rem = divmod(n_c, G)
        self.cells = [base + (1 if g < rem else 0) for g in range(G)]
        self.offsets = [sum(self.cells[:g]) for g in range(G)]
        self.n_c, self.m = n_c, int(X.shape[1])
        comms = DeviceThreadComm.make(G)
        self.pool = ThreadPoolExecutor(max_workers=G, thread_name_prefix="omb-dev")
        X = np.ascontiguousarray(X, dtype=np.float64) if (X.dtype != np.float64 or not X.flags.c_contiguous) else X
        Xh = torch.from_numpy(X)
        xyz_a = None if xyz is None else np.asarray(xyz)

        def build(g):
            dev = torch.device("cuda", self.devices[g])
            torch.cuda.set_device(dev)
            c0, ncl = self.offsets[g], self.cells[g]
            Xd = torch.empty(F * ncl, self.m, dtype=torch.float64, device=dev)
            for f in range(F):                                  # this device's cells of every feature
                Xd[f * ncl:(f + 1) * ncl].copy_(Xh[f * n_c + c0:f * n_c + c0 + ncl], non_blocking=True)
            xl = None if xyz_a is None or xyz_a.ndim != 2 or xyz_a.shape[0] != n_c else xyz_a[c0:c0 + ncl]
            return cls.from_device(Xd, F, xl, comm=comms[g])

        self.subs = self.run(build, with_sub=False)

    # ------------------------------------------------------------------ plumbing
    def run(self, fn, with_sub=True):
        """fn(rank) or fn(sub, rank) on every device's thread, concurrently; returns the list of results."""
        def task(g):
            torch.cuda.set_device(self.devices[g])
            try:
                return fn(self.subs[g], g) if with_sub else fn(g)
            except BaseException:
                sub0 = self.subs[0] if with_sub else None
                try:                                            # wake the peers waiting on this rank
                    (sub0._eng.comm if sub0 is not None else None).shared.barrier.abort()
                except Exception:
                    pass
                raise
        futs = [self.pool.submit(task, g) for g in range(self.G)]
        out, err = [], None
        for f in futs:
            try:
                out.append(f.result())
            except BaseException as e:          # noqa: BLE001  (first real error wins over BrokenBarrierError)
                if err is None or isinstance(err, __import__("threading").BrokenBarrierError):
                    err = e
                out.append(None)
        if err is not None:
            for s in (self.subs if with_sub else []):
                try:
                    s._eng.comm.shared.barrier.reset()
                except Exception:
                    pass
            raise err
        return out

    def call(self, name, *args, **kw):
        return self.run(lambda s, g: getattr(s, name)(*args, **kw))

    def global_rows(self, g):
        """Global row index of every local row of device g (feature-major on both sides)."""
        ncl, c0 = self.cells[g], self.offsets[g]
        f = np.repeat(np.arange(self.F), ncl)
        return f * self.n_c + c0 + np.tile(np.arange(ncl), self.F)

    def assemble(self, parts):
        """Row-distributed per-device arrays (F * n_c_loc, ...) -> the reference's global (F * n_c, ...) array."""
        first = np.asarray(parts[0])
        out = np.empty((self.F * self.n_c,) + first.shape[1:], dtype=first.dtype)
        for g, p in enumerate(parts):
            out[self.global_rows(g)] = np.asarray(p)
        return out

    def shard(self, a, g):
        """The rows of a global row-indexed array that live on device g."""
        return np.ascontiguousarray(np.asarray(a)[self.global_rows(g)])

    def close(self):
        self.pool.shutdown(wait=False)
