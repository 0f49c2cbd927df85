"""Device engine: owns the HBM-resident state of one ROM/SPR object and drives libomb200.so.

PyTorch is plumbing only (device allocations, the current CUDA stream, the m x m eigensolve /
s x r pseudo-inverse on device, torch.distributed for the tiny cross-rank exchanges); every pass
over the n-row data is a hand-written sm_100a kernel called through the C ABI (include/omb200.h).

Data layout in HBM (FP64 throughout):
    X      (F * n_c_loc, m) C-order snapshot shard, feature-major, exactly the reference's layout
    cnt    (F * n_c_loc,)   centring value per row          (reference X_cnt[:, 0])
    scl    (F,)             scale per feature block         (reference X_scl[f * n_points, 0])
    Ut     (ntiles, r, 128) tiled mode-major basis, Ut[i // 128, q, i % 128] = U_r[i, q]
    work   (ntiles, r, 128) trailing matrix of the pivoted QR
    X0c    like X           X - cnt, m > 64 only: the centred copy the FP64 tensor-core passes read (a DADD in their
                            inner loop shares the pipe with DMMA); from the Gram pass to the back-projection, pooled
With torch.distributed initialised, every rank holds the cells [c0, c0 + n_c_loc) of EVERY
feature; only F*4 statistics, the m x m Gram and (per pivot step) one small record cross NVLink.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import comm as _comm

SCALE_CODES = {"std": 0, "none": 1, "pareto": 2, "vast": 3, "range": 4, "level": 5, "max": 6,
               "variance": 7, "poisson": 8, "l2-norm": 9}
NEEDS_SQDEV = {"std", "pareto", "vast", "variance", "l2-norm"}
EPS = float(np.finfo(np.float64).eps) / 2  # unit roundoff 2^-53


def require_cuda():
    _lib.load()
    if not torch.cuda.is_available():
        raise _lib.OmbError("openmeasure_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


TB = 128    # candidates per basis tile (OMB_TB in csrc/common.cuh)

# the framework's dense linear algebra initialises its backend lazily and not thread-safely; the multi-device mode
# (multi.py) and the thread-emulated ranks of the tests call it from several threads
import threading as _threading
LINALG_LOCK = _threading.Lock()

# One spare buffer per device for the centred copy of X that the many-snapshot tensor-core passes read: it is
# as large as X itself (33 GB at config 3, 68.7 GB per GPU at config 5), lives from the Gram pass to the
# back-projection, and is handed from fit to fit instead of going through the caching allocator each time
# (splitting and re-growing a block of that size cost 100+ ms per fit).  release_scratch() returns it.
_SCRATCH_POOL = {}


def _scratch_take(dev, nbytes):
    buf = _SCRATCH_POOL.get(dev)
    if buf is not None and buf.numel() >= nbytes:
        del _SCRATCH_POOL[dev]
        return buf
    return None


def _scratch_give(dev, buf):
    old = _SCRATCH_POOL.get(dev)
    if old is None or old.numel() < buf.numel():
        _SCRATCH_POOL[dev] = buf


def release_scratch():
    """Return the pooled centred-copy buffers to the allocator."""
    _SCRATCH_POOL.clear()



def tiles_for(n):
    return (int(n) + TB - 1) // TB


def _staged(name):
    """Record CUDA events around a stage when Engine.trace is on (bench.py: per-stage device times of the TIMED
    steps themselves, no host synchronisation added)."""
    def deco(fn):
        def wrapped(self, *a, **k):
            if not Engine.trace:
                return fn(self, *a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(self, *a, **k)
            e1.record()
            self.marks.append((name, e0, e1))
            return out
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


class Engine:
    trace = False

    def __init__(self, X_dev, n_features, group=None, comm=None):
        require_cuda()
        assert X_dev.is_cuda and X_dev.dtype == torch.float64 and X_dev.dim() == 2
        if not X_dev.is_contiguous():
            X_dev = X_dev.contiguous()
        self.X = X_dev
        self.dev = X_dev.device
        self.F = int(n_features)
        self.n_loc, self.m = (int(v) for v in X_dev.shape)
        assert self.n_loc % self.F == 0
        self.n_c_loc = self.n_loc // self.F
        if comm is None:
            comm = _comm.SingleComm()
            if group is not False and torch.distributed.is_available() and torch.distributed.is_initialized() \
                    and torch.distributed.get_world_size(group) > 1:
                comm = _comm.TorchDistComm(group)
        self.comm = comm
        self.world, self.rank = comm.world, comm.rank
        if self.world > 1:
            mine = torch.tensor([self.n_c_loc], dtype=torch.int64, device=self.dev)
            cells = [int(c) for c in comm.allgather(mine).view(-1).cpu()]
        else:
            cells = [self.n_c_loc]
        self.layout = _comm.ShardLayout(self.F, cells, self.rank)
        self.n_c = self.layout.n_c                      # global cells per feature
        self.cell0 = self.layout.cell0
        self._cnt = None
        self._cnt_pending = False                        # row means still to be produced by the Gram pass
        self.scl = None
        self.Ut = None
        self.vn = None
        self.r = None
        self.ntiles = tiles_for(self.n_loc)
        self.marks = []
        self._X0c = None                                 # centred copy of X for the many-snapshot tensor-core passes

    # ------------------------------------------------------------------------------------ centred copy
    def _centred_copy_wanted(self):
        """m > 64 (the DMMA-bound kernels), even m, aligned X, and room for a second copy of X next to the basis
        and the placement workspace.  OMB_CENTRED_COPY=0 forces the in-kernel centring path."""
        import os
        if self.m <= 64 or os.environ.get("OMB_CENTRED_COPY", "1") == "0":
            return False
        mp = self.m + (self.m & 1)                       # an odd snapshot count gets an even row pitch (a zero snapshot)
        pooled = _SCRATCH_POOL.get(self.dev)
        if pooled is not None and pooled.numel() >= 8 * self.n_loc * mp:
            return True
        free, _ = torch.cuda.mem_get_info(self.dev)
        free += torch.cuda.memory_reserved(self.dev) - torch.cuda.memory_allocated(self.dev)
        need = 8 * self.n_loc * mp
        return free - need >= 16 * self.n_loc * min(self.m, 128) + (1 << 30)

    def _new_centred(self):
        """Uninitialised buffer shaped like X from the scratch pool (or the allocator); None when out of memory."""
        mp = self.m + (self.m & 1)
        nbytes = 8 * self.n_loc * mp
        buf = _scratch_take(self.dev, nbytes)
        if buf is None:
            try:
                buf = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            except torch.OutOfMemoryError:
                return None
        self._X0c_buf = buf
        return buf[:nbytes].view(torch.float64).view(self.n_loc, mp)

    def _drop_centred(self):
        """The centred copy's last reader has been queued on the current stream: hand the buffer on."""
        buf = getattr(self, "_X0c_buf", None)
        if buf is not None:
            _scratch_give(self.dev, buf)
        self._X0c, self._X0c_buf = None, None

    def _centre(self, X, cnt, compute_means, out=None):
        """X0c = X - cnt (row means computed in the same pass when asked); None when there is no room."""
        if out is None:
            out = self._new_centred()
            if out is None:
                return None
        _lib.call("omb_center_rows_padded", _p(X), int(X.shape[0]), self.m, int(out.shape[1]), 1 if compute_means else 0,
                  _p(cnt), _p(out), _stream())
        return out

    def _gram_of_copy(self, X0c, F, ncl, Gf_out):
        """Per-feature Grams of a centred copy (row pitch mp >= m) into Gf_out (F * m * m)."""
        m, mp = self.m, int(X0c.shape[1])
        ws = _ws(_lib.load().omb_gram_ws_bytes(F, ncl, mp), self.dev)
        if mp == m:
            _lib.call("omb_gram", _p(X0c), F, ncl, m, None, _p(Gf_out), _p(ws), _stream())
            return
        Gp = torch.empty(F * mp * mp, dtype=torch.float64, device=self.dev)
        _lib.call("omb_gram", _p(X0c), F, ncl, mp, None, _p(Gp), _p(ws), _stream())
        Gf_out.view(F, m, m).copy_(Gp.view(F, mp, mp)[:, :m, :m])      # drop the zero snapshot

    def check_p2p(self):
        """Raise if a peer never answered during a peer-memory exchange since the last check (the pivot
        exchange of the placement, the small collectives of the fit).  Synchronises with the device."""
        self.comm.check()
        st = getattr(self, "_p2p_err", None)
        if st is not None:
            word = st[0][st[1]:st[1] + 1].view(torch.int64)
            if word.item() != 0:
                word.zero_()
                raise _lib.OmbError("multi-rank placement: a peer rank did not publish its record (10 s timeout)")

    @property
    def cnt(self):
        """Centring value per local row; flushes row means deferred to the Gram pass."""
        if self._cnt_pending:
            self._wait_arrival()
            _lib.call("omb_row_means", _p(self.X), self.n_loc, self.m, _p(self._cnt), _stream())
            self._cnt_pending = False
        return self._cnt

    @cnt.setter
    def cnt(self, v):
        self._cnt, self._cnt_pending = v, False

    def _wait_arrival(self):
        """X may still be in flight from the host (ROM._engine uploads it block by block): make the
        current stream wait for all of it.  stats() consumes the blocks one by one instead."""
        arrival = getattr(self, "_arrival", None)
        if arrival is not None:
            for ev in arrival:
                torch.cuda.current_stream().wait_event(ev)
            self._arrival = None

    def _shard_args(self):
        return (self.n_c_loc, self.n_c, self.cell0, self.rank, self.world)

    # ------------------------------------------------------------------------------------ K1
    @_staged("stats")
    def stats(self, scale_type="std", axis_cnt=1, defer_row_means=False):
        """Centring vector and per-feature scale (sparse_sensing.py:106-167).  defer_row_means: the
        caller runs gram() next, whose single read of X also yields the row means (m <= 64)."""
        if scale_type not in SCALE_CODES:
            raise NotImplementedError("The scaling method selected has not been implemented yet")
        F, ncl, m = self.F, self.n_c_loc, self.m
        st = _stream()
        cnt = torch.empty(self.n_loc, dtype=torch.float64, device=self.dev)
        stats = torch.zeros(F * 4, dtype=torch.float64, device=self.dev)
        blk = ncl * m
        ws = _ws(_lib.load().omb_block_stats_ws_bytes(F, blk), self.dev)
        pending = False
        count = self.n_c * m
        self._Gf_cached = None
        arrival = getattr(self, "_arrival", None)
        if arrival is not None:
            # X is still on its way from the host, one feature block at a time (ROM._engine / from_host):
            # run every pass of this stage on a block as soon as it has landed -- the statistics and, when
            # the caller defers the row means, the Gram pass hide behind the PCIe copy of the next blocks.
            # Multi-rank: the second-moment pass needs the GLOBAL block sums, so it runs over the whole
            # shard after the first exchange (the Gram, the bulk of the work, still overlaps the upload).
            self._arrival = None
            L = _lib.load()
            ws1 = _ws(L.omb_block_stats_ws_bytes(1, blk), self.dev)
            fuse_gram = defer_row_means and axis_cnt == 1
            X0c = None
            if fuse_gram:
                Gf = torch.empty(F * m * m, dtype=torch.float64, device=self.dev)
                gws = _ws(L.omb_gram_ws_bytes(1, ncl, m), self.dev)
                if self._centred_copy_wanted():
                    X0c = self._new_centred()
            cur = torch.cuda.current_stream()
            for f in range(F):
                cur.wait_event(arrival[f])
                Xf, sf, cf = self.X[f * ncl:(f + 1) * ncl], stats[4 * f:4 * f + 4], cnt[f * ncl:(f + 1) * ncl]
                _lib.call("omb_block_stats", _p(Xf), 1, blk, 0, count, _p(sf), _p(ws1), st)
                if scale_type in NEEDS_SQDEV and self.world == 1:
                    _lib.call("omb_block_stats", _p(Xf), 1, blk, 1, count, _p(sf), _p(ws1), st)
                if fuse_gram and X0c is not None:
                    X0f = X0c[f * ncl:(f + 1) * ncl]
                    self._centre(Xf, cf, True, out=X0f)
                    self._gram_of_copy(X0f, 1, ncl, Gf[f * m * m:(f + 1) * m * m])
                elif fuse_gram:
                    _lib.call("omb_gram_rowmeans", _p(Xf), 1, ncl, m, _p(cf), _p(Gf[f * m * m:(f + 1) * m * m]), _p(gws), st)
                elif axis_cnt == 1:
                    _lib.call("omb_row_means", _p(Xf), ncl, m, _p(cf), st)
            if fuse_gram:
                self._Gf_cached = Gf
                self._X0c = X0c
            if self.world > 1:
                stats = self.comm.combine_stats(stats, F, sq=False)
                if scale_type in NEEDS_SQDEV:
                    _lib.call("omb_block_stats", _p(self.X), F, blk, 1, count, _p(stats), _p(ws), st)
                    stats = self.comm.combine_stats(stats, F, sq=True)
        else:
            if axis_cnt == 1:
                if defer_row_means:
                    pending = True
                else:
                    _lib.call("omb_row_means", _p(self.X), self.n_loc, m, _p(cnt), st)
            _lib.call("omb_block_stats", _p(self.X), F, blk, 0, count, _p(stats), _p(ws), st)
            if self.world > 1:
                stats = self.comm.combine_stats(stats, F, sq=False)
            if scale_type in NEEDS_SQDEV:
                _lib.call("omb_block_stats", _p(self.X), F, blk, 1, count, _p(stats), _p(ws), st)
                if self.world > 1:
                    stats = self.comm.combine_stats(stats, F, sq=True)
        scl = torch.empty(F, dtype=torch.float64, device=self.dev)
        _lib.call("omb_finalize_scale", _p(stats), F, count, SCALE_CODES[scale_type], _p(scl),
                  1 if axis_cnt is None else 0, _p(cnt), ncl, st)
        if arrival is None:
            self._drop_centred()                        # a copy left over from an earlier fit
        self.cnt, self.scl, self.block_stats = cnt, scl, stats
        self._cnt_pending = pending
        return cnt, scl

    def set_scale_feature(self, f, value):
        self.scl[f] = value

    # ------------------------------------------------------------------------------------ K3
    @_staged("gram")
    def gram(self, centred=True, scaled=True):
        """G = X0^T X0 (m x m) from per-feature Grams of the centred rows."""
        self._wait_arrival()
        F, ncl, m = self.F, self.n_c_loc, self.m
        st = _stream()
        Gf = torch.empty(F * m * m, dtype=torch.float64, device=self.dev)
        ws = _ws(_lib.load().omb_gram_ws_bytes(F, ncl, m), self.dev)
        if centred and getattr(self, "_Gf_cached", None) is not None:   # produced block by block behind the H2D copy
            Gf, self._Gf_cached = self._Gf_cached, None
        elif centred and self._centred_copy_wanted() and \
                (X0c := self._centre(self.X, self._cnt, self._cnt_pending)) is not None:
            # many snapshots: the row means (when still due) and a centred copy of X from one pass, then the
            # tensor-core Gram on the copy -- no FP64 add left in its inner loop; back-projection reuses the copy
            self._cnt_pending = False
            self._X0c = X0c
            if Engine.trace:                             # bench.py: the centred-copy pass timed on its own
                ec = torch.cuda.Event(enable_timing=True)
                ec.record()
                self.marks.append(("centre_end", ec, ec))
            self._gram_of_copy(X0c, F, ncl, Gf)
        elif centred and self._cnt_pending:             # row means + centred Grams from one read of X
            _lib.call("omb_gram_rowmeans", _p(self.X), F, ncl, m, _p(self._cnt), _p(Gf), _p(ws), st)
            self._cnt_pending = False
        else:
            _lib.call("omb_gram", _p(self.X), F, ncl, m, _p(self.cnt if centred else None), _p(Gf), _p(ws), st)
        G = torch.empty(m, m, dtype=torch.float64, device=self.dev)
        _lib.call("omb_gram_combine", _p(Gf), F, m, _p(self.scl if scaled else None), _p(G), st)
        if self.world > 1:                          # fixed order: identical bits on every rank
            G = self.comm.sum_ordered(G)
        return G

    @_staged("eigh")
    def eig_pod(self, G):
        """m x m eigensolve -> singular values (descending) and right singular vectors (the
        largest-magnitude component of each vector positive)."""
        m = int(G.shape[0])
        SV = None
        if m <= int(_lib.load().omb_eigh_max_m()):
            w = torch.empty(m, dtype=torch.float64, device=self.dev)
            SV = torch.empty(m + m * m, dtype=torch.float64, device=self.dev)    # sigma | V: one D2H for the host side
            V = SV[m:].view(m, m)
            _lib.call("omb_eigh_jacobi", _p(G.contiguous()), m, _p(w), _p(V), None, _stream())
        elif self.world > 1 and self.rank != 0:
            # library eigensolver: rank 0 solves, everybody receives its result below
            w = torch.empty(m, dtype=torch.float64, device=self.dev)
            V = torch.empty(m, m, dtype=torch.float64, device=self.dev)
        else:
            with LINALG_LOCK:
                w, V = torch.linalg.eigh(G)
            w = torch.flip(w, dims=(0,))
            V = torch.flip(V, dims=(1,))
            idx = torch.argmax(V.abs(), dim=0)
            sgn = torch.sign(V[idx, torch.arange(m, device=V.device)])
            V = (V * torch.where(sgn == 0, torch.ones_like(sgn), sgn)).contiguous()
        if self.world > 1 and m > int(_lib.load().omb_eigh_max_m()):
            # library eigensolver: every rank must rotate with the very same V.  (The one-CTA Jacobi
            # kernel is deterministic and its input G is bit-identical on every rank: no broadcast.)
            w = self.comm.bcast(w.contiguous(), 0)
            V = self.comm.bcast(V.contiguous(), 0)
        # sigma and the back-projection weights V diag(1/sigma) (zero for numerically-zero modes): one launch
        if SV is None:
            SV = torch.empty(m + m * m, dtype=torch.float64, device=self.dev)
            SV[m:].copy_(V.reshape(-1))
            V = SV[m:].view(m, m)
        S = SV[:m]
        Wfull = torch.empty(m, m, dtype=torch.float64, device=self.dev)
        # a mode is "resolved" by the Gram route when lambda > m eps lambda_1, i.e. sigma > sqrt(m eps) sigma_1: below
        # that the eigenvalue is rounding noise of G (a true null direction shows up as +-m eps lambda_1) and the
        # mode gets a zero weight instead of noise amplified by 1/sigma (ROM._pod then takes the full-width route)
        _lib.call("omb_pod_weights", _p(w.contiguous()), _p(V), m, C.c_double(float(np.sqrt(m * EPS))), _p(S), _p(Wfull), _stream())
        self.pod_weights, self.pod_sv = Wfull, SV
        return S, V

    # ------------------------------------------------------------------------------------ K5
    @_staged("backproject")
    def backproject(self, W, centred=True, scaled=True, norms=True):
        """Ut (r x ld, mode-major) = (X0 W)^T, plus the initial pivoted-QR norms."""
        self._wait_arrival()
        W = W.contiguous()
        m, r = (int(v) for v in W.shape)
        assert m == self.m
        Ut = self._new_basis(r)
        vn = torch.zeros(self.ntiles * TB, dtype=torch.float64, device=self.dev) if norms else None
        src, cnt, mk = self.X, (self.cnt if centred else None), self.m
        if centred and self._X0c is not None:           # the centred copy the Gram pass left behind
            src, cnt, mk = self._X0c, None, int(self._X0c.shape[1])
            if mk != m:                                 # odd snapshot count: the copy carries a zero snapshot
                W = torch.cat([W, torch.zeros(mk - m, r, dtype=torch.float64, device=self.dev)], dim=0).contiguous()
        if src is self._X0c and (r & 1) and (mk > 64 or r > 64):
            # odd mode count: back-project one zero mode more (the tensor-core kernel wants even r), then drop it
            Wp = torch.cat([W, torch.zeros(mk, 1, dtype=torch.float64, device=self.dev)], dim=1).contiguous()
            Utp = self._new_basis(r + 1)
            _lib.call("omb_backproject", _p(src), self.F, self.n_c_loc, mk, None, _p(self.scl if scaled else None),
                      _p(Wp), r + 1, _p(Utp), _p(vn), _stream())
            _lib.call("omb_copy_modes", _p(Utp), r + 1, _p(Ut), r, self.n_loc, _stream())
            del Utp
        else:
            _lib.call("omb_backproject", _p(src), self.F, self.n_c_loc, mk, _p(cnt), _p(self.scl if scaled else None),
                      _p(W), r, _p(Ut), _p(vn), _stream())
        self._drop_centred()                            # its last reader has been queued: the memory can be reused
        self.Ut, self.vn, self.r = Ut, vn, r
        return Ut

    def _new_basis(self, r):
        Ut = torch.empty(self.ntiles, r, TB, dtype=torch.float64, device=self.dev)
        if self.n_loc % TB:
            Ut[-1].zero_()           # padding candidates of the last tile
        return Ut

    def set_basis_rows(self, Ur_dev):
        """Install a user-supplied basis (n_loc, r) (fit(basis=...), attribute assignment)."""
        Ur_dev = Ur_dev.contiguous()
        n, r = (int(v) for v in Ur_dev.shape)
        assert n == self.n_loc
        Ut = self._new_basis(r)
        vn = torch.zeros(self.ntiles * TB, dtype=torch.float64, device=self.dev)
        _lib.call("omb_rows_to_modes", _p(Ur_dev), n, r, _p(Ut), _p(vn), _stream())
        self.Ut, self.vn, self.r = Ut, vn, r

    def tiled_copy(self, Ur_dev):
        """Tiled mode-major copy of an (n_loc, r) matrix (does not touch the installed basis)."""
        Ur_dev = Ur_dev.contiguous()
        Ut = self._new_basis(int(Ur_dev.shape[1]))
        _lib.call("omb_rows_to_modes", _p(Ur_dev), self.n_loc, int(Ur_dev.shape[1]), _p(Ut), None, _stream())
        return Ut

    def basis_rows(self):
        """(n_loc, r) C-order copy of the basis on device."""
        out = torch.empty(self.n_loc, self.r, dtype=torch.float64, device=self.dev)
        _lib.call("omb_modes_to_rows", _p(self.Ut), self.n_loc, self.r, _p(out), _stream())
        return out

    def basis_gram(self):
        """H = U_r^T U_r (r x r) of the installed basis, identical on every rank: the basis is laid out as
        rows once and goes through the same Gram kernels as the snapshots."""
        U1 = self.basis_rows()
        r = self.r
        Gf = torch.empty(r * r, dtype=torch.float64, device=self.dev)
        ws = _ws(_lib.load().omb_gram_ws_bytes(1, self.n_loc, r), self.dev)
        _lib.call("omb_gram", _p(U1), 1, self.n_loc, r, None, _p(Gf), _p(ws), _stream())
        H = Gf.view(r, r)
        if self.world > 1:
            H = self.comm.sum_ordered(H.contiguous())
        return H, U1

    def basis_rotate(self, U1, M, norms=True):
        """Install U_r <- U1 M (U1: n_loc x k rows, M: k x r) as the basis, with fresh placement norms (the
        back-projection kernel with U1 in the role of the snapshots)."""
        M = M.contiguous()
        k, r = (int(v) for v in M.shape)
        assert int(U1.shape[1]) == k
        self.Ut = None                                # U1 holds the data: the old tiles can go before the new ones come
        Ut = self._new_basis(r)
        vn = torch.zeros(self.ntiles * TB, dtype=torch.float64, device=self.dev) if norms else None
        _lib.call("omb_backproject", _p(U1), 1, self.n_loc, k, None, None, _p(M), r, _p(Ut), _p(vn), _stream())
        self.Ut, self.vn, self.r = Ut, vn, r

    def mask_rows(self, mask_dev):
        """optimal_placement(mask=...): zero the excluded rows of the basis in place (:737-738)."""
        keep = torch.zeros(self.ntiles * TB, dtype=torch.float64, device=self.dev)
        keep[: self.n_loc] = mask_dev.to(torch.float64)
        self.Ut.mul_(keep.view(self.ntiles, 1, TB))
        self.vn = None      # recomputed by the placement

    # ------------------------------------------------------------------------------------ K6
    @_staged("qrcp")
    def qrcp(self, s=None, block=8):
        """Pivoted QR over the candidate rows; returns (piv, rdiag, gap) as device tensors.
        Multi-rank: piv holds GLOBAL row indices and is identical on every rank."""
        r = self.r
        s = r if s is None else int(s)
        block = max(1, min(8, int(block)))
        work = torch.empty(self.ntiles, r, TB, dtype=torch.float64, device=self.dev)
        ws = _ws(_lib.load().omb_qrcp_ws_bytes(self.n_loc, r), self.dev)
        self._qr_ws = ws                              # kept for qr_stats()
        out = torch.empty(3 * s, dtype=torch.float64, device=self.dev)      # one buffer: one D2H for all three
        piv, rdiag, gap = out[:s].view(torch.int64), out[s:2 * s], out[2 * s:]
        self.qr_out = out
        if self.world == 1:
            _lib.call("omb_qrcp", _p(self.Ut), self.n_loc, r, s, _p(self.vn), _p(work), _p(ws),
                      block, 0, _p(piv), _p(rdiag), _p(gap), _stream())
            return piv, rdiag, gap
        L = _lib.load()
        p2p = _comm.p2p_state(self.comm, self.dev, int(L.omb_qrcp_p2p_buffer_doubles(self.world)))
        self.qr_exchange = "p2p" if p2p is not None else "allgather: " + _comm.p2p_why_not(self.comm)
        if p2p is not None:
            # the kernels exchange the per-step records themselves over NVLink peer memory
            p2p["epoch"] += 1
            _lib.call("omb_qrcp_p2p", _p(self.Ut), self.n_loc, r, s, _p(self.vn), _p(work), _p(ws), block,
                      *self._shard_args(), C.c_void_p(p2p["peers_dev"]), _p(p2p["buf"]), p2p["epoch"],
                      _p(piv), _p(rdiag), _p(gap), _stream())
            self._p2p_err = (p2p["buf"], int(L.omb_qrcp_p2p_error_index(self.world)))
            return piv, rdiag, gap
        nrec = int(L.omb_qrcp_record_doubles())
        rec = torch.zeros(nrec, dtype=torch.float64, device=self.dev)
        sh = self._shard_args()
        _lib.call("omb_qrcp_mr_start", _p(self.Ut), self.n_loc, r, s, _p(self.vn), _p(ws), *sh, _stream())
        for i in range(s):
            _lib.call("omb_qrcp_mr_local", _p(self.Ut), _p(work), self.n_loc, r, _p(ws), block, i, *sh,
                      _p(rec), _stream())
            recs = self.comm.allgather(rec)           # one small record per rank per pivot step
            _lib.call("omb_qrcp_mr_step", _p(self.Ut), _p(work), self.n_loc, r, s, _p(ws), block, i, *sh,
                      _p(recs), _p(piv), _p(rdiag), _p(gap), _stream())
        return piv, rdiag, gap

    def qr_stats(self):
        """Executed schedule of the last placement's read-only passes (lazy norm down-dates,
        include/omb200.h `omb_qrcp_stats`): dict(seg_rows, seg_visits, retries, lazy, alpha).  Synchronises."""
        ws = getattr(self, "_qr_ws", None)
        if ws is None:
            return None
        out = (C.c_int64 * 6)()
        _lib.call("omb_qrcp_stats", _p(ws), self.n_loc, C.cast(out, C.c_void_p), _stream())
        return {"seg_rows": int(out[0]), "seg_visits": int(out[1]), "retries": int(out[2]), "lazy": bool(out[3]),
                "alpha": out[4] / 1.0e6}

    # ------------------------------------------------------------------------------------ GEM
    def gem(self, n_sensors, mask_dev=None, xyz_dev=None, d_min=0.0, Ut=None, normal=None, verbose=False):
        """Greedy entropy-maximisation placement (sparse_sensing.py:586-698): one streaming pass over
        the basis per sensor; the k x k covariance inverse (with the reference's random jitter, drawn
        through `normal` exactly where the reference calls np.random.normal) is formed on the host
        from the chosen rows.  Returns the sensor row indices (numpy int64)."""
        L = _lib.load()
        if n_sensors > int(L.omb_gem_max_sensors()):
            raise ValueError("n_sensors exceeds the supported %d" % int(L.omb_gem_max_sensors()))
        if normal is None:
            normal = lambda size: np.random.normal(size=size)
        Ut = self.Ut if Ut is None else Ut
        r = int(Ut.shape[1])
        n, st = self.n_loc, _stream()
        multi = self.world > 1
        var = torch.empty(n, dtype=torch.float64, device=self.dev)
        _lib.call("omb_gem_variance", _p(Ut), n, r, _p(var), st)
        alive = torch.ones(n, dtype=torch.uint8, device=self.dev) if mask_dev is None \
            else mask_dev.to(torch.uint8).contiguous()
        smax = torch.where(alive.bool(), var, torch.full_like(var, -1.0)).max().reshape(1)
        if multi:                                          # the largest variance over ALL ranks' live candidates
            smax = self.comm.allgather(smax).max().reshape(1)
        sigma_max = float(smax)
        coef = 1 / np.sqrt(sigma_max) * 2                  # :622
        ws = _ws(L.omb_gem_ws_bytes(), self.dev)
        idx = torch.empty(1, dtype=torch.int64, device=self.dev)
        val = torch.empty(1, dtype=torch.float64, device=self.dev)
        row = torch.empty(r, dtype=torch.float64, device=self.dev)
        rec = torch.empty(5 + r, dtype=torch.float64, device=self.dev) if multi else None
        sensors, rows, H_tot = [], [], 0.0
        if verbose:
            print(f"{'-'*70} \n {'# sensors':^10} {'sigma^2 y':^10} {'sigma^2 y|a':^10} {'Htot':^10} \n ")
        for s in range(int(n_sensors)):
            k, Zd, Bd = len(sensors), None, None
            if k > 0:
                A = np.asarray(rows) * coef                # Ur_scl[sensor_list_glb, :]
                Sigma_aa = np.cov(A, ddof=1)               # :660
                if s == 1:
                    B = np.atleast_2d(1 / Sigma_aa)        # :663
                else:
                    noise = 1e-5 * normal(Sigma_aa.shape[0])           # :667
                    if multi:                              # one jitter for all ranks: rank 0's draw
                        noise = self.comm.allgather(torch.from_numpy(noise).to(self.dev))[0].cpu().numpy()
                    B = np.linalg.inv(Sigma_aa + np.diag(noise))
                Zd = torch.from_numpy(np.ascontiguousarray(A - A.mean(axis=1, keepdims=True))).to(self.dev)
                Bd = torch.from_numpy(np.ascontiguousarray(B)).to(self.dev)
            _lib.call("omb_gem_step", _p(Ut), n, r, C.c_double(coef), k, _p(Zd), _p(Bd), _p(var), _p(alive), _p(ws),
                      _p(idx), _p(val), _p(row), st)
            if not multi:
                if d_min > 0.0:                            # :646-649 (d_min == 0 keeps every candidate)
                    _lib.call("omb_gem_exclude", _p(xyz_dev), self.n_c_loc, n, _p(idx), C.c_double(d_min), _p(alive), st)
                i = int(idx.item())
                if i < 0:
                    break                                  # every candidate excluded
                v, var_i, row_h = float(val.item()) if verbose else 0.0, None, row.cpu().numpy().copy()
            else:
                # row-sharded: every rank offers its local winner (value, GLOBAL index, coordinates, basis row); the
                # largest value wins, ties go to the lowest global index like np.argmax over the whole basis (:681)
                il = int(idx.item())
                rec.zero_()
                rec[0] = val[0] if il >= 0 else float("-inf")
                rec[1] = float(self.layout.to_global(torch.tensor([max(il, 0)])).item()) if il >= 0 else -1.0
                if il >= 0:
                    if xyz_dev is not None:
                        rec[2:5] = xyz_dev[il % self.n_c_loc]
                    rec[5:] = row
                allr = self.comm.allgather(rec).cpu().numpy()
                live = allr[:, 1] >= 0
                if not live.any():
                    break
                best = max((g for g in range(self.world) if live[g]), key=lambda g: (allr[g, 0], -allr[g, 1]))
                i, v, row_h = int(allr[best, 1]), float(allr[best, 0]), allr[best, 5:].copy()
                if d_min > 0.0:
                    _lib.call("omb_gem_exclude_point", _p(xyz_dev), self.n_c_loc, n, C.c_double(allr[best, 2]),
                              C.c_double(allr[best, 3]), C.c_double(allr[best, 4]), C.c_double(d_min), _p(alive), st)
            sensors.append(i)
            rows.append(row_h)
            if verbose:
                if s == 0:
                    print(f"{s+1:^10} {v:^10.2e} {'  -':^10} {'  -':^10}")
                else:
                    H_tot += 0.5 * np.log(v) + 0.5 * (np.log(2 * np.pi) + 1)
                    var_s = float(np.var(row_h * coef, ddof=1))
                    print(f"{s+1:^10} {var_s:^10.2e} {v:^10.2e} {H_tot:^10.2e}")
        return np.asarray(sensors, dtype=np.int64)

    # ------------------------------------------------------------------------------- K8 - K11
    def gather(self, piv_dev):
        """Theta = rows `piv` (global indices) of U_r, and the centring values at those rows."""
        s = int(piv_dev.numel())
        Theta = torch.empty(s, self.r, dtype=torch.float64, device=self.dev)
        cnt_s = torch.empty(s, dtype=torch.float64, device=self.dev)
        if self.world == 1:
            _lib.call("omb_gather_rows", _p(self.Ut), self.r, _p(piv_dev), s, _p(Theta),
                      _p(self.cnt), _p(cnt_s), _stream())
            return Theta, cnt_s
        owner, local = self.layout.owner_and_local(piv_dev)
        mine = owner == self.rank
        loc = torch.where(mine, local, torch.zeros_like(local)).contiguous()
        _lib.call("omb_gather_rows", _p(self.Ut), self.r, _p(loc), s, _p(Theta), _p(self.cnt), _p(cnt_s),
                  _stream())
        both = torch.cat([Theta * mine.unsqueeze(1), (cnt_s * mine).unsqueeze(1)], dim=1)
        both = self.comm.sum_ordered(both.contiguous())                  # one non-zero term per row
        return both[:, : self.r].contiguous(), both[:, self.r].contiguous()

    def local_columns(self, M):
        """The columns of a (s x n_global) scipy-sparse / dense matrix that fall on this rank's rows, as a scipy CSR
        matrix with LOCAL column indices (single rank: the matrix itself in CSR form)."""
        import scipy.sparse as sp
        M = M.tocsr() if sp.issparse(M) else sp.csr_matrix(np.asarray(M, dtype=np.float64))
        if self.world > 1:
            cols = self.layout.to_global(torch.arange(self.n_loc)).numpy()
            M = M.tocsc()[:, cols].tocsr()
        M.sort_indices()
        return M

    def csr_apply(self, M, want_basis=True):
        """(M U_r, M cnt, M scl_rows) for a general matrix M (s x n_global; sparse never densified): the CSR kernel on
        this rank's columns, partial results summed over the ranks in rank order.  M U_r is None without a basis."""
        M = self.local_columns(M)
        s = int(M.shape[0])
        r = int(self.r) if (want_basis and self.Ut is not None) else 0
        indptr = torch.from_numpy(M.indptr.astype(np.int64)).to(self.dev)
        indices = torch.from_numpy(M.indices.astype(np.int64)).to(self.dev)
        data = torch.from_numpy(M.data.astype(np.float64)).to(self.dev)
        max_nnz = int(np.diff(M.indptr).max()) if s else 0
        if self.world > 1:                               # the same launch geometry is not required, only the sums
            pass
        out = torch.zeros(s, r + 2, dtype=torch.float64, device=self.dev)
        Theta = torch.empty(s, max(r, 1), dtype=torch.float64, device=self.dev)
        cs = torch.empty(s, dtype=torch.float64, device=self.dev)
        ss = torch.empty(s, dtype=torch.float64, device=self.dev)
        ws = _ws(_lib.load().omb_csr_ws_bytes(s, max_nnz, r), self.dev)
        _lib.call("omb_csr_times_basis", _p(indptr), _p(indices if indices.numel() else None), _p(data if data.numel() else None),
                  s, max_nnz, _p(self.Ut if r else None), self.n_loc, r, _p(self.cnt), _p(self.scl), self.n_c_loc,
                  _p(Theta if r else None), _p(cs), _p(ss), _p(ws), _stream())
        if r:
            out[:, :r] = Theta
        out[:, r] = cs
        out[:, r + 1] = ss
        if self.world > 1:
            out = self.comm.sum_ordered(out.contiguous())
        return (out[:, :r].contiguous() if r else None), out[:, r].contiguous(), out[:, r + 1].contiguous()

    def ols_predict(self, Y_dev, cnt_s, scl_s, PinvT):
        N, s = (int(v) for v in Y_dev.shape)
        r = int(PinvT.shape[1])
        A = torch.empty(N, r, dtype=torch.float64, device=self.dev)
        _lib.call("omb_ols_predict", _p(Y_dev.contiguous()), _p(cnt_s), _p(scl_s), _p(PinvT.contiguous()),
                  N, s, r, _p(A), _stream())
        return A

    def wols_predict(self, Theta, y0v, y0s):
        """Weighted OLS for Nw vectors (y0v, y0s: (Nw, s) scaled values / uncertainties): (Ar, Ar_sigma), each
        (Nw, r).  One CTA per vector (Householder QR of diag(1/y0s) Theta in shared memory); vectors whose weighted
        Theta is not of full column rank -- and shapes beyond the shared-memory factorisation -- take the batched
        pseudo-inverse the reference's np.linalg.pinv defines."""
        Theta = Theta.contiguous()
        s, r = (int(v) for v in Theta.shape)
        Nw = int(y0v.shape[0])
        Ar = torch.zeros(Nw, r, dtype=torch.float64, device=self.dev)
        As = torch.zeros(Nw, r, dtype=torch.float64, device=self.dev)
        todo = torch.arange(Nw, device=self.dev)
        if s >= r and int(_lib.load().omb_wols_smem_bytes(s, r)) <= 227 * 1024:
            flag = torch.zeros(Nw, dtype=torch.int32, device=self.dev)
            _lib.call("omb_wols_predict", _p(Theta), s, r, _p(y0v), _p(y0s), Nw, C.c_double(1e-13), _p(Ar), _p(As), _p(flag),
                      _stream())
            todo = torch.nonzero(flag).flatten()
        if todo.numel():
            Wt = (1.0 / y0s[todo]).unsqueeze(2) * Theta.unsqueeze(0)           # diag(1/sigma) Theta
            with LINALG_LOCK:
                P = torch.linalg.pinv(Wt, rtol=1e-15)
            Ar[todo] = torch.bmm(P, (y0v[todo] / y0s[todo]).unsqueeze(2)).squeeze(2)
            As[todo] = torch.bmm(P, y0s[todo].unsqueeze(2)).squeeze(2).abs()
        return Ar, As

    def reconstruct(self, A_dev, row0=0, nrows=None, out=None):
        """rows [row0, row0+nrows) of scl * (U_r A^T) + cnt, as a (nrows, N) device tensor."""
        A_dev = A_dev.contiguous()
        N, r = (int(v) for v in A_dev.shape)
        assert r == self.r
        nrows = self.n_loc - row0 if nrows is None else int(nrows)
        if out is None:
            out = torch.empty(nrows, N, dtype=torch.float64, device=self.dev)
        _lib.call("omb_reconstruct", _p(self.Ut), self.n_loc, r, _p(A_dev), N, _p(self.cnt), _p(self.scl),
                  self.n_c_loc, int(row0), nrows, _p(out), _stream())
        return out

    def scaled_matrix(self):
        """X0 = (X - cnt)/scl materialised on device (only when user code asks for .X0)."""
        self._wait_arrival()
        X0 = torch.empty_like(self.X)
        _lib.call("omb_scale_rows", _p(self.X), self.F, self.n_c_loc, self.m, _p(self.cnt), _p(self.scl),
                  _p(X0), _stream())
        return X0

    def unscale(self, x0_dev):
        out = torch.empty_like(x0_dev)
        _lib.call("omb_unscale", _p(x0_dev.contiguous()), _p(self.cnt), _p(self.scl), self.n_c_loc,
                  int(x0_dev.numel()), _p(out), _stream())
        return out
