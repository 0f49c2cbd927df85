"""Tiny communicator abstraction for the row-sharded path (one process per GPU).

Only small, fixed-size objects ever cross ranks (F*4 block statistics, the m x m Gram, one
record per pivot step, the s x r Theta), always as an all-gather followed by a fixed-order
combination, so every rank computes bit-identical results:

    SingleComm      world = 1, no communication
    TorchDistComm   torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests)
    ThreadComm      G ranks emulated by G threads of one process (single-GPU tests of the
                    multi-rank numerics; kernels of all ranks are issued on the same stream, so
                    no kernel ever waits on another rank's kernel)
"""
import threading

import torch


class SingleComm:
    world = 1
    rank = 0

    def allgather(self, t):
        """(world, numel) tensor holding every rank's flattened t, in rank order."""
        return t.reshape(1, -1)

    def bcast(self, t, src=0):
        return t

    def check(self):
        pass

    def combine_stats(self, stats, F, sq):
        """Block statistics (F*4: sum, min, max, sqdev per feature) combined over the ranks in rank
        order -- identical bits on every rank."""
        if self.world == 1:
            return stats
        return combine_block_stats(self.allgather(stats), F, sq)

    def sum_ordered(self, t):
        """Element-wise sum of t over the ranks, added in rank order (identical bits on every rank)."""
        if self.world == 1:
            return t
        return ordered_sum(self.allgather(t)).view_as(t)


class TorchDistComm(SingleComm):
    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    P2P_CAPACITY = 8192          # doubles per rank and call served by the peer-memory collectives

    def _p2p(self, flat):
        """Peer-memory state when `flat` qualifies for the one-kernel collectives, else None."""
        if flat.is_cuda and flat.dtype == torch.float64 and 0 < flat.numel() <= self.P2P_CAPACITY:
            return p2p_coll_state(self, flat.device, self.P2P_CAPACITY)
        return None

    def _p2p_launch(self, name, st, flat, out, *extra):
        import ctypes as C
        from . import _lib
        st["seq"] += 1
        _lib.call(name, C.c_void_p(flat.data_ptr()), flat.numel(), C.c_void_p(out.data_ptr()),
                  C.c_void_p(st["peers_dev"]), C.c_void_p(st["buf"].data_ptr()), self.P2P_CAPACITY,
                  st["seq"], self.rank, self.world, *extra, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.p2p_collectives = getattr(self, "p2p_collectives", 0) + 1
        self._p2p_used = st

    def allgather(self, t):
        flat = t.contiguous().reshape(-1)
        out = torch.empty(self.world * flat.numel(), dtype=flat.dtype, device=flat.device)
        st = self._p2p(flat)
        if st is not None:
            # small FP64 payload: one kernel stores it into every peer's symmetric buffer (tagged
            # low-latency words) and polls its own -- a few microseconds instead of an NCCL launch
            self._p2p_launch("omb_p2p_allgather", st, flat, out)
            return out.view(self.world, flat.numel())
        self.dist.all_gather_into_tensor(out, flat, group=self.group)
        return out.view(self.world, flat.numel())

    def _allreduce(self, t, op):
        flat = t.contiguous().reshape(-1)
        st = self._p2p(flat)
        if st is None:
            return None
        out = torch.empty_like(flat)
        self._p2p_launch("omb_p2p_allreduce", st, flat, out, op)      # exchange + rank-ordered combine: one kernel
        return out

    def combine_stats(self, stats, F, sq):
        out = self._allreduce(stats, 2 if sq else 1)
        return out if out is not None else combine_block_stats(self.allgather(stats), F, sq)

    def sum_ordered(self, t):
        out = self._allreduce(t, 0)
        return out.view_as(t) if out is not None else ordered_sum(self.allgather(t)).view_as(t)

    def check(self):
        """Raise if a peer never published during a peer-memory collective since the last check (the
        kernels give up after 10 s and set an error word; the payload they returned is then garbage).
        Host-synchronising: call it where the host waits for results anyway."""
        st = getattr(self, "_p2p_used", None)
        if st is None:
            return
        from . import _lib
        idx = int(_lib.load().omb_p2p_allgather_error_index(self.world, self.P2P_CAPACITY))
        word = st["buf"][idx:idx + 1].view(torch.int64)
        if int(word.item()) != 0:
            word.zero_()
            raise _lib.OmbError("peer-memory collective: a peer rank did not publish its payload (10 s timeout); "
                                "statistics / Gram / sensor rows of this step are invalid")

    def bcast(self, t, src=0):
        gsrc = src if self.group is None else self.dist.get_global_rank(self.group, src)
        self.dist.broadcast(t, gsrc, group=self.group)
        return t


_P2P_CACHE = {}


def _one_rank_per_gpu(comm):
    """The peer-memory kernels spin on words another rank's kernel writes: every rank's kernel must be
    resident at the same time, i.e. every rank needs a GPU of its own (two ranks sharing one GPU would
    wait on each other until the time-out).  Raises -> the caller falls back to NCCL."""
    if comm.world > torch.cuda.device_count() and not _all_ranks_distinct_devices(comm):
        raise RuntimeError("world size %d exceeds the %d visible GPUs: ranks would share a GPU"
                           % (comm.world, torch.cuda.device_count()))


def _all_ranks_distinct_devices(comm):
    """Multi-node or masked-visibility layouts: accept when every rank of this host reports a distinct
    device UUID."""
    try:
        uuid = str(torch.cuda.get_device_properties(torch.cuda.current_device()).uuid)
        mine = torch.tensor([hash(uuid) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda")
        allv = torch.empty(comm.world, dtype=torch.int64, device="cuda")
        comm.dist.all_gather_into_tensor(allv, mine, group=comm.group)
        return len(set(allv.tolist())) == comm.world
    except Exception:
        return False


def p2p_state(comm, device, ndoubles):
    """Symmetric (peer-mapped over NVLink) FP64 buffer shared by the ranks of `comm`, created once
    per process group: dict(buf, peers_dev, epoch) or None when unavailable / disabled
    (OMB_QR_EXCHANGE=nccl).  Collective: every rank must call it at the same point."""
    import os
    if not isinstance(comm, TorchDistComm) or os.environ.get("OMB_QR_EXCHANGE", "p2p") != "p2p":
        return None
    key = (id(comm.group) if comm.group is not None else 0, comm.world)
    st = _P2P_CACHE.get(key)
    if st is None or st["buf"].numel() < ndoubles:
        try:
            _one_rank_per_gpu(comm)
            import torch.distributed._symmetric_memory as symm_mem
            grp = comm.group if comm.group is not None else comm.dist.group.WORLD
            buf = symm_mem.empty(int(ndoubles), dtype=torch.float64, device=device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, group=grp)
            torch.cuda.synchronize(device)
            comm.dist.barrier(group=comm.group)
            st = {"buf": buf, "hdl": hdl, "peers_dev": int(hdl.buffer_ptrs_dev), "epoch": 0}
        except Exception as e:          # symmetric memory not available on this system
            st = {"buf": torch.empty(0), "hdl": None, "peers_dev": 0, "epoch": 0, "error": repr(e)}
        _P2P_CACHE[key] = st
    return st if st["hdl"] is not None else None


_P2P_COLL_CACHE = {}


def p2p_coll_state(comm, device, capacity):
    """Symmetric buffer of the peer-memory all-gather for `comm` (created once per process group):
    dict(buf, peers_dev, seq) or None when unavailable / disabled (OMB_SMALL_ALLGATHER=nccl).
    Collective: the first call must happen at the same point on every rank."""
    import os
    if not isinstance(comm, TorchDistComm) or os.environ.get("OMB_SMALL_ALLGATHER", "p2p") != "p2p":
        return None
    key = (id(comm.group) if comm.group is not None else 0, comm.world, int(capacity))
    st = _P2P_COLL_CACHE.get(key)
    if st is None:
        try:
            _one_rank_per_gpu(comm)
            import torch.distributed._symmetric_memory as symm_mem
            from . import _lib
            nd = int(_lib.load().omb_p2p_allgather_buffer_doubles(comm.world, int(capacity)))
            buf = symm_mem.empty(nd, dtype=torch.float64, device=device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, group=comm.group if comm.group is not None else comm.dist.group.WORLD)
            torch.cuda.synchronize(device)
            comm.dist.barrier(group=comm.group)
            st = {"buf": buf, "hdl": hdl, "peers_dev": int(hdl.buffer_ptrs_dev), "seq": 0}
        except Exception as e:          # symmetric memory not available on this system
            st = {"buf": None, "hdl": None, "peers_dev": 0, "seq": 0, "error": repr(e)}
        _P2P_COLL_CACHE[key] = st
    return st if st["hdl"] is not None else None


def p2p_why_not(comm):
    """Why the peer-memory exchange is not in use for `comm` (diagnostics for bench.py / logs)."""
    import os
    if not isinstance(comm, TorchDistComm):
        return "communicator is not torch.distributed"
    if os.environ.get("OMB_QR_EXCHANGE", "p2p") != "p2p":
        return "OMB_QR_EXCHANGE=" + os.environ["OMB_QR_EXCHANGE"]
    key = (id(comm.group) if comm.group is not None else 0, comm.world)
    return str(_P2P_CACHE.get(key, {}).get("error", "unknown"))


class _ThreadShared:
    def __init__(self, world):
        self.slots = [None] * world
        self.barrier = threading.Barrier(world)


class ThreadComm(SingleComm):
    """Create with ThreadComm.make(world) -> list of per-rank communicators."""

    def __init__(self, shared, rank, world):
        self.shared, self.rank, self.world = shared, rank, world

    @classmethod
    def make(cls, world):
        shared = _ThreadShared(world)
        return [cls(shared, r, world) for r in range(world)]

    def allgather(self, t):
        sh = self.shared
        sh.slots[self.rank] = t.contiguous().reshape(-1)
        sh.barrier.wait()
        out = torch.stack(list(sh.slots))
        sh.barrier.wait()
        return out

    def bcast(self, t, src=0):
        g = self.allgather(t)
        t.copy_(g[src].view_as(t))
        return t


def combine_block_stats(gathered, F, sq):
    """Fixed-order (rank 0..G-1) combination of per-rank block statistics (world, F*4):
    columns {sum, min, max, sqdev} per feature.  sq selects which columns are combined."""
    g = gathered.view(gathered.shape[0], F, 4)
    out = g[0].clone()
    if not sq:
        acc = g[0, :, 0].clone()
        for k in range(1, g.shape[0]):
            acc = acc + g[k, :, 0]
        out[:, 0] = acc
        out[:, 1] = g[:, :, 1].min(dim=0).values
        out[:, 2] = g[:, :, 2].max(dim=0).values
    else:
        acc = g[0, :, 3].clone()
        for k in range(1, g.shape[0]):
            acc = acc + g[k, :, 3]
        out[:, 3] = acc
    return out.reshape(-1).contiguous()


def ordered_sum(gathered):
    """Sum over ranks in rank order (identical bits on every rank)."""
    acc = gathered[0].clone()
    for k in range(1, gathered.shape[0]):
        acc = acc + gathered[k]
    return acc


class ShardLayout:
    """Cells [cell0, cell0 + n_c_loc) of every one of the F features live on this rank."""

    def __init__(self, F, cells_per_rank, rank):
        self.F = int(F)
        self.cells = [int(c) for c in cells_per_rank]
        self.rank = int(rank)
        self.n_c = sum(self.cells)
        self.offsets = [sum(self.cells[:k]) for k in range(len(self.cells))]
        self.n_c_loc = self.cells[self.rank]
        self.cell0 = self.offsets[self.rank]

    def to_global(self, local_rows):
        f = local_rows // self.n_c_loc
        return f * self.n_c + self.cell0 + (local_rows - f * self.n_c_loc)

    def owner_and_local(self, global_rows):
        """(owner rank, local row on the owner) of global rows (tensors or numpy arrays)."""
        f = global_rows // self.n_c
        c = global_rows - f * self.n_c
        owner = (c * 0)
        local = (c * 0)
        for k, (off, cnt) in enumerate(zip(self.offsets, self.cells)):
            inside = (c >= off) & (c < off + cnt)
            owner = owner + inside * k
            local = local + inside * (f * cnt + (c - off))
        return owner, local
