"""CPU tests of the host-side logic of the row-sharded (N > 1) path: shard layout / index maps,
fixed-order combination of the small cross-rank objects, and the communicators (world_size-2
gloo processes and the in-process thread emulation)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from openmeasure_b200 import comm as C


def test_shard_layout_round_trip():
    F, cells = 3, [5, 7, 4]
    n_c = sum(cells)
    for rank in range(3):
        lay = C.ShardLayout(F, cells, rank)
        loc = torch.arange(F * cells[rank])
        glob = lay.to_global(loc)
        f, c = loc // cells[rank], loc % cells[rank]
        assert torch.equal(glob, f * n_c + lay.cell0 + c)
        owner, local = lay.owner_and_local(glob)
        assert torch.all(owner == rank) and torch.equal(local, loc)
    lay = C.ShardLayout(F, cells, 0)
    owner, local = lay.owner_and_local(torch.tensor([0, 4, 5, 11, 12, 15, 16 + 6, 2 * 16 + 15]))
    assert owner.tolist() == [0, 0, 1, 1, 2, 2, 1, 2]
    assert local.tolist() == [0, 4, 0, 6, 0, 3, 7 + 1, 2 * 4 + 3]


def test_combine_block_stats_fixed_order():
    F = 2
    g = torch.tensor([[1.0, -2.0, 5.0, 10.0, 3.0, 0.5, 9.0, 1.0],
                      [1e-17, -3.0, 4.0, 20.0, 4.0, 0.25, 11.0, 2.0],
                      [2.0, -1.0, 7.0, 30.0, 5.0, 1.0, 2.0, 3.0]], dtype=torch.float64)
    out = C.combine_block_stats(g, F, sq=False).view(F, 4)
    assert out[0, 0] == (1.0 + 1e-17) + 2.0 and out[1, 0] == 12.0
    assert out[:, 1].tolist() == [-3.0, 0.25] and out[:, 2].tolist() == [7.0, 11.0]
    out = C.combine_block_stats(g, F, sq=True).view(F, 4)
    assert out[:, 3].tolist() == [60.0, 6.0]
    assert torch.equal(C.ordered_sum(g), (g[0] + g[1]) + g[2])


def test_thread_comm_allgather_and_bcast():
    import threading
    comms = C.ThreadComm.make(3)
    res = [None] * 3

    def run(rk):
        t = torch.full((4,), float(rk))
        g = comms[rk].allgather(t)
        b = comms[rk].bcast(torch.full((2,), float(rk + 10)), src=1)
        res[rk] = (g.clone(), b.clone())

    th = [threading.Thread(target=run, args=(k,)) for k in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    for rk in range(3):
        assert res[rk][0].shape == (3, 4)
        assert res[rk][0][:, 0].tolist() == [0.0, 1.0, 2.0]
        assert res[rk][1].tolist() == [11.0, 11.0]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = C.TorchDistComm(None)
        # per-rank block statistics of a sharded block -> every rank combines to the same result
        rng = np.random.default_rng(7)
        X = rng.random((2 * 12, 5))                              # F=2, 12 cells, 6 per rank
        lay = C.ShardLayout(2, [6, 6], rank)
        rows = lay.to_global(torch.arange(12)).numpy()
        Xl = X[rows]
        stats = torch.zeros(2, 4, dtype=torch.float64)
        for f in range(2):
            blk = Xl[f * 6:(f + 1) * 6]
            stats[f] = torch.tensor([blk.sum(), blk.min(), blk.max(), 0.0])
        comb = C.combine_block_stats(comm.allgather(stats.reshape(-1)), 2, sq=False).view(2, 4)
        G = torch.from_numpy(Xl.T @ Xl)
        Gs = C.ordered_sum(comm.allgather(G)).view(5, 5)
        v = comm.bcast(torch.full((3,), float(rank + 1), dtype=torch.float64), src=0)
        q.put((rank, comb.numpy(), Gs.numpy(), v.numpy(), X))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_combination_is_identical_on_every_rank():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    (_, c0, G0, v0, X), (_, c1, G1, v1, _) = out
    np.testing.assert_array_equal(c0, c1)
    np.testing.assert_array_equal(G0, G1)
    np.testing.assert_array_equal(v0, [1.0, 1.0, 1.0])
    np.testing.assert_array_equal(v1, [1.0, 1.0, 1.0])
    for f in range(2):
        blk = X[f * 12:(f + 1) * 12]
        np.testing.assert_allclose(c0[f, 0], blk.sum(), rtol=1e-14)
        assert c0[f, 1] == blk.min() and c0[f, 2] == blk.max()
    np.testing.assert_allclose(G0, X.T @ X, rtol=1e-13)
