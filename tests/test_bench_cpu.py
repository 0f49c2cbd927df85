"""CPU tests of bench.py's arithmetic (algorithmic work per stage, shard split) and of the multi-device facade's
row bookkeeping -- no GPU, no compute calls."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_qrcp_schedule_bytes_counts_the_executed_passes():
    b = _bench()
    n, r = 1000, 40
    # block = 1: dlaqp2 itself -- every step reads L rows and writes L - 1 (+ the norm arrays), SURVEY 8(d)'s Q row
    tot, launches = b.qrcp_schedule_bytes(n, r, r, 1)
    expect = 8 * n + sum(8 * n * (r - i) + 24 * n + 8 * n * (r - i - 1) for i in range(r - 1))
    assert (tot, launches) == (expect, r)
    # blocked: fewer bytes, same number of pass launches; b = 8 ~ sqrt(r) is near the optimum
    t8, l8 = b.qrcp_schedule_bytes(n, r, r, 8)
    assert l8 == r and t8 < 0.70 * tot
    assert t8 <= min(b.qrcp_schedule_bytes(n, r, r, k)[0] for k in (2, 4, 16, 32)) * 1.02


def test_qrcp_schedule_bytes_of_a_lazy_run_charge_what_the_passes_visited():
    """Lazy norm down-dates: the block-closing apply passes touch every column, the read-only passes only the
    segments the kernels counted (omb_qrcp_stats): 512 bytes per (segment, row), 1.5 KB of norms per visit."""
    b = _bench()
    n, r, blk = 64 * 1000, 40, 8
    eager, launches = b.qrcp_schedule_bytes(n, r, r, blk)
    nseg = n // 64
    # every segment visited by every read-only pass = the eager schedule
    rows = sum((r - (i // blk) * blk) * nseg for i in range(r - 1) if i % blk != blk - 1)
    visits = sum(nseg for i in range(r - 1) if i % blk != blk - 1)
    full, l2 = b.qrcp_schedule_bytes(n, r, r, blk, {"lazy": True, "seg_rows": rows, "seg_visits": visits})
    assert (full, l2) == (eager, launches)
    # nothing visited: only the apply passes (and the step-0 argmax) are left
    none, _ = b.qrcp_schedule_bytes(n, r, r, blk, {"lazy": True, "seg_rows": 0, "seg_visits": 0})
    apply_only = 8 * n + sum(8 * n * (r - i0) + 24 * n + 8 * n * (r - i0 - blk) for i0 in range(0, r - blk, blk))
    assert none == apply_only and none < 0.3 * eager
    # stats of an eager run (lazy flag off) are ignored
    assert b.qrcp_schedule_bytes(n, r, r, blk, {"lazy": False, "seg_rows": 1, "seg_visits": 1})[0] == eager


def test_stage_rooflines_use_algorithmic_work():
    b = _bench()
    n, m, r = 16_200_000, 256, 100
    stages = {"stats": 14.0, "centre": 11.5, "gram": 33.0, "eigh": 2.9, "backproject": 26.5, "qrcp": 128.0}
    st, comp, qb, ql = b.stage_rooflines(n, m, r, stages, 216.0, 6467.4, 37.2, {})
    assert st["gram"]["bound"] == "fp64" and abs(st["gram"]["algorithmic_flop"] - n * m * (m + 1.0)) < 1
    assert st["backproject"]["bound"] == "fp64" and st["backproject"]["algorithmic_flop"] == 2.0 * n * m * r
    assert st["stats"]["bound"] == "hbm" and st["qrcp"]["algorithmic_bytes"] == float(qb)
    assert abs(st["gram"]["frac"] - (n * m * (m + 1.0) / 33e-3 / 1e12) / 37.2) < 1e-12
    # the centred-copy pass is reported but is not one of SURVEY 8(d)'s floors
    floors = sum(st[k]["floor_ms"] for k in ("stats", "gram", "backproject", "qrcp"))
    assert abs(comp["sum_floor_ms"] - floors) < 1e-9 and abs(comp["frac"] - floors / 216.0) < 1e-12
    # few snapshots: the same stages are HBM-bound
    st2, _, _, _ = b.stage_rooflines(1_652_580, 41, 40, {"gram": 0.25, "backproject": 0.3}, 4.2, 6467.4, 37.2, {})
    assert st2["gram"]["bound"] == "hbm" and st2["backproject"]["bound"] == "hbm"


def test_shard_cells_partitions_every_cell_once():
    b = _bench()
    for n_c, world in [(1_800_000, 8), (1_800_001, 8), (10, 3), (7, 7)]:
        spans = [b.shard_cells(n_c, world, rk) for rk in range(world)]
        assert spans[0][0] == 0 and sum(k for _, k in spans) == n_c
        for (c0, k), (c1, _) in zip(spans, spans[1:]):
            assert c0 + k == c1
        assert max(k for _, k in spans) - min(k for _, k in spans) <= 1


def test_multi_device_row_bookkeeping():
    """_MultiROM._pieces / MultiDevice.assemble: local rows of a device <-> the reference's global row order."""
    from openmeasure_b200 import multi, sparse_sensing as sps

    class FakeMD(multi.MultiDevice):
        def __init__(self, F, cells):
            self.F, self.cells, self.G = F, cells, len(cells)
            self.offsets = [sum(cells[:g]) for g in range(self.G)]
            self.n_c = sum(cells)

    F, cells = 3, [5, 4, 4]
    md = FakeMD(F, cells)
    n = F * md.n_c
    glob = np.arange(n * 2, dtype=np.float64).reshape(n, 2)
    parts = [md.shard(glob, g) for g in range(md.G)]
    assert [p.shape[0] for p in parts] == [F * c for c in cells]
    np.testing.assert_array_equal(md.assemble(parts), glob)
    rom = object.__new__(sps._MultiROM)
    rom._md = md
    for g in range(md.G):
        rows = md.global_rows(g)
        ncl = cells[g]
        for row0, cnt in [(0, F * ncl), (2, ncl), (ncl - 1, 3), (F * ncl - 2, 2)]:
            pieces = rom._pieces(g, row0, cnt)
            assert sum(k for _, k, _ in pieces) == cnt
            for off, k, g0 in pieces:                      # every piece is a contiguous run of global rows
                np.testing.assert_array_equal(rows[row0 + off:row0 + off + k], np.arange(g0, g0 + k))


def test_resolve_devices_parsing(monkeypatch):
    from openmeasure_b200 import multi
    monkeypatch.delenv("OMB_DEVICES", raising=False)
    assert multi.resolve_devices(None) == []
    import torch
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 4)
    assert multi.resolve_devices("all") == [0, 1, 2, 3]
    assert multi.resolve_devices("0, 2") == [0, 2]
    assert multi.resolve_devices([1, 3]) == [1, 3]
    monkeypatch.setenv("OMB_DEVICES", "1,2")
    assert multi.resolve_devices(None) == [1, 2]
