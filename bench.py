#!/usr/bin/env python
"""bench.py -- snapshot GB/s through POD + pivoted-QR placement (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" = one pass of the hot path (fit: centring/scaling statistics, POD; optimal_placement:
pivoted QR) over one batch of synthetic snapshots.  At N = 1 the workload is BASELINE.json
configs[1]: 9 features x 183 620 cells = 1 652 580 rows x 41 snapshots, std scaling, r = 40 modes
(r = m includes the rounding-noise mode of row-centred data) + 40-sensor QR
placement.  For N > 1 every rank holds the same number of cells of every feature (weak scaling).

  value       8*n*m bytes / device time, X resident in HBM, CUDA events, max over ranks
  e2e         the same metric through the reference-facing API SPR(X_host).fit().optimal_placement()
              with X in pinned host memory: H2D of X and D2H of the pivots inside the timed region
  roofline    the pivoted-QR pass kernels (the dominant kernels): algorithmic bytes of the schedule
              actually executed / their CUDA-event time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle port (numpy/scipy = the reference's own library calls) on the host cores

--impl reference times the reference's CPU implementation of the path (the oracle port: the
reference is pure Python and /root/reference does not exist on the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="configs[1]: synthetic 1.65M x 41 snapshots, std scaling, POD r=40 + 40-sensor QR placement",
                F=9, n_c=183620, m=41, r=40, scale_type="std")
QR_BLOCK = int(os.environ.get("OMB_QR_BLOCK", "8"))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=None, help="override cells per feature per GPU")
    ap.add_argument("--snapshots", type=int, default=None)
    ap.add_argument("--modes", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_step(X, F, r):
    from oracle import pod_oracle as po
    t0 = time.perf_counter()
    po.placement_pipeline(X, F, r)
    return time.perf_counter() - t0


def cpu_sample(w, frac):
    """Rows subsample of the workload for the CPU arm (every stage is O(n))."""
    from oracle import synth as osynth
    n_c = max(int(w["n_c"] * frac), 4 * w["m"])
    X = osynth.snapshots(w["F"], n_c, w["m"], w["r"])
    return X, n_c


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # all the host threads the box has: torchrun exports OMP_NUM_THREADS=1 to its children, so the limit is
    # lifted before numpy / scipy (OpenBLAS) are first imported, and again through threadpoolctl
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(os.cpu_count() or 1)
    import numpy as np  # noqa: F401
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass
    frac = 0.25
    X, n_c = cpu_sample(w, frac)
    for _ in range(min(args.warmup, 1)):
        cpu_step(X, w["F"], w["r"])
    steps = max(1, min(args.steps, 5))
    ts = [cpu_step(X, w["F"], w["r"]) for _ in range(steps)]
    t = sum(ts) / len(ts)
    gbs = 8.0 * X.shape[0] * X.shape[1] / t / 1e9
    sample = f"{w['F']}x{n_c} rows x {w['m']} snapshots ({frac:.2f} of the workload's rows; every stage is O(n))"
    out = {
        "impl": "reference", "metric": "snapshot GB/s through POD+pivoted-QR placement",
        "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "rows": w["F"] * w["n_c"], "snapshots": w["m"], "modes": w["r"],
                   "scale_type": w["scale_type"], "cpu_sample": sample},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cpu_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class Clocks:
    """SM clock / throttle-reason sampler (NVML in-process: a polling nvidia-smi child perturbs the
    very launches being timed).  Falls back to one nvidia-smi query if NVML is unavailable."""

    def __init__(self, index, period=0.02):
        self.samples = []
        self.stop_flag = False
        self.h = None
        self.period = period
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None
        self.index = index
        self.th = None

    def start(self):
        if self.h is None:
            return
        self.th = threading.Thread(target=self._poll, daemon=True)
        self.th.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), sm, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.th is not None:
            self.th.join(timeout=1.0)
        if self.h is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=20).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [],
                        "note": "NVML unavailable: single nvidia-smi query after the timed region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"]}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        inside = [(sm, rs) for ts, sm, rs in self.samples if t0 <= ts <= t1] or [(sm, rs) for _, sm, rs in self.samples[-1:]]
        sms = sorted(sm for sm, _ in inside)
        reasons = sorted(nm for nm, bit in names.items() if any(rs & bit for _, rs in inside))
        return {"sm_mhz": float(sms[len(sms) // 2]) if sms else None, "sm_max_mhz": float(self.max_sm),
                "reasons": reasons, "samples": len(inside)}


# ------------------------------------------------------------------------------------------------
# algorithmic bytes of the pivoted-QR schedule that omb_qrcp executes (DESIGN.md "Roofline")
# ------------------------------------------------------------------------------------------------
def qrcp_schedule_bytes(n, r, s, block):
    """(bytes, launches) of the pass kernels: read-only GEMV passes read L rows, block-closing
    apply passes read L rows and write L-t-1; every pass reads vn1/vn2 and writes vn1."""
    total, launches = 8 * n, 1            # step-0 argmax pass reads vn1
    i0 = 0
    for i in range(s - 1):
        t, L = i - i0, r - i0
        total += 8 * n * L + 24 * n
        launches += 1
        if t == block - 1:
            total += 8 * n * (L - t - 1)
            i0 = i + 1
    return total, launches


def main():
    args = parse()
    w = dict(WORKLOAD)
    if args.cells:
        w["n_c"] = args.cells
    if args.snapshots:
        w["m"] = args.snapshots
    if args.modes:
        w["r"] = args.modes
    if args.cells or args.snapshots or args.modes:
        w["name"] = f"custom: {w['F']}x{w['n_c']} rows x {w['m']} snapshots, r={w['r']}"
    if args.impl == "reference":
        return run_reference(args, w)

    import numpy as np
    import torch
    import torch.distributed as dist
    from openmeasure_b200 import _lib, build, synth as gsynth
    from openmeasure_b200.sparse_sensing import SPR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    L = _lib.load()

    F, n_c_loc, m, r = w["F"], w["n_c"], w["m"], w["r"]
    n_c = n_c_loc * world
    n_glob = F * n_c
    Xd = gsynth.snapshots(F, n_c, m, r, cell0=rank * n_c_loc, ncell_loc=n_c_loc)
    torch.cuda.synchronize()
    x_bytes_glob = 8.0 * n_glob * m
    group = None if world > 1 else False

    qr_events = []

    def step(timed_qr=False):
        spr = SPR.from_device(Xd, F, group=group)
        spr.fit(scale_type=w["scale_type"], select_modes="number", n_modes=r)
        if timed_qr:                                 # events only: no host sync inside the timed region
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        C = spr.optimal_placement(block=QR_BLOCK)
        if timed_qr:
            e1.record()
            qr_events.append((e0, e1))
        return spr, C

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the sampler starts BEFORE the warm-up: the first NVML queries of a process take milliseconds inside
    # the driver and stall kernel launches (measured: one 10 ms step right after the sampler's first poll)
    clocks = Clocks(local_rank) if rank == 0 and not os.environ.get("OMB_BENCH_NO_CLOCKS") else None
    if clocks:
        clocks.start()
    # warm-up: at least W steps AND at least 0.25 s, identical to the timed steps.  The time floor lets the
    # NVML sampler get several polls done under load: its first polls while kernels are in flight block
    # the driver for 10-70 ms (measured as one 72 ms "step" when they fell into the timed region).
    n_warm = max(args.warmup, 3)
    t_w = time.perf_counter()
    k_w = 0
    while True:
        spr, C = step(timed_qr=True)              # same object lifetimes as the timed loop (the caching
        torch.cuda.Event(enable_timing=True).record()   # allocator reaches its steady state here)
        k_w += 1
        go_on = 1 if (k_w < n_warm or time.perf_counter() - t_w < 0.25) else 0
        if world > 1:                             # every rank must run the same number of steps
            flag = torch.tensor([go_on], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            go_on = int(flag.item())
        if not go_on:
            break
    sync_all()
    qr_events.clear()
    L.omb_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_wall0 = time.perf_counter()
    e0.record()
    step_marks = [e0]
    for _ in range(args.steps):
        spr, C = step(timed_qr=True)
        mk = torch.cuda.Event(enable_timing=True)
        mk.record()
        step_marks.append(mk)
    e1.record()
    sync_all()
    t_wall1 = time.perf_counter()
    launches = int(L.omb_launch_count())
    ms = e0.elapsed_time(e1)
    qr_ms = [a.elapsed_time(b) for a, b in qr_events]
    step_ms = [step_marks[k].elapsed_time(step_marks[k + 1]) for k in range(len(step_marks) - 1)]
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    ms_per_step = ms / args.steps
    value = x_bytes_glob / (ms_per_step * 1e-3) / 1e9

    # ---- per-stage device times (diagnostic pass, not part of the timed region) ----
    stages = {}
    if rank == 0 or world > 1:
        from openmeasure_b200 import engine as eng_mod
        ev = lambda: torch.cuda.Event(enable_timing=True)
        eng = eng_mod.Engine(Xd, F, group=group)
        for rep in range(2):                         # first repetition warms the allocator
            marks = [ev() for _ in range(6)]
            marks[0].record()
            eng.stats(w["scale_type"], 1, defer_row_means=True)
            marks[1].record()
            G = eng.gram()
            marks[2].record()
            S, V = eng.eig_pod(G)
            marks[3].record()
            eng.backproject((V[:, :r] / S[:r]).contiguous())
            marks[4].record()
            if world == 1:
                eng.qrcp(block=QR_BLOCK)
            marks[5].record()
            torch.cuda.synchronize()
        names = ["stats", "gram", "eigh", "backproject", "qrcp"]
        stages = {nm: marks[i].elapsed_time(marks[i + 1]) for i, nm in enumerate(names)}

    # ---- roofline of the dominant kernels: the pivoted-QR passes ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    n_loc = F * n_c_loc
    qbytes, qlaunch = qrcp_schedule_bytes(n_loc, r, r, QR_BLOCK)
    qr_avg_ms = sum(qr_ms) / max(len(qr_ms), 1)
    achieved = qbytes / (qr_avg_ms * 1e-3) / 1e9 if qr_avg_ms > 0 else 0.0
    traffic = None                                # measured DRAM bytes per launch (ncu), default workload only
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if not (args.cells or args.snapshots or args.modes) and tj.get("qr_block") == QR_BLOCK:
            traffic = tj["qr_passes"]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_source": "profiles/r01_traffic.json (ncu dram__bytes_read+write per pass launch)" if traffic else None,
                "algorithmic_bytes_per_launch": qbytes / qlaunch,
                "kernel": "qr_gemv_kernel + qr_apply_kernel (pivoted-QR passes, block=%d)" % QR_BLOCK,
                "algorithmic_bytes_per_step": qbytes, "launches_per_step": qlaunch,
                "avg_launch_us": qr_avg_ms * 1e3 / qlaunch, "qrcp_ms_per_step": qr_avg_ms,
                "peak_source": peak_src,
                "note": "bytes = schedule actually executed (blocked QRCP); includes the 1-CTA panel kernels' time"}

    # ---- e2e through the reference-facing API with HOST buffers (rank-local shard) ----
    e2e = None
    if not args.no_e2e:
        Xh_t = torch.empty(Xd.shape, dtype=torch.float64, pin_memory=True)
        Xh_t.copy_(Xd)
        torch.cuda.synchronize()
        Xh = Xh_t.numpy()
        xyz = np.zeros((n_c_loc, 3))

        def e2e_step():
            if world == 1:
                s = SPR(Xh, F, xyz)
            else:                                   # host shard -> device shard, then the same API
                xd = torch.empty(Xd.shape, dtype=torch.float64, device="cuda")
                xd.copy_(Xh_t, non_blocking=True)
                s = SPR.from_device(xd, F, group=group)
            s.fit(scale_type=w["scale_type"], select_modes="number", n_modes=r)
            Cq = s.optimal_placement(block=QR_BLOCK)
            return Cq.pivots

        for _ in range(2):
            e2e_step()
        sync_all()
        k = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k):
            piv = e2e_step()
        sync_all()
        dt = (time.perf_counter() - t0) / k
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": x_bytes_glob / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(8 * n_loc * m * world),
               "d2h_bytes_per_step": int((3 * r * 8 + (m + m * m) * 8) * world),   # pivots|rdiag|gaps + sigma|V "ms_per_step": dt * 1e3,
               "api": "SPR(X_host, F, xyz).fit(select_modes='number', n_modes=r); optimal_placement()"}
        del Xh_t

    # ---- reconstructions / s (second half of the BASELINE metric), small diagnostic ----
    recon = None
    if rank == 0 and world == 1:
        spr.train(C)
        Nvec = 128
        Y = torch.rand(Nvec, r, dtype=torch.float64, device="cuda")
        eng = spr._eng
        scl_s = eng.scl[torch.from_numpy(C.pivots // n_c_loc).cuda()].contiguous()
        out = torch.empty(eng.n_loc, Nvec, dtype=torch.float64, device="cuda")
        for _ in range(2):
            A = eng.ols_predict(Y, spr._cnt_s, scl_s, spr._PinvT)
            eng.reconstruct(A, out=out)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        reps = 3
        for _ in range(reps):
            A = eng.ols_predict(Y, spr._cnt_s, scl_s, spr._PinvT)
            eng.reconstruct(A, out=out)
        a1.record()
        torch.cuda.synchronize()
        rms = a0.elapsed_time(a1) / reps
        recon = {"value": Nvec / (rms * 1e-3), "unit": "reconstructions/s", "vectors": Nvec, "rows": eng.n_loc,
                 "modes": r, "ms": rms, "fp64_tflops": 2.0 * eng.n_loc * r * Nvec / (rms * 1e-3) / 1e12}
        del out

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0, N = 1 only ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        frac = 0.25
        Xs, ncs = cpu_sample(w, frac)
        t = cpu_step(Xs, F, r)
        cpu = {"value": 8.0 * Xs.shape[0] * m / t / 1e9, "unit": "GB/s", "cores": cpu_threads(), "kind": "port",
               "sample": f"{F}x{ncs} rows x {m} snapshots ({frac:.2f} of the rows; every stage is O(n)), "
                         f"{t:.1f} s of numpy svd + scipy qr(pivoting=True)"}

    if rank == 0:
        out = {
            "metric": "snapshot GB/s through POD+pivoted-QR placement", "value": value, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": k_w, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "rows_per_gpu": n_loc, "rows": n_glob, "snapshots": m, "modes": r,
                       "sensors": r, "scale_type": w["scale_type"], "qr_block": QR_BLOCK,
                       "l2": "inputs (%.0f MB per GPU) exceed the 126 MB L2" % (8.0 * n_loc * m / 1e6),
                       "parallelism": "cells sharded over %d rank(s)" % world,
                       "qr_exchange": getattr(spr._eng, "qr_exchange", "none (single rank)")},
            "e2e": e2e, "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "stages_ms": stages, "step_ms": [round(v, 3) for v in step_ms], "reconstruct": recon,
            "pivots_head": [int(p) for p in spr.qr_pivots[:8]], "min_pivot_gap": float(spr.qr_gap.min()),
            "sigma_r_over_sigma_1": float(spr.Sigma_r[-1] / spr.Sigma_r[0]),
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
