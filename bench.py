#!/usr/bin/env python
"""bench.py -- snapshot GB/s through POD + pivoted-QR placement; reconstructions/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c3|c2|c5]

A "step" = one pass of the hot path (fit: centring/scaling statistics, POD; optimal_placement:
pivoted QR) over one batch of synthetic snapshots.

Workload.  N = 1: BASELINE.json configs[2], the largest single-GPU configuration: 9 features x
1 800 000 cells = 16 200 000 rows x 256 snapshots (33.2 GB), std scaling, r = 100 modes + 100-sensor QR
placement.  N > 1: the SAME global problem, cells sharded over the N ranks (strong scaling).  At
N = 8 the line also carries `north_star`: configs[4], 8 x 2^23 cells x 1024 snapshots (550 GB, 68.7 GB
per GPU), r = 100 -- the run BASELINE.json's target sentence is about.  configs[1] (1.65M x 41, the
round-1 headline) is kept as `extra.configs1`; configs[3] (batched OLS predict + reconstruct against an
r = 256 basis on the 16.2M-row mesh) is the `reconstruct` block, the second half of the metric.

  value        8*n*m bytes / device time, X resident in HBM, CUDA events, max over ranks
  e2e          the same metric through the reference-facing API with X in pinned HOST memory:
               H2D of X and D2H of the results inside the timed region
  roofline     the dominant kernels (pivoted-QR passes) against the measured HBM peak, plus one entry per
               stage in roofline.stages (HBM stages against MEASURED_PEAKS.json hbm_gbs; FP64 stages on
               SURVEY 8(d)'s ALGORITHMIC flops against the DMMA peak measured in this same run) and the
               composite sum(stage floors) / step
  cpu_baseline the oracle port (numpy/scipy = the reference's own library calls) on the host cores, on a
               row subsample (every stage is O(n)), with the linear extrapolation labelled as such

--impl reference times the reference's CPU implementation of the path (the oracle port: the reference is
pure Python and /root/reference does not exist on the GPU box) on the host cores.
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

WORKLOADS = {
    "c3": dict(key="c3", name="configs[2]: synthetic 3D 16.2M rows (9 features x 1.8M cells) x 256 snapshots, std scaling, "
               "POD r=100 + 100-sensor QR placement", F=9, n_c=1_800_000, m=256, r=100, scale_type="std",
               cpu_cells=50_000),
    "c2": dict(key="c2", name="configs[1]: synthetic 1.65M x 41 snapshots, std scaling, POD r=40 + 40-sensor QR placement",
               F=9, n_c=183_620, m=41, r=40, scale_type="std", cpu_cells=45_905),
    "c5": dict(key="c5", name="configs[4]: 64M rows (8 features x 2^23 cells) x 1024 snapshots FP64 (550 GB), std scaling, "
               "POD r=100 + 100-sensor QR placement", F=8, n_c=1 << 23, m=1024, r=100, scale_type="std",
               cpu_cells=8_192),
}
RECON = dict(name="configs[3]: batched OLS predict + reconstruct against an r=256 basis on the 16.2M-row mesh",
             r=256, n_dev=1024, n_e2e=256)
QR_BLOCK = int(os.environ.get("OMB_QR_BLOCK", "8"))
METRIC = "snapshot GB/s through POD+pivoted-QR placement"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--cells", type=int, default=None, help="override GLOBAL cells per feature")
    ap.add_argument("--snapshots", type=int, default=None)
    ap.add_argument("--modes", type=int, default=None)
    ap.add_argument("--shard", action="store_true", help="--cells is per GPU (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip configs[1], reconstruct and north_star blocks")
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


def all_host_threads():
    # torchrun exports OMP_NUM_THREADS=1 to its children: lift it before numpy / scipy (OpenBLAS) start
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(os.cpu_count() or 1)
    import numpy as np  # noqa: F401
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def cpu_step(X, F, r):
    from oracle import pod_oracle as po
    t0 = time.perf_counter()
    po.placement_pipeline(X, F, r)
    return time.perf_counter() - t0


def cpu_sample(w):
    """Row subsample of the workload for the CPU arm (every stage of the path is O(n))."""
    from oracle import synth as osynth
    n_c = min(w["n_c"], max(int(w["cpu_cells"]), 4 * w["m"]))
    X = osynth.snapshots(w["F"], n_c, w["m"], w["r"])
    return X, n_c


def cpu_describe(w, n_c, t):
    frac = n_c / w["n_c"]
    return (f"{w['F']}x{n_c} rows x {w['m']} snapshots = {frac:.4f} of the workload's rows, {t:.1f} s per step of numpy svd + "
            f"scipy qr(pivoting=True); every stage is O(n): linear extrapolation to the full workload = {t / frac:.0f} s per step")


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    all_host_threads()
    X, n_c = cpu_sample(w)
    warm = min(args.warmup, 1)
    for _ in range(warm):
        cpu_step(X, w["F"], w["r"])
    steps = max(1, min(args.steps, 3))
    ts = [cpu_step(X, w["F"], w["r"]) for _ in range(steps)]
    t = sum(ts) / len(ts)
    gbs = 8.0 * X.shape[0] * X.shape[1] / t / 1e9
    sample = cpu_describe(w, n_c, t)
    out = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "rows": w["F"] * w["n_c"], "snapshots": w["m"], "modes": w["r"],
                   "sensors": w["r"], "scale_type": w["scale_type"], "cpu_sample": sample},
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

    def window(self, t0, t1):
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

    def stop(self):
        self.stop_flag = True
        if self.th is not None:
            self.th.join(timeout=1.0)


# ------------------------------------------------------------------------------------------------
# algorithmic work per stage; QR: bytes of the schedule omb_qrcp actually executes
# ------------------------------------------------------------------------------------------------
def qrcp_schedule_bytes(n, r, s, block, stats=None):
    """(bytes, launches) of the pass kernels: read-only GEMV passes read L rows, block-closing
    apply passes read L rows and write L-t-1; every pass reads vn1/vn2 and writes vn1.
    stats (Engine.qr_stats() of a run with lazy norm down-dates): the read-only passes are charged
    what they VISITED -- 512 bytes per (64-candidate segment, row) and 64 x 24 bytes of norms per
    segment visit, counted by the kernels themselves -- instead of every row of every column."""
    lazy = bool(stats and stats.get("lazy"))
    total, launches = 8 * n, 1            # step-0 argmax pass reads vn1
    i0 = 0
    for i in range(s - 1):
        t, L = i - i0, r - i0
        launches += 1
        if t == block - 1:
            total += 8 * n * L + 24 * n + 8 * n * (L - t - 1)
            i0 = i + 1
        elif not lazy:
            total += 8 * n * L + 24 * n
    if lazy:
        total += 512 * int(stats["seg_rows"]) + 64 * 24 * int(stats["seg_visits"])
    return total, launches


def stage_rooflines(n_loc, m, r, stages_ms, step_ms, hbm_gbs, fp64_tflops, peak_src, qr_stats=None):
    """One roofline entry per stage, per GPU (n_loc rows), on ALGORITHMIC work, plus the composite."""
    qbytes, qlaunch = qrcp_schedule_bytes(n_loc, r, r, QR_BLOCK, qr_stats)
    gram_flop = float(n_loc) * m * (m + 1)
    gram_bytes = 8.0 * n_loc * m + 8.0 * n_loc
    bp_flop = 2.0 * n_loc * m * r
    bp_bytes = 8.0 * n_loc * m + 16.0 * n_loc + 8.0 * n_loc * r
    spec = {
        # S1 + second-moment pass: two reads of X (np.std is two-pass: bit-exact replay of numpy's tree)
        "stats": ("hbm", 2 * (8.0 * n_loc * m) + 8.0 * n_loc, None),
        # row means + centred copy for the tensor-core passes (m > 64): one read, one write of X
        "centre": ("hbm", 2 * (8.0 * n_loc * m) + 8.0 * n_loc, None),
        "gram": ("fp64" if gram_flop / (fp64_tflops * 1e12) > gram_bytes / (hbm_gbs * 1e9) else "hbm", gram_bytes, gram_flop),
        "backproject": ("fp64" if bp_flop / (fp64_tflops * 1e12) > bp_bytes / (hbm_gbs * 1e9) else "hbm", bp_bytes, bp_flop),
        "qrcp": ("hbm", float(qbytes), None),
    }
    out, floors = {}, 0.0
    for nm, (bound, nbytes, flop) in spec.items():
        ms = stages_ms.get(nm)
        if not ms:
            continue
        if bound == "hbm":
            ach, peak, unit = nbytes / (ms * 1e-3) / 1e9, hbm_gbs, "GB/s"
            floor = nbytes / (hbm_gbs * 1e9) * 1e3
        else:
            ach, peak, unit = flop / (ms * 1e-3) / 1e12, fp64_tflops, "TFLOP/s"
            floor = flop / (fp64_tflops * 1e12) * 1e3
        if nm != "centre":                           # not one of SURVEY 8(d)'s stages: reported, not counted as a floor
            floors += floor
        out[nm] = {"bound": bound, "ms": ms, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                   "floor_ms": floor, "algorithmic_bytes": nbytes, "algorithmic_flop": flop}
    if "eigh" in stages_ms:
        out["eigh"] = {"bound": "latency (m x m eigensolve, off the n-row path)", "ms": stages_ms["eigh"]}
    comp = {"sum_floor_ms": floors, "step_ms": step_ms, "frac": floors / step_ms if step_ms else None,
            "note": "sum of SURVEY 8(d)'s per-stage floors (stats, gram, backproject, qrcp; the m x m eigensolve and the "
                    "centred-copy pass have none) / measured step"}
    return out, comp, qbytes, qlaunch


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    w = dict(WORKLOADS[args.workload])
    if args.cells:
        w["n_c"] = args.cells * (args.gpus if args.shard else 1)
    if args.snapshots:
        w["m"] = args.snapshots
    if args.modes:
        w["r"] = args.modes
    if args.cells or args.snapshots or args.modes:
        w["name"] = f"custom: {w['F']}x{w['n_c']} rows x {w['m']} snapshots, r={w['r']}"
        w["key"] = "custom"
    if args.impl == "reference":
        return run_reference(args, w)

    import numpy as np
    import torch
    import torch.distributed as dist
    from openmeasure_b200 import _lib, build

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    numa = pin_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    L = _lib.load()
    ctx = Ctx(args, world, rank, local_rank, torch, dist, np, L)
    ctx.numa = numa

    peaks = measure_peaks(ctx)
    clocks = Clocks(local_rank) if rank == 0 and not os.environ.get("OMB_BENCH_NO_CLOCKS") else None
    if clocks:
        clocks.start()       # before any warm-up: the first NVML polls of a process stall kernel launches for ms
    ctx.clocks = clocks

    main_res = run_workload(ctx, w, args.steps, args.warmup, peaks, e2e=not args.no_e2e, parity=world > 1,
                            keep_for_recon=False)
    extras = {}
    recon = None
    north = None
    if not args.no_extras and w["key"] == "c3":
        recon = run_reconstruct(ctx, WORKLOADS["c3"], peaks)
        if world == 1:
            extras["configs1"] = run_workload(ctx, WORKLOADS["c2"], max(args.steps, 10), max(args.warmup, 5), peaks,
                                              e2e=not args.no_e2e, parity=False, brief=True)
        if world == 8:
            north = run_workload(ctx, WORKLOADS["c5"], min(args.steps, 3), 1, peaks, e2e=False, parity=False, brief=True)
            north["target"] = "full POD + 100-sensor QR placement at >= 70 % of the aggregate roofline on 8 x B200 (BASELINE.json north_star)"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        all_host_threads()
        Xs, ncs = cpu_sample(w)
        t = cpu_step(Xs, w["F"], w["r"])
        cpu = {"value": 8.0 * Xs.shape[0] * w["m"] / t / 1e9, "unit": "GB/s", "cores": cpu_threads(), "kind": "port",
               "sample": cpu_describe(w, ncs, t)}
    if clocks:
        clocks.stop()

    if rank == 0:
        out = {
            "metric": METRIC, "value": main_res["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": main_res["warmup"], "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak" if args.shard else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": main_res["config"], "e2e": main_res["e2e"], "gpu_launches": main_res["gpu_launches"],
            "clocks": main_res["clocks"], "roofline": main_res["roofline"], "cpu_baseline": cpu,
            "stages_ms": main_res["stages_ms"], "step_ms": main_res["step_ms"], "peaks": peaks,
            "reconstruct": recon, "multigpu_parity": main_res.get("multigpu_parity"),
            "pivots_head": main_res["pivots_head"], "min_pivot_gap": main_res["min_pivot_gap"],
            "sigma_r_over_sigma_1": main_res["sigma_r_over_sigma_1"], "pod_rel_err_bound": main_res["pod_rel_err_bound"],
            "north_star": north, "extra": extras or None,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


class Ctx:
    def __init__(self, args, world, rank, local_rank, torch, dist, np, L):
        self.args, self.world, self.rank, self.local_rank = args, world, rank, local_rank
        self.torch, self.dist, self.np, self.L = torch, dist, np, L
        self.clocks = None
        self.numa = None

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def pin_to_gpu_numa_node(index):
    """Bind this rank's host threads (and therefore its pinned staging buffers: first touch) to the cores next
    to its GPU, so that N ranks uploading at once do not all cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1 and 64 * i + b < ncpu]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} GPU-local cores"
    except Exception as e:
        return f"not applied ({type(e).__name__})"
    return "not applied"


def measure_peaks(ctx):
    """Roofline denominators.  HBM: MEASURED_PEAKS.json (driver-written).  FP64: measured HERE, in this run:
    a register-only DMMA.8x8x4 stream on every SM (omb_fp64_peak) as a burst (~5 ms) and sustained
    (~300 ms, the length of a config-3 step) figure, next to cuBLAS DGEMM 8192^3 for reference."""
    torch, L = ctx.torch, ctx.L
    import ctypes as C
    peaks = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        mp = {}
    peaks["hbm_gbs"] = float(mp.get("hbm_gbs", 6650.0))
    peaks["hbm_source"] = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in mp else "6650 GB/s (of fallback)"
    scratch = torch.empty(2 * 148 * 256 * 2, dtype=torch.float64, device="cuda")
    tf, ms = C.c_double(0.0), C.c_double(0.0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for label, target in (("burst", 5.0), ("sustained", 300.0)):
        _lib_call(ctx, "omb_fp64_peak", C.c_double(target), C.c_void_p(scratch.data_ptr()), scratch.numel(),
                  C.byref(tf), C.byref(ms), st)
        peaks[f"fp64_dmma_tflops_{label}"] = tf.value
        peaks[f"fp64_dmma_{label}_ms"] = ms.value
    try:
        a = torch.rand(8192, 8192, dtype=torch.float64, device="cuda")
        b = torch.rand(8192, 8192, dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        peaks["fp64_cublas_dgemm_8192_tflops"] = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a, b
    except Exception as e:
        peaks["fp64_cublas_dgemm_8192_tflops"] = None
        peaks["fp64_cublas_error"] = repr(e)
    peaks["fp64_tflops"] = peaks["fp64_dmma_tflops_sustained"]
    peaks["fp64_source"] = ("omb_fp64_peak measured in this run: register-only DMMA.8x8x4 stream on all SMs for ~300 ms "
                            "(sustained; the FP64 stages run inside steps of that length); burst and cuBLAS figures beside it")
    torch.cuda.empty_cache()
    return peaks


def _lib_call(ctx, name, *a):
    from openmeasure_b200 import _lib
    _lib.call(name, *a)


def shard_cells(n_c, world, rank):
    base, rem = divmod(n_c, world)
    mine = base + (1 if rank < rem else 0)
    c0 = rank * base + min(rank, rem)
    return c0, mine


def run_workload(ctx, w, steps, warmup, peaks, e2e=True, parity=False, brief=False, keep_for_recon=False):
    """Timed fit + optimal_placement steps on the workload, rows sharded over the ranks of ctx."""
    torch, dist, np, L = ctx.torch, ctx.dist, ctx.np, ctx.L
    from openmeasure_b200 import synth as gsynth, engine as eng_mod
    from openmeasure_b200.sparse_sensing import SPR
    world, rank = ctx.world, ctx.rank
    F, n_c, m, r = w["F"], w["n_c"], w["m"], w["r"]
    cell0, n_c_loc = shard_cells(n_c, world, rank)
    n_glob, n_loc = F * n_c, F * n_c_loc
    Xd = gsynth.snapshots(F, n_c, m, r, cell0=cell0, ncell_loc=n_c_loc)
    torch.cuda.synchronize()
    x_bytes_glob = 8.0 * n_glob * m
    group = None if world > 1 else False
    qr_events = []

    step_marks_all = []

    def step():
        spr = SPR.from_device(Xd, F, group=group)
        spr.fit(scale_type=w["scale_type"], select_modes="number", n_modes=r)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()                                   # events only: no host sync inside the timed region
        C = spr.optimal_placement(block=QR_BLOCK)
        e1.record()
        qr_events.append((e0, e1))
        step_marks_all.append(spr._eng.marks)
        return spr, C

    # warm-up: W steps (>= 3) and at least 0.25 s, identical to the timed steps (same object lifetimes: the caching
    # allocator reaches its steady state; the NVML sampler gets its slow first polls done under load)
    eng_mod.Engine.trace = True
    n_warm = max(warmup, 3)
    t_w, k_w = time.perf_counter(), 0
    while True:
        spr, C = step()
        k_w += 1
        go_on = 1 if (k_w < n_warm or time.perf_counter() - t_w < 0.25) else 0
        if world > 1:                             # every rank must run the same number of steps
            flag = torch.tensor([go_on], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            go_on = int(flag.item())
        if not go_on:
            break
    ctx.sync_all()
    qr_events.clear()
    step_marks_all.clear()
    L.omb_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync_all()
    t_wall0 = time.perf_counter()
    e0.record()
    marks = [e0]
    for _ in range(steps):
        spr, C = step()
        mk = torch.cuda.Event(enable_timing=True)
        mk.record()
        marks.append(mk)
    e1.record()
    ctx.sync_all()
    t_wall1 = time.perf_counter()
    launches = int(L.omb_launch_count())
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    qr_ms = [a.elapsed_time(b) for a, b in qr_events]
    step_ms = [marks[k].elapsed_time(marks[k + 1]) for k in range(len(marks) - 1)]
    clk = ctx.clocks.window(t_wall0, t_wall1) if ctx.clocks else None
    ms_per_step = ms / steps
    value = x_bytes_glob / (ms_per_step * 1e-3) / 1e9

    # ---- per-stage device times of the TIMED steps themselves (Engine.trace: events around every stage) ----
    acc = {}
    for marks_k in step_marks_all:
        d = {}
        for nm, m0, m1 in marks_k:
            if nm == "centre_end":
                d["_centre_end"] = m0
            else:
                d[nm] = (m0, m1)
        for nm, (m0, m1) in [(k2, v2) for k2, v2 in d.items() if not k2.startswith("_")]:
            if nm == "gram" and "_centre_end" in d:
                acc.setdefault("centre", []).append(m0.elapsed_time(d["_centre_end"]))
                acc.setdefault("gram", []).append(d["_centre_end"].elapsed_time(m1))
            else:
                acc.setdefault(nm, []).append(m0.elapsed_time(m1))
    order = ["stats", "centre", "gram", "eigh", "backproject", "qrcp"]
    stages = {nm: ctx.max_over_ranks(sum(acc[nm]) / len(acc[nm])) if nm in acc else None for nm in order}
    stages = {k2: v2 for k2, v2 in stages.items() if v2 is not None}

    # ---- rooflines ----
    hbm, fp64 = peaks["hbm_gbs"], peaks["fp64_tflops"]
    n_max = F * shard_cells(n_c, world, 0)[1]          # the largest shard bounds the step
    # the read-only QR passes report what they visited (lazy norm down-dates); the largest count bounds the step
    qst = spr._eng.qr_stats() or {"seg_rows": 0, "seg_visits": 0, "retries": 0, "lazy": False, "alpha": 0.0}
    qst = {"seg_rows": int(ctx.max_over_ranks(float(qst["seg_rows"]))), "seg_visits": int(ctx.max_over_ranks(float(qst["seg_visits"]))),
           "retries": int(qst["retries"]), "lazy": bool(qst["lazy"]), "alpha": float(qst["alpha"])}
    st_roof, comp, qbytes, qlaunch = stage_rooflines(n_max, m, r, stages, ms_per_step, hbm, fp64, peaks, qst)
    qbytes_eager, _ = qrcp_schedule_bytes(n_max, r, r, QR_BLOCK)
    qr_lazy = {"on": qst["lazy"], "alpha": qst["alpha"], "catch_up_rounds": qst["retries"],
               "read_only_pass_bytes": 512 * qst["seg_rows"] + 1536 * qst["seg_visits"],
               "eager_schedule_bytes_per_step": qbytes_eager, "executed_over_eager": qbytes / qbytes_eager,
               "note": "partial column norms only shrink: segments of 64 candidates whose largest norm at a block start is "
                       "below alpha x the pivot norm sit the block's read-only passes out (exact; pivots are those of the "
                       "eager schedule, parity-tested); bytes counted by the kernels"}
    qr_avg = ctx.max_over_ranks(sum(qr_ms) / max(len(qr_ms), 1))
    achieved = qbytes / (qr_avg * 1e-3) / 1e9 if qr_avg > 0 else 0.0
    traffic, tsrc = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        ent = tj.get(w["key"])
        if ent and ent.get("qr_block") == QR_BLOCK and world == 1 and bool(ent.get("qr_lazy")) == qst["lazy"]:
            traffic = ent["qr_passes"]["dram_bytes_per_launch"]
            tsrc = "profiles/r02_traffic.json (ncu dram__bytes_read+write per pass launch, same command)"
            for nm, t_ent in (ent.get("stages") or {}).items():      # measured DRAM bytes of the other stages' kernels
                if nm in st_roof:
                    st_roof[nm]["traffic"] = t_ent["dram_bytes_per_step"]
            if "qrcp" in st_roof:
                st_roof["qrcp"]["traffic"] = ent["qr_passes"]["dram_bytes_per_step"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": traffic, "traffic_source": tsrc,
                "kernel": "qr_gemv_kernel + qr_apply_mma_kernel (pivoted-QR passes, block=%d): the dominant kernels, %.0f %% of the step"
                          % (QR_BLOCK, 100.0 * qr_avg / ms_per_step),
                "algorithmic_bytes_per_launch": qbytes / qlaunch, "algorithmic_bytes_per_step": qbytes,
                "launches_per_step": qlaunch, "avg_launch_us": qr_avg * 1e3 / qlaunch, "qrcp_ms_per_step": qr_avg,
                "peak_source": peaks["hbm_source"],
                "note": "bytes = schedule actually executed (blocked QRCP with lazy norm down-dates: see `lazy`), per GPU; "
                        "time = CUDA events around optimal_placement in the timed steps, incl. the 1-CTA panel kernels",
                "lazy": qr_lazy,
                "stages": st_roof, "composite": comp,
                "fp64_peak": {"tflops": fp64, "source": peaks["fp64_source"]}}

    res = {"value": value, "ms_per_step": ms_per_step, "warmup": k_w, "steps": steps, "gpu_launches": launches,
           "clocks": clk, "roofline": roofline, "stages_ms": stages, "step_ms": [round(v, 3) for v in step_ms],
           "pivots_head": [int(p) for p in spr.qr_pivots[:8]], "min_pivot_gap": float(spr.qr_gap.min()),
           "sigma_r_over_sigma_1": float(spr.Sigma_r[-1] / spr.Sigma_r[0]),
           "pod_rel_err_bound": float(spr.pod_rel_err_bound),
           "config": {"workload": w["name"], "rows": n_glob, "rows_per_gpu": n_loc, "snapshots": m, "modes": r, "sensors": r,
                      "scale_type": w["scale_type"], "qr_block": QR_BLOCK,
                      "l2": "inputs (%.1f GB per GPU) exceed the 126 MB L2" % (8.0 * n_loc * m / 1e9),
                      "parallelism": "cells sharded over %d rank(s), one process per GPU" % world,
                      "warmup_note": "%d warm-up steps ran (requested %d; at least 3 and at least 0.25 s)" % (k_w, warmup),
                      "qr_exchange": getattr(spr._eng, "qr_exchange", "none (single rank)"),
                      "host_affinity": ctx.numa}}

    # ---- multi-GPU parity: the sharded run against ONE GPU running the same global problem ----
    if parity and world > 1:
        res["multigpu_parity"] = multigpu_parity(ctx, w, spr)
    piv_sharded = spr.qr_pivots.copy()
    del spr, C

    # ---- e2e through the reference-facing API with HOST buffers ----
    if e2e:
        res["e2e"] = run_e2e(ctx, w, Xd, n_loc, x_bytes_glob, steps, piv_sharded)
    else:
        res["e2e"] = None
    del Xd
    eng_mod.release_scratch()                        # the pooled centred-copy buffer of this workload
    torch.cuda.empty_cache()
    if brief:
        for k in ("step_ms",):
            res.pop(k, None)
    return res


def multigpu_parity(ctx, w, spr):
    """Rank 0 runs the whole problem on its own GPU (it fits for configs[2]) and compares."""
    torch, np = ctx.torch, ctx.np
    from openmeasure_b200 import synth as gsynth
    from openmeasure_b200.sparse_sensing import SPR
    out = None
    if ctx.rank == 0:
        try:
            Xg = gsynth.snapshots(w["F"], w["n_c"], w["m"], w["r"])
            one = SPR.from_device(Xg, w["F"], group=False)
            one.fit(scale_type=w["scale_type"], select_modes="number", n_modes=w["r"])
            one.optimal_placement(block=QR_BLOCK)
            same = bool(np.array_equal(one.qr_pivots, spr.qr_pivots))
            ds = float(np.max(np.abs(spr.Sigma_r - one.Sigma_r) / one.Sigma_r))
            out = {"reference": "the same global problem on one GPU (rank 0)", "pivots_identical": same,
                   "max_rel_dsigma": ds, "ok": bool(same and ds < 1e-12),
                   "p2p_collectives": int(getattr(spr._eng.comm, "p2p_collectives", 0))}
            del one, Xg
        except torch.OutOfMemoryError:
            out = {"skipped": "the global problem does not fit one GPU"}
        torch.cuda.empty_cache()
    ctx.sync_all()
    if out is not None and out.get("ok") is False:
        print(json.dumps({"error": "multi-GPU parity failed", **out}), file=sys.stderr, flush=True)
    return out


def run_e2e(ctx, w, Xd, n_loc, x_bytes_glob, steps, piv_expect):
    torch, np = ctx.torch, ctx.np
    from openmeasure_b200.sparse_sensing import SPR
    world = ctx.world
    F, m, r = w["F"], w["m"], w["r"]
    Xh_t = torch.empty(Xd.shape, dtype=torch.float64, pin_memory=True)
    Xh_t.copy_(Xd)
    torch.cuda.synchronize()
    Xh = Xh_t.numpy()
    xyz = np.zeros((n_loc // F, 3))

    def e2e_step():
        if world == 1:
            s = SPR(Xh, F, xyz)
        else:
            s = SPR.from_host(Xh, F, xyz)              # this rank's host shard, same upload path
        s.fit(scale_type=w["scale_type"], select_modes="number", n_modes=r)
        return s.optimal_placement(block=QR_BLOCK).pivots

    for _ in range(2):
        piv = e2e_step()
    ctx.sync_all()
    k = max(2, min(steps, 5 if 8.0 * n_loc * m < 4e9 else 3))
    t0 = time.perf_counter()
    for _ in range(k):
        piv = e2e_step()
    ctx.sync_all()
    dt = ctx.max_over_ranks((time.perf_counter() - t0) / k)
    ok = bool(np.array_equal(piv, piv_expect))
    del Xh_t
    api = ("SPR(X_host, n_features, xyz).fit(select_modes='number', n_modes=r); optimal_placement()" if world == 1 else
           "SPR.from_host(X_host_shard, n_features, xyz).fit(select_modes='number', n_modes=r); optimal_placement() on every rank")
    return {"value": x_bytes_glob / dt / 1e9, "unit": "GB/s", "ms_per_step": dt * 1e3, "steps": k,
            "h2d_bytes_per_step": int(x_bytes_glob),
            "d2h_bytes_per_step": int((3 * r * 8 + (m + m * m) * 8) * world),   # pivots|rdiag|gaps + sigma|V per rank
            "api": api, "pivots_match_device_resident_run": ok,
            "note": "wall clock over %d steps, max over ranks; X in pinned host memory, feature blocks uploaded on a copy "
                    "stream while the statistics / Gram passes consume the blocks that have landed" % k}


def run_reconstruct(ctx, w, peaks):
    """Second half of the metric: reconstructions/s = N / t(predict(N) + reconstruct(N)) against an r = 256
    basis on the 16.2M-row mesh (configs[3]); rows sharded over the ranks, every rank writes its own rows."""
    torch, np, dist = ctx.torch, ctx.np, ctx.dist
    from openmeasure_b200 import synth as gsynth
    from openmeasure_b200.sparse_sensing import SPR
    world, rank = ctx.world, ctx.rank
    F, n_c, m = w["F"], w["n_c"], w["m"]
    r = RECON["r"]
    cell0, n_c_loc = shard_cells(n_c, world, rank)
    n_loc = F * n_c_loc
    Xd = gsynth.snapshots(F, n_c, m, r, cell0=cell0, ncell_loc=n_c_loc)
    spr = SPR.from_device(Xd, F, group=None if world > 1 else False)
    spr.pod_refine = False            # any full-rank basis serves here; the correction would double the basis memory
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spr.fit(scale_type="std", select_modes="number", n_modes=r)      # r = m: includes the null mode of row-centred data
    C = spr.optimal_placement(block=QR_BLOCK)
    spr.train(C)
    eng = spr._eng
    s = r
    piv = C.pivots
    feat = (piv // n_c).astype(np.float64)
    rng = np.random.default_rng(7)

    def measurements(N):
        Y = np.zeros((N, s, 3))
        Y[:, :, 0] = spr._cnt_s.cpu().numpy()[None, :] + rng.standard_normal((N, s))
        Y[:, :, 2] = feat[None, :]
        return Y

    # ---- device-resident: coefficients and output chunks stay in HBM (output overwritten chunk by chunk) ----
    N = RECON["n_dev"]
    Yd = torch.from_numpy(measurements(N)[:, :, 0].copy()).cuda()
    scl_s = eng.scl[torch.from_numpy(piv // n_c).cuda()].contiguous()
    rows_chunk = min(n_loc, max(128, ((2 << 30) // (8 * N)) // 128 * 128))
    out = torch.empty(rows_chunk, N, dtype=torch.float64, device="cuda")

    def dev_pass():
        A = eng.ols_predict(Yd, spr._cnt_s, scl_s, spr._PinvT)
        for row0 in range(0, n_loc, rows_chunk):
            eng.reconstruct(A, row0=row0, nrows=min(rows_chunk, n_loc - row0), out=out)

    dev_pass()
    ctx.sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dev_pass()
    e1.record()
    ctx.sync_all()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    n_max = F * shard_cells(n_c, world, 0)[1]
    flop = 2.0 * n_max * r * N + 2.0 * N * s * r
    fp64 = peaks["fp64_tflops"]
    dev = {"value": N / (ms * 1e-3), "unit": "reconstructions/s", "vectors": N, "ms": ms,
           "roofline": {"bound": "fp64", "achieved": flop / (ms * 1e-3) / 1e12, "peak": fp64, "unit": "TFLOP/s",
                        "frac": flop / (ms * 1e-3) / 1e12 / fp64, "algorithmic_flop": flop,
                        "written_bytes": 8.0 * n_max * N, "ceiling_recon_per_s_per_gpu": fp64 * 1e12 / (2.0 * F * n_c * r),
                        "peak_source": peaks["fp64_source"]}}
    del out, Yd

    # ---- e2e through predict(list of (s,3) arrays) + reconstruct(Ar, out=sink): host in, host out ----
    N2 = RECON["n_e2e"]
    Ylist = list(measurements(N2))
    sunk = [0.0, 0]

    def sink(row0, block):
        sunk[0] += float(block[0, 0])              # the bytes are on the host; a consumer would write them out here
        sunk[1] += block.size

    def e2e_pass():
        Ar, _ = spr.predict(Ylist)
        spr.reconstruct(Ar, out=sink)

    e2e_pass()
    ctx.sync_all()
    sunk[1] = 0
    t0 = time.perf_counter()
    e2e_pass()
    ctx.sync_all()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": N2 / dt, "unit": "reconstructions/s", "vectors": N2, "ms": dt * 1e3,
           "h2d_bytes": int(N2 * s * 3 * 8 * world), "d2h_bytes": int(8 * F * n_c * N2),
           "api": "spr.predict(list of (s,3) arrays); spr.reconstruct(Ar, out=callable) -- row chunks stream through two "
                  "pinned host buffers; PCIe-bound: every reconstruction is 8 n = %.0f MB of output" % (8.0 * F * n_c / 1e6)}

    # ---- CPU baseline on a row subsample: the oracle port's predict + reconstruct ----
    cpu = None
    if rank == 0 and world == 1 and not ctx.args.no_cpu_baseline:
        try:
            all_host_threads()
            from oracle import pod_oracle as po
            ncs, Nc = 20_000, 32
            nrow = F * ncs
            Ur = rng.standard_normal((nrow, r))
            cnt = rng.standard_normal((nrow, 1))
            scl = np.ones((nrow, 1))
            pv = rng.choice(nrow, s, replace=False)
            Cm = po.one_hot(pv, nrow)
            Th = po.theta(Cm, Ur)
            ys = [np.column_stack([rng.standard_normal(s), np.zeros(s), np.zeros(s)]) for _ in range(Nc)]
            t0 = time.perf_counter()
            A, _ = po.predict_ols(Th, ys, Cm, cnt, scl, ncs)      # the reference's per-vector loop (pinv + C.X_cnt each)
            po.reconstruct(Ur, A, cnt, scl)
            t = time.perf_counter() - t0
            frac = ncs / n_c
            cpu = {"value": Nc / t * frac, "unit": "reconstructions/s", "cores": cpu_threads(), "kind": "port",
                   "sample": f"{Nc} vectors on {nrow} rows ({frac:.4f} of the mesh) in {t:.2f} s = {Nc / t:.1f} recon/s at that size; "
                             f"O(n) per vector: linear extrapolation to 16.2M rows = {Nc / t * frac:.2f} recon/s"}
        except Exception as e:
            cpu = {"error": repr(e)}
    del spr, C, Xd
    from openmeasure_b200 import engine as eng_mod
    eng_mod.release_scratch()
    torch.cuda.empty_cache()
    return {"workload": RECON["name"], "rows": F * n_c, "modes": r, "sensors": s, "device_resident": dev, "e2e": e2e,
            "cpu_baseline": cpu}


if __name__ == "__main__":
    main()
