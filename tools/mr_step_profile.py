"""Where does a multi-rank fit + optimal_placement step spend its time?  (developer tool)
torchrun --nproc-per-node N tools/mr_step_profile.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from openmeasure_b200 import synth
from openmeasure_b200.sparse_sensing import SPR

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
F, n_c, m, r = 9, 183620, 41, 40
Xd = synth.snapshots(F, n_c * world, m, r, cell0=rank * n_c, ncell_loc=n_c)
torch.cuda.synchronize()


def seg(name, fn, log):
    t0 = time.perf_counter()
    out = fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    log.append(f"{name}: issue {1e3*(t1-t0):.3f} +sync {1e3*(t2-t0):.3f}")
    return out


for it in range(5):
    log = []
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    spr = seg("ctor", lambda: SPR.from_device(Xd, F, group=None), log)
    eng = spr._eng
    seg("stats", lambda: spr._scale_stats("std", 1, defer_row_means=True), log)
    G = seg("gram", lambda: eng.gram(), log)
    SV = seg("eig", lambda: eng.eig_pod(G), log)
    seg("bp", lambda: eng.backproject((SV[1][:, :r] / SV[0][:r]).contiguous()), log)
    seg("place", lambda: spr.optimal_placement(block=8), log)
    t1 = time.perf_counter()
    # the same step without intermediate syncs
    dist.barrier(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    spr = SPR.from_device(Xd, F, group=None)
    spr.fit(select_modes="number", n_modes=r)
    t3 = time.perf_counter()
    spr.optimal_placement(block=8)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    if it >= 3:
        print(f"[rank {rank}] " + " | ".join(log) + f" || step: fit {1e3*(t3-t2):.3f} place {1e3*(t4-t3):.3f} total {1e3*(t4-t2):.3f} ms", flush=True)
dist.destroy_process_group()
