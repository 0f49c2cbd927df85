"""Multi-rank placement timing (developer tool): torchrun --nproc-per-node N tools/mr_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from openmeasure_b200 import synth, engine as E, comm as Cm

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
F, n_c, m, r = 9, 183620, 41, 40
Xd = synth.snapshots(F, n_c * world, m, r, cell0=rank * n_c, ncell_loc=n_c)
eng = E.Engine(Xd, F, group=None)
eng.stats("std", 1, defer_row_means=True)
S, V = eng.eig_pod(eng.gram())
eng.backproject((V[:, :r] / S[:r]).contiguous())
torch.cuda.synchronize()
for mode in ("p2p", "nccl", "p2p"):
    os.environ["OMB_QR_EXCHANGE"] = mode
    for _ in range(2):
        piv, rd, gap = eng.qrcp(block=8)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        piv, rd, gap = eng.qrcp(block=8)
    e1.record(); torch.cuda.synchronize()
    eng.check_p2p()
    if rank == 0:
        print(f"exchange={mode:5s} used={eng.qr_exchange[:60]:60s} {e0.elapsed_time(e1)/5:8.3f} ms  piv[:4]={piv[:4].tolist()}", flush=True)
dist.destroy_process_group()
