"""Small all-gather latency: peer-memory kernel vs NCCL (developer tool, torchrun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from openmeasure_b200 import comm as Cm, _lib

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
_lib.load()
c = Cm.TorchDistComm(None)
for n in (36, 1681, 4096):
    x = torch.full((n,), float(rank + 1), dtype=torch.float64, device="cuda")
    for mode in ("p2p", "nccl"):
        os.environ["OMB_SMALL_ALLGATHER"] = mode
        for _ in range(5):
            out = c.allgather(x)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            out = c.allgather(x)
        e1.record(); torch.cuda.synchronize()
        ok = bool((out == torch.arange(1, world + 1, device="cuda", dtype=torch.float64)[:, None]).all())
        if rank == 0:
            print(f"n={n:5d} {mode:5s} {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us/call  ok={ok}", flush=True)
dist.destroy_process_group()
