"""Multi-GPU parity check (torchrun --nproc-per-node N tools/mr_check.py): the row-sharded path over
real NVLink peers (peer-memory all-gathers + in-kernel pivot exchange) must reproduce the single-GPU
result of the same global problem: pivots identical, sigma to 1e-12, Theta and reconstructions to 1e-9."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from openmeasure_b200 import synth
from openmeasure_b200.sparse_sensing import SPR

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n_c_loc = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 41
r = int(sys.argv[3]) if len(sys.argv) > 3 else 40
F = 9 if m != 1024 else 8
n_c = n_c_loc * world
Xl = synth.snapshots(F, n_c, m, r, cell0=rank * n_c_loc, ncell_loc=n_c_loc)
spr = SPR.from_device(Xl, F, group=None)
spr.fit(select_modes="number", n_modes=r)
C = spr.optimal_placement(block=8)
spr.train(C)
y = np.zeros((r, 3))
ok = True
qst = spr._eng.qr_stats()
if rank == 0:
    from openmeasure_b200 import _lib
    _lib.load().omb_qrcp_set_lazy(0.0)                      # the single-GPU run is the EAGER schedule
    Xg = synth.snapshots(F, n_c, m, r)                      # the whole problem on one GPU
    one = SPR.from_device(Xg, F, group=False)
    one.fit(select_modes="number", n_modes=r)
    C1 = one.optimal_placement(block=8)
    one.train(C1)
    same_piv = bool(np.array_equal(C.pivots, C1.pivots))
    ds = float(np.max(np.abs(spr.Sigma_r - one.Sigma_r) / one.Sigma_r))
    dth = float(np.max(np.abs(np.abs(spr.Theta) - np.abs(one.Theta))))
    ok = same_piv and ds < 1e-12 and dth < 1e-9
    print(f"world={world} rows={F*n_c} m={m} r={r} exchange={spr._eng.qr_exchange} p2p_allgathers={getattr(spr._eng.comm, 'p2p_collectives', 0)} "
          f"pivots_identical={same_piv} max_rel_dsigma={ds:.2e} max_dTheta={dth:.2e} min_gap={spr.qr_gap.min():.2e} "
          f"lazy={qst['lazy']} catch_up_rounds={qst['retries']} -> {'OK' if ok else 'FAIL'}",
          flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
