"""Time the pivoted-QR placement alone (developer tool):
python tools/qr_profile.py [n_c] [m] [r] [blocks] [lazy alphas]      e.g.  1800000 256 100 8 0,0.9,0.94,0.97
Prints, per (block, alpha): ms, pivots equal to the first run, the bytes the read-only passes visited."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openmeasure_b200 import _lib, synth, engine as E

F = 9
n_c = int(sys.argv[1]) if len(sys.argv) > 1 else 183620
m = int(sys.argv[2]) if len(sys.argv) > 2 else 41
r = int(sys.argv[3]) if len(sys.argv) > 3 else 40
blocks = [int(b) for b in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 4, 8]
alphas = [float(a) for a in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0.0, 0.94]
from openmeasure_b200.sparse_sensing import SPR
spr = SPR.from_device(synth.snapshots(F, n_c, m, r), F, group=False)
spr.fit(select_modes="number", n_modes=r)
eng = spr._eng
torch.cuda.synchronize()
L = _lib.load()
ref = None
n = F * n_c
for b in blocks:
    for a in (alphas if b > 1 else [0.0]):
        L.omb_qrcp_set_lazy(a)
        for _ in range(2):
            piv, rd, gap = eng.qrcp(block=b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            piv, rd, gap = eng.qrcp(block=b)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        st = eng.qr_stats()
        p = piv.cpu()
        if ref is None: ref = p
        gb = (512 * st["seg_rows"] + 1536 * st["seg_visits"]) / 1e9
        print(f"block={b:2d} alpha={a:6.4f}  {ms:8.3f} ms  pivots_equal_first={bool((p == ref).all())}  min gap {gap.min().item():.2e}  "
              f"read-only passes {gb:8.2f} GB  catch-up rounds {st['retries']}", flush=True)
