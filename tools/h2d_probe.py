"""Host-to-device upload rate of a pinned buffer with 1, 2 and 4 concurrent copy streams (developer tool):
is one cudaMemcpyAsync stream enough to saturate the PCIe link the e2e number is bound by?"""
import sys, time
import torch
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
n = int(gb * (1 << 30) // 8)
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device="cuda")
for ns in (1, 2, 4, 1):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        per = n // ns
        for k, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[k * per:(k + 1) * per].copy_(h[k * per:(k + 1) * per], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{ns} stream(s): {8.0 * n / dt / 1e9:.1f} GB/s")
# device -> host
for ns in (1, 2):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    per = n // ns
    for k, s in enumerate(streams):
        with torch.cuda.stream(s):
            h[k * per:(k + 1) * per].copy_(d[k * per:(k + 1) * per], non_blocking=True)
    torch.cuda.synchronize()
    print(f"D2H {ns} stream(s): {8.0 * n / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
