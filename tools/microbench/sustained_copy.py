#!/usr/bin/env python
"""How much HBM bandwidth does a B200 sustain once the board sits at its power cap?

MEASURED_PEAKS.json's hbm_gbs is a burst figure (best of 10 copies).  The PCG loop runs for
seconds at the 1 kW cap; this measures a pure device-to-device copy (read + write bytes) as a
burst and back to back for a few seconds, alone and interleaved with an FP64-heavy kernel
that heats the SMs the way the operator kernel does."""
import subprocess
import time

import torch

dev = torch.device("cuda")
n = 1 << 29  # 4 GiB of doubles
a = torch.rand(n, dtype=torch.float64, device=dev)
b = torch.empty_like(a)


def smi():
    try:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw", "--format=csv,noheader,nounits"],
                           stdout=subprocess.PIPE, text=True, timeout=5).stdout.strip().splitlines()[0]
        return o
    except Exception:
        return "n/a"


def copies(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    return 2 * n * 8 * k / (e0.elapsed_time(e1) * 1e-3) / 1e9


for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize()
print("burst (best of 10 single copies): %.0f GB/s" % max(copies(1) for _ in range(10)), smi())
t0 = time.time()
while time.time() - t0 < 4.0:
    g = copies(40)
    print("  sustained window (40 copies): %.0f GB/s  [%s]" % (g, smi()), flush=True)
# SM-heavy phase in between: FP64 matmuls (DFMA pipe busy), then copies again
m = torch.rand(8192, 8192, dtype=torch.float64, device=dev)
t0 = time.time()
while time.time() - t0 < 4.0:
    for _ in range(4):
        torch.matmul(m, m)
    g = copies(20)
    print("  copies between FP64 GEMMs: %.0f GB/s  [%s]" % (g, smi()), flush=True)
