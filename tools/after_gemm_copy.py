#!/usr/bin/env python
"""What the power-capped state after a long tensor-core kernel does to HBM-bound work (GPU box):
a 4 GiB device copy (torch, read + write) timed cold, and timed immediately behind ~150 ms of
back-to-back fp16 GEMMs (the clock / power state the stem steps of a slice run in)."""
import torch

dev = torch.device("cuda:0")
a = torch.empty(1 << 30, dtype=torch.float32, device=dev).normal_()
b = torch.empty_like(a)
x = torch.randn(8192, 8192, device=dev, dtype=torch.float16)
y = torch.randn(8192, 8192, device=dev, dtype=torch.float16)


def copy_ms(pre_gemms, n_copies=4):
    torch.cuda.synchronize()
    for _ in range(pre_gemms):
        x @ y
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_copies + 1)]
    evs[0].record()
    for i in range(n_copies):
        b.copy_(a)
        evs[i + 1].record()
    torch.cuda.synchronize()
    return [evs[i].elapsed_time(evs[i + 1]) for i in range(n_copies)]


for _ in range(2):
    copy_ms(0)
gb = 2 * a.numel() * 4 / 1e9
for pre in (0, 50, 200, 400):
    ms = copy_ms(pre)
    print(f"copy of {gb:.1f} GB behind {pre:3d} GEMMs (8192^3 fp16): " + "  ".join(f"{gb / (t * 1e-3) / 1e3:.2f} TB/s" for t in ms), flush=True)
