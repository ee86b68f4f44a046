#!/usr/bin/env python3
"""Raw pinned-memory PCIe bandwidth of this box (one GPU): H2D alone, D2H alone, both at once on two streams - the ceiling of bench.py's
e2e leg, which moves 480 MB each way per step.  Run on the GPU box: python tools/pcie_check.py"""
import time

import torch

n = 480 * 1000 * 1000 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return n * 8 * reps / (time.perf_counter() - t0) / 1e9


run(True, True, 2)
print(f"H2D alone {run(True, False):.1f} GB/s, D2H alone {run(False, True):.1f} GB/s, both at once {run(True, True):.1f} GB/s each way")
