#!/usr/bin/env python3
"""Measures pinned H2D / D2H / bidirectional copy bandwidth of the box (context for bench e2e)."""
import time
import torch

n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


print(f"H2D  {n / timed(h2d) / 1e9:.1f} GB/s")
print(f"D2H  {n / timed(d2h) / 1e9:.1f} GB/s")
print(f"both {2 * n / timed(both) / 1e9:.1f} GB/s aggregate")
