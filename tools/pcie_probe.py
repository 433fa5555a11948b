#!/usr/bin/env python
"""Raw PCIe bandwidth of the box with pinned buffers (one direction, both at once, 4 MB pieces): the ceiling of the
host-buffer entry points.  B200 box of this pool: H2D 55, D2H 57, both at once 46.6 each, 4 MB pieces 54 GB/s."""
import time

import torch

n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(f, reps=5):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def both():
    h2d()
    d2h()


m = 4 << 20


def d2h_pieces():
    with torch.cuda.stream(s2):
        for k in range(n // m):
            h2[k * m:(k + 1) * m].copy_(d2[k * m:(k + 1) * m], non_blocking=True)


print("H2D GB/s", n / timed(h2d) / 1e9)
print("D2H GB/s", n / timed(d2h) / 1e9)
print("both at once, each GB/s", n / timed(both) / 1e9)
print("D2H in 4 MB pieces GB/s", n / timed(d2h_pieces, 3) / 1e9)
