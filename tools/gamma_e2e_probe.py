#!/usr/bin/env python
"""g4hb200_gamma_step_host over a pinned 1M-photon batch (BASELINE configs[1] through host buffers)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from g4hepem_b200 import batches, engine as eng, tables
ft = tables.load_state_json("tests/golden/hepem_state.json")
e = eng.Engine(ft, 0)
n = 1 << 20
g = batches.make_gamma_batch(n, ft.num_matcut, seed=5, pinned=True)
work = batches.GammaHostBatch(n, pinned=True)
q = batches.SecondaryHostQueue(2 * n, pinned=True)
ts = []
for i in range(6):
    for grp in g.groups() + ("meta", "winner"):
        getattr(work, grp)[...] = getattr(g, grp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e.gamma_step_host(work, q, 2026)
    ts.append(time.perf_counter() - t0)
print("gamma_step_host 1M:", min(ts[2:]) * 1e3, "ms", n / min(ts[2:]) / 1e6, "M gamma-steps/s")
