#!/usr/bin/env python
"""Time g4hb200_electron_step_host (pinned host batch in/out) for the chunk size in G4HB200_HOST_CHUNK."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from g4hepem_b200 import batches, engine as eng, tables
n = 1 << 20
ft = tables.load_state_json(os.path.join(ROOT, "tests", "golden", "hepem_state.json"))
e = eng.Engine(ft, 0)
pristine = batches.make_electron_batch(n, ft.num_matcut, seed=2026, pinned=True)
work = batches.ElectronHostBatch(n, pinned=True)
hsec = batches.SecondaryHostQueue(2 * n, pinned=True)
groups = batches.ElectronHostBatch.PAIR_GROUPS + ("meta",)
ts = []
for i in range(8):
    for g in groups:
        getattr(work, g)[...] = getattr(pristine, g)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e.electron_step_host(work, hsec, 2026)
    ts.append(time.perf_counter() - t0)
print("chunk", os.environ.get("G4HB200_HOST_CHUNK", "default"), "ms/step", 1e3 * min(ts[2:]), "M steps/s", n / min(ts[2:]) / 1e6, "nsec", int(hsec.count[0]))
