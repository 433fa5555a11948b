import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from g4hepem_b200 import batches, engine as eng, tables
ft = tables.load_state_json("tests/golden/hepem_state.json")
e = eng.Engine(ft, 0)
n = 1 << 20
g = batches.make_gamma_batch(n, ft.num_matcut, seed=2026)
gd = eng.GammaDeviceBatch(n); gs = eng.SecondaryDeviceQueue(2 * n)
for _ in range(2):
    gd.upload(g); gs.reset(); eng.GammaManager.Step(e, gd, gs, 2026); torch.cuda.synchronize()
print("done")
