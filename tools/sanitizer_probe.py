#!/usr/bin/env python
"""Small runs of every pipeline (fused step, HowFar + Perform, gamma step, showers with and without Woodcock tracking)
for `compute-sanitizer --tool memcheck|racecheck python tools/sanitizer_probe.py` (r01c on the B200: 0 errors, 0 hazards)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from g4hepem_b200 import batches, engine as eng, tables, shower, _capi
ft = tables.load_state_json("tests/golden/hepem_state.json")
e = eng.Engine(ft, 0)
n = 20000
host = batches.make_electron_batch(n, ft.num_matcut, seed=3)
dev = eng.ElectronDeviceBatch(n); sec = eng.SecondaryDeviceQueue(2 * n)
dev.upload(host); eng.ElectronManager.Step(e, dev, sec, 2026); torch.cuda.synchronize()
dev.upload(host); sec.reset(); eng.ElectronManager.HowFar(e, dev, 2026); eng.ElectronManager.Perform(e, dev, sec, 2026); torch.cuda.synchronize()
g = batches.make_gamma_batch(n, ft.num_matcut, seed=4)
gd = eng.GammaDeviceBatch(n); gs = eng.SecondaryDeviceQueue(2 * n)
gd.upload(g); eng.GammaManager.Step(e, gd, gs, 2026); torch.cuda.synchronize()
for w in (False, True):
    r = shower.run(e, shower.SlabCalorimeter(woodcock=w), 4, 200.0, 2026, capacity=1 << 15)
    print("shower", w, r.stats["num_steps"], float(r.edep.sum()))
imc = torch.from_numpy(np.random.default_rng(1).integers(1, ft.num_matcut, n).astype(np.int32)).cuda()
ek = torch.from_numpy(np.exp(np.random.default_rng(2).uniform(np.log(1e-4), np.log(1e8), n))).cuda()
r64 = e.electron_lookups(imc, ek, torch.log(ek), True)
r32 = e.electron_lookups_f32(imc, ek.float(), torch.log(ek).float(), False)
torch.cuda.synchronize()
print("lookups", float(r64[0].sum()), float(r32[0].double().sum()))
# the offered / opt-in paths: single-precision SampleMSC, lane-refill samplers (second engine: the switch is read at creation)
e.set_msc_precision(32)
dev.upload(host); sec.reset(); eng.ElectronManager.Step(e, dev, sec, 2026); torch.cuda.synchronize()
e.set_msc_precision(64)
os.environ["G4HB200_REFILL"] = "2"
e2 = eng.Engine(ft, 0)
del os.environ["G4HB200_REFILL"]
dev.upload(host); sec.reset(); eng.ElectronManager.Step(e2, dev, sec, 2026); torch.cuda.synchronize()
gd.upload(g); gs.reset(); eng.GammaManager.Step(e2, gd, gs, 2026); torch.cuda.synchronize()
print("variants", int(sec.count[0].item()), int(gs.count[0].item()))
print("done")
