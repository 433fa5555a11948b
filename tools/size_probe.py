#!/usr/bin/env python
"""Time of one fused e-/e+ step and one gamma step against the batch size, for the fused (one persistent launch) and the
staged (pipeline of stage kernels) paths: where the cross-over is.  usage: python tools/size_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from g4hepem_b200 import batches, engine as eng, tables  # noqa: E402

ft = tables.load_state_json(os.path.join(ROOT, "tests", "golden", "hepem_state.json"))
SEED = 2026
engines = {}
for mode in ("0", "1"):
    os.environ["G4HB200_FUSED"] = mode
    engines[mode] = eng.Engine(ft, 0)
del os.environ["G4HB200_FUSED"]
print(f"{'n':>9s} {'e staged':>10s} {'e fused':>10s} {'g staged':>10s} {'g fused':>10s}   (us per step, best of 7)")
for n in (256, 1024, 4096, 16384, 32768, 65536, 131072, 262144, 524288, 1048576):
    host = batches.make_electron_batch(n, ft.num_matcut, seed=SEED)
    ghost = batches.make_gamma_batch(n, ft.num_matcut, seed=SEED + 1)
    dev, gdev = eng.ElectronDeviceBatch(n), eng.GammaDeviceBatch(n)
    sec = eng.SecondaryDeviceQueue(2 * n)
    row = []
    for kind in ("e", "g"):
        for mode in ("0", "1"):
            e = engines[mode]
            ts = []
            for _ in range(7):
                if kind == "e":
                    dev.upload(host)
                else:
                    gdev.upload(ghost)
                sec.reset()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                if kind == "e":
                    eng.ElectronManager.Step(e, dev, sec, SEED)
                else:
                    eng.GammaManager.Step(e, gdev, sec, SEED)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            row.append(min(ts))
    print(f"{n:9d} {row[0]:10.1f} {row[1]:10.1f} {row[2]:10.1f} {row[3]:10.1f}", flush=True)
