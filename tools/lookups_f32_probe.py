#!/usr/bin/env python
"""FP64 against FP32 look-up sets (BASELINE configs[0]) on the same 1M inputs: CUDA-event time per call."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from g4hepem_b200 import engine as eng, tables
ft = tables.load_state_json("tests/golden/hepem_state.json")
e = eng.Engine(ft, 0)
n = 1 << 20
rng = np.random.default_rng(0)
imc = torch.from_numpy(rng.integers(1, ft.num_matcut, n).astype(np.int32)).cuda()
ek = np.exp(rng.uniform(np.log(0.95e-4), np.log(1.02e8), n))
ek64 = torch.from_numpy(ek).cuda(); lek64 = torch.from_numpy(np.log(ek)).cuda()
ek32 = ek64.float(); lek32 = lek64.float()
def timed(f):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 50
t64 = timed(lambda: e.electron_lookups(imc, ek64, lek64, True))
t32 = timed(lambda: e.electron_lookups_f32(imc, ek32, lek32, True))
print("lookups f64 %.4f ms (%.3g /s)  f32 %.4f ms (%.3g /s)" % (t64, n / t64 * 1e3, t32, n / t32 * 1e3))
