#!/usr/bin/env python
"""Per output of the FP32 look-up kernel: largest relative / absolute deviation from the FP64 entry point (the numbers
behind the bound stated in include/g4hepem_b200.h)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from g4hepem_b200 import engine as eng, tables
ft = tables.load_state_json("tests/golden/hepem_state.json")
e = eng.Engine(ft, 0)
n = 1 << 20
rng = np.random.default_rng(0)
imcn = rng.integers(1, ft.num_matcut, n).astype(np.int32)
imc = torch.from_numpy(imcn).cuda()
ek32n = np.exp(rng.uniform(np.log(0.95e-4), np.log(1.02e8), n)).astype(np.float32)
ek64 = torch.from_numpy(ek32n.astype(np.float64)).cuda(); lek64 = torch.log(ek64)
want = e.electron_lookups(imc, ek64, lek64, True).cpu().numpy()
got = e.electron_lookups_f32(imc, torch.from_numpy(ek32n).cuda(), lek64.float(), True).cpu().numpy().astype(np.float64)
names = ["range", "dedx", "invrange", "ioni", "brem", "nuc", "tr1mfp"]
for k in range(7):
    w, g = want[k], got[k]
    scale = np.abs(w).max()
    rel = np.abs(g - w) / np.maximum(np.abs(w), 1e-300)
    big = np.abs(w) >= 1e-4 * scale
    i = np.argmax(np.where(big, rel, 0))
    print("%-9s scale %.3e  max rel (|w|>=1e-4 scale) %.3e at E=%.4e imc=%d w=%.6e g=%.6e ; nonfinite %d ; small-part max abs/scale %.2e ; p99.99 rel %.2e" % (
        names[k], scale, np.where(big, rel, 0).max(), ek32n[i], imcn[i], w[i], g[i], int((~np.isfinite(g)).sum()),
        (np.abs(g - w)[~big].max() / scale) if (~big).any() else 0.0, np.percentile(rel[big], 99.99)))
