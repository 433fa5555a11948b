#!/usr/bin/env python
"""The offered single-precision SampleMSC (g4hb200_set_msc_precision(h, 32)) against the FP64 path on the same batch:
how many tracks keep their discrete outcome, how far direction and displacement move, and what the step costs.
usage: python tools/msc_f32_probe.py [n_tracks]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from g4hepem_b200 import batches, engine as eng, tables  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
ft = tables.load_state_json(os.path.join(ROOT, "tests", "golden", "hepem_state.json"))
e = eng.Engine(ft, 0)
SEED = 2026
host = batches.make_electron_batch(n, ft.num_matcut, seed=31)
dev, sec = eng.ElectronDeviceBatch(n), eng.SecondaryDeviceQueue(2 * n)


def run(bits):
    e.set_msc_precision(bits)
    ts = []
    for _ in range(5):
        dev.upload(host)
        sec.reset()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.ElectronManager.Step(e, dev, sec, SEED)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return dev.download(), min(ts)


d64, t64 = run(64)
d32, t32 = run(32)
e.set_msc_precision(64)
out = compare = {}
same = (d64.meta == d32.meta).all(axis=1) & (d64.winner == d32.winner)
scattered = np.abs(d64.dirx_diry - host.dirx_diry).max(axis=1) > 0
out["tracks"] = n
out["ms_f64"], out["ms_f32"] = t64, t32
out["same_discrete_fraction"] = float(same.mean())
out["different_discrete"] = int((~same).sum())
for g in ("ekin_logekin", "gstep_pstep", "nia01", "nia23"):
    out["exact_" + g] = bool(np.array_equal(getattr(d64, g)[same], getattr(d32, g)[same], equal_nan=True))
dir64 = np.concatenate([d64.dirx_diry, d64.dirz_safety[:, :1]], axis=1)
dir32 = np.concatenate([d32.dirx_diry, d32.dirz_safety[:, :1]], axis=1)
dir0 = np.concatenate([host.dirx_diry, host.dirz_safety[:, :1]], axis=1)
dd = np.abs(dir64 - dir32).max(axis=1)[same]
out["dir_abs_max"], out["dir_abs_p999"], out["dir_abs_median"] = float(dd.max()), float(np.quantile(dd, 0.999)), float(np.median(dd))
# the deflection itself: 1 - cos between pre- and post-step direction, relative difference of the two precisions
omc64 = 1.0 - np.clip((dir64 * dir0).sum(axis=1), -1, 1)
omc32 = 1.0 - np.clip((dir32 * dir0).sum(axis=1), -1, 1)
norm32 = np.abs((dir32 ** 2).sum(axis=1) - 1.0)
out["unit_norm_err_max"] = float(norm32.max())
out["nan_f32"] = {g: int(np.isnan(getattr(d32, g)).sum()) for g in ("dirx_diry", "dirz_safety", "edep_dispx", "dispy_dispz")}
out["nan_f64"] = {g: int(np.isnan(getattr(d64, g)).sum()) for g in ("dirx_diry", "dirz_safety", "edep_dispx", "dispy_dispz")}
flagdiff = (d64.meta[:, 1] != d32.meta[:, 1])
drawdiff = (d64.meta[:, 3] != d32.meta[:, 3])
out["flag_diff"], out["draw_diff"], out["noscatter_diff"] = int(flagdiff.sum()), int(drawdiff.sum()), int((((d64.meta[:, 1] ^ d32.meta[:, 1]) & 0x20) != 0).sum())
disp64 = np.concatenate([d64.edep_dispx[:, 1:], d64.dispy_dispz], axis=1)
disp32 = np.concatenate([d32.edep_dispx[:, 1:], d32.dispy_dispz], axis=1)
dl = np.linalg.norm(disp64, axis=1)
has = same & (dl > 0) & np.isfinite(dl)
rel = np.linalg.norm(disp64 - disp32, axis=1)[has] / dl[has]
out["disp_rel_max"], out["disp_rel_p999"] = float(rel.max()), float(np.quantile(rel, 0.999))
ang = np.sqrt(np.maximum(2 * omc64, 0))
big = same & (ang > 1e-3)
rel_ang = np.abs(np.sqrt(np.maximum(2 * omc32[big], 0)) - ang[big]) / ang[big]
out["angle_rel_max_above_1mrad"], out["angle_rel_p999"] = float(rel_ang.max()), float(np.quantile(rel_ang, 0.999))
bad = np.where(np.isnan(d32.edep_dispx[:, 1]))[0][:6]
for k in bad:
    print("nan track", int(k), "pstep", d64.gstep_pstep[k], "tz", d64.tstep_zpath[k], "flags64", hex(d64.meta[k, 1]), "disp64", d64.edep_dispx[k, 1], d64.dispy_dispz[k],
          "disp32", d32.edep_dispx[k, 1], d32.dispy_dispz[k], "dir0", dir0[k], "dir64", dir64[k], "dir32", dir32[k], "ekin", host.ekin_logekin[k, 0], d64.ekin_logekin[k, 0], file=sys.stderr)
print(json.dumps(out))
