#!/usr/bin/env python
"""Accuracy and speed of one build flavour of the library (G4HB200_LIB) against the compiled reference.

For `batches` seeded batches of n e-/e+ (fused step) and n gammas (fused step): the number of flipped discrete
outcomes (winner process, flags, draw counts, secondary counts / kinds / parents) and the largest relative error
of every real field, against the reference run on the same batch with the same uniform stream.
usage: G4HB200_LIB=... python tools/flavour_probe.py [n] [batches] [--json out.json]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from g4hepem_b200 import _capi, batches, engine as eng, tables  # noqa: E402
from oracle import ref as oref  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if len(args) > 0 else 1 << 20
nb = int(args[1]) if len(args) > 1 else 4
out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
STATE = os.path.join(ROOT, "tests", "golden", "hepem_state.json")
ft = tables.load_state_json(STATE)
e = eng.Engine(ft, 0)
ref = oref.Reference(STATE)
threads = ref.hardware_threads()
SEED = 2026


def relerr(a, b, scale=None):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.abs(a - b)
    s = np.maximum(np.abs(a), np.abs(b)) if scale is None else scale
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(d == 0, 0.0, d / s)
    r = np.where(np.isnan(a) & np.isnan(b), 0.0, r)
    return float(np.nanmax(r)) if r.size else 0.0


def sec_cmp(qa, qb, acc):
    ra, rb = qa.sorted_records(), qb.sorted_records()
    if len(ra["ekin"]) != len(rb["ekin"]):
        acc["flips.sec_count"] = acc.get("flips.sec_count", 0) + abs(len(ra["ekin"]) - len(rb["ekin"]))
        return
    for k in ("parent_index", "slot", "parent_id", "kind"):
        acc["flips.sec_" + k] = acc.get("flips.sec_" + k, 0) + int((ra[k] != rb[k]).sum())
    acc["rel.sec_ekin"] = max(acc.get("rel.sec_ekin", 0.0), relerr(ra["ekin"], rb["ekin"]))
    acc["abs.sec_dir"] = max(acc.get("abs.sec_dir", 0.0), float(np.abs(ra["dir"] - rb["dir"]).max()) if len(ra["ekin"]) else 0.0)
    acc["secondaries"] = acc.get("secondaries", 0) + len(ra["ekin"])


acc_e, acc_g = {}, {}
t_gpu_e = []
t_gpu_g = []
dev = eng.ElectronDeviceBatch(n)
gdev = eng.GammaDeviceBatch(n)
sec = eng.SecondaryDeviceQueue(2 * n)
t0 = time.time()
for k in range(nb):
    host = batches.make_electron_batch(n, ft.num_matcut, seed=1000 + k)
    want = host.copy()
    qwant = batches.SecondaryHostQueue(2 * n)
    ref.electron_step(want, qwant, SEED, threads)
    dev.upload(host)
    sec.reset()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.ElectronManager.Step(e, dev, sec, SEED)
    b.record()
    torch.cuda.synchronize()
    t_gpu_e.append(a.elapsed_time(b))
    got = dev.download()
    for col, nm in enumerate(("imc", "flags", "id", "draws")):
        acc_e["flips.meta." + nm] = acc_e.get("flips.meta." + nm, 0) + int((want.meta[:, col] != got.meta[:, col]).sum())
    acc_e["flips.winner"] = acc_e.get("flips.winner", 0) + int((want.winner != got.winner).sum())
    same = (want.meta[:, 3] == got.meta[:, 3]) & (want.winner == got.winner) & (want.meta[:, 1] == got.meta[:, 1])
    for g in ("ekin_logekin", "nia01", "nia23", "msc_irange_dynrf", "gstep_pstep"):
        for c in (0, 1):
            if g == "ekin_logekin" and c == 1:
                continue
            key = f"rel.{g}[{c}]"
            acc_e[key] = max(acc_e.get(key, 0.0), relerr(getattr(want, g)[same, c], getattr(got, g)[same, c]))
    acc_e["rel.edep"] = max(acc_e.get("rel.edep", 0.0), relerr(want.edep_dispx[same, 0], got.edep_dispx[same, 0]))
    acc_e["rel.tlimitmin"] = max(acc_e.get("rel.tlimitmin", 0.0), relerr(want.msc_tlimmin_gauss[same, 0], got.msc_tlimmin_gauss[same, 0]))
    da = np.stack([want.dirx_diry[:, 0], want.dirx_diry[:, 1], want.dirz_safety[:, 0]], 1)[same]
    db = np.stack([got.dirx_diry[:, 0], got.dirx_diry[:, 1], got.dirz_safety[:, 0]], 1)[same]
    acc_e["abs.dir"] = max(acc_e.get("abs.dir", 0.0), float(np.abs(da - db).max()))
    pa = np.stack([want.edep_dispx[:, 1], want.dispy_dispz[:, 0], want.dispy_dispz[:, 1]], 1)[same]
    pb = np.stack([got.edep_dispx[:, 1], got.dispy_dispz[:, 0], got.dispy_dispz[:, 1]], 1)[same]
    nrm = np.maximum(np.linalg.norm(pa, axis=1), 1e-300)
    acc_e["rel.displacement"] = max(acc_e.get("rel.displacement", 0.0), float((np.abs(pa - pb).max(axis=1) / nrm).max()))
    sec_cmp(qwant, sec.download(), acc_e)

    ghost = batches.make_gamma_batch(n, ft.num_matcut, seed=5000 + k)
    gwant = ghost.copy()
    gq = batches.SecondaryHostQueue(2 * n)
    ref.gamma_step(gwant, gq, SEED, threads)
    gdev.upload(ghost)
    sec.reset()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.GammaManager.Step(e, gdev, sec, SEED)
    b.record()
    torch.cuda.synchronize()
    t_gpu_g.append(a.elapsed_time(b))
    ggot = gdev.download()
    for col, nm in enumerate(("imc", "flags", "id", "draws")):
        acc_g["flips.meta." + nm] = acc_g.get("flips.meta." + nm, 0) + int((gwant.meta[:, col] != ggot.meta[:, col]).sum())
    acc_g["flips.winner"] = acc_g.get("flips.winner", 0) + int((gwant.winner != ggot.winner).sum())
    same = (gwant.meta[:, 3] == ggot.meta[:, 3]) & (gwant.winner == ggot.winner)
    for g, c in (("ekin_logekin", 0), ("dirz_nia0", 1), ("gstep_mfp0", 0), ("gstep_mfp0", 1), ("edep_pemxsec", 0)):
        key = f"rel.{g}[{c}]"
        acc_g[key] = max(acc_g.get(key, 0.0), relerr(getattr(gwant, g)[same, c], getattr(ggot, g)[same, c]))
    da = np.stack([gwant.dirx_diry[:, 0], gwant.dirx_diry[:, 1], gwant.dirz_nia0[:, 0]], 1)[same]
    db = np.stack([ggot.dirx_diry[:, 0], ggot.dirx_diry[:, 1], ggot.dirz_nia0[:, 0]], 1)[same]
    acc_g["abs.dir"] = max(acc_g.get("abs.dir", 0.0), float(np.abs(da - db).max()))
    sec_cmp(gq, sec.download(), acc_g)

rep = dict(
    lib=os.path.basename(_capi.LIB_PATH), tracks_per_batch=n, batches=nb, track_steps=n * nb,
    electron=dict(acc_e, ms_per_batch_min=min(t_gpu_e), ms_per_batch_median=float(np.median(t_gpu_e))),
    gamma=dict(acc_g, ms_per_batch_min=min(t_gpu_g), ms_per_batch_median=float(np.median(t_gpu_g))),
    wall_s=time.time() - t0,
)
print(json.dumps(rep, indent=1))
if out_json:
    with open(out_json, "w") as f:
        json.dump(rep, f, indent=1)
