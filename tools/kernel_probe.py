#!/usr/bin/env python
"""Launch every stepping kernel once on a 1M batch (for `ncu` captures) and print CUDA-event timings.
usage: python tools/kernel_probe.py [n_tracks] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from g4hepem_b200 import batches, engine as eng, tables  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ft = tables.load_state_json(os.path.join(ROOT, "tests", "golden", "hepem_state.json"))
e = eng.Engine(ft, 0)
SEED = 2026


def timed(name, fn, prep):
    ts = []
    for _ in range(reps):
        prep()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{name:28s} {min(ts):9.3f} ms   {n / min(ts) / 1e3:10.1f} M track-steps/s", flush=True)


host = batches.make_electron_batch(n, ft.num_matcut, seed=SEED)
if os.environ.get("PROBE_SORT"):
    # what a (couple, charge, energy)-sorted queue would give the kernels
    mode = os.environ["PROBE_SORT"]
    ebin = np.floor(np.log(host.ekin_logekin[:, 0]) * float(os.environ.get("PROBE_EBINS", "4"))).astype(np.int64)
    pos = (host.meta[:, 1] & 1).astype(np.int64)
    imc = host.meta[:, 0].astype(np.int64)
    key = {"full": (imc * 2 + pos) * 4096 + (ebin + 2048), "energy": ebin, "couple": imc * 2 + pos,
           "exact": None}[mode]
    order = np.argsort(host.ekin_logekin[:, 0] + 1e9 * (imc * 2 + pos), kind="stable") if key is None else np.argsort(key, kind="stable")
    for g in host.groups() + ("meta", "winner"):
        getattr(host, g)[...] = getattr(host, g)[order]
    print("sorted by", mode, flush=True)
dev = eng.ElectronDeviceBatch(n)
sec = eng.SecondaryDeviceQueue(2 * n)


def prep_fresh():
    dev.upload(host)
    sec.reset()


timed("electron_step (fused)", lambda: eng.ElectronManager.Step(e, dev, sec, SEED), prep_fresh)
if os.environ.get("PROBE_STAGES"):
    prep_fresh()
    torch.cuda.synchronize()
    e.set_kernel_timing(True)
    e.kernel_times()
    eng.ElectronManager.Step(e, dev, sec, SEED)
    torch.cuda.synchronize()
    for name, (ms, nl, items) in e.kernel_times().items():
        if nl:
            print(f"   stage {name:32s} {ms * 1000:8.1f} us  {items:9d} tracks", flush=True)
    e.set_kernel_timing(False)
timed("electron_howfar", lambda: eng.ElectronManager.HowFar(e, dev, SEED), prep_fresh)


def prep_after_howfar():
    prep_fresh()
    eng.ElectronManager.HowFar(e, dev, SEED)


timed("electron_perform", lambda: eng.ElectronManager.Perform(e, dev, sec, SEED), prep_after_howfar)

ghost = batches.make_gamma_batch(n, ft.num_matcut, seed=SEED + 1)
gdev = eng.GammaDeviceBatch(n)


def gprep():
    gdev.upload(ghost)
    sec.reset()


timed("gamma_step (fused)", lambda: eng.GammaManager.Step(e, gdev, sec, SEED), gprep)
timed("gamma_howfar", lambda: eng.GammaManager.HowFar(e, gdev, SEED), gprep)

rng = np.random.default_rng(0)
imc = torch.from_numpy(rng.integers(0, ft.num_matcut, n).astype(np.int32)).cuda()
ek = np.exp(rng.uniform(np.log(0.95e-4), np.log(1.02e8), n))
ekin = torch.from_numpy(ek).cuda()
lek = torch.from_numpy(np.log(ek)).cuda()
out = torch.empty((7, n), dtype=torch.float64, device="cuda")
timed("electron_lookups (config 1)", lambda: e.electron_lookups_into(imc, ekin, lek, out), lambda: None)
x = torch.from_numpy(ek).cuda()
timed("vdt_log_exp", lambda: e.vdt_log_exp(x), lambda: None)
