#!/usr/bin/env python
"""Side measurements of the other BASELINE.json configs (bench.py keeps the headline, configs[2]):

  configs[0]  1M e- look-ups (range, dE/dx, inverse range, sigma_ioni, sigma_brem, sigma_nuc, lambda_1)
  configs[1]  1M gamma full step (HowFar + SelectInteraction + Perform: Compton / conversion / photoelectric)

One JSON line per config: device-resident rate (CUDA events, inputs rotated through a ring larger than L2),
the HBM roofline of the call (algorithmic bytes / measured time / measured peak) and the compiled reference
(oracle/_ref) on all host threads over the same inputs.   usage: python tools/bench_configs.py [--steps K]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STATE_JSON = os.path.join(ROOT, "tests", "golden", "hepem_state.json")
SEED = 2026


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def main():
    import torch

    from g4hepem_b200 import batches, engine as eng, tables
    from oracle import checker

    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--tracks", type=int, default=1 << 20)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    n, steps = args.tracks, args.steps
    ft = tables.load_state_json(STATE_JSON)
    e = eng.Engine(ft, 0)
    ora = None if args.no_cpu else checker.best_available(STATE_JSON)
    threads = ora.hardware_threads() if ora is not None else 0
    hbm = peak()

    def timed(fn, reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # ---- configs[0]: look-ups (testing/ElectronXSections, testing/ElectronEnergyLoss style inputs) ----------------
    rng = np.random.default_rng(0)
    ring = 6  # 6 x (20 MB in + 56 MB out) > L2
    sets = []
    for _ in range(ring):
        imc = rng.integers(0, ft.num_matcut, n).astype(np.int32)
        ek = np.exp(rng.uniform(np.log(0.95e-4), np.log(1.02e8), n))
        sets.append((imc, ek, np.log(ek)))
    dsets = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(c).cuda(),
              torch.empty((7, n), dtype=torch.float64, device="cuda")) for a, b, c in sets]
    for i in range(3):
        e.electron_lookups_into(*dsets[i][:3], dsets[i][3])
    ms = timed(lambda i: e.electron_lookups_into(*dsets[i % ring][:3], dsets[i % ring][3]), steps)
    bytes_per = 4 + 8 + 8 + 7 * 8
    line = {"config": "BASELINE configs[0]: 1M e- look-ups (range, dE/dx, inv-range, sigma ioni/brem/nuclear, lambda_1)",
            "metric": "e- look-up sets/s", "value": n / (ms * 1e-3), "ms_per_call": ms, "n": n, "dtype": "f64",
            "roofline": {"bound": "hbm", "achieved": n * bytes_per / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": n * bytes_per / (ms * 1e-3) / 1e9 / hbm, "bytes_per_track": bytes_per}}
    if ora is not None:
        t0 = time.perf_counter()
        ora.electron_lookups(*sets[0], True)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": "look-up sets/s", "cores": 1, "kind": ora.kind,
                                "sample": f"one pass over the same {n} inputs, single thread"}
    print(json.dumps(line), flush=True)

    # ---- configs[1]: gamma step ----------------------------------------------------------------------------------------
    pristine = batches.make_gamma_batch(n, ft.num_matcut, seed=SEED + 1)
    ring = steps + 3
    devs = []
    for _ in range(ring):
        d = eng.GammaDeviceBatch(n)
        d.upload(pristine)
        devs.append(d)
    sec = eng.SecondaryDeviceQueue(2 * n)
    for i in range(3):
        sec.reset()
        eng.GammaManager.Step(e, devs[i], sec, SEED)

    def gstep(i):
        sec.reset()
        eng.GammaManager.Step(e, devs[3 + i], sec, SEED)

    ms = timed(gstep, steps)
    n_sec = int(sec.count[0].item())
    # read 3 groups + meta (64 B); written: 5 groups + meta + winner (100 B); 48 B per secondary
    alg = n * (64 + 100) + n_sec * 48
    e.set_kernel_timing(True)
    e.kernel_times()
    d = eng.GammaDeviceBatch(n)
    d.upload(pristine)
    sec.reset()
    eng.GammaManager.Step(e, d, sec, SEED)
    torch.cuda.synchronize()
    stages = {k: {"ms": v[0] / v[1], "tracks": v[2] / v[1]} for k, v in e.kernel_times().items() if v[1]}
    e.set_kernel_timing(False)
    line = {"config": "BASELINE configs[1]: 1M gamma step (macroscopic xs, element selector, Klein-Nishina / Bethe-Heitler / "
                      "photoelectric), E log-uniform 100 eV-100 TeV",
            "metric": "gamma track-steps/s", "value": n / (ms * 1e-3), "ms_per_step": ms, "n": n, "dtype": "f64",
            "secondaries_per_step": n_sec, "stages": stages,
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes_per_step": alg}}
    if ora is not None:
        work = pristine.copy()
        hsec = batches.SecondaryHostQueue(2 * n)
        t0 = time.perf_counter()
        ora.gamma_step(work, hsec, SEED, threads)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": "track-steps/s", "cores": threads, "kind": ora.kind,
                                "sample": f"one pass of HowFar+SelectInteraction+Perform over the same {n} gammas, std::thread x {threads}"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
