#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (contract: one JSON line on stdout from rank 0).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 5 --warmup 1

Workload (BASELINE.json configs[2], the one the north-star target "e- HowFar+Perform track-steps/s per B200"
is quoted on): a batch of 1M e-/e+ tracks, 50/50, E log-uniform 1 keV-100 GeV, couples uniform over the table
set, isotropic directions, safety U[0,1mm], 10 % on a boundary, first-step state.  One "step" = one fused
HowFar + Perform pass (g4hb200_electron_step) over one such batch = 1M track-steps.  Every step of the timed
region gets its own pristine device batch out of a ring (K x 276 MB >> the 126 MB L2, so inputs are always
cold and every step does the same work); at N GPUs every rank owns its own ring (weak scaling, no data-path
collective; one tiny allreduce of the deposited-energy sum after the timed region, as TestEm3's Run::Merge).

`value`   : device-resident track-steps/s (CUDA events on the launch stream, max over ranks).
`e2e`     : the same metric through the host-buffer C-ABI call g4hb200_electron_step_host with pinned HOST
            batches: H2D of the persistent groups, the kernel, D2H of state + results + secondaries, per step.
`roofline`: HBM; algorithmic bytes = 128 B read + 180 B written per track-step + 48 B per secondary.
`cpu_baseline` / `--impl reference`: the unmodified reference (oracle/_ref) on the host cores, same batch.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STATE_JSON = os.path.join(ROOT, "tests", "golden", "hepem_state.json")
SEED = 2026
METRIC = "e-/e+ HowFar+Perform track-steps/s"
UNIT = "track-steps/s"
READ_BYTES = 7 * 16 + 16           # 7 persistent pair groups + meta
WRITE_BYTES = 10 * 16 + 16 + 4     # persistent + 3 result groups + meta + winner
SEC_BYTES = 16 + 16 + 8 + 8        # one secondary record
# algorithmic bytes per track of every pipeline stage: (read, written, secondaries appended); DESIGN.md par. 4
STAGE_BYTES = {
    # (bytes read, bytes written, secondaries) per track a stage processes; 16 B per {a,b} group, 16 B meta, 4 B winner
    "ElHowFarXSKernel": (3 * 16 + 16, 7 * 16 + 16 + 4, 0),
    "ElHowFarMSCKernel": (7 * 16 + 16 + 4, 8 * 16 + 16 + 4, 0),
    "ElHowFarMSCRangeKernel": (4 + 3 * 16 + 16 + 4, 4 * 16 + 4, 0),
    "ElAlongStepKernel": (11 * 16 + 16 + 4, 7 * 16, 0),
    # fused step: 6 persistent groups + meta in; those + 3 result groups + winner + 4 hand-over groups + pre-step out
    "ElStepHeadKernel": (6 * 16 + 16, 5 * 16 + 16 + 3 * 16 + 4 + 4 * 16 + 16, 0),
    "ElMSCSampleKernel<e->": (4 + 10 * 16 + 16 + 4, 4 * 16 + 16, 0),
    "ElMSCSampleKernel<e+>": (4 + 10 * 16 + 16 + 4, 4 * 16 + 16, 0),
    "ElFluctuationKernel": (4 + 16 + 3 * 16 + 4, 3 * 16 + 16, 0),
    "ElDiscreteKernel": (4 + 16 + 16 + 4 + 16 + 16, 3 * 16, 0),
    "ElSamplerKernel<Moller>": (4 + 16 + 3 * 16, 3 * 16 + 16, 1),
    "ElSamplerKernel<Bhabha>": (4 + 16 + 3 * 16, 3 * 16 + 16, 1),
    "ElSamplerKernel<SeltzerBerger>": (4 + 16 + 3 * 16, 3 * 16 + 16, 1),
    "ElSamplerKernel<RelBrem>": (4 + 16 + 3 * 16, 3 * 16 + 16, 1),
    "ElSamplerKernel<Annihilation>": (4 + 16 + 3 * 16, 3 * 16 + 16, 2),
    "ElSamplerKernel<AtRest>": (4 + 16, 16, 2),
    # the whole step as one persistent launch (g4h_fused.cuh): the caller-visible state in and out; secondaries are added
    # from the queue counter of the run
    "ElFusedStepKernel": (READ_BYTES, WRITE_BYTES, 0),
}


def _workload_config(n_tracks):
    return {
        "workload": "BASELINE configs[2]: e-/e+ fused HowFar+Perform step (eloss fluctuation, Urban MSC, "
                    "Moller/Bhabha, SB+RB brem, annihilation), 50/50 e-/e+, E log-uniform 1 keV-100 GeV",
        "tracks_per_step_per_gpu": n_tracks,
        "couples": "synthetic ATLASbar-shaped set (Galactic, Pb, lAr x2 regions) + PbWO4 + water, "
                   "tests/golden/hepem_state.json",
        "cache": "every timed step works on its own pristine batch out of a ring of batches (inputs larger than L2); "
                 "the reference arm: its own host copy of the batch per step",
        "seed": SEED,
    }


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.tmp.close()
        os.unlink(self.tmp.name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(smax))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


class NvmlClockSampler:
    """SM clock and throttle reasons polled in-process through NVML every ~1 ms: the timed region of the default run is
    ~10 ms, which `nvidia-smi -lms 100` (ClockSampler, the fallback) sees once at best."""

    def __init__(self, index):
        import threading

        import pynvml

        self.nv = pynvml
        pynvml.nvmlInit()
        self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        self.samples = []
        self.reasons = 0
        self.running = True
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.running:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.reasons |= int(get_reasons(self.handle))
            except nv.NVMLError:
                pass
            time.sleep(0.001)

    def stop(self):
        self.running = False
        self.thread.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.smax,
               "reasons": sorted(k for k, bit in names.items() if self.reasons & bit), "samples": len(self.samples),
               "source": "nvml, 1 ms poll over the timed region"}
        try:
            nv.nvmlShutdown()
        except nv.NVMLError:
            pass
        return out


_REAL_STDOUT = None


def _emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def make_clock_sampler(index):
    try:
        return NvmlClockSampler(index)
    except Exception:  # no pynvml / no NVML: fall back to the nvidia-smi poller
        return ClockSampler(index)


def _spread_over_cores(local, world):
    """One slice of the allowed cores per rank (rank r of N gets the r-th N-th of them): eight ranks left on the same cores
    -- and their pinned staging buffers on the memory of one socket -- is what held the 8-GPU host-buffer number at 1.9x
    one GPU in round 1 (SCALE_r01).  Called before anything allocates pinned memory (first touch decides its placement)."""
    if world <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return [mine[0], mine[-1]]
    except OSError:
        return None


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _cpu_reference_rate(n_tracks, reps, id_offset=0):
    """The unmodified reference's HowFar+Perform over the same batch on all host threads; best of `reps`."""
    from g4hepem_b200 import batches, tables
    from oracle import checker

    ft = tables.load_state_json(STATE_JSON)
    ora = checker.best_available(STATE_JSON)
    threads = ora.hardware_threads() if hasattr(ora, "hardware_threads") else (os.cpu_count() or 1)
    pristine = batches.make_electron_batch(n_tracks, ft.num_matcut, seed=SEED, id_offset=id_offset)
    times = []
    for _ in range(reps):
        work = pristine.copy()
        sec = batches.SecondaryHostQueue(2 * n_tracks)
        t0 = time.perf_counter()
        ora.electron_step(work, sec, SEED, threads)
        times.append(time.perf_counter() - t0)
    return n_tracks / min(times), threads, ora.kind, times


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, rank 0 only."""
    rank, world, _ = _dist_env()
    if rank != 0:
        return 0
    n = args.tracks
    from g4hepem_b200 import batches, tables
    from oracle import checker

    ft = tables.load_state_json(STATE_JSON)
    ora = checker.best_available(STATE_JSON)
    threads = ora.hardware_threads()
    pristine = batches.make_electron_batch(n, ft.num_matcut, seed=SEED)
    works = [pristine.copy() for _ in range(args.warmup + args.steps)]
    secs = [batches.SecondaryHostQueue(2 * n) for _ in range(args.warmup + args.steps)]
    for i in range(args.warmup):
        ora.electron_step(works[i], secs[i], SEED, threads)
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        ora.electron_step(works[i], secs[i], SEED, threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _workload_config(n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": ora.kind,
                         "sample": f"{args.steps} passes of G4HepEmElectronManager::HowFar+Perform over the full {n}-track batch, "
                                   f"std::thread x {threads}, wall clock"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


def run_gpu(args):
    import torch

    from g4hepem_b200 import batches, engine as eng, tables

    rank, world, local = _dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: g4hepem_b200 has no CPU fallback")
    core_range = None if args.no_affinity else _spread_over_cores(local, world)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.tracks
    steps, warmup = args.steps, max(args.warmup, 3)
    ft = tables.load_state_json(STATE_JSON)
    engine = eng.Engine(ft, device=local)
    pristine = batches.make_electron_batch(n, ft.num_matcut, seed=SEED, id_offset=rank * n, pinned=True)

    # ---- device-resident ring: one pristine batch per step (warm-up steps get their own) ------------------
    per_batch = n * (16 * 16 + 16 + 4)
    ring_n = warmup + steps
    if ring_n * per_batch > 100e9:
        ring_n = max(4, int(100e9 // per_batch))
    ring = []
    for i in range(ring_n):
        d = eng.ElectronDeviceBatch(n, device=local)
        d.upload(pristine, groups=batches.ElectronHostBatch.PAIR_GROUPS + ("meta",))
        ring.append(d)
    sec = eng.SecondaryDeviceQueue(2 * n, device=local)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        sec.reset()
        eng.ElectronManager.Step(engine, ring[i % ring_n], sec, SEED)
    barrier()
    launches0 = engine.launch_count
    sampler = make_clock_sampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    ev0.record()
    for i in range(steps):
        b = ring[(warmup + i) % ring_n]
        sec.reset()
        kev[i][0].record()
        eng.ElectronManager.Step(engine, b, sec, SEED)
        kev[i][1].record()
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler is not None else None
    elapsed_ms = ev0.elapsed_time(ev1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    n_sec = int(sec.count[0].item())
    launches = engine.launch_count - launches0
    # per-kernel durations: the same steps again over re-initialised ring batches, this time with CUDA events
    # around every kernel of the pipeline (on the launch stream).  Kept out of the region `value` is timed over:
    # an event between two small kernels opens a ~10 us gap that back-to-back launches do not have.
    t_steps = min(steps, ring_n, 10)
    for i in range(t_steps):
        ring[i].upload(pristine, groups=batches.ElectronHostBatch.PAIR_GROUPS + ("meta",))
    torch.cuda.synchronize()
    engine.set_kernel_timing(True)
    engine.kernel_times()
    for i in range(t_steps):
        sec.reset()
        eng.ElectronManager.Step(engine, ring[i], sec, SEED)
    torch.cuda.synchronize()
    stage_times = engine.kernel_times()
    engine.set_kernel_timing(False)
    if dist is not None:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * n * steps / (elapsed_ms * 1e-3)

    # ---- the one collective of the path: per-couple sums of the deposited energy over all ranks (cf. Run::Merge) ----
    from g4hepem_b200 import sharding

    last = ring[(warmup + steps - 1) % ring_n]
    hist = torch.zeros(ft.num_matcut, dtype=torch.float64, device="cuda")
    hist.index_add_(0, last.t["meta"][:n, 0].long(), last.t["edep_dispx"][:n, 0])
    tot, _ = sharding.allreduce_scores(hist.cpu().numpy(), [n], dist, device=torch.device("cuda", local))
    edep_sum = float(tot.sum())

    # ---- sustained load: the same step back to back for >= 1 s (the 20-step region above is a 10 ms burst) -------------------
    sustained = None
    if args.sustained_seconds > 0:
        est = max(elapsed_ms / steps, 1e-3)
        s_steps = int(args.sustained_seconds * 1e3 / est) + 1
        in_groups_dev = batches.ElectronHostBatch.PAIR_GROUPS + ("meta",)
        pristine_dev = eng.ElectronDeviceBatch(n, device=local)
        pristine_dev.upload(pristine, groups=in_groups_dev)
        barrier()
        s_sampler = make_clock_sampler(local) if rank == 0 else None
        sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(s_steps)]
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(s_steps):
            b = ring[i % ring_n]
            for g in in_groups_dev:  # device-to-device restore of the pristine inputs (outside the per-step events)
                b.t[g][:n].copy_(pristine_dev.t[g][:n], non_blocking=True)
            sec.reset()
            sev[i][0].record()
            eng.ElectronManager.Step(engine, b, sec, SEED)
            sev[i][1].record()
        s1.record()
        barrier()
        s_clocks = s_sampler.stop() if s_sampler is not None else None
        s_wall_ms = s0.elapsed_time(s1)
        s_ms = float(sum(a.elapsed_time(b) for a, b in sev))
        if dist is not None:
            t = torch.tensor([s_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s_ms = float(t.item())
        sustained = {"value": world * n * s_steps / (s_ms * 1e-3), "unit": UNIT, "loop_seconds": s_wall_ms * 1e-3,
                     "step_seconds": s_ms * 1e-3, "steps": s_steps, "clocks": s_clocks,
                     "note": "the step of `value` back to back for >= 1 s, every step on pristine inputs restored device-to-device "
                             "between the per-step CUDA events; value = tracks / summed step time"}
        del pristine_dev

    # ---- offered variant: SampleMSC in single precision (g4hb200_set_msc_precision(h, 32), tests/test_msc_f32.py states the
    # bound); the same timed region as `value`, a record beside it -- the headline is the FP64 drop-in -----------------------------
    variants = None
    if args.variants:
        in_groups_dev = batches.ElectronHostBatch.PAIR_GROUPS + ("meta",)
        for i in range(ring_n):
            ring[i].upload(pristine, groups=in_groups_dev)
        engine.set_msc_precision(32)
        try:
            for i in range(warmup):
                sec.reset()
                eng.ElectronManager.Step(engine, ring[i % ring_n], sec, SEED)
            barrier()
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record()
            for i in range(steps):
                sec.reset()
                eng.ElectronManager.Step(engine, ring[(warmup + i) % ring_n], sec, SEED)
            v1.record()
            barrier()
        finally:
            engine.set_msc_precision(64)
        v_ms = v0.elapsed_time(v1)
        if dist is not None:
            t = torch.tensor([v_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            v_ms = float(t.item())
        variants = {"msc_f32": {"value": world * n * steps / (v_ms * 1e-3), "unit": UNIT, "ms_per_step": v_ms / steps, "dtype": "f64 + f32 SampleMSC",
                                "bound": "discrete outcomes, energies, step lengths identical to the f64 path; direction within 5e-4 "
                                         "absolute (99.9 %: 1e-5), displacement within 2e-4 of its length (tests/test_msc_f32.py)"}}

    # ---- BASELINE configs[4]: TestEm3 ATLASbar 10 GeV e- showers, primaries sharded over the ranks, the per-layer deposits
    # summed over ranks by ONE NCCL all_reduce on the device -- inside the timed region ---------------------------------------
    shower_rec = None
    if args.shower_primaries > 0:
        from g4hepem_b200 import shower

        calo = shower.SlabCalorimeter()
        prim = args.shower_primaries
        cap = max(1 << 18, prim * 1536)
        shower.run(engine, calo, min(8, prim), 1000.0, SEED, capacity=1 << 18)  # warm-up: streams, workspaces, kernels
        hist_dev = torch.zeros(calo.num_cells + 4, dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(hist_dev)  # warm-up of the communicator
        barrier()
        res = shower.run(engine, calo, prim, 10000.0, SEED, first_track_id=rank * prim, capacity=cap)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = res.stats
        a0.record()
        hist_dev[:calo.num_cells] = torch.from_numpy(res.edep.ravel()).cuda()
        hist_dev[calo.num_cells:] = torch.tensor([st["electron_track_steps"], st["gamma_track_steps"], st["leak_electron"],
                                                  st["leak_gamma"]], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(hist_dev)
        a1.record()
        torch.cuda.synchronize()
        # device time of the loop (CUDA events on the library's stream) + of the collective, max over ranks
        t_ms = torch.tensor([st["device_ms"] + a0.elapsed_time(a1)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        tot = hist_dev.cpu().numpy()
        sh_ms = float(t_ms.item())
        track_steps = float(tot[calo.num_cells] + tot[calo.num_cells + 1])
        edep_total = float(tot[:calo.num_cells].sum())
        shower_rec = {
            "workload": "BASELINE configs[4]: TestEm3 ATLASbar (50 x (2.3 mm Pb + 5.7 mm lAr)) 10 GeV e- showers, stepped until no track "
                        "is left; primaries sharded over the ranks, per-(layer, absorber) deposits summed by one NCCL all_reduce inside "
                        "the timed region",
            "primaries_per_gpu": prim, "value": track_steps / (sh_ms * 1e-3), "unit": "track-steps/s",
            "showers_per_s": world * prim / (sh_ms * 1e-3), "ms": sh_ms, "allreduce_ms_rank0": a0.elapsed_time(a1),
            "loop_iterations_rank0": int(st["num_steps"]), "kernel_launches_rank0": int(st["kernel_launches"]),
            "energy_balance": (edep_total + float(tot[calo.num_cells + 2] + tot[calo.num_cells + 3])) / (world * prim * 10000.0),
            "scaling": "weak", "n_gpus": world,
        }

    # ---- e2e: host buffers through the C-ABI ---------------------------------------------------------------
    e2e_steps = min(steps, args.e2e_steps)
    work = batches.ElectronHostBatch(n, pinned=True)
    hsec = batches.SecondaryHostQueue(2 * n, pinned=True)
    in_groups = batches.ElectronHostBatch.PAIR_GROUPS + ("meta",)

    def restore():
        for g in in_groups:
            getattr(work, g)[...] = getattr(pristine, g)

    e2e_times = []
    for i in range(2 + e2e_steps):
        restore()
        barrier()
        t0 = time.perf_counter()
        engine.electron_step_host(work, hsec, SEED)
        dt = time.perf_counter() - t0
        if i >= 2:
            e2e_times.append(dt)
    e2e_dt = float(np.sum(e2e_times))
    if dist is not None:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = world * n * e2e_steps / e2e_dt
    n_sec_host = int(hsec.count[0])
    h2d = n * READ_BYTES
    d2h = n * WRITE_BYTES + n_sec_host * SEC_BYTES + 4

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = _peaks()
    step_bytes = n * (READ_BYTES + WRITE_BYTES) + n_sec * SEC_BYTES
    stages = []
    for name, (ms_sum, nl, items) in stage_times.items():
        if nl == 0:
            continue
        rd, wr, nsec = STAGE_BYTES[name]
        per_launch_bytes = (items / nl) * (rd + wr + nsec * SEC_BYTES)
        ms = ms_sum / nl
        stages.append({"kernel": name, "ms": ms, "tracks": items / nl, "bytes": per_launch_bytes,
                       "gbs": per_launch_bytes / (ms * 1e-3) / 1e9 if ms > 0 else None})
    dom = max(stages, key=lambda x: x["ms"])
    step_gbs = step_bytes / (kernel_ms * 1e-3) / 1e9
    step_traffic, traffic_src = _traffic_from_profile([x["kernel"] for x in stages])
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": elapsed_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": _workload_config(n), "ring_batches": ring_n,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "g4hb200_electron_step_host (pinned host batch in, host batch + secondaries out)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # the roofline of the WHOLE step (all its kernels): algorithmic bytes of the caller-visible state over the CUDA-event
        # time of a step; the kernel with the longest duration is a sub-record
        "roofline": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                     "traffic": step_traffic, "traffic_source": traffic_src, "scope": "whole step (every kernel of g4hb200_electron_step)",
                     "algorithmic_bytes_per_step": step_bytes, "step_ms": kernel_ms, "secondaries_per_step": n_sec,
                     "peak_source": peak_src,
                     "dominant_kernel": {"kernel": dom["kernel"], "kernel_ms": dom["ms"], "achieved": dom["gbs"],
                                         "frac": dom["gbs"] / peak, "algorithmic_bytes_per_launch": dom["bytes"],
                                         "share_of_step": dom["ms"] / sum(x["ms"] for x in stages)},
                     "stages": stages},
        "edep_sum_mev_last_step_allreduced": edep_sum,
        "cpu_cores_of_rank0": core_range,
    }
    if sustained is not None:
        line["sustained"] = sustained
    if shower_rec is not None:
        line["shower"] = shower_rec
    if variants is not None:
        line["variants"] = variants
    if not args.no_cpu_baseline and world == 1:
        try:
            v, threads, kind, times = _cpu_reference_rate(min(n, args.cpu_sample), 3)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": kind,
                "sample": f"best of 3 passes of HowFar+Perform over {min(n, args.cpu_sample)} tracks of the same batch "
                          f"({sum(times):.2f} s wall in total), std::thread x {threads}"}
        except Exception as exc:  # the checker library is test infrastructure; its absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(exc)}
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


NCU_NAMES = {"ElMSCSampleKernel<e->": "ElMSCSampleKernel<0>", "ElMSCSampleKernel<e+>": "ElMSCSampleKernel<1>",
             "ElSamplerKernel<AtRest>": "ElSamplerKernel<4>", "ElSamplerKernel<Moller>": "ElSamplerKernel<5>",
             "ElSamplerKernel<Bhabha>": "ElSamplerKernel<6>", "ElSamplerKernel<SeltzerBerger>": "ElSamplerKernel<7>",
             "ElSamplerKernel<RelBrem>": "ElSamplerKernel<8>", "ElSamplerKernel<Annihilation>": "ElSamplerKernel<9>",
             "ElFusedStepKernel": "ElFusedStepKernel<0>"}


def _traffic_from_profile(kernels):
    """dram bytes read + written by one 1M-track step = the sum over its kernels in the committed `ncu --set full` summary
    (unsplit 1M-track launches).  Returns (bytes or None, file)."""
    for name in ("r02b_pipeline_full.json", "r02_pipeline_full.json", "r01c_pipeline_full.json"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        try:
            with open(path) as f:
                prof = json.load(f)
            total = 0.0
            for k in kernels:
                total += prof[NCU_NAMES.get(k, k)]["dram_bytes_per_launch"]
            return total, "profiles/" + name
        except (OSError, ValueError, KeyError):
            continue
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tracks", type=int, default=1 << 20, help="tracks per step per GPU (BASELINE: 1M)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=1 << 20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-affinity", action="store_true", help="leave the CPU affinity of the ranks alone")
    ap.add_argument("--sustained-seconds", type=float, default=1.0, help="length of the sustained-load loop (0: skip)")
    ap.add_argument("--no-variants", dest="variants", action="store_false", help="skip the offered-variant records (f32 SampleMSC)")
    ap.add_argument("--shower-primaries", type=int, default=4096, help="BASELINE configs[4] record: primaries per GPU (0: skip)")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: whatever a library prints there while the bench runs (NCCL writes its version
    # line to stdout on the 8-GPU box) goes to stderr instead; the line itself is written to the real stdout at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
