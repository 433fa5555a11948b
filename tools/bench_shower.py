#!/usr/bin/env python
"""BASELINE configs[3] and [4] on the device-resident stepping loops (side measurements; bench.py keeps the headline).

  --config 4 : TestEm3 ATLASbar calorimeter (50 x (2.3 mm Pb + 5.7 mm lAr)), P primaries of E MeV per GPU, stepped
               until no track is left; primaries are sharded over the ranks (weak scaling: P per rank), the
               per-(layer, absorber) deposits are summed with ONE all_reduce (NCCL) after the loops.
  --config 3 : mixed e-/e+/gamma population in queue order, k consecutive fused steps, secondaries fed back.

    python tools/bench_shower.py --config 4 --primaries 256 --ekin 10000
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_shower.py --config 4 --gpus N
One JSON line from rank 0: value = (e-/e+ + gamma track-steps of all ranks) / (max over ranks of the CUDA-event time).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STATE_JSON = os.path.join(ROOT, "tests", "golden", "hepem_state.json")


def main():
    import torch

    from g4hepem_b200 import engine as eng, shower, tables

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--primaries", type=int, default=256, help="per GPU")
    ap.add_argument("--ekin", type=float, default=10000.0)
    ap.add_argument("--tracks", type=int, default=16 << 20, help="config 3: total tracks per GPU")
    ap.add_argument("--steps", type=int, default=4, help="config 3: consecutive steps")
    ap.add_argument("--capacity", type=int, default=0)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--woodcock", action="store_true", help="config 4: Woodcock tracking of the gammas (TestEm3's default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ft = tables.load_state_json(STATE_JSON)
    e = eng.Engine(ft, device=local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == 4:
        calo = shower.SlabCalorimeter(woodcock=args.woodcock)
        cap = args.capacity or max(1 << 18, args.primaries * 1536)
        shower.run(e, calo, min(8, args.primaries), min(args.ekin, 1000.0), args.seed, capacity=1 << 18)  # warm-up
        barrier()
        res = shower.run(e, calo, args.primaries, args.ekin, args.seed, first_track_id=rank * args.primaries, capacity=cap)
        barrier()
        st = res.stats
        vec = np.array([st["electron_track_steps"], st["gamma_track_steps"], st["secondaries"], st["leak_electron"], st["leak_gamma"]],
                       dtype=np.float64)
        from g4hepem_b200 import sharding

        hist, cnt = sharding.allreduce_scores(res.edep.ravel(), vec, dist, torch.device("cuda", local))
        ms = torch.tensor([st["device_ms"]], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        if rank == 0:
            steps = cnt[0] + cnt[1]
            total_e = world * args.primaries * args.ekin
            line = {"config": f"BASELINE configs[4]: TestEm3 ATLASbar (50 x (2.3 mm Pb + 5.7 mm lAr)) {args.ekin / 1000:g} GeV e- showers, "
                              f"{args.primaries} primaries per GPU, stepped until no track is left"
                              + (", Woodcock tracking of gammas" if args.woodcock else ""),
                    "metric": "e-/e+/gamma track-steps/s", "value": steps / (ms * 1e-3), "unit": "track-steps/s", "n_gpus": world,
                    "scaling": "weak", "ms": ms, "primaries_total": world * args.primaries,
                    "primaries_per_s": world * args.primaries / (ms * 1e-3),
                    "electron_track_steps": cnt[0], "gamma_track_steps": cnt[1], "tracks_created": cnt[2],
                    "loop_iterations_rank0": st["num_steps"], "peak_electrons_rank0": st["peak_electrons"],
                    "peak_gammas_rank0": st["peak_gammas"], "kernel_launches_rank0": st["kernel_launches"],
                    "edep_mev": float(hist.sum()), "leak_mev": float(cnt[3] + cnt[4]),
                    "energy_balance": float((hist.sum() + cnt[3] + cnt[4]) / total_e),
                    "edep_fraction_absorbers": [float(x) for x in hist.reshape(calo.num_layers, -1).sum(axis=0) / total_e],
                    "collective": "one all_reduce (sum) of the per-(layer, absorber) histogram + 5 counters after the loops",
                    "dtype": "f64", "data": "synthetic tables (tests/golden/hepem_state.json)"}
            print(json.dumps(line), flush=True)
    else:
        n_el = (2 * args.tracks) // 3
        n_gm = args.tracks - n_el
        cap = args.capacity or 7 * max(n_el, n_gm) + (1 << 16)
        shower.run_mixed(e, 1 << 16, 1 << 15, 2, args.seed)  # warm-up
        barrier()
        edep, st = shower.run_mixed(e, n_el, n_gm, args.steps, args.seed + rank, capacity=cap)
        barrier()
        vec = torch.tensor([st["electron_track_steps"], st["gamma_track_steps"], st["secondaries"], edep], dtype=torch.float64, device="cuda")
        ms = torch.tensor([st["device_ms"]], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(vec)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        vec, ms = vec.cpu().numpy(), float(ms.item())
        if rank == 0:
            line = {"config": f"BASELINE configs[3]: mixed e-/e+/gamma {args.tracks} tracks per GPU in queue order (particle, couple), "
                              f"E log-uniform 1 keV-100 GeV, all couples of the table set, {args.steps} consecutive steps, secondaries fed back",
                    "metric": "e-/e+/gamma track-steps/s", "value": (vec[0] + vec[1]) / (ms * 1e-3), "unit": "track-steps/s",
                    "n_gpus": world, "scaling": "weak", "ms": ms, "electron_track_steps": vec[0], "gamma_track_steps": vec[1],
                    "tracks_created": vec[2], "edep_mev": vec[3], "peak_electrons_rank0": st["peak_electrons"],
                    "peak_gammas_rank0": st["peak_gammas"], "kernel_launches_rank0": st["kernel_launches"], "dtype": "f64"}
            print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
