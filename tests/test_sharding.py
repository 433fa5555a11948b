"""Host side of the N>1 path on CPU: sharding is a partition, results do not depend on it, and the one collective
(sum of the per-layer deposits) gives the single-process answer -- world size 2 over gloo."""
import os
import socket
import subprocess
import sys

import numpy as np

from g4hepem_b200 import batches, sharding
from tests.conftest import ROOT

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["G4H_ROOT"])
import torch.distributed as dist
from g4hepem_b200 import batches, sharding, tables
from oracle import checker
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
J = os.path.join(os.environ["G4H_ROOT"], "tests", "golden", "hepem_state.json")
ft = tables.load_state_json(J)
ora = checker.best_available(J)
n = 6000
full = batches.make_electron_batch(n, ft.num_matcut, seed=77)
mine = sharding.shard_host_batch(full, rank, world)
sec = batches.SecondaryHostQueue(2 * mine.n)
ora.electron_step(mine, sec, 2026, 1)      # the CPU checker stands in for the device step: the host logic is what is tested
layers = mine.meta[:, 2] % 50
hist = sharding.layer_histogram(mine.edep_dispx[:, 0], layers, 50)
tot, cnt = sharding.allreduce_scores(hist, [mine.n, int(sec.count[0])], dist)
if rank == 0:
    np.savez(os.environ["G4H_OUT"], hist=tot, counters=cnt)
dist.destroy_process_group()
'''


def test_shard_bounds_partition():
    for n in (0, 1, 7, 1000, 1 << 20):
        for w in (1, 2, 3, 8):
            b = [sharding.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo_matches_single_process(tmp_path, reference, flat_tables):
    out = str(tmp_path / "scores.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, G4H_ROOT=ROOT, G4H_OUT=out, OMP_NUM_THREADS="1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)], env=env, timeout=600)
    got = np.load(out)
    n = 6000
    full = batches.make_electron_batch(n, flat_tables.num_matcut, seed=77)
    sec = batches.SecondaryHostQueue(2 * n)
    reference.electron_step(full, sec, 2026, 1)
    want = sharding.layer_histogram(full.edep_dispx[:, 0], full.meta[:, 2] % 50, 50)
    assert np.allclose(got["hist"], want, rtol=1e-12, atol=0)
    assert got["counters"].tolist() == [n, int(sec.count[0])]


SHOWER_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["G4H_ROOT"])
import torch.distributed as dist
from g4hepem_b200 import sharding, shower
from oracle import checker
from tests import shower_oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ora = checker.best_available(os.path.join(os.environ["G4H_ROOT"], "tests", "golden", "hepem_state.json"))
calo = shower.SlabCalorimeter(num_layers=20)
lo, hi = sharding.shard_bounds(6, rank, world)
# the CPU loop stands in for the device loop: the host logic (slice of the primaries, ids, the one collective) is tested
hist, st = shower_oracle.run(ora, calo, hi - lo, 120.0, 2026, first_track_id=lo, threads=1)
tot, cnt = sharding.allreduce_scores(hist.ravel(), [st["electron_track_steps"], st["gamma_track_steps"], st["secondaries"]], dist)
if rank == 0:
    np.savez(os.environ["G4H_OUT"], hist=tot, counters=cnt)
dist.destroy_process_group()
'''


def test_sharded_showers_world_size_2_gloo(tmp_path, reference):
    from g4hepem_b200 import shower
    from tests import shower_oracle

    out = str(tmp_path / "shower.npz")
    script = tmp_path / "shower_worker.py"
    script.write_text(SHOWER_WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, G4H_ROOT=ROOT, G4H_OUT=out, OMP_NUM_THREADS="1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)], env=env, timeout=600)
    got = np.load(out)
    whole, st = shower_oracle.run(reference, shower.SlabCalorimeter(num_layers=20), 6, 120.0, 2026, threads=1)
    np.testing.assert_allclose(got["hist"], whole.ravel(), rtol=1e-12, atol=1e-12)
    assert list(got["counters"]) == [st["electron_track_steps"], st["gamma_track_steps"], st["secondaries"]]
