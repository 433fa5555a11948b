"""The stepping loop over a TestEm3-style slab calorimeter (BASELINE configs[4]): host-side mirror of
g4hb200_shower_run (include/g4hepem_b200.h).  The loop itself runs in the C++/CUDA library; this module only
describes the geometry, calls it, and shards primaries over ranks (tracks are independent, so every rank runs its own
slice of the primaries against its replica of the tables and the per-layer histograms are summed with one
all_reduce -- TestEm3's Run::Merge, apps/examples/TestEm3/src/Run.cc:146-190)."""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi, sharding


@dataclass
class SlabCalorimeter:
    """apps/examples/TestEm3/src/DetectorConstruction.cc:72-84: the ATLASbar defaults."""
    num_layers: int = 50
    absorber_thickness: tuple = (2.3, 5.7)   # mm: G4_Pb, G4_lAr
    absorber_couple: tuple = (1, 2)          # material-cuts couple of each absorber in the table set
    half_yz: float = 200.0                   # mm
    # Woodcock tracking of gammas in the calorimeter (PhysListHepEmTracking.cc:42); couple = the densest absorber's
    woodcock: bool = False
    woodcock_couple: int = 1
    woodcock_ekin_min: float = 0.2           # MeV, G4HepEmConfig::fWDTEnergyLimit

    def as_struct(self):
        g = _capi.SlabGeometry()
        g.num_layers = self.num_layers
        g.num_absorbers = len(self.absorber_thickness)
        for k, (t, c) in enumerate(zip(self.absorber_thickness, self.absorber_couple)):
            g.absorber_thickness[k] = t
            g.absorber_couple[k] = c
        g.half_yz = self.half_yz
        g.woodcock_on = 1 if self.woodcock else 0
        g.woodcock_couple = self.woodcock_couple
        g.woodcock_ekin_min = self.woodcock_ekin_min
        return g

    @property
    def num_cells(self):
        return self.num_layers * len(self.absorber_thickness)


@dataclass
class ShowerResult:
    edep: np.ndarray                 # [num_layers, num_absorbers] MeV
    stats: dict = field(default_factory=dict)


def run(engine, calo, num_primaries, primary_ekin, seed, kind=_capi.SEC_ELECTRON, first_track_id=0, capacity=None,
        max_steps=0):
    """num_primaries showers of `kind` at primary_ekin [MeV] through `calo` on the engine's GPU."""
    if capacity is None:
        capacity = max(1 << 16, int(num_primaries) * 4096)
    g = calo.as_struct()
    edep = np.zeros(calo.num_cells, dtype=np.float64)
    st = _capi.ShowerStats()
    _capi.check(engine.lib.g4hb200_shower_run(engine.handle, C.byref(g), int(num_primaries), int(kind), float(primary_ekin),
                                              int(seed), int(first_track_id), int(capacity), int(max_steps),
                                              edep.ctypes.data, C.byref(st)), "shower_run")
    stats = {name: getattr(st, name) for name, _ in _capi.ShowerStats._fields_}
    return ShowerResult(edep.reshape(calo.num_layers, -1), stats)


def run_sharded(engine, calo, total_primaries, primary_ekin, seed, rank, world, dist=None, device=None, **kw):
    """Rank `rank` of `world` runs its contiguous slice of the primaries (ids = global primary index, so the result does
    not depend on the sharding) and the histograms / counters are summed over ranks."""
    lo, hi = sharding.shard_bounds(total_primaries, rank, world)
    res = run(engine, calo, hi - lo, primary_ekin, seed, first_track_id=lo, **kw)
    keys = ("electron_track_steps", "gamma_track_steps", "secondaries", "leak_electron", "leak_gamma")
    hist, cnt = sharding.allreduce_scores(res.edep.ravel(), [res.stats[k] for k in keys], dist, device)
    total = dict(res.stats)
    total.update({k: float(v) for k, v in zip(keys, cnt)})
    return ShowerResult(hist.reshape(res.edep.shape), total), res


def run_mixed(engine, num_electrons, num_gammas, num_steps, seed, emin=1.0e-3, emax=1.0e5, capacity=None):
    """BASELINE configs[3]: mixed e-/e+/gamma population in queue order, `num_steps` consecutive fused steps with the
    secondaries fed back (g4hb200_mixed_run).  Returns (deposited energy [MeV], stats)."""
    if capacity is None:
        capacity = 7 * max(int(num_electrons), int(num_gammas)) + (1 << 16)
    edep = np.zeros(1, dtype=np.float64)
    st = _capi.ShowerStats()
    _capi.check(engine.lib.g4hb200_mixed_run(engine.handle, int(num_electrons), int(num_gammas), float(emin), float(emax),
                                             int(seed), int(capacity), int(num_steps), edep.ctypes.data, C.byref(st)), "mixed_run")
    return float(edep[0]), {name: getattr(st, name) for name, _ in _capi.ShowerStats._fields_}
