"""Host-side track batches (structure of paired arrays) and synthetic batch generators.

A host batch owns numpy arrays laid out exactly like the device batch (include/g4hepem_b200.h):
every double group is an (n, 2) float64 array, `meta` is (n, 4) int32.  `as_struct()` gives the
ctypes view passed through the C-ABI.
"""
import ctypes as C

import numpy as np

from . import _capi


class _HostBatch:
    PAIR_GROUPS = ()
    RESULT_GROUPS = ()
    HANDOVER_GROUPS = ()
    STRUCT = None

    def __init__(self, n, pinned=False):
        self.n = int(n)
        self._pinned = pinned
        self._torch = []
        for g in self.PAIR_GROUPS + self.RESULT_GROUPS + self.HANDOVER_GROUPS:
            setattr(self, g, self._alloc((self.n, 2), np.float64))
        self.meta = self._alloc((self.n, 4), np.int32)
        self.winner = self._alloc((self.n,), np.int32)

    def _alloc(self, shape, dtype):
        if self._pinned:
            import torch

            tdt = torch.float64 if dtype == np.float64 else torch.int32
            t = torch.zeros(shape, dtype=tdt).pin_memory()
            self._torch.append(t)
            return t.numpy()
        return np.zeros(shape, dtype=dtype)

    def groups(self):
        return self.PAIR_GROUPS + self.RESULT_GROUPS + self.HANDOVER_GROUPS

    def as_struct(self):
        s = self.STRUCT()
        s.n = self.n
        for g in self.groups():
            setattr(s, g, getattr(self, g).ctypes.data_as(_capi.c_dp))
        s.meta = self.meta.ctypes.data_as(_capi.c_ip)
        s.winner = self.winner.ctypes.data_as(_capi.c_ip)
        return s

    def copy(self):
        o = type(self)(self.n)
        for g in self.groups() + ("meta", "winner"):
            getattr(o, g)[...] = getattr(self, g)
        return o

    def nbytes(self, groups):
        return sum(getattr(self, g).nbytes for g in groups)


class ElectronHostBatch(_HostBatch):
    PAIR_GROUPS = _capi.ELECTRON_PAIR_GROUPS
    RESULT_GROUPS = _capi.ELECTRON_RESULT_GROUPS
    HANDOVER_GROUPS = _capi.ELECTRON_HANDOVER_GROUPS
    STRUCT = _capi.ElectronBatch


class GammaHostBatch(_HostBatch):
    PAIR_GROUPS = _capi.GAMMA_PAIR_GROUPS
    RESULT_GROUPS = _capi.GAMMA_RESULT_GROUPS
    STRUCT = _capi.GammaBatch


class SecondaryHostQueue:
    def __init__(self, capacity, pinned=False):
        self.capacity = int(capacity)
        self._torch = []
        self.dirx_diry = self._alloc((self.capacity, 2), np.float64, pinned)
        self.dirz_ekin = self._alloc((self.capacity, 2), np.float64, pinned)
        self.parent_kind = self._alloc((self.capacity, 2), np.int32, pinned)
        self.parent_slot = self._alloc((self.capacity, 2), np.int32, pinned)
        self.count = self._alloc((4,), np.int32, pinned)[:1]

    def _alloc(self, shape, dtype, pinned):
        if pinned:
            import torch

            t = torch.zeros(shape, dtype=torch.float64 if dtype == np.float64 else torch.int32).pin_memory()
            self._torch.append(t)
            return t.numpy()
        return np.zeros(shape, dtype=dtype)

    def as_struct(self):
        s = _capi.SecondaryQueue()
        s.capacity = self.capacity
        s.dirx_diry = self.dirx_diry.ctypes.data_as(_capi.c_dp)
        s.dirz_ekin = self.dirz_ekin.ctypes.data_as(_capi.c_dp)
        s.parent_kind = self.parent_kind.ctypes.data_as(_capi.c_ip)
        s.parent_slot = self.parent_slot.ctypes.data_as(_capi.c_ip)
        s.count = self.count.ctypes.data_as(_capi.c_ip)
        return s

    def sorted_records(self):
        """Records ordered by (parent batch index, slot): the order is launch-shape independent."""
        n = int(self.count[0])
        key = self.parent_slot[:n, 0].astype(np.int64) * 4 + self.parent_slot[:n, 1]
        order = np.argsort(key, kind="stable")
        return dict(
            parent_index=self.parent_slot[:n, 0][order], slot=self.parent_slot[:n, 1][order],
            parent_id=self.parent_kind[:n, 0][order], kind=self.parent_kind[:n, 1][order],
            dir=np.concatenate([self.dirx_diry[:n], self.dirz_ekin[:n, :1]], axis=1)[order],
            ekin=self.dirz_ekin[:n, 1][order],
        )


def _isotropic(rng, n):
    cost = rng.uniform(-1.0, 1.0, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    sint = np.sqrt((1.0 - cost) * (1.0 + cost))
    return sint * np.cos(phi), sint * np.sin(phi), cost


def make_electron_batch(n, num_couples, seed=2026, emin=1.0e-3, emax=1.0e5, positron_fraction=0.5,
                        boundary_fraction=0.1, couples=None, id_offset=0, pinned=False):
    """BASELINE config 3 inputs (SURVEY.md par. 8d): e-/e+ 50/50, E log-uniform 1 keV-100 GeV, couple
    uniform, isotropic direction, safety ~ U[0, 1 mm], onBoundary ~ Bernoulli(0.1), first-step flag set,
    numIALeft = -1, MSC data as after G4HepEmMSCTrackData::ReSet()."""
    rng = np.random.Generator(np.random.PCG64(seed))
    b = ElectronHostBatch(n, pinned=pinned)
    ekin = np.exp(rng.uniform(np.log(emin), np.log(emax), n))
    b.ekin_logekin[:, 0] = ekin
    b.ekin_logekin[:, 1] = 100.0  # not cached: the kernels evaluate the (VDT) log themselves
    dx, dy, dz = _isotropic(rng, n)
    b.dirx_diry[:, 0], b.dirx_diry[:, 1], b.dirz_safety[:, 0] = dx, dy, dz
    onb = rng.uniform(size=n) < boundary_fraction
    b.dirz_safety[:, 1] = np.where(onb, 0.0, rng.uniform(0.0, 1.0, n))
    b.nia01[...] = -1.0
    b.nia23[...] = -1.0
    b.msc_irange_dynrf[:, 0] = 1.0e21
    b.msc_irange_dynrf[:, 1] = 0.04
    b.msc_tlimmin_gauss[:, 0] = 1.0e-7
    b.msc_tlimmin_gauss[:, 1] = 0.0
    if couples is None:
        couples = np.arange(num_couples)
    b.meta[:, 0] = rng.choice(np.asarray(couples, dtype=np.int32), n)
    flags = np.full(n, _capi.F_MSC_FIRST_STEP, dtype=np.int32)
    flags |= np.where(rng.uniform(size=n) < positron_fraction, _capi.F_POSITRON, 0).astype(np.int32)
    flags |= np.where(onb, _capi.F_ON_BOUNDARY, 0).astype(np.int32)
    b.meta[:, 1] = flags
    b.meta[:, 2] = np.arange(n, dtype=np.int32) + id_offset
    b.meta[:, 3] = 0
    b.winner[...] = -1
    return b


def make_gamma_batch(n, num_couples, seed=2027, emin=0.98e-4, emax=1.02e8, boundary_fraction=0.0, couples=None,
                     id_offset=0, pinned=False):
    """BASELINE config 2 inputs: E log-uniform 100 eV*0.98 ... 100 TeV*1.02
    (testing/GammaXSections/src/Implementation.cc:46-56), couple uniform, isotropic direction."""
    rng = np.random.Generator(np.random.PCG64(seed))
    b = GammaHostBatch(n, pinned=pinned)
    b.ekin_logekin[:, 0] = np.exp(rng.uniform(np.log(emin), np.log(emax), n))
    b.ekin_logekin[:, 1] = 100.0
    dx, dy, dz = _isotropic(rng, n)
    b.dirx_diry[:, 0], b.dirx_diry[:, 1], b.dirz_nia0[:, 0] = dx, dy, dz
    b.dirz_nia0[:, 1] = -1.0
    if couples is None:
        couples = np.arange(num_couples)
    b.meta[:, 0] = rng.choice(np.asarray(couples, dtype=np.int32), n)
    onb = rng.uniform(size=n) < boundary_fraction
    b.meta[:, 1] = np.where(onb, _capi.F_ON_BOUNDARY, 0).astype(np.int32)
    b.meta[:, 2] = np.arange(n, dtype=np.int32) + id_offset
    b.meta[:, 3] = 0
    b.winner[...] = -1
    return b
