"""Python host side of the drop-in boundary: mirrors the reference's manager interface over batches.

`Engine` owns the C-ABI handle (device arena with the flattened tables).  `ElectronManager` /
`GammaManager` expose the reference's entry points -- `HowFar`, `Perform` (for gammas including
`SelectInteraction`) -- with the same meaning, but over a whole SoA batch per call instead of the
single primary track of a `G4HepEmTLData` (G4HepEmRun/include/G4HepEmElectronManager.hh:71,222,
G4HepEmGammaManager.hh:32-51), plus the fused `Step` used when no geometry sits between the two.

torch supplies device memory and streams only: device batches are torch CUDA tensors whose pointers
are handed to the C-ABI, and kernels are launched on torch's current stream so that torch CUDA events
time them.  There is no CPU path: without the CUDA extension or a GPU every call raises.
"""
import ctypes as C

import numpy as np

from . import _capi
from .batches import ElectronHostBatch, GammaHostBatch, SecondaryHostQueue


def _torch():
    import torch

    return torch


def _stream_ptr(stream=None):
    torch = _torch()
    st = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(st.cuda_stream)


class Engine:
    """Device-resident table set (replaces CopyG4HepEmDataToGPU, G4HepEmData/src/G4HepEmData.cc:78-101)."""

    def __init__(self, flat_tables, device=0):
        self.lib = _capi.load_library()
        torch = _torch()
        if not torch.cuda.is_available() or self.lib.g4hb200_device_count() <= 0:
            raise RuntimeError("g4hepem_b200 needs a CUDA device (there is no CPU fallback)")
        self.device = int(device)
        torch.cuda.set_device(self.device)
        torch.cuda.init()
        self.tables = flat_tables
        self.handle = C.c_void_p()
        _capi.check(self.lib.g4hb200_create(C.byref(flat_tables.desc), self.device, C.byref(self.handle)), "g4hb200_create")

    def close(self):
        if self.handle:
            self.lib.g4hb200_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self.lib.g4hb200_launch_count(self.handle))

    def set_msc_precision(self, bits):
        """64: the drop-in path (default); 32: the offered single-precision SampleMSC (csrc/g4h_msc_f32.cuh)."""
        _capi.check(self.lib.g4hb200_set_msc_precision(self.handle, int(bits)), "set_msc_precision")

    def set_kernel_timing(self, enable=True):
        _capi.check(self.lib.g4hb200_set_kernel_timing(self.handle, int(enable)), "set_kernel_timing")

    def kernel_times(self):
        """Per pipeline stage since the last call: {name: (ms_sum, launches, items)}."""
        ms = np.zeros(_capi.NUM_STAGES)
        ln = np.zeros(_capi.NUM_STAGES, dtype=np.int64)
        it = np.zeros(_capi.NUM_STAGES, dtype=np.int64)
        _capi.check(self.lib.g4hb200_kernel_times(self.handle, ms.ctypes.data, ln.ctypes.data, it.ctypes.data), "kernel_times")
        return {self.lib.g4hb200_stage_name(k).decode(): (float(ms[k]), int(ln[k]), int(it[k])) for k in range(_capi.NUM_STAGES)}

    # ---- look-ups on torch CUDA tensors ------------------------------------------------------------------
    def electron_lookups(self, imc, ekin, logekin, is_electron=True):
        torch = _torch()
        n = imc.numel()
        out = torch.empty((7, n), dtype=torch.float64, device=imc.device)
        _capi.check(self.lib.g4hb200_electron_lookups(self.handle, n, imc.data_ptr(), ekin.data_ptr(), logekin.data_ptr(),
                                                      int(is_electron), out.data_ptr(), _stream_ptr()), "electron_lookups")
        return out

    def electron_lookups_f32(self, imc, ekin, logekin, is_electron=True):
        """single precision variant (float32 tensors in, (7, n) float32 out): stated bound 2e-5 relative"""
        torch = _torch()
        n = imc.numel()
        out = torch.empty((7, n), dtype=torch.float32, device=imc.device)
        _capi.check(self.lib.g4hb200_electron_lookups_f32(self.handle, n, imc.data_ptr(), ekin.data_ptr(), logekin.data_ptr(),
                                                          int(is_electron), out.data_ptr(), _stream_ptr()), "electron_lookups_f32")
        return out

    def electron_lookups_into(self, imc, ekin, logekin, out, is_electron=True):
        _capi.check(self.lib.g4hb200_electron_lookups(self.handle, imc.numel(), imc.data_ptr(), ekin.data_ptr(), logekin.data_ptr(),
                                                      int(is_electron), out.data_ptr(), _stream_ptr()), "electron_lookups")

    def electron_stepping_xsecs(self, imc, ekin, logekin, is_electron=True):
        torch = _torch()
        n = imc.numel()
        out = torch.empty((4, n), dtype=torch.float64, device=imc.device)
        _capi.check(self.lib.g4hb200_electron_stepping_xsecs(self.handle, n, imc.data_ptr(), ekin.data_ptr(), logekin.data_ptr(),
                                                             int(is_electron), out.data_ptr(), _stream_ptr()), "stepping_xsecs")
        return out

    def gamma_lookups(self, imc, ekin, logekin, urnd):
        torch = _torch()
        n = imc.numel()
        mx = torch.empty(n, dtype=torch.float64, device=imc.device)
        pid = torch.empty(n, dtype=torch.int32, device=imc.device)
        _capi.check(self.lib.g4hb200_gamma_lookups(self.handle, n, imc.data_ptr(), ekin.data_ptr(), logekin.data_ptr(),
                                                   urnd.data_ptr(), mx.data_ptr(), pid.data_ptr(), _stream_ptr()), "gamma_lookups")
        return mx, pid

    def select_target_element(self, kind, is_electron, imc, ekin, logekin, urnd):
        torch = _torch()
        n = imc.numel()
        out = torch.empty(n, dtype=torch.int32, device=imc.device)
        _capi.check(self.lib.g4hb200_select_target_element(self.handle, kind, int(is_electron), n, imc.data_ptr(), ekin.data_ptr(),
                                                           logekin.data_ptr(), urnd.data_ptr(), out.data_ptr(), _stream_ptr()),
                    "select_target_element")
        return out

    def vdt_log_exp(self, x):
        torch = _torch()
        lo = torch.empty_like(x)
        ex = torch.empty_like(x)
        _capi.check(self.lib.g4hb200_vdt_log_exp(self.handle, x.numel(), x.data_ptr(), lo.data_ptr(), ex.data_ptr(), _stream_ptr()),
                    "vdt_log_exp")
        return lo, ex

    def rng_uniforms(self, seed, track_id, ndraw):
        torch = _torch()
        out = torch.empty((track_id.numel(), ndraw), dtype=torch.float64, device=track_id.device)
        _capi.check(self.lib.g4hb200_rng_uniforms(self.handle, seed, track_id.numel(), track_id.data_ptr(), ndraw, out.data_ptr(),
                                                  _stream_ptr()), "rng_uniforms")
        return out

    # ---- host-buffer entry points (the e2e path) -----------------------------------------------------------------
    def electron_step_host(self, host_batch, host_sec, seed):
        s = host_batch.as_struct()
        q = host_sec.as_struct()
        _capi.check(self.lib.g4hb200_electron_step_host(self.handle, C.byref(s), C.byref(q), seed), "electron_step_host")

    def gamma_step_host(self, host_batch, host_sec, seed):
        s = host_batch.as_struct()
        q = host_sec.as_struct()
        _capi.check(self.lib.g4hb200_gamma_step_host(self.handle, C.byref(s), C.byref(q), seed), "gamma_step_host")


class _DeviceBatch:
    HOST = None
    STRUCT = None

    def __init__(self, capacity, device=0):
        torch = _torch()
        self.capacity = int(capacity)
        self.n = 0
        dev = torch.device("cuda", device)
        self.t = {}
        for g in self.HOST.PAIR_GROUPS + self.HOST.RESULT_GROUPS + self.HOST.HANDOVER_GROUPS:
            self.t[g] = torch.zeros((self.capacity, 2), dtype=torch.float64, device=dev)
        self.t["meta"] = torch.zeros((self.capacity, 4), dtype=torch.int32, device=dev)
        self.t["winner"] = torch.zeros((self.capacity,), dtype=torch.int32, device=dev)

    def as_struct(self):
        s = self.STRUCT()
        s.n = self.n
        for g, t in self.t.items():
            setattr(s, g, C.cast(t.data_ptr(), _capi.c_ip if g in ("meta", "winner") else _capi.c_dp))
        return s

    def upload(self, host, groups=None, non_blocking=False):
        torch = _torch()
        self.n = host.n
        for g in (groups or (host.groups() + ("meta", "winner"))):
            self.t[g][: host.n].copy_(torch.from_numpy(getattr(host, g)), non_blocking=non_blocking)

    def download(self, host=None, groups=None):
        if host is None:
            host = self.HOST(self.n)
        for g in (groups or (host.groups() + ("meta", "winner"))):
            getattr(host, g)[...] = self.t[g][: self.n].cpu().numpy()
        return host


class ElectronDeviceBatch(_DeviceBatch):
    HOST = ElectronHostBatch
    STRUCT = _capi.ElectronBatch


class GammaDeviceBatch(_DeviceBatch):
    HOST = GammaHostBatch
    STRUCT = _capi.GammaBatch


class SecondaryDeviceQueue:
    def __init__(self, capacity, device=0):
        torch = _torch()
        dev = torch.device("cuda", device)
        self.capacity = int(capacity)
        self.dirx_diry = torch.zeros((self.capacity, 2), dtype=torch.float64, device=dev)
        self.dirz_ekin = torch.zeros((self.capacity, 2), dtype=torch.float64, device=dev)
        self.parent_kind = torch.zeros((self.capacity, 2), dtype=torch.int32, device=dev)
        self.parent_slot = torch.zeros((self.capacity, 2), dtype=torch.int32, device=dev)
        self.count = torch.zeros(4, dtype=torch.int32, device=dev)

    def as_struct(self):
        s = _capi.SecondaryQueue()
        s.capacity = self.capacity
        s.dirx_diry = C.cast(self.dirx_diry.data_ptr(), _capi.c_dp)
        s.dirz_ekin = C.cast(self.dirz_ekin.data_ptr(), _capi.c_dp)
        s.parent_kind = C.cast(self.parent_kind.data_ptr(), _capi.c_ip)
        s.parent_slot = C.cast(self.parent_slot.data_ptr(), _capi.c_ip)
        s.count = C.cast(self.count.data_ptr(), _capi.c_ip)
        return s

    def reset(self):
        self.count.zero_()

    def download(self):
        n = int(self.count[0].item())
        if n > self.capacity:
            raise _capi.G4HB200Error(f"secondary queue overflow: {n} > {self.capacity}")
        q = SecondaryHostQueue(max(n, 1))
        q.count[0] = n
        q.dirx_diry[:n] = self.dirx_diry[:n].cpu().numpy()
        q.dirz_ekin[:n] = self.dirz_ekin[:n].cpu().numpy()
        q.parent_kind[:n] = self.parent_kind[:n].cpu().numpy()
        q.parent_slot[:n] = self.parent_slot[:n].cpu().numpy()
        return q


class ElectronManager:
    """Batch counterpart of G4HepEmElectronManager (G4HepEmRun/include/G4HepEmElectronManager.hh:50-275)."""

    @staticmethod
    def HowFar(engine, batch, seed, stream=None):
        s = batch.as_struct()
        _capi.check(engine.lib.g4hb200_electron_howfar(engine.handle, C.byref(s), seed, _stream_ptr(stream)), "electron_howfar")

    @staticmethod
    def Perform(engine, batch, secondaries, seed, stream=None):
        s = batch.as_struct()
        q = secondaries.as_struct()
        _capi.check(engine.lib.g4hb200_electron_perform(engine.handle, C.byref(s), C.byref(q), seed, _stream_ptr(stream)),
                    "electron_perform")

    @staticmethod
    def Step(engine, batch, secondaries, seed, stream=None):
        s = batch.as_struct()
        q = secondaries.as_struct()
        _capi.check(engine.lib.g4hb200_electron_step(engine.handle, C.byref(s), C.byref(q), seed, _stream_ptr(stream)),
                    "electron_step")


    # ---- the track-level statics the production caller drives one by one (G4HepEmElectronManager.hh:90-206) -------------
    @staticmethod
    def _op(engine, op, batch, secondaries=None, seed=0, flags=None, stream=None):
        s = batch.as_struct()
        q = secondaries.as_struct() if secondaries is not None else None
        _capi.check(engine.lib.g4hb200_electron_track_op(engine.handle, op, C.byref(s), C.byref(q) if q is not None else None, seed,
                                                         flags.data_ptr() if flags is not None else None, _stream_ptr(stream)),
                    f"electron_track_op({op})")

    @staticmethod
    def ResampleNumIALeft(engine, batch, seed, stream=None):
        ElectronManager._op(engine, _capi.OP_RESAMPLE_NIA, batch, seed=seed, stream=stream)

    @staticmethod
    def HowFarToDiscreteInteraction(engine, batch, stream=None):
        ElectronManager._op(engine, _capi.OP_HOWFAR_DISCRETE, batch, stream=stream)

    @staticmethod
    def HowFarToMSC(engine, batch, seed, stream=None):
        ElectronManager._op(engine, _capi.OP_HOWFAR_MSC, batch, seed=seed, stream=stream)

    @staticmethod
    def UpdatePStepLength(engine, batch, stream=None):
        ElectronManager._op(engine, _capi.OP_UPDATE_PSTEP, batch, stream=stream)

    @staticmethod
    def UpdateNumIALeft(engine, batch, stream=None):
        ElectronManager._op(engine, _capi.OP_UPDATE_NIA, batch, stream=stream)

    @staticmethod
    def ApplyMeanEnergyLoss(engine, batch, stopped=None, stream=None):
        ElectronManager._op(engine, _capi.OP_MEAN_ELOSS, batch, flags=stopped, stream=stream)

    @staticmethod
    def SampleMSC(engine, batch, seed, stream=None):
        ElectronManager._op(engine, _capi.OP_SAMPLE_MSC, batch, seed=seed, stream=stream)

    @staticmethod
    def SampleLossFluctuations(engine, batch, seed, stopped=None, stream=None):
        ElectronManager._op(engine, _capi.OP_LOSS_FLUCT, batch, seed=seed, flags=stopped, stream=stream)

    @staticmethod
    def PerformContinuous(engine, batch, seed, stopped=None, stream=None):
        ElectronManager._op(engine, _capi.OP_PERFORM_CONTINUOUS, batch, seed=seed, flags=stopped, stream=stream)

    @staticmethod
    def PerformDiscrete(engine, batch, secondaries, seed, stream=None):
        ElectronManager._op(engine, _capi.OP_DISCRETE, batch, secondaries, seed, stream=stream)

    @staticmethod
    def AnnihilateAtRest(engine, batch, secondaries, seed, stream=None):
        ElectronManager._op(engine, _capi.OP_ANNIHILATE_AT_REST, batch, secondaries, seed, stream=stream)

    @staticmethod
    def CheckDelta(engine, batch, urnd, flags, stream=None):
        s = batch.as_struct()
        _capi.check(engine.lib.g4hb200_electron_check_delta(engine.handle, C.byref(s), urnd.data_ptr(), flags.data_ptr(),
                                                            _stream_ptr(stream)), "electron_check_delta")


class GammaManager:
    """Batch counterpart of G4HepEmGammaManager (G4HepEmRun/include/G4HepEmGammaManager.hh:21-57)."""

    @staticmethod
    def HowFar(engine, batch, seed, stream=None):
        s = batch.as_struct()
        _capi.check(engine.lib.g4hb200_gamma_howfar(engine.handle, C.byref(s), seed, _stream_ptr(stream)), "gamma_howfar")

    @staticmethod
    def Perform(engine, batch, secondaries, seed, stream=None):
        """SelectInteraction (when not on boundary) + Perform, as the reference's callers sequence them."""
        s = batch.as_struct()
        q = secondaries.as_struct()
        _capi.check(engine.lib.g4hb200_gamma_perform(engine.handle, C.byref(s), C.byref(q), seed, _stream_ptr(stream)),
                    "gamma_perform")

    @staticmethod
    def Step(engine, batch, secondaries, seed, stream=None):
        s = batch.as_struct()
        q = secondaries.as_struct()
        _capi.check(engine.lib.g4hb200_gamma_step(engine.handle, C.byref(s), C.byref(q), seed, _stream_ptr(stream)),
                    "gamma_step")

    # ---- the track-level statics (G4HepEmGammaManager.hh:34-51) --------------------------------------------------------------
    @staticmethod
    def _op(engine, op, batch, secondaries=None, seed=0, stream=None):
        s = batch.as_struct()
        q = secondaries.as_struct() if secondaries is not None else None
        _capi.check(engine.lib.g4hb200_gamma_track_op(engine.handle, op, C.byref(s), C.byref(q) if q is not None else None, seed,
                                                      _stream_ptr(stream)), f"gamma_track_op({op})")

    @staticmethod
    def HowFarTrack(engine, batch, stream=None):
        GammaManager._op(engine, _capi.GOP_HOWFAR_TRACK, batch, stream=stream)

    @staticmethod
    def UpdateNumIALeft(engine, batch, stream=None):
        GammaManager._op(engine, _capi.GOP_UPDATE_NIA, batch, stream=stream)

    @staticmethod
    def SelectInteraction(engine, batch, seed, stream=None):
        GammaManager._op(engine, _capi.GOP_SELECT_INTERACTION, batch, seed=seed, stream=stream)

    @staticmethod
    def PerformSelected(engine, batch, secondaries, seed, stream=None):
        GammaManager._op(engine, _capi.GOP_PERFORM_SELECTED, batch, secondaries, seed, stream=stream)
