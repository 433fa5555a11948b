"""ctypes binding of oracle/_ref/libg4hepem_ref.so (TEST INFRASTRUCTURE).

The library is the unmodified reference (mnovak42/g4hepem) + oracle/ref_shim.cc, built by
oracle/Makefile.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg import this module; the product (g4hepem_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

from g4hepem_b200 import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libg4hepem_ref.so")


def available():
    return os.path.exists(REF_LIB)


_vp = C.c_void_p
_PROTOS = {
    "g4href_load_state": (_vp, [C.c_char_p]),
    "g4href_free_state": (None, [_vp]),
    "g4href_save_state": (C.c_int, [_vp, C.c_char_p]),
    "g4href_flat_tables": (C.POINTER(_capi.Tables), [_vp]),
    "g4href_hardware_threads": (C.c_int, []),
    "g4href_electron_lookups": (None, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int, _vp]),
    "g4href_electron_stepping_xsecs": (None, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int, _vp]),
    "g4href_gamma_lookups": (None, [_vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4href_select_target_element": (None, [_vp, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "g4href_vdt_log_exp": (None, [C.c_int64, _vp, _vp, _vp]),
    "g4href_rng_uniforms": (None, [C.c_uint64, C.c_int64, _vp, C.c_int32, _vp]),
    "g4href_electron_howfar": (C.c_int, [_vp, C.POINTER(_capi.ElectronBatch), C.c_uint64, C.c_int]),
    "g4href_electron_perform": (C.c_int, [_vp, C.POINTER(_capi.ElectronBatch), C.POINTER(_capi.SecondaryQueue), C.c_uint64, C.c_int]),
    "g4href_electron_step": (C.c_int, [_vp, C.POINTER(_capi.ElectronBatch), C.POINTER(_capi.SecondaryQueue), C.c_uint64, C.c_int]),
    "g4href_gamma_howfar": (C.c_int, [_vp, C.POINTER(_capi.GammaBatch), C.c_uint64, C.c_int]),
    "g4href_gamma_perform": (C.c_int, [_vp, C.POINTER(_capi.GammaBatch), C.POINTER(_capi.SecondaryQueue), C.c_uint64, C.c_int]),
    "g4href_electron_track_op": (C.c_int, [_vp, C.c_int, C.POINTER(_capi.ElectronBatch), C.POINTER(_capi.SecondaryQueue), C.c_uint64, _vp]),
    "g4href_electron_check_delta": (C.c_int, [_vp, C.POINTER(_capi.ElectronBatch), _vp, _vp]),
    "g4href_gamma_track_op": (C.c_int, [_vp, C.c_int, C.POINTER(_capi.GammaBatch), C.POINTER(_capi.SecondaryQueue), C.c_uint64]),
    "g4href_gamma_step": (C.c_int, [_vp, C.POINTER(_capi.GammaBatch), C.POINTER(_capi.SecondaryQueue), C.c_uint64, C.c_int]),
}


def _p(a):
    return a.ctypes.data_as(_vp)


class Reference:
    """The reference's CPU stepping functions on host batches."""

    def __init__(self, json_path):
        if not available():
            raise RuntimeError(f"{REF_LIB} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(REF_LIB)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(self.lib, name)
            fn.restype = res
            fn.argtypes = args
        self.state = self.lib.g4href_load_state(json_path.encode())
        if not self.state:
            raise RuntimeError(f"reference failed to load {json_path}")

    def close(self):
        if self.state:
            self.lib.g4href_free_state(self.state)
            self.state = None

    def flat_tables(self):
        return self.lib.g4href_flat_tables(self.state).contents

    def save_state(self, path):
        return self.lib.g4href_save_state(self.state, path.encode())

    def hardware_threads(self):
        return int(self.lib.g4href_hardware_threads())

    # ---- look-ups
    def electron_lookups(self, imc, ekin, lekin, is_electron=True):
        n = len(imc)
        out = np.zeros((7, n))
        self.lib.g4href_electron_lookups(self.state, n, _p(imc), _p(ekin), _p(lekin), int(is_electron), _p(out))
        return out

    def electron_stepping_xsecs(self, imc, ekin, lekin, is_electron=True):
        n = len(imc)
        out = np.zeros((4, n))
        self.lib.g4href_electron_stepping_xsecs(self.state, n, _p(imc), _p(ekin), _p(lekin), int(is_electron), _p(out))
        return out

    def gamma_lookups(self, imc, ekin, lekin, urnd):
        n = len(imc)
        mx = np.zeros(n)
        pid = np.zeros(n, dtype=np.int32)
        self.lib.g4href_gamma_lookups(self.state, n, _p(imc), _p(ekin), _p(lekin), _p(urnd), _p(mx), _p(pid))
        return mx, pid

    def select_target_element(self, kind, is_electron, imc, ekin, lekin, urnd):
        n = len(imc)
        out = np.zeros(n, dtype=np.int32)
        self.lib.g4href_select_target_element(self.state, kind, int(is_electron), n, _p(imc), _p(ekin), _p(lekin), _p(urnd), _p(out))
        return out

    def vdt_log_exp(self, x):
        lo = np.zeros_like(x)
        ex = np.zeros_like(x)
        self.lib.g4href_vdt_log_exp(len(x), _p(x), _p(lo), _p(ex))
        return lo, ex

    def rng_uniforms(self, seed, track_id, ndraw):
        out = np.zeros((len(track_id), ndraw))
        self.lib.g4href_rng_uniforms(seed, len(track_id), _p(track_id), ndraw, _p(out))
        return out

    # ---- stepping (in place on the host batch)
    def _run(self, fn, batch, sec, seed, nthreads):
        s = batch.as_struct()
        if sec is None:
            rc = fn(self.state, C.byref(s), seed, nthreads)
        else:
            q = sec.as_struct()
            rc = fn(self.state, C.byref(s), C.byref(q), seed, nthreads)
        if rc != 0:
            raise RuntimeError(f"reference driver returned {rc}")

    def electron_howfar(self, batch, seed, nthreads=1):
        self._run(self.lib.g4href_electron_howfar, batch, None, seed, nthreads)

    def electron_perform(self, batch, sec, seed, nthreads=1):
        self._run(self.lib.g4href_electron_perform, batch, sec, seed, nthreads)

    def electron_step(self, batch, sec, seed, nthreads=1):
        self._run(self.lib.g4href_electron_step, batch, sec, seed, nthreads)

    def gamma_howfar(self, batch, seed, nthreads=1):
        self._run(self.lib.g4href_gamma_howfar, batch, None, seed, nthreads)

    def gamma_perform(self, batch, sec, seed, nthreads=1):
        self._run(self.lib.g4href_gamma_perform, batch, sec, seed, nthreads)

    def gamma_step(self, batch, sec, seed, nthreads=1):
        self._run(self.lib.g4href_gamma_step, batch, sec, seed, nthreads)

    # ---- the track-level statics, one at a time (op codes: _capi.OP_*, _capi.GOP_*)
    def electron_track_op(self, op, batch, seed=0, sec=None, flags=None):
        s = batch.as_struct()
        q = sec.as_struct() if sec is not None else None
        rc = self.lib.g4href_electron_track_op(self.state, op, C.byref(s), C.byref(q) if q is not None else None, seed,
                                               _p(flags) if flags is not None else None)
        if rc != 0:
            raise RuntimeError(f"reference driver returned {rc}")

    def electron_check_delta(self, batch, urnd, flags):
        s = batch.as_struct()
        self.lib.g4href_electron_check_delta(self.state, C.byref(s), _p(urnd), _p(flags))

    def gamma_track_op(self, op, batch, seed=0, sec=None):
        s = batch.as_struct()
        q = sec.as_struct() if sec is not None else None
        rc = self.lib.g4href_gamma_track_op(self.state, op, C.byref(s), C.byref(q) if q is not None else None, seed)
        if rc != 0:
            raise RuntimeError(f"reference driver returned {rc}")
