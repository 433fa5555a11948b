"""ctypes binding of the pre-flight host build of the device functions (tests only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from g4hepem_b200 import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_hostsim.so")


def build():
    src = os.path.join(_HERE, "hostsim.cc")
    deps = [src] + [os.path.join(_HERE, "../../g4hepem_b200/csrc", f) for f in os.listdir(os.path.join(_HERE, "../../g4hepem_b200/csrc")) if f.endswith(".cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) > os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-x", "c++", "-shared",
                           "-I" + os.path.join(_HERE, "../../include"), src, "-o", LIB, "-lm"])


_vp = C.c_void_p


def _p(a):
    return a.ctypes.data_as(_vp)


class HostSim:
    def __init__(self, flat_tables):
        build()
        self.lib = C.CDLL(LIB)
        self.ft = flat_tables
        self.t = C.byref(flat_tables.desc)

    def electron_lookups(self, imc, ekin, lekin, is_electron=True):
        out = np.zeros((7, len(imc)))
        self.lib.g4hsim_electron_lookups(self.t, C.c_int64(len(imc)), _p(imc), _p(ekin), _p(lekin), int(is_electron), _p(out))
        return out

    def electron_stepping_xsecs(self, imc, ekin, lekin, is_electron=True):
        out = np.zeros((4, len(imc)))
        self.lib.g4hsim_electron_stepping_xsecs(self.t, C.c_int64(len(imc)), _p(imc), _p(ekin), _p(lekin), int(is_electron), _p(out))
        return out

    def gamma_lookups(self, imc, ekin, lekin, urnd):
        mx = np.zeros(len(imc))
        pid = np.zeros(len(imc), dtype=np.int32)
        self.lib.g4hsim_gamma_lookups(self.t, C.c_int64(len(imc)), _p(imc), _p(ekin), _p(lekin), _p(urnd), _p(mx), _p(pid))
        return mx, pid

    def select_target_element(self, kind, is_electron, imc, ekin, lekin, urnd):
        out = np.zeros(len(imc), dtype=np.int32)
        self.lib.g4hsim_select_target_element(self.t, kind, int(is_electron), C.c_int64(len(imc)), _p(imc), _p(ekin), _p(lekin), _p(urnd), _p(out))
        return out

    def vdt_log_exp(self, x):
        lo, ex = np.zeros_like(x), np.zeros_like(x)
        self.lib.g4hsim_vdt_log_exp(C.c_int64(len(x)), _p(x), _p(lo), _p(ex))
        return lo, ex

    def rng_uniforms(self, seed, track_id, ndraw):
        out = np.zeros((len(track_id), ndraw))
        self.lib.g4hsim_rng_uniforms(C.c_uint64(seed), C.c_int64(len(track_id)), _p(track_id), C.c_int32(ndraw), _p(out))
        return out

    def _run(self, fn, batch, sec, seed, mode):
        s = batch.as_struct()
        q = sec.as_struct() if sec is not None else _capi.SecondaryQueue()
        fn(self.t, C.byref(s), C.byref(q), C.c_uint64(seed), mode)

    def electron_howfar(self, b, seed):
        self._run(self.lib.g4hsim_electron, b, None, seed, 0)

    def electron_howfar_staged(self, b, seed):
        s = b.as_struct()
        self.lib.g4hsim_electron_howfar_staged(self.t, C.byref(s), C.c_uint64(seed))

    def electron_perform(self, b, sec, seed):
        self._run(self.lib.g4hsim_electron, b, sec, seed, 1)

    def electron_perform_staged(self, b, sec, seed):
        s, q = b.as_struct(), sec.as_struct()
        self.lib.g4hsim_electron_perform_staged(self.t, C.byref(s), C.byref(q), C.c_uint64(seed), 0)

    def electron_step_staged(self, b, sec, seed):
        s, q = b.as_struct(), sec.as_struct()
        self.lib.g4hsim_electron_perform_staged(self.t, C.byref(s), C.byref(q), C.c_uint64(seed), 1)

    def electron_step(self, b, sec, seed):
        self._run(self.lib.g4hsim_electron, b, sec, seed, 2)

    def gamma_howfar(self, b, seed):
        self._run(self.lib.g4hsim_gamma, b, None, seed, 0)

    def gamma_perform(self, b, sec, seed):
        self._run(self.lib.g4hsim_gamma, b, sec, seed, 1)

    def gamma_step(self, b, sec, seed):
        self._run(self.lib.g4hsim_gamma, b, sec, seed, 2)

    def gamma_perform_staged(self, b, sec, seed):
        self._run(self.lib.g4hsim_gamma_staged, b, sec, seed, 1)

    def gamma_step_staged(self, b, sec, seed):
        self._run(self.lib.g4hsim_gamma_staged, b, sec, seed, 2)
