"""The C++ host layer (g4hepem_b200/host/G4HepEmB200Managers.hh) as a drop-in for the reference's managers:
the same G4HepEmElectronTrack / G4HepEmGammaTrack objects go through G4HepEmElectronManager / G4HepEmGammaManager
(CPU, reference) and through G4HepEmB200Session (C++ -> C-ABI -> CUDA) for several steps with a geometry stub in
between; tests/dropin/dropin_test.cc counts the tracks whose public state differs (1e-12 relative, ids exact)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.conftest import ROOT, STATE_JSON

DROPIN = os.path.join(ROOT, "oracle", "_ref", "libg4hepem_dropin.so")


def test_dropin_library_is_built_from_the_host_headers():
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libg4hepem_dropin.so not built (needs /root/reference)")
    lib = C.CDLL(DROPIN)
    assert hasattr(lib, "g4hdropin_run")
    assert hasattr(lib, "g4hdropin_track_level")


@pytest.mark.gpu
def test_cpp_session_matches_reference_managers(engine, flat_tables):
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libg4hepem_dropin.so not built")
    lib = C.CDLL(DROPIN)
    n, nsteps = 20000, 4
    rng = np.random.default_rng(5)
    ekin = np.exp(rng.uniform(np.log(1e-3), np.log(1e4), n))
    imc = rng.integers(0, flat_tables.num_matcut, n).astype(np.int32)
    pos = (rng.uniform(size=n) < 0.5).astype(np.int32)
    cost = rng.uniform(-1, 1, n)
    phi = rng.uniform(0, 2 * np.pi, n)
    sint = np.sqrt(1 - cost * cost)
    d = np.ascontiguousarray(np.stack([sint * np.cos(phi), sint * np.sin(phi), cost], axis=1))
    safety = rng.uniform(0, 1, n)
    report = np.zeros(8, dtype=np.int64)
    vp = C.c_void_p
    lib.g4hdropin_run.restype = C.c_int
    lib.g4hdropin_run.argtypes = [C.c_char_p, C.c_int64, C.c_uint64, C.c_int, vp, vp, vp, vp, vp, vp]
    rc = lib.g4hdropin_run(STATE_JSON.encode(), n, 2026, nsteps, ekin.ctypes.data, imc.ctypes.data, pos.ctypes.data,
                           d.ctypes.data, safety.ctypes.data, report.ctypes.data)
    assert rc == 0, (rc, report)
    assert report[6] > n // 4, "the run produced hardly any secondaries"
    assert report[:6].tolist() == [0, 0, 0, 0, 0, 0], report


@pytest.mark.gpu
def test_dropin_statics_follow_the_tracking_manager_order(engine, flat_tables):
    """G4HepEmB200ElectronManager / G4HepEmB200GammaManager (g4hepem_b200/host/G4HepEmB200DropIn.hh: the reference's static
    signatures, secondaries into G4HepEmTLData) against G4HepEmElectronManager / G4HepEmGammaManager, driven one track at a
    time in the order of G4HepEmTrackingManager::TrackElectron / TrackGamma (G4HepEmTrackingManager.cc:408-665, 985-1140)
    with the MSC sub-step loop and a geometry stub: the same template function instantiated with either manager class."""
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libg4hepem_dropin.so not built")
    lib = C.CDLL(DROPIN)
    n, nsteps = 400, 5
    rng = np.random.default_rng(15)
    ekin = np.exp(rng.uniform(np.log(1e-3), np.log(1e4), n))
    imc = rng.integers(0, flat_tables.num_matcut, n).astype(np.int32)
    pos = (rng.uniform(size=n) < 0.5).astype(np.int32)
    cost = rng.uniform(-1, 1, n)
    phi = rng.uniform(0, 2 * np.pi, n)
    sint = np.sqrt(1 - cost * cost)
    d = np.ascontiguousarray(np.stack([sint * np.cos(phi), sint * np.sin(phi), cost], axis=1))
    safety = rng.uniform(0, 1, n)
    report = np.zeros(8, dtype=np.int64)
    vp = C.c_void_p
    lib.g4hdropin_track_level.restype = C.c_int
    lib.g4hdropin_track_level.argtypes = [C.c_char_p, C.c_int64, C.c_uint64, C.c_int, vp, vp, vp, vp, vp, vp]
    rc = lib.g4hdropin_track_level(STATE_JSON.encode(), n, 2026, nsteps, ekin.ctypes.data, imc.ctypes.data, pos.ctypes.data,
                                   d.ctypes.data, safety.ctypes.data, report.ctypes.data)
    assert rc == 0, (rc, report)
    assert report[5] > n // 2, "the run produced hardly any secondaries"
    assert report[2] > 0, "no step went through the MSC sub-step loop more than once"
    assert [report[k] for k in (0, 1, 3, 4, 6)] == [0, 0, 0, 0, 0], report
