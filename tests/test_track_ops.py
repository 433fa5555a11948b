"""The track-level entry points (g4hb200_electron_track_op / g4hb200_gamma_track_op / g4hb200_electron_check_delta:
include/g4hepem_b200.h) against the reference's statics of the same name, piece by piece in the order
G4HepEmTrackingManager::TrackElectron / TrackGamma call them (G4HepEmTrackingManager.cc:428-665, 985-1140).

Before every op the device gets the reference's (already compared) state, so differences do not accumulate; an op is
compared on the tracks the production caller would hand it (e.g. no SampleMSC for a track ApplyMeanEnergyLoss stopped)."""
import numpy as np
import pytest

from g4hepem_b200 import _capi, batches
from tests import compare

SEED = 2026
pytestmark = pytest.mark.gpu


def _subset(b, mask):
    o = type(b)(int(mask.sum()))
    for g in b.groups() + ("meta", "winner"):
        getattr(o, g)[...] = getattr(b, g)[mask]
    return o


def _check_electron(want, got, mask, what):
    rep = compare.compare_electron_batches(_subset(want, mask), _subset(got, mask), handover=True)
    compare.compare_group("prestep", _subset(want, mask).prestep, _subset(got, mask).prestep, ("rel", "rel"), rep)
    assert compare.total_bad(rep) == 0, what + "\n" + compare.format_report(rep, True)


def test_electron_track_level_ops_in_tracking_manager_order(engine, reference, flat_tables):
    import torch

    from g4hepem_b200 import engine as eng

    n = 200000
    a = batches.make_electron_batch(n, flat_tables.num_matcut, seed=61)
    a.prestep[...] = 0.0
    dev = eng.ElectronDeviceBatch(n)
    sec = eng.SecondaryDeviceQueue(2 * n)
    flags = torch.zeros(n, dtype=torch.int32, device="cuda")
    everyone = np.ones(n, dtype=bool)
    EM = eng.ElectronManager

    def both(op, gpu_call, mask, what, use_sec=False, want_flags=False):
        dev.upload(a)
        qa = batches.SecondaryHostQueue(2 * n) if use_sec else None
        fa = np.zeros(n, dtype=np.int32) if want_flags else None
        reference.electron_track_op(op, a, SEED, qa, fa)
        sec.reset()
        flags.zero_()
        gpu_call()
        torch.cuda.synchronize()
        got = dev.download()
        _check_electron(a, got, mask, what)
        if want_flags:
            assert np.array_equal(fa[mask], flags.cpu().numpy()[mask]), what + ": returned bools differ"
        if use_sec:
            qg = sec.download()
            ra, rg = qa.sorted_records(), qg.sorted_records()
            keep_a, keep_g = mask[ra["parent_index"]], mask[rg["parent_index"]]
            assert int(keep_a.sum()) == int(keep_g.sum()), what + ": secondary counts differ"
            for k in ("parent_index", "slot", "parent_id", "kind"):
                assert np.array_equal(ra[k][keep_a], rg[k][keep_g]), (what, k)
            assert compare.rel_close(ra["ekin"][keep_a], rg["ekin"][keep_g]).all(), what
            assert compare.rel_close(ra["dir"][keep_a], rg["dir"][keep_g], compare.REL, compare.ABS_DIR).all(), what
            return fa, int(keep_a.sum())
        return fa, 0

    rng = np.random.default_rng(4)
    total_sec = 0
    for step in range(3):
        both(_capi.OP_RESAMPLE_NIA, lambda: EM.ResampleNumIALeft(engine, dev, SEED), everyone, "ResampleNumIALeft")
        both(_capi.OP_HOWFAR_DISCRETE, lambda: EM.HowFarToDiscreteInteraction(engine, dev), everyone, "HowFarToDiscreteInteraction")
        # SavePreStepEKin (G4HepEmTrackingManager.cc:445): the caller's own statement
        a.prestep[:, 0] = a.ekin_logekin[:, 0]
        a.prestep[:, 1] = a.ekin_logekin[:, 1]
        both(_capi.OP_HOWFAR_MSC, lambda: EM.HowFarToMSC(engine, dev, SEED), everyone, "HowFarToMSC")
        # geometry stub: 20 % of the steps are cut short and end on a boundary
        cut = rng.uniform(size=n) < 0.2
        a.gstep_pstep[cut, 0] *= rng.uniform(0.3, 1.0, n)[cut]
        a.meta[:, 1] = np.where(cut, a.meta[:, 1] | _capi.F_ON_BOUNDARY, a.meta[:, 1] & ~_capi.F_ON_BOUNDARY)
        moving = a.gstep_pstep[:, 0] > 0.0
        both(_capi.OP_UPDATE_PSTEP, lambda: EM.UpdatePStepLength(engine, dev), moving, "UpdatePStepLength")
        moving &= a.gstep_pstep[:, 1] > 0.0
        both(_capi.OP_UPDATE_NIA, lambda: EM.UpdateNumIALeft(engine, dev), moving, "UpdateNumIALeft")
        stopped, _ = both(_capi.OP_MEAN_ELOSS, lambda: EM.ApplyMeanEnergyLoss(engine, dev, flags), moving, "ApplyMeanEnergyLoss",
                          want_flags=True)
        alive = moving & (stopped == 0)
        both(_capi.OP_SAMPLE_MSC, lambda: EM.SampleMSC(engine, dev, SEED), alive, "SampleMSC")
        stopped2, _ = both(_capi.OP_LOSS_FLUCT, lambda: EM.SampleLossFluctuations(engine, dev, SEED, flags), alive,
                           "SampleLossFluctuations", want_flags=True)
        is_pos = (a.meta[:, 1] & _capi.F_POSITRON) != 0
        at_rest = moving & is_pos & ((stopped != 0) | (alive & (stopped2 != 0)))
        alive &= stopped2 == 0
        # CheckDelta with caller supplied uniforms (the lepto-nuclear branch of the caller, .cc:655-665)
        has_winner = alive & (a.winner >= 0)
        u = rng.uniform(size=n)
        fa = np.zeros(n, dtype=np.int32)
        dev.upload(a)
        reference.electron_check_delta(a, u, fa)
        ud = torch.from_numpy(u).cuda()
        EM.CheckDelta(engine, dev, ud, flags)
        torch.cuda.synchronize()
        assert np.array_equal(fa[has_winner], flags.cpu().numpy()[has_winner]), "CheckDelta"
        _, ns = both(_capi.OP_DISCRETE, lambda: EM.PerformDiscrete(engine, dev, sec, SEED), alive, "PerformDiscrete", use_sec=True)
        total_sec += ns
        _, ns = both(_capi.OP_ANNIHILATE_AT_REST, lambda: EM.AnnihilateAtRest(engine, dev, sec, SEED), at_rest, "annihilation at rest",
                     use_sec=True)
        total_sec += ns
        # next step: stopped tracks are re-born
        dead = a.ekin_logekin[:, 0] <= 0
        a.ekin_logekin[dead, 0] = 1.0
        a.ekin_logekin[dead, 1] = 100.0
    assert total_sec > n // 4


def test_perform_continuous_op(engine, reference, flat_tables):
    """PerformContinuous(data, pars, elTrack, rng) (G4HepEmElectronManager.hh:184) as one op, after HowFar."""
    import torch

    from g4hepem_b200 import engine as eng

    n = 100000
    a = batches.make_electron_batch(n, flat_tables.num_matcut, seed=62)
    a.prestep[...] = 0.0
    reference.electron_howfar(a, SEED, 4)
    dev = eng.ElectronDeviceBatch(n)
    dev.upload(a)
    fa = np.zeros(n, dtype=np.int32)
    reference.electron_track_op(_capi.OP_PERFORM_CONTINUOUS, a, SEED, None, fa)
    flags = torch.zeros(n, dtype=torch.int32, device="cuda")
    eng.ElectronManager.PerformContinuous(engine, dev, SEED, flags)
    torch.cuda.synchronize()
    _check_electron(a, dev.download(), np.ones(n, dtype=bool), "PerformContinuous")
    assert np.array_equal(fa, flags.cpu().numpy())


def test_gamma_track_level_ops_in_tracking_manager_order(engine, reference, flat_tables):
    import torch

    from g4hepem_b200 import engine as eng

    n = 200000
    g = batches.make_gamma_batch(n, flat_tables.num_matcut, seed=63)
    # the interaction length is the caller's to sample for HowFar(data, pars, gammaTrack)
    g.dirz_nia0[:, 1] = -np.log(np.random.default_rng(1).uniform(size=n))
    dev = eng.GammaDeviceBatch(n)
    sec = eng.SecondaryDeviceQueue(2 * n)
    GM = eng.GammaManager
    pe_mask = g.ekin_logekin[:, 0] <= flat_tables.desc.gm_emax1
    everyone = np.ones(n, dtype=bool)

    def both(op, gpu_call, mask, what, use_sec=False):
        dev.upload(g)
        qa = batches.SecondaryHostQueue(2 * n) if use_sec else None
        reference.gamma_track_op(op, g, SEED, qa)
        sec.reset()
        gpu_call()
        torch.cuda.synchronize()
        got = dev.download()
        rep = compare.compare_gamma_batches(_subset(g, mask), _subset(got, mask), pe_mask=pe_mask[mask])
        assert compare.total_bad(rep) == 0, what + "\n" + compare.format_report(rep, True)
        if use_sec:
            ra, rg = qa.sorted_records(), sec.download().sorted_records()
            keep_a, keep_g = mask[ra["parent_index"]], mask[rg["parent_index"]]
            assert int(keep_a.sum()) == int(keep_g.sum()), what
            for k in ("parent_index", "slot", "parent_id", "kind"):
                assert np.array_equal(ra[k][keep_a], rg[k][keep_g]), (what, k)
            assert compare.rel_close(ra["ekin"][keep_a], rg["ekin"][keep_g]).all(), what
            assert compare.rel_close(ra["dir"][keep_a], rg["dir"][keep_g], compare.REL, compare.ABS_DIR).all(), what
            return int(keep_a.sum())
        return 0

    both(_capi.GOP_HOWFAR_TRACK, lambda: GM.HowFarTrack(engine, dev), everyone, "HowFar(track)")
    rng = np.random.default_rng(8)
    onb = rng.uniform(size=n) < 0.3
    g.gstep_mfp0[onb, 0] *= rng.uniform(0.1, 1.0, n)[onb]
    g.meta[:, 1] = np.where(onb, _capi.F_ON_BOUNDARY, 0).astype(np.int32)
    both(_capi.GOP_UPDATE_NIA, lambda: GM.UpdateNumIALeft(engine, dev), onb, "UpdateNumIALeft")
    both(_capi.GOP_SELECT_INTERACTION, lambda: GM.SelectInteraction(engine, dev, SEED), ~onb, "SelectInteraction")
    nsec = both(_capi.GOP_PERFORM_SELECTED, lambda: GM.PerformSelected(engine, dev, sec, SEED), ~onb, "Perform", use_sec=True)
    assert nsec > n // 4
