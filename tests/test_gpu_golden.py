"""GPU parity against the committed golden vectors (tests/golden/*.npz, generated from the unmodified reference):
needs neither /root/reference nor the oracle library at run time."""
import numpy as np
import pytest

from g4hepem_b200 import batches
from tests import compare, golden_io

pytestmark = pytest.mark.gpu
SEED = 2026


def _cuda(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_lookups_against_golden(engine):
    z = golden_io.load("lookups.npz")
    assert np.array_equal(engine.vdt_log_exp(_cuda(z["vdt_x"]))[0].cpu().numpy(), z["vdt_log"])
    assert np.array_equal(engine.vdt_log_exp(_cuda(z["vdt_xe"]))[1].cpu().numpy(), z["vdt_exp"])
    imc, ek, lek, u = (_cuda(z[k]) for k in ("lk_imc", "lk_ekin", "lk_lekin", "lk_u"))
    for isel, tag in ((True, "em"), (False, "ep")):
        assert np.array_equal(engine.electron_lookups(imc, ek, lek, isel).cpu().numpy(), z[f"lk_{tag}"])
        assert np.array_equal(engine.electron_stepping_xsecs(imc, ek, lek, isel).cpu().numpy(), z[f"sx_{tag}"])
    mx, pid = engine.gamma_lookups(imc, ek, lek, u)
    assert np.array_equal(mx.cpu().numpy(), z["gm_mxsec"]) and np.array_equal(pid.cpu().numpy(), z["gm_pid"])
    for kind, key in ((0, "sel_couples"), (1, "sel_couples"), (2, "sel_mats")):
        for isel, tag in ((True, "em"), (False, "ep")):
            got = engine.select_target_element(kind, isel, _cuda(z[key]), ek, lek, u).cpu().numpy()
            assert np.array_equal(got, z[f"sel_{kind}_{tag}"])


def _run_electron(engine, host, mode):
    import torch

    from g4hepem_b200 import engine as eng

    dev = eng.ElectronDeviceBatch(host.n)
    sec = eng.SecondaryDeviceQueue(2 * host.n)
    dev.upload(host)
    {"howfar": lambda: eng.ElectronManager.HowFar(engine, dev, SEED),
     "perform": lambda: eng.ElectronManager.Perform(engine, dev, sec, SEED),
     "step": lambda: eng.ElectronManager.Step(engine, dev, sec, SEED)}[mode]()
    torch.cuda.synchronize()
    return dev.download(), sec.download()


def test_electron_steps_against_golden(engine):
    z = golden_io.load("electron_steps.npz")
    got, _ = _run_electron(engine, golden_io.electron_batch(z, "in_"), "howfar")
    rep = compare.compare_electron_batches(golden_io.electron_batch(z, "howfar_"), got)
    assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
    got, q = _run_electron(engine, golden_io.electron_batch(z, "geom_"), "perform")
    rep = compare.compare_electron_batches(golden_io.electron_batch(z, "perform_"), got)
    assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
    assert compare.total_bad(compare.compare_secondaries(golden_io.GoldenSecondaries(z, "perform_sec_"), q)) == 0
    got, q = _run_electron(engine, golden_io.electron_batch(z, "in_"), "step")
    want = golden_io.electron_batch(z, "step_")
    rep = compare.compare_electron_batches(want, got, handover=False)
    assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
    assert compare.total_bad(compare.compare_secondaries(golden_io.GoldenSecondaries(z, "step_sec_"), q)) == 0
    # discrete outcomes bit exact
    assert np.array_equal(want.winner, got.winner) and np.array_equal(want.meta, got.meta)


def test_gamma_step_against_golden(engine):
    import torch

    from g4hepem_b200 import engine as eng

    z = golden_io.load("gamma_steps.npz")
    g = golden_io.gamma_batch(z, "in_")
    dev = eng.GammaDeviceBatch(g.n)
    sec = eng.SecondaryDeviceQueue(2 * g.n)
    dev.upload(g)
    eng.GammaManager.Step(engine, dev, sec, SEED)
    torch.cuda.synchronize()
    got = dev.download()
    want = golden_io.gamma_batch(z, "step_")
    rep = compare.compare_gamma_batches(want, got)
    assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
    assert compare.total_bad(compare.compare_secondaries(golden_io.GoldenSecondaries(z, "step_sec_"), sec.download())) == 0
    assert np.array_equal(want.winner, got.winner) and np.array_equal(want.meta, got.meta)
