"""Pre-flight (CPU): the per-track device functions built for the host agree with the reference.

This is a transcription check that runs without a GPU; the parity claims rest on tests/test_gpu_*.py.
"""
import numpy as np
import pytest

from g4hepem_b200 import _capi, batches
from tests import compare


@pytest.fixture(scope="module")
def sim(flat_tables):
    from tests.hostsim.hostsim import HostSim

    return HostSim(flat_tables)


def test_vdt_and_lookups_bit_exact(sim, reference, flat_tables):
    rng = np.random.default_rng(1)
    n = 50000
    x = np.exp(rng.uniform(np.log(1e-300), np.log(1e300), n))
    assert np.array_equal(reference.vdt_log_exp(x)[0], sim.vdt_log_exp(x)[0])
    xe = rng.uniform(-720, 720, n)
    assert np.array_equal(reference.vdt_log_exp(xe)[1], sim.vdt_log_exp(xe)[1])
    imc = rng.integers(0, flat_tables.num_matcut, n).astype(np.int32)
    ek = np.exp(rng.uniform(np.log(0.5e-4), np.log(2e8), n))
    lek = np.log(ek)
    for isel in (True, False):
        assert np.array_equal(reference.electron_lookups(imc, ek, lek, isel), sim.electron_lookups(imc, ek, lek, isel))
        assert np.array_equal(reference.electron_stepping_xsecs(imc, ek, lek, isel), sim.electron_stepping_xsecs(imc, ek, lek, isel))
    u = rng.uniform(size=n)
    a, b = reference.gamma_lookups(imc, ek, lek, u), sim.gamma_lookups(imc, ek, lek, u)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("staged", [False, True])
def test_electron_multi_step_with_geometry_stub(sim, reference, flat_tables, staged):
    n = 20000
    a = batches.make_electron_batch(n, flat_tables.num_matcut, seed=5)
    b = a.copy()
    rng = np.random.default_rng(3)
    for _ in range(4):
        qa, qb = batches.SecondaryHostQueue(2 * n), batches.SecondaryHostQueue(2 * n)
        reference.electron_howfar(a, 2026, 4)
        (sim.electron_howfar_staged if staged else sim.electron_howfar)(b, 2026)
        rep = compare.compare_electron_batches(a, b)
        assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
        cut = rng.uniform(size=n) < 0.2
        f = rng.uniform(0.3, 1.0, n)
        for x in (a, b):
            x.gstep_pstep[cut, 0] *= f[cut]
            x.meta[:, 1] = np.where(cut, x.meta[:, 1] | _capi.F_ON_BOUNDARY, x.meta[:, 1] & ~_capi.F_ON_BOUNDARY)
        reference.electron_perform(a, qa, 2026, 4)
        (sim.electron_perform_staged if staged else sim.electron_perform)(b, qb, 2026)
        rep = compare.compare_electron_batches(a, b)
        assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
        assert compare.total_bad(compare.compare_secondaries(qa, qb)) == 0
        dead = a.ekin_logekin[:, 0] <= 0
        for x in (a, b):
            x.ekin_logekin[dead, 0] = 1.0
            x.ekin_logekin[dead, 1] = 100.0


@pytest.mark.parametrize("staged", [False, True])
def test_gamma_step(sim, reference, flat_tables, staged):
    n = 30000
    g = batches.make_gamma_batch(n, flat_tables.num_matcut, boundary_fraction=0.1)
    h = g.copy()
    qa, qb = batches.SecondaryHostQueue(2 * n), batches.SecondaryHostQueue(2 * n)
    reference.gamma_step(g, qa, 2026, 4)
    (sim.gamma_step_staged if staged else sim.gamma_step)(h, qb, 2026)
    rep = compare.compare_gamma_batches(g, h)
    assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
    assert compare.total_bad(compare.compare_secondaries(qa, qb)) == 0


def test_electron_fused_step_staged(sim, reference, flat_tables):
    """StageStepHead (HowFar + along-step in one pass) + the queue stages over consecutive steps."""
    n = 30000
    a = batches.make_electron_batch(n, flat_tables.num_matcut, seed=9)
    b = a.copy()
    for _ in range(4):
        qa, qb = batches.SecondaryHostQueue(2 * n), batches.SecondaryHostQueue(2 * n)
        reference.electron_step(a, qa, 2026, 4)
        sim.electron_step_staged(b, qb, 2026)
        rep = compare.compare_electron_batches(a, b, handover=False)
        assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
        assert compare.total_bad(compare.compare_secondaries(qa, qb)) == 0
        dead = a.ekin_logekin[:, 0] <= 0
        for x in (a, b):
            x.ekin_logekin[dead, 0] = 1.0
            x.ekin_logekin[dead, 1] = 100.0


def test_fused_step_with_odd_draw_counters(sim, reference, flat_tables):
    """Fresh tracks (all four interaction lengths resampled) whose streams stand at an odd draw: the head stage then
    needs the sixth uniform of its window from a fourth Philox block (DrawWindow::Sixth)."""
    n = 20000
    a = batches.make_electron_batch(n, flat_tables.num_matcut, seed=21)
    a.meta[:, 3] = 2 * np.arange(n, dtype=np.int32) % 1000 + 1
    b = a.copy()
    qa, qb = batches.SecondaryHostQueue(2 * n), batches.SecondaryHostQueue(2 * n)
    reference.electron_step(a, qa, 2026, 4)
    sim.electron_step_staged(b, qb, 2026)
    rep = compare.compare_electron_batches(a, b, handover=False)
    assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
    assert compare.total_bad(compare.compare_secondaries(qa, qb)) == 0
