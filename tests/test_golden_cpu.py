"""CPU checks against the committed golden vectors (generated from the unmodified reference by
tests/golden/make_golden.py): (1) the oracle library built in this checkout reproduces them bit for bit (the oracle
is pinned), (2) the host build of the product's device functions (tests/hostsim) reproduces them."""
import numpy as np
import pytest

from g4hepem_b200 import batches
from tests import compare, golden_io

SEED = 2026


@pytest.fixture(scope="module")
def sim(flat_tables):
    from tests.hostsim.hostsim import HostSim

    return HostSim(flat_tables)


def _check_lookups(impl):
    z = golden_io.load("lookups.npz")
    assert np.array_equal(impl.vdt_log_exp(z["vdt_x"])[0], z["vdt_log"])
    assert np.array_equal(impl.vdt_log_exp(z["vdt_xe"])[1], z["vdt_exp"])
    imc, ek, lek, u = z["lk_imc"], z["lk_ekin"], z["lk_lekin"], z["lk_u"]
    for isel, tag in ((True, "em"), (False, "ep")):
        assert np.array_equal(impl.electron_lookups(imc, ek, lek, isel), z[f"lk_{tag}"])
        assert np.array_equal(impl.electron_stepping_xsecs(imc, ek, lek, isel), z[f"sx_{tag}"])
    mx, pid = impl.gamma_lookups(imc, ek, lek, u)
    assert np.array_equal(mx, z["gm_mxsec"]) and np.array_equal(pid, z["gm_pid"])
    for kind, idx in ((0, z["sel_couples"]), (1, z["sel_couples"]), (2, z["sel_mats"])):
        for isel, tag in ((True, "em"), (False, "ep")):
            assert np.array_equal(impl.select_target_element(kind, isel, idx, ek, lek, u), z[f"sel_{kind}_{tag}"])


def _check_steps(impl, exact):
    z = golden_io.load("electron_steps.npz")
    b = golden_io.electron_batch(z, "in_")
    impl.electron_howfar(b, SEED)
    rep = compare.compare_electron_batches(golden_io.electron_batch(z, "howfar_"), b)
    assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
    b = golden_io.electron_batch(z, "geom_")
    q = batches.SecondaryHostQueue(2 * b.n)
    impl.electron_perform(b, q, SEED)
    want = golden_io.electron_batch(z, "perform_")
    rep = compare.compare_electron_batches(want, b)
    assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
    assert compare.total_bad(compare.compare_secondaries(golden_io.GoldenSecondaries(z, "perform_sec_"), q)) == 0
    b = golden_io.electron_batch(z, "in_")
    q = batches.SecondaryHostQueue(2 * b.n)
    impl.electron_step(b, q, SEED)
    want = golden_io.electron_batch(z, "step_")
    rep = compare.compare_electron_batches(want, b, handover=False)
    assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
    assert compare.total_bad(compare.compare_secondaries(golden_io.GoldenSecondaries(z, "step_sec_"), q)) == 0
    if exact:
        assert np.array_equal(want.ekin_logekin, b.ekin_logekin) and np.array_equal(want.winner, b.winner)
    z = golden_io.load("gamma_steps.npz")
    g = golden_io.gamma_batch(z, "in_")
    q = batches.SecondaryHostQueue(2 * g.n)
    impl.gamma_step(g, q, SEED)
    rep = compare.compare_gamma_batches(golden_io.gamma_batch(z, "step_"), g)
    assert compare.total_bad(rep) == 0, compare.format_report(rep, True)
    assert compare.total_bad(compare.compare_secondaries(golden_io.GoldenSecondaries(z, "step_sec_"), q)) == 0


def test_oracle_reproduces_golden_lookups(reference):
    _check_lookups(reference)


def test_oracle_reproduces_golden_steps(reference):
    _check_steps(reference, exact=True)


def test_hostsim_reproduces_golden_lookups(sim):
    _check_lookups(sim)


def test_hostsim_reproduces_golden_steps(sim):
    _check_steps(sim, exact=False)
