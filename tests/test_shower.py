"""The slab-calorimeter stepping loop (BASELINE configs[4]): CPU restatement sanity (energy bookkeeping, independence of
the sharding) without a GPU; the device loop against the CPU restatement, track population by track population, on the
GPU."""
import numpy as np
import pytest

from g4hepem_b200 import _capi, shower
from tests import shower_oracle

SEED = 2026


def test_oracle_loop_conserves_energy(reference, flat_tables):
    calo = shower.SlabCalorimeter()
    hist, st = shower_oracle.run(reference, calo, 6, 300.0, SEED, tables=flat_tables)
    total = hist.sum() + st["leak_electron"] + st["leak_gamma"]
    # kinetic energy in = deposits + leakage (every e+ of these showers annihilates inside: 2 m_e c^2 taken by the
    # conversion come back as the two annihilation photons)
    assert abs(total - 6 * 300.0) < 1e-6 * 6 * 300.0
    assert st["num_steps"] > 20 and st["secondaries"] > 100
    # lead (absorber 0) takes more than liquid argon per layer around the shower maximum
    assert hist[:10, 0].sum() > hist[:10, 1].sum()


def test_oracle_loop_does_not_depend_on_the_sharding(reference, flat_tables):
    """Streams are keyed by track ids derived from the parent: two halves of the primaries give the whole."""
    calo = shower.SlabCalorimeter(num_layers=20)
    whole, st = shower_oracle.run(reference, calo, 8, 150.0, SEED, tables=flat_tables)
    a, sa = shower_oracle.run(reference, calo, 4, 150.0, SEED, first_track_id=0, tables=flat_tables)
    b, sb = shower_oracle.run(reference, calo, 4, 150.0, SEED, first_track_id=4, tables=flat_tables)
    np.testing.assert_allclose(a + b, whole, rtol=1e-12, atol=1e-12)
    for k in ("electron_track_steps", "gamma_track_steps", "secondaries"):
        assert sa[k] + sb[k] == st[k]


def test_child_streams_are_distinct():
    ids = np.arange(20000, dtype=np.int32)
    draws = (np.arange(20000, dtype=np.int32) * 7) % 1000
    i0, d0 = shower_oracle.child_stream(SEED, ids, draws, np.zeros(20000, dtype=np.int32))
    i1, d1 = shower_oracle.child_stream(SEED, ids, draws, np.ones(20000, dtype=np.int32))
    keys = np.concatenate([i0.astype(np.int64) << 32 | d0, i1.astype(np.int64) << 32 | d1])
    assert len(np.unique(keys)) == keys.size
    assert (d0 % 2 == 0).all() and (d0 >= 0).all() and (d0 < (1 << 30)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,nprim,ekin", [(_capi.SEC_ELECTRON, 12, 400.0), (_capi.SEC_GAMMA, 8, 250.0),
                                             (_capi.SEC_POSITRON, 5, 100.0)])
def test_device_shower_matches_cpu_loop(engine, reference, flat_tables, kind, nprim, ekin):
    calo = shower.SlabCalorimeter()
    want, wst = shower_oracle.run(reference, calo, nprim, ekin, SEED, kind=kind, first_track_id=100, tables=flat_tables)
    got = shower.run(engine, calo, nprim, ekin, SEED, kind=kind, first_track_id=100, capacity=1 << 16)
    for k in ("num_steps", "electron_track_steps", "gamma_track_steps", "secondaries", "peak_electrons", "peak_gammas"):
        assert got.stats[k] == wst[k], (k, got.stats[k], wst[k])
    # same tracks, same deposits; the sums differ by the summation order only
    np.testing.assert_allclose(got.edep, want, rtol=1e-9, atol=1e-9)
    assert abs(got.stats["leak_electron"] - wst["leak_electron"]) <= 1e-9 * max(1.0, wst["leak_electron"])
    assert abs(got.stats["leak_gamma"] - wst["leak_gamma"]) <= 1e-9 * max(1.0, wst["leak_gamma"])
    total = got.edep.sum() + got.stats["leak_electron"] + got.stats["leak_gamma"]
    assert abs(total - nprim * ekin) < 1e-6 * nprim * ekin or kind == _capi.SEC_POSITRON


@pytest.mark.gpu
def test_device_shower_sharding_and_capacity(engine):
    calo = shower.SlabCalorimeter(num_layers=20)
    whole = shower.run(engine, calo, 64, 200.0, SEED)
    a = shower.run(engine, calo, 32, 200.0, SEED, first_track_id=0)
    b = shower.run(engine, calo, 32, 200.0, SEED, first_track_id=32)
    np.testing.assert_allclose(a.edep + b.edep, whole.edep, rtol=1e-9, atol=1e-9)
    assert a.stats["secondaries"] + b.stats["secondaries"] == whole.stats["secondaries"]
    with pytest.raises(_capi.G4HB200Error):
        shower.run(engine, calo, 64, 2000.0, SEED, capacity=128)


@pytest.mark.gpu
def test_mixed_population_steps(engine):
    """configs[3] driver: reproducible, populations evolve, every step's e-/e+ and gamma pipelines ran."""
    e1, s1 = shower.run_mixed(engine, 40000, 20000, 3, SEED)
    e2, s2 = shower.run_mixed(engine, 40000, 20000, 3, SEED)
    assert s1["num_steps"] == 3
    for k in ("electron_track_steps", "gamma_track_steps", "secondaries", "peak_electrons", "peak_gammas"):
        assert s1[k] == s2[k]
    assert abs(e1 - e2) <= 1e-9 * abs(e1)
    assert s1["electron_track_steps"] > 40000 and s1["gamma_track_steps"] > 20000
    assert s1["secondaries"] > 10000 and e1 > 0.0


def test_oracle_loop_with_woodcock_tracking(reference, flat_tables):
    """Woodcock tracking of the gammas (G4HepEmWoodcockHelper): the energy bookkeeping holds, the gammas take far
    fewer steps (they no longer stop on every slab boundary) and the shower deposits its energy in the same place
    within the statistics of a few showers."""
    plain = shower.SlabCalorimeter()
    wdt = shower.SlabCalorimeter(woodcock=True)
    mat = flat_tables.couple_material()
    h0, s0 = shower_oracle.run(reference, plain, 6, 300.0, SEED, tables=flat_tables)
    h1, s1 = shower_oracle.run(reference, wdt, 6, 300.0, SEED, couple_material=mat, tables=flat_tables)
    total = h1.sum() + s1["leak_electron"] + s1["leak_gamma"]
    # kinetic energy in = deposits + leakage, up to 2 m_e c^2 for every e+ that leaves the calorimeter
    missing = (6 * 300.0 - total) / (2 * 0.51099891)
    assert missing > -1e-6 and abs(missing - round(missing)) < 1e-5 and round(missing) <= 3, missing
    assert s1["gamma_track_steps"] < 0.6 * s0["gamma_track_steps"]
    assert abs(h1[:, 0].sum() / h1.sum() - h0[:, 0].sum() / h0.sum()) < 0.08


@pytest.mark.gpu
@pytest.mark.parametrize("kind,nprim,ekin", [(_capi.SEC_ELECTRON, 10, 400.0), (_capi.SEC_GAMMA, 12, 250.0)])
def test_device_shower_with_woodcock_tracking_matches_cpu_loop(engine, reference, flat_tables, kind, nprim, ekin):
    calo = shower.SlabCalorimeter(woodcock=True)
    want, wst = shower_oracle.run(reference, calo, nprim, ekin, SEED, kind=kind, first_track_id=300,
                                  couple_material=flat_tables.couple_material(), tables=flat_tables)
    got = shower.run(engine, calo, nprim, ekin, SEED, kind=kind, first_track_id=300, capacity=1 << 16)
    for k in ("num_steps", "electron_track_steps", "gamma_track_steps", "secondaries", "peak_electrons", "peak_gammas"):
        assert got.stats[k] == wst[k], (k, got.stats[k], wst[k])
    np.testing.assert_allclose(got.edep, want, rtol=1e-9, atol=1e-9)
    assert abs(got.stats["leak_electron"] - wst["leak_electron"]) <= 1e-9 * max(1.0, wst["leak_electron"])
    assert abs(got.stats["leak_gamma"] - wst["leak_gamma"]) <= 1e-9 * max(1.0, wst["leak_gamma"])


@pytest.mark.gpu
@pytest.mark.parametrize("woodcock", [False, True])
def test_device_showers_at_scale_keep_the_energy_balance(engine, woodcock):
    """512 x 1 GeV showers (populations of ~1e5 tracks: the half-batch pipelines, the two-stream loop): kinetic energy
    in = deposits + leakage, up to 2 m_e c^2 per e+ that leaves the calorimeter."""
    calo = shower.SlabCalorimeter(woodcock=woodcock)
    nprim, ekin = 512, 1000.0
    res = shower.run(engine, calo, nprim, ekin, SEED, capacity=1 << 21)
    total = res.edep.sum() + res.stats["leak_electron"] + res.stats["leak_gamma"]
    missing = (nprim * ekin - total) / (2 * 0.51099891)
    assert missing > -1e-3 and abs(missing - round(missing)) < 1e-3 and round(missing) < 200, missing
    assert res.stats["peak_electrons"] > 20000
    # lead takes ~78 % of the deposit, liquid argon ~21 % (sampling fraction of the ATLASbar stack)
    frac = res.edep[:, 0].sum() / res.edep.sum()
    assert 0.7 < frac < 0.85, frac


@pytest.mark.gpu
@pytest.mark.parametrize("kind,nprim,ekin", [(_capi.SEC_ELECTRON, 10, 400.0), (_capi.SEC_GAMMA, 16, 250.0)])
def test_device_woodcock_passes_equal_the_uncut_reference_loop(engine, reference, flat_tables, kind, nprim, ekin):
    """The device loop ends a Woodcock pass after 4 virtual steps at a fictitious interaction point and carries on in the next
    iteration with the next uniform of the track's stream.  The oracle here does NOT: it runs the reference's KeepTracking loop
    (G4HepEmWoodcockHelper.cc:150-300) until the gamma interacts or reaches the surface.  The showers must be the same:
    deposits per cell, leakage, the secondaries created and the e-/e+ population (the number of gamma passes differs)."""
    calo = shower.SlabCalorimeter(woodcock=True)
    want, wst = shower_oracle.run(reference, calo, nprim, ekin, SEED, kind=kind, first_track_id=700,
                                  couple_material=flat_tables.couple_material(), tables=flat_tables, wdt_max_virtual_steps=None)
    got = shower.run(engine, calo, nprim, ekin, SEED, kind=kind, first_track_id=700, capacity=1 << 16)
    for k in ("electron_track_steps", "secondaries"):
        assert got.stats[k] == wst[k], (k, got.stats[k], wst[k])
    assert got.stats["gamma_track_steps"] >= wst["gamma_track_steps"]  # a cut pass counts as a track-step of its own
    np.testing.assert_allclose(got.edep, want, rtol=1e-9, atol=1e-9)
    assert abs(got.stats["leak_electron"] - wst["leak_electron"]) <= 1e-9 * max(1.0, wst["leak_electron"])
    assert abs(got.stats["leak_gamma"] - wst["leak_gamma"]) <= 1e-9 * max(1.0, wst["leak_gamma"])


@pytest.mark.gpu
def test_mixed_run_matches_cpu_driver_step_by_step(engine, reference, flat_tables):
    """BASELINE configs[3] (g4hb200_mixed_run: mixed e-/e+/gamma population, consecutive fused steps, secondaries fed back)
    against a CPU loop around the reference's managers (tests/shower_oracle.py: run_mixed -- the same population generated
    operation by operation, the same child streams): the populations, the number of secondaries and the deposited energy
    after every step."""
    n_el, n_gm = 60000, 30000
    for steps in (1, 2, 3, 4):
        want_e, want = shower_oracle.run_mixed(reference, n_el, n_gm, steps, SEED, flat_tables.num_matcut)
        got_e, got = shower.run_mixed(engine, n_el, n_gm, steps, SEED)
        for k in ("num_steps", "electron_track_steps", "gamma_track_steps", "secondaries", "peak_electrons", "peak_gammas"):
            assert got[k] == want[k], (steps, k, got[k], want[k])
        assert abs(got_e - want_e) <= 1e-9 * want_e, (steps, got_e, want_e)


@pytest.mark.gpu
def test_loop_stopped_on_max_steps_reports_what_is_left(engine):
    """A loop that stops on max_steps says so: the populations and the kinetic energy still alive are in the stats, and the
    energy balance closes with them (up to 2 m_e c^2 per e+ created so far)."""
    calo = shower.SlabCalorimeter()
    res = shower.run(engine, calo, 16, 500.0, SEED, max_steps=12, capacity=1 << 16)
    st = res.stats
    assert st["num_steps"] == 12
    assert st["remaining_electrons"] > 0 and st["remaining_gammas"] > 0 and st["remaining_ekin"] > 0.0
    total = res.edep.sum() + st["leak_electron"] + st["leak_gamma"] + st["remaining_ekin"]
    assert abs(total - 16 * 500.0) < 2 * 0.51099891 * st["secondaries"] + 1e-6
    done = shower.run(engine, calo, 16, 500.0, SEED, capacity=1 << 16)
    assert done.stats["remaining_electrons"] == 0 and done.stats["remaining_gammas"] == 0 and done.stats["remaining_ekin"] == 0.0
