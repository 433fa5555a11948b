"""CPU restatement of the slab-calorimeter stepping loop (TEST INFRASTRUCTURE): the reference's managers
(oracle/_ref through the batch shim) drive the physics, numpy does what the device kernels of
g4hepem_b200/csrc/g4h_shower.cuh do around them -- geometry step, MSC displacement, scoring, relocation,
secondaries -> tracks -- with the same floating point operations in the same order, so that a shower is the same
track by track.  Follows G4HepEmTrackingManager::TrackElectron / TrackGamma
(G4HepEm/G4HepEm/src/G4HepEmTrackingManager.cc:428-705,985-1140) and TestEm3's geometry / scoring
(apps/examples/TestEm3/src/DetectorConstruction.cc:281-384, SteppingAction.cc:83-84)."""
import numpy as np

from g4hepem_b200 import _capi, batches

M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c, k):
    """c: (n,4) uint32 counters, k: (2,) uint32 key -> (n,4) uint32 (Random123 philox4x32-10)."""
    x = c.astype(np.uint64)
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * x[:, 0]
        p1 = np.uint64(0xCD9E8D57) * x[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & M32
        x = np.stack([(hi1 ^ x[:, 1] ^ k0) & M32, lo1, (hi0 ^ x[:, 3] ^ k1) & M32, lo0], axis=1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return x.astype(np.uint32)


def child_stream(seed, parent_id, parent_draw, slot):
    """(id, first draw) of a secondary: ChildStream of g4h_shower.cuh."""
    n = len(parent_id)
    c = np.zeros((n, 4), dtype=np.uint32)
    c[:, 0] = parent_draw.astype(np.uint32)
    c[:, 2] = parent_id.astype(np.uint32)
    key = np.array([(seed & 0xFFFFFFFF) ^ 0x5EC0DA2A, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    r = philox4x32_10(c, key)
    a = np.where(slot == 0, r[:, 0], r[:, 2])
    b = np.where(slot == 0, r[:, 1], r[:, 3])
    return a.astype(np.uint32).view(np.int32), (b & np.uint32(0x3FFFFFFE)).astype(np.int32)


def uniform_at(seed, ids, draw):
    """draw `draw[k]` of the counter based stream of track ids[k] (g4h_rng.cuh / oracle/g4h_rng_host.h)."""
    n = len(ids)
    c = np.zeros((n, 4), dtype=np.uint32)
    c[:, 0] = (draw.astype(np.int64) >> 1).astype(np.uint32)
    c[:, 2] = ids.astype(np.uint32)
    r = philox4x32_10(c, np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)).astype(np.uint64)
    odd = (draw.astype(np.int64) & 1) == 1
    lo = np.where(odd, r[:, 2], r[:, 0])
    hi = np.where(odd, r[:, 3], r[:, 1])
    bits = ((hi << np.uint64(32)) | lo) >> np.uint64(12)
    d = (bits | np.uint64(0x3FF0000000000000)).view(np.float64)
    return d - 0.99999999999999988897769753748


DBL_MAX = np.finfo(np.float64).max
WDT_MAX_VIRTUAL_STEPS = 4  # kWdtMaxVirtualSteps of g4h_shower.cuh


class Slab:
    def __init__(self, calo):
        self.nl = calo.num_layers
        self.na = len(calo.absorber_thickness)
        self.front = np.concatenate([[0.0], np.cumsum(np.asarray(calo.absorber_thickness, dtype=np.float64))])
        self.couple = np.asarray(calo.absorber_couple, dtype=np.int32)
        self.half = float(calo.half_yz)
        self.xfront = -0.5 * (self.nl * self.front[self.na])

    def bounds(self, vol):
        layer = vol // self.na
        iabs = vol - layer * self.na
        lf = self.xfront + layer * self.front[self.na]
        return lf + self.front[iabs], lf + self.front[iabs + 1]

    @staticmethod
    def along(p, d, lo, hi):
        with np.errstate(divide="ignore", invalid="ignore"):
            pos = np.maximum(0.0, (hi - p) / d)
            neg = np.maximum(0.0, (lo - p) / d)
        return np.where(d > 0, pos, np.where(d < 0, neg, 1.0e30))

    def distance(self, vol, pos, dirs):
        xlo, xhi = self.bounds(vol)
        dx = self.along(pos[:, 0], dirs[:, 0], xlo, xhi)
        dy = self.along(pos[:, 1], dirs[:, 1], -self.half, self.half)
        dz = self.along(pos[:, 2], dirs[:, 2], -self.half, self.half)
        nv = np.where(dirs[:, 0] > 0, vol + 1, vol - 1)
        nv = np.where(nv >= self.nl * self.na, -1, nv)
        d = dx.copy()
        m = dy < d
        d = np.where(m, dy, d)
        nv = np.where(m, -1, nv)
        m = dz < d
        d = np.where(m, dz, d)
        nv = np.where(m, -1, nv)
        return d, nv.astype(np.int32)

    def distance_out(self, pos, dirs):
        """along dirs to the surface of the whole calorimeter (DistanceToCalorimeterOut)"""
        dx = self.along(pos[:, 0], dirs[:, 0], self.xfront, -self.xfront)
        dy = self.along(pos[:, 1], dirs[:, 1], -self.half, self.half)
        dz = self.along(pos[:, 2], dirs[:, 2], -self.half, self.half)
        return np.minimum(dx, np.minimum(dy, dz))

    def locate(self, x):
        """LocateSlab of g4h_shower.cuh"""
        layer_t = self.front[self.na]
        t = x - self.xfront
        layer = np.clip((t / layer_t).astype(np.int64), 0, self.nl - 1)
        u = t - layer * layer_t
        iabs = np.full(len(x), self.na - 1, dtype=np.int64)
        for k in range(self.na - 1, 0, -1):
            iabs = np.where(u < self.front[k], k - 1, iabs)
        return (layer * self.na + iabs).astype(np.int32)

    def safety(self, vol, pos):
        xlo, xhi = self.bounds(vol)
        s = np.minimum(pos[:, 0] - xlo, xhi - pos[:, 0])
        s = np.minimum(s, self.half - np.abs(pos[:, 1]))
        s = np.minimum(s, self.half - np.abs(pos[:, 2]))
        return np.maximum(0.0, s)


def _new_electrons(n):
    b = batches.ElectronHostBatch(n)
    b.ekin_logekin[:, 1] = 100.0
    b.nia01[...] = -1.0
    b.nia23[...] = -1.0
    b.msc_irange_dynrf[:, 0] = 1.0e21
    b.msc_irange_dynrf[:, 1] = 0.04
    b.msc_tlimmin_gauss[:, 0] = 1.0e-7
    b.winner[...] = -1
    return b


def _new_gammas(n):
    b = batches.GammaHostBatch(n)
    b.ekin_logekin[:, 1] = 100.0
    b.dirz_nia0[:, 1] = -1.0
    b.winner[...] = -1
    return b


def _take(b, idx, cls):
    o = cls(len(idx))
    for g in b.groups() + ("meta", "winner"):
        getattr(o, g)[...] = getattr(b, g)[idx]
    return o


def _concat(parts, cls):
    n = sum(p.n for p in parts)
    o = cls(n)
    for g in o.groups() + ("meta", "winner"):
        if parts:
            getattr(o, g)[...] = np.concatenate([getattr(p, g) for p in parts], axis=0)
    return o


def _total_mxsec(reference, seed, imc, ekin, lekin, pe_prev):
    """G4HepEmGammaManager::GetTotalMacXSec for (couple, energy): (total macroscopic cross section, fPEmxSec after the
    call -- the previous value where the function does not set it), both from the reference."""
    mx, _ = reference.gamma_lookups(imc.astype(np.int32), ekin, lekin, np.zeros(len(imc)))
    scratch = _new_gammas(len(imc))
    scratch.ekin_logekin[:, 0] = ekin
    scratch.ekin_logekin[:, 1] = lekin
    scratch.dirz_nia0[:, 1] = 1.0  # no resampling, no draw
    scratch.meta[:, 0] = imc
    scratch.edep_pemxsec[:, 1] = pe_prev
    reference.gamma_howfar(scratch, seed, 1)
    return mx, scratch.edep_pemxsec[:, 1].copy()


def _woodcock(reference, slab, calo, couple_material, gm, gm_pos, dirs, seed, max_virtual=WDT_MAX_VIRTUAL_STEPS):
    """G4HepEmWoodcockHelper::KeepTracking (G4HepEmWoodcockHelper.cc:150-300) with the decisions around it in
    G4HepEmTrackingManager::TrackGamma (G4HepEmTrackingManager.cc:955-1032), for the gammas it applies to; the same
    operations in the same order as SlabGammaGeometryStep of g4h_shower.cuh.
    max_virtual: virtual steps per pass (the device loop ends a pass after kWdtMaxVirtualSteps at a fictitious interaction
    point); None: the reference's own, uncut loop.
    Returns (wdt mask, physical step, positions, volumes of the wdt tracks after the Woodcock part, cut mask)."""
    n = gm.n
    ekin = gm.ekin_logekin[:, 0]
    lim = calo.woodcock_ekin_min
    dout = np.maximum(slab.distance_out(gm_pos, dirs) - 1.0e-3, 0.0)
    on = (gm.meta[:, 1] & _capi.F_WDT_ON) != 0
    on = np.where(on, True, ~(ekin < lim) & ~(dout < 1.0e-6))
    on = on & (ekin > lim)
    idx = np.flatnonzero(on)
    phys = np.zeros(n)
    pos = gm_pos.copy()
    vol = np.zeros(n, dtype=np.int32)
    gm.meta[:, 1] = np.where(on, gm.meta[:, 1], gm.meta[:, 1] & ~_capi.F_WDT_ON)
    if len(idx) == 0:
        return on, phys, pos, vol, np.zeros(n, dtype=bool)
    m = len(idx)
    ek = ekin[idx]
    lek = gm.ekin_logekin[idx, 1]
    lek = np.where(lek > 99.0, reference.vdt_log_exp(ek)[0], lek)
    gm.ekin_logekin[idx, 1] = lek
    ids = gm.meta[idx, 2]
    draw = gm.meta[idx, 3].astype(np.int64)
    wimc = np.full(m, calo.woodcock_couple, dtype=np.int32)
    wmat = couple_material[calo.woodcock_couple]
    wmx, wpe = _total_mxsec(reference, seed, wimc, ek, lek, gm.edep_pemxsec[idx, 1])
    with np.errstate(divide="ignore"):
        wmfp = np.where(wmx > 0.0, 1.0 / wmx, DBL_MAX)
    pe = wpe.copy()
    dist = dout[idx].copy()
    steplen = np.zeros(m)
    reached = np.zeros(m, dtype=bool)
    stop = np.zeros(m, dtype=bool)
    mx = np.zeros(m)
    prev = np.full(m, -1, dtype=np.int32)
    mfp0 = np.full(m, -1.0)
    d0 = dirs[idx, 0]
    x0 = gm_pos[idx, 0]
    passes = 0
    while not stop.all() and (max_virtual is None or passes < max_virtual):
        passes += 1
        a = np.flatnonzero(~stop)
        need = wmfp[a] < DBL_MAX
        u = np.ones(len(a))
        if need.any():
            u[need] = uniform_at(seed, ids[a][need], draw[a][need])
            draw[a[need]] += 1
        pstep = np.where(need, -reference.vdt_log_exp(u)[0] * wmfp[a], DBL_MAX)
        hit = dist[a] < pstep
        steplen[a] += np.where(hit, dist[a], pstep)
        reached[a[hit]] = True
        stop[a[hit]] = True
        b = a[~hit]
        if len(b) == 0:
            continue
        dist[b] -= pstep[~hit]
        pvol = slab.locate(x0[b] + steplen[b] * d0[b])
        pimc = slab.couple[pvol % slab.na]
        same = couple_material[pimc] == wmat
        bs = b[same]
        stop[bs] = True
        mfp0[bs] = wmfp[bs]
        pe[bs] = wpe[bs]
        bd = b[~same]
        if len(bd) > 0:
            pimc_d = pimc[~same]
            changed = pimc_d != prev[bd]
            if changed.any():
                c = bd[changed]
                prev[c] = pimc_d[changed]
                mx[c], pe[c] = _total_mxsec(reference, seed, pimc_d[changed], ek[c], lek[c], pe[c])
            u2 = uniform_at(seed, ids[bd], draw[bd])
            draw[bd] += 1
            st = mx[bd] * wmfp[bd] > u2
            s_idx = bd[st]
            stop[s_idx] = True
            with np.errstate(divide="ignore"):
                mfp0[s_idx] = np.where(mx[s_idx] > 0.0, 1.0 / mx[s_idx], DBL_MAX)
    pos[idx] = gm_pos[idx] + steplen[:, None] * dirs[idx]
    vol[idx] = slab.locate(pos[idx, 0])
    gm.meta[idx, 0] = slab.couple[vol[idx] % slab.na]
    gm.meta[idx, 3] = draw.astype(np.int32)
    gm.dirz_nia0[idx, 1] = -1.0
    gm.gstep_mfp0[idx, 1] = mfp0
    gm.edep_pemxsec[idx, 1] = pe
    phys[idx] = np.where(reached, 10.0, 0.0)
    gm.meta[idx, 1] = np.where(reached, gm.meta[idx, 1] & ~_capi.F_WDT_ON, gm.meta[idx, 1] | _capi.F_WDT_ON)
    cut = np.zeros(n, dtype=bool)
    cut[idx] = ~stop  # the pass was cut after WDT_MAX_VIRTUAL_STEPS virtual steps: a fictitious interaction point
    return on, phys, pos, vol, cut


def _put(b, idx, sub):
    for g in b.groups() + ("meta", "winner"):
        getattr(b, g)[idx] = getattr(sub, g)


def _op(reference, op, b, mask, seed, sec=None, flags=None):
    """One track-level static of the reference (oracle/ref.py: electron_track_op) on the tracks of `b` selected by mask.
    sec: secondaries are appended with parent indices of the full batch; returns the returned bools (full size) if flags."""
    idx = np.flatnonzero(mask)
    out = np.zeros(b.n, dtype=np.int32)
    if len(idx) == 0:
        return out
    sub = _take(b, idx, batches.ElectronHostBatch)
    fl = np.zeros(len(idx), dtype=np.int32) if flags else None
    q = batches.SecondaryHostQueue(2 * len(idx)) if sec is not None else None
    reference.electron_track_op(op, sub, seed, q, fl)
    _put(b, idx, sub)
    if fl is not None:
        out[idx] = fl
    if q is not None:
        n = int(q.count[0])
        k = int(sec.count[0])
        sec.dirx_diry[k:k + n] = q.dirx_diry[:n]
        sec.dirz_ekin[k:k + n] = q.dirz_ekin[:n]
        sec.parent_kind[k:k + n] = q.parent_kind[:n]
        sec.parent_slot[k:k + n, 0] = idx[q.parent_slot[:n, 0]]
        sec.parent_slot[k:k + n, 1] = q.parent_slot[:n, 1]
        sec.count[0] = k + n
    return out


F_MSC_SUBSTEP = 0x100


def _electron_round(reference, slab, el, el_pos, el_vol, sub, multi, seed):
    """One round of G4HepEmTrackingManager::TrackElectron (G4HepEmTrackingManager.cc:408-665) for every e-/e+ of the batch,
    piece by piece through the reference's statics: a whole step, or -- where fIsMultipleStepsInMSCTrans of the region is set and
    MSC limited the step -- one MSC sub-step of it (the track then carries F_MSC_SUBSTEP and the caller's local variables in `sub`
    into the next round, like the device loop).  Returns (post-step positions, on-boundary mask, next volumes, secondaries)."""
    n = el.n
    everyone = np.ones(n, dtype=bool)
    resume = (el.meta[:, 1] & F_MSC_SUBSTEP) != 0
    fresh = ~resume
    _op(reference, _capi.OP_RESAMPLE_NIA, el, fresh, seed)
    _op(reference, _capi.OP_HOWFAR_DISCRETE, el, fresh, seed)
    sub["proc"] = np.where(fresh, el.winner, sub["proc"]).astype(np.int32)
    sub["left"] = np.where(fresh, el.gstep_pstep[:, 1], sub["left"])
    sub["eloss"] = np.where(fresh, 0.0, sub["eloss"])
    sub["pre_e"] = np.where(fresh, el.ekin_logekin[:, 0], sub["pre_e"])
    sub["pre_le"] = np.where(fresh, el.ekin_logekin[:, 1], sub["pre_le"])
    # a resumed step: the rest of the step limit, the reduced range, the winner MSC replaced (.cc:574-596)
    el.gstep_pstep[resume, 0] = sub["left"][resume]
    el.gstep_pstep[resume, 1] = sub["left"][resume]
    el.range_lambtr1[resume, 0] = sub["range"][resume]
    el.winner[resume] = sub["proc"][resume]
    _op(reference, _capi.OP_HOWFAR_MSC, el, everyone, seed)
    cont = multi[el.meta[:, 0]] & (el.winner == -2)
    # geometry
    dirs = np.stack([el.dirx_diry[:, 0], el.dirx_diry[:, 1], el.dirz_safety[:, 0]], axis=1)
    dist, nv = slab.distance(el_vol, el_pos, dirs)
    onb = dist < el.gstep_pstep[:, 0]
    step = np.where(onb, dist, el.gstep_pstep[:, 0])
    el_pos = el_pos + step[:, None] * dirs
    el.gstep_pstep[:, 0] = step
    el.meta[:, 1] = np.where(onb, el.meta[:, 1] | _capi.F_ON_BOUNDARY, el.meta[:, 1] & ~_capi.F_ON_BOUNDARY)
    cont &= ~onb
    # along the step
    moving = step > 0.0
    el.edep_dispx[:, 0] = 0.0
    el.gstep_pstep[~moving, 1] = step[~moving]  # a step of zero length: G4HepEmElectronManager::Perform (.icc:461-470)
    # SavePreStepEKin of this round (.cc:445,581): the logarithm is cached by HowFarToMSC
    el.prestep[:, 0] = el.ekin_logekin[:, 0]
    el.prestep[:, 1] = el.ekin_logekin[:, 1]
    _op(reference, _capi.OP_UPDATE_PSTEP, el, moving, seed)
    going = moving & (el.gstep_pstep[:, 1] > 0.0)
    _op(reference, _capi.OP_UPDATE_NIA, el, going, seed)
    stopped = _op(reference, _capi.OP_MEAN_ELOSS, el, going, seed, flags=True) != 0
    sub["eloss"] = np.where(going, sub["eloss"] + el.edep_dispx[:, 0], sub["eloss"])
    cont &= going & ~stopped
    alive = going & ~stopped
    _op(reference, _capi.OP_SAMPLE_MSC, el, alive, seed)
    # between two sub-steps (.cc:574-596)
    pstep = el.gstep_pstep[:, 1]
    sub["left"] = np.where(cont, sub["left"] - pstep, sub["left"])
    sub["range"] = np.where(cont, el.range_lambtr1[:, 0] - pstep, sub["range"])
    el.edep_dispx[cont, 0] = 0.0
    el.winner[cont] = sub["proc"][cont]
    el.meta[:, 1] = np.where(cont, el.meta[:, 1] | F_MSC_SUBSTEP, el.meta[:, 1] & ~F_MSC_SUBSTEP)
    # the end of the step (.cc:599-665)
    done = moving & ~cont
    el.edep_dispx[done, 0] = sub["eloss"][done]
    fluct = alive & ~cont
    el.prestep[fluct, 0] = sub["pre_e"][fluct]
    el.prestep[fluct, 1] = sub["pre_le"][fluct]
    stopped2 = _op(reference, _capi.OP_LOSS_FLUCT, el, fluct, seed, flags=True) != 0
    sec = batches.SecondaryHostQueue(2 * n)
    is_pos = (el.meta[:, 1] & _capi.F_POSITRON) != 0
    _op(reference, _capi.OP_ANNIHILATE_AT_REST, el, is_pos & ((going & stopped) | (fluct & stopped2)), seed, sec=sec)
    # PerformDiscrete returns at once without a winner or on a boundary; a step whose true length came out zero goes there too
    discrete = (fluct & ~stopped2) | (moving & ~going)
    _op(reference, _capi.OP_DISCRETE, el, discrete, seed, sec=sec)
    return el_pos, onb, nv, sec


ELECTRON_MASS_C2 = 5.1099890999999997e-01


def run(reference, calo, num_primaries, primary_ekin, seed, kind=_capi.SEC_ELECTRON, first_track_id=0, max_steps=0, threads=4,
        couple_material=None, tables=None, wdt_max_virtual_steps=WDT_MAX_VIRTUAL_STEPS, mixed=None):
    """Returns (edep[num_layers, num_absorbers], stats) like g4hepem_b200.shower.run.
    couple_material: material index of every couple (needed with calo.woodcock).
    tables: the FlatTables of the run: secondary production cuts of the caller (StackSecondaries(..., isApplyCuts),
    G4HepEmTrackingManager.cc:1254-1326) are applied where the region of the parent's couple asks for them; None: no cuts.
    wdt_max_virtual_steps: None = the reference's uncut Woodcock loop (KeepTracking runs until the gamma interacts or reaches
    the surface); the default mirrors the device loop's passes of at most kWdtMaxVirtualSteps virtual steps, which makes the
    two loops agree iteration by iteration.
    mixed: (e-/e+ batch, gamma batch) = the loop of g4hb200_mixed_run (BASELINE configs[3]) instead: that population, no geometry
    (fused steps: the proposed step is accepted, nothing moves but by the MSC displacement), one scoring cell, a secondary
    inherits its parent's couple, no production cuts; calo is then the one-cell geometry of capi_shower.inl."""
    slab = Slab(calo)
    multi = None
    if tables is not None:
        multi = ((tables.region_pars()[:, 7].astype(np.int64) & 1) != 0)[tables.couple_region()]
    if tables is not None:
        couple_cuts = tables.couple_cuts()
        apply_cuts = (tables.region_pars()[:, 7].astype(np.int64) & 2) != 0
        apply_cuts = apply_cuts[tables.couple_region()]
    hist = np.zeros(slab.nl * slab.na)
    stats = dict(num_steps=0, electron_track_steps=0, gamma_track_steps=0, secondaries=0, peak_electrons=0, peak_gammas=0,
                 leak_electron=0.0, leak_gamma=0.0)
    ids = first_track_id + np.arange(num_primaries, dtype=np.int32)
    if mixed is not None:
        el, gm = mixed
        tables = None
    elif kind == _capi.SEC_GAMMA:
        el, gm = _new_electrons(0), _new_gammas(num_primaries)
        gm.ekin_logekin[:, 0] = primary_ekin
        gm.dirx_diry[:, 0] = 1.0
        gm.meta[:, 0] = slab.couple[0]
        gm.meta[:, 1] = _capi.F_ON_BOUNDARY
        gm.meta[:, 2] = ids
    else:
        el, gm = _new_electrons(num_primaries), _new_gammas(0)
        el.ekin_logekin[:, 0] = primary_ekin
        el.dirx_diry[:, 0] = 1.0
        el.meta[:, 0] = slab.couple[0]
        el.meta[:, 1] = _capi.F_MSC_FIRST_STEP | _capi.F_ON_BOUNDARY | (_capi.F_POSITRON if kind == _capi.SEC_POSITRON else 0)
        el.meta[:, 2] = ids
    el_pos = np.zeros((el.n, 3)); el_pos[:, 0] = 0.0 if mixed is not None else slab.xfront
    gm_pos = np.zeros((gm.n, 3)); gm_pos[:, 0] = 0.0 if mixed is not None else slab.xfront
    el_vol = np.zeros(el.n, dtype=np.int32)
    gm_vol = np.zeros(gm.n, dtype=np.int32)

    def new_sub(k):
        return dict(left=np.zeros(k), eloss=np.zeros(k), pre_e=np.zeros(k), pre_le=np.zeros(k), range=np.zeros(k),
                    proc=np.full(k, -1, dtype=np.int32))

    el_sub = new_sub(el.n)

    def children(sec, parent_meta, parent_pos, parent_vol):
        r = sec  # raw queue order is irrelevant: everything is derived per record
        n = int(sec.count[0])
        if n == 0:
            return (_new_electrons(0), np.zeros((0, 3)), np.zeros(0, dtype=np.int32)), (_new_gammas(0), np.zeros((0, 3)), np.zeros(0, dtype=np.int32))
        p = sec.parent_slot[:n, 0]
        slot = sec.parent_slot[:n, 1]
        knd = sec.parent_kind[:n, 1]
        cid, cdraw = child_stream(seed, sec.parent_kind[:n, 0], parent_meta[p, 3], slot)
        pos = parent_pos[p]
        vol = parent_vol[p]
        imc = parent_meta[p, 0] if mixed is not None else slab.couple[vol % slab.na]
        keep = np.ones(n, dtype=bool)
        if tables is not None:
            pimc = parent_meta[p, 0]
            ek = sec.dirz_ekin[:n, 1]
            cuts = couple_cuts[pimc]
            on = apply_cuts[pimc]
            drop_e = on & (knd == _capi.SEC_ELECTRON) & (ek < cuts[:, 0])
            drop_p = on & (knd == _capi.SEC_POSITRON) & (ELECTRON_MASS_C2 < cuts[:, 2]) & (ek < cuts[:, 1])
            drop_g = on & (knd == _capi.SEC_GAMMA) & (ek < cuts[:, 2])
            cut_edep = np.where(drop_e | drop_g, ek, np.where(drop_p, ek + 2 * ELECTRON_MASS_C2, 0.0))
            drop = drop_e | drop_p | drop_g
            np.add.at(hist, vol[drop], cut_edep[drop])
            keep = ~drop
        ie = np.flatnonzero((knd != _capi.SEC_GAMMA) & keep)
        ig = np.flatnonzero((knd == _capi.SEC_GAMMA) & keep)
        ne = _new_electrons(len(ie))
        ne.ekin_logekin[:, 0] = sec.dirz_ekin[:n, 1][ie]
        ne.dirx_diry[...] = sec.dirx_diry[:n][ie]
        ne.dirz_safety[:, 0] = sec.dirz_ekin[:n, 0][ie]
        ne.dirz_safety[:, 1] = slab.safety(vol[ie], pos[ie])
        ne.meta[:, 0] = imc[ie]
        ne.meta[:, 1] = _capi.F_MSC_FIRST_STEP | np.where(knd[ie] == _capi.SEC_POSITRON, _capi.F_POSITRON, 0)
        ne.meta[:, 2] = cid[ie]
        ne.meta[:, 3] = cdraw[ie]
        ng = _new_gammas(len(ig))
        ng.ekin_logekin[:, 0] = sec.dirz_ekin[:n, 1][ig]
        ng.dirx_diry[...] = sec.dirx_diry[:n][ig]
        ng.dirz_nia0[:, 0] = sec.dirz_ekin[:n, 0][ig]
        ng.meta[:, 0] = imc[ig]
        ng.meta[:, 2] = cid[ig]
        ng.meta[:, 3] = cdraw[ig]
        return (ne, pos[ie], vol[ie]), (ng, pos[ig], vol[ig])

    while el.n > 0 or gm.n > 0:
        if max_steps and stats["num_steps"] >= max_steps:
            break
        stats["num_steps"] += 1
        stats["electron_track_steps"] += el.n
        stats["gamma_track_steps"] += gm.n
        stats["peak_electrons"] = max(stats["peak_electrons"], el.n)
        stats["peak_gammas"] = max(stats["peak_gammas"], gm.n)
        next_el, next_gm, next_sub = [], [], []
        # ---- e-/e+ ------------------------------------------------------------------------------------------------
        if el.n > 0 and mixed is not None:
            sec = batches.SecondaryHostQueue(2 * el.n)
            reference.electron_step(el, sec, seed, threads)
            onb = np.zeros(el.n, dtype=bool)
            nv = el_vol.copy()
        elif el.n > 0 and multi is not None:
            el_pos, onb, nv, sec = _electron_round(reference, slab, el, el_pos, el_vol, el_sub, multi, seed)
        elif el.n > 0:
            reference.electron_howfar(el, seed, threads)
            dirs = np.stack([el.dirx_diry[:, 0], el.dirx_diry[:, 1], el.dirz_safety[:, 0]], axis=1)
            dist, nv = slab.distance(el_vol, el_pos, dirs)
            onb = dist < el.gstep_pstep[:, 0]
            step = np.where(onb, dist, el.gstep_pstep[:, 0])
            el_pos = el_pos + step[:, None] * dirs
            el.gstep_pstep[:, 0] = step
            el.meta[:, 1] = np.where(onb, el.meta[:, 1] | _capi.F_ON_BOUNDARY, el.meta[:, 1] & ~_capi.F_ON_BOUNDARY)
            sec = batches.SecondaryHostQueue(2 * el.n)
            reference.electron_perform(el, sec, seed, threads)
        if el.n > 0:
            # MSC displacement
            disp = np.stack([el.edep_dispx[:, 1], el.dispy_dispz[:, 0], el.dispy_dispz[:, 1]], axis=1)
            d2 = disp[:, 0] * disp[:, 0] + disp[:, 1] * disp[:, 1] + disp[:, 2] * disp[:, 2]
            kmin = 5.0e-8
            cand = (~onb) & (d2 > kmin * kmin)
            dr = np.sqrt(d2)
            ps = 0.99 * slab.safety(el_vol, el_pos)
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = ps / dr
            scale = np.where((ps > 0.0) & (dr <= ps), 1.0, np.where(dr < ps, 1.0, np.where(ps > kmin, ratio, 0.0)))
            scale = np.where(cand, scale, 0.0)
            moved = scale > 0.0
            el_pos = np.where(moved[:, None], el_pos + disp * scale[:, None], el_pos)
            np.add.at(hist, el_vol, el.edep_dispx[:, 0])
            new_vol = np.where(onb, nv, el_vol)
            ekin = el.ekin_logekin[:, 0]
            stats["leak_electron"] += float(ekin[(ekin > 0) & (new_vol < 0)].sum())
            (ce, cepos, cevol), (cg, cgpos, cgvol) = children(sec, el.meta, el_pos, el_vol)
            stats["secondaries"] += int(sec.count[0])
            alive = np.flatnonzero((ekin > 0) & (new_vol >= 0))
            surv = _take(el, alive, batches.ElectronHostBatch)
            spos, svol = el_pos[alive], new_vol[alive].astype(np.int32)
            surv.meta[:, 0] = np.where(onb[alive], slab.couple[svol % slab.na], surv.meta[:, 0])
            in_step = (surv.meta[:, 1] & F_MSC_SUBSTEP) != 0  # the caller sets the safety once per step (.cc:419-423)
            surv.dirz_safety[:, 1] = np.where(in_step, surv.dirz_safety[:, 1], np.where(onb[alive], 0.0, slab.safety(svol, spos)))
            surv.edep_dispx[...] = 0.0
            surv.winner[...] = -1
            next_el += [(surv, spos, svol), (ce, cepos, cevol)]
            next_gm += [(cg, cgpos, cgvol)]
            next_sub += [{k: v[alive] for k, v in el_sub.items()}, new_sub(ce.n)]
        # ---- gamma ------------------------------------------------------------------------------------------------------
        if gm.n > 0 and mixed is not None:
            sec = batches.SecondaryHostQueue(2 * gm.n)
            reference.gamma_step(gm, sec, seed, threads)
            onb = np.zeros(gm.n, dtype=bool)
            nv = gm_vol.copy()
        elif gm.n > 0:
            dirs = np.stack([gm.dirx_diry[:, 0], gm.dirx_diry[:, 1], gm.dirz_nia0[:, 0]], axis=1)
            if getattr(calo, "woodcock", False):
                wdt, phys, gm_pos, wvol, cut = _woodcock(reference, slab, calo, np.asarray(couple_material), gm, gm_pos, dirs, seed,
                                                         wdt_max_virtual_steps)
                gm_vol = np.where(wdt, wvol, gm_vol).astype(np.int32)
                normal = np.flatnonzero(~wdt)
                if len(normal) > 0:
                    sub = _take(gm, normal, batches.GammaHostBatch)
                    reference.gamma_howfar(sub, seed, threads)
                    for g_ in sub.groups() + ("meta", "winner"):
                        getattr(gm, g_)[normal] = getattr(sub, g_)
                    phys[normal] = sub.gstep_mfp0[:, 0]
            else:
                wdt = np.zeros(gm.n, dtype=bool)
                cut = np.zeros(gm.n, dtype=bool)
                reference.gamma_howfar(gm, seed, threads)
                phys = gm.gstep_mfp0[:, 0].copy()
            dist, nv = slab.distance(gm_vol, gm_pos, dirs)
            onb = dist < phys
            step = np.where(onb, dist, phys)
            gm_pos = gm_pos + step[:, None] * dirs
            gm.gstep_mfp0[:, 0] = np.where(wdt, 0.0, step)
            gm.meta[:, 1] = np.where(onb, gm.meta[:, 1] | _capi.F_ON_BOUNDARY, gm.meta[:, 1] & ~_capi.F_ON_BOUNDARY)
            sec = batches.SecondaryHostQueue(2 * gm.n)
            if cut.any():
                # a Woodcock pass that was cut: the track waits at a fictitious interaction point, nothing happens
                go = np.flatnonzero(~cut)
                sub = _take(gm, go, batches.GammaHostBatch)
                reference.gamma_perform(sub, sec, seed, threads)
                for g_ in sub.groups() + ("meta", "winner"):
                    getattr(gm, g_)[go] = getattr(sub, g_)
                nsec = int(sec.count[0])
                sec.parent_slot[:nsec, 0] = go[sec.parent_slot[:nsec, 0]]
                gm.edep_pemxsec[cut, 0] = 0.0
            else:
                reference.gamma_perform(gm, sec, seed, threads)
        if gm.n > 0:
            np.add.at(hist, gm_vol, gm.edep_pemxsec[:, 0])
            new_vol = np.where(onb, nv, gm_vol)
            ekin = gm.ekin_logekin[:, 0]
            stats["leak_gamma"] += float(ekin[(ekin > 0) & (new_vol < 0)].sum())
            (ce, cepos, cevol), (cg, cgpos, cgvol) = children(sec, gm.meta, gm_pos, gm_vol)
            stats["secondaries"] += int(sec.count[0])
            alive = np.flatnonzero((ekin > 0) & (new_vol >= 0))
            surv = _take(gm, alive, batches.GammaHostBatch)
            spos, svol = gm_pos[alive], new_vol[alive].astype(np.int32)
            surv.meta[:, 0] = np.where(onb[alive], slab.couple[svol % slab.na], surv.meta[:, 0])
            surv.edep_pemxsec[:, 0] = 0.0
            next_el += [(ce, cepos, cevol)]
            next_gm += [(surv, spos, svol), (cg, cgpos, cgvol)]
            next_sub += [new_sub(ce.n)]
        el = _concat([p[0] for p in next_el], batches.ElectronHostBatch)
        el_pos = np.concatenate([p[1] for p in next_el], axis=0) if next_el else np.zeros((0, 3))
        el_vol = np.concatenate([p[2] for p in next_el]).astype(np.int32) if next_el else np.zeros(0, dtype=np.int32)
        el_sub = {k: np.concatenate([d[k] for d in next_sub]) for k in new_sub(0)} if next_sub else new_sub(0)
        gm = _concat([p[0] for p in next_gm], batches.GammaHostBatch)
        gm_pos = np.concatenate([p[1] for p in next_gm], axis=0) if next_gm else np.zeros((0, 3))
        gm_vol = np.concatenate([p[2] for p in next_gm]).astype(np.int32) if next_gm else np.zeros(0, dtype=np.int32)
    return hist.reshape(slab.nl, slab.na), stats


class MixedCell:
    """The one-cell geometry g4hb200_mixed_run runs its loop in (capi_shower.inl)."""
    num_layers = 1
    absorber_thickness = (1.0,)
    absorber_couple = (0,)
    half_yz = 1.0
    woodcock = False


def mixed_population(reference, seed, n_el, n_gm, num_couples, emin, emax):
    """The population of g4hb200_mixed_run: MixedPopulationKernel of g4h_shower.cuh, operation by operation."""
    import math

    n = n_el + n_gm
    t = np.arange(n, dtype=np.int64)
    ids = t.astype(np.int32)
    gseed = seed ^ 0x1A2B3C4D
    lmin, lrange = math.log(emin), math.log(emax / emin)

    def u(k):
        return uniform_at(gseed, ids, np.full(n, k, dtype=np.int64))

    ekin = reference.vdt_log_exp(lmin + u(0) * lrange)[1]
    cost = 2.0 * u(1) - 1.0
    safety = u(2)
    sint = np.sqrt((1.0 - cost) * (1.0 + cost))
    vx, vy, r2 = np.zeros(n), np.zeros(n), np.full(n, 2.0)
    draw = np.full(n, 3, dtype=np.int64)
    todo = np.ones(n, dtype=bool)
    while todo.any():
        k = np.flatnonzero(todo)
        a = 2.0 * uniform_at(gseed, ids[k], draw[k]) - 1.0
        b = 2.0 * uniform_at(gseed, ids[k], draw[k] + 1) - 1.0
        draw[k] += 2
        vx[k], vy[k] = a, b
        r2[k] = a * a + b * b
        todo[k] = (r2[k] > 1.0) | (r2[k] == 0.0)
    cphi = (vx * vx - vy * vy) / r2
    sphi = 2.0 * vx * vy / r2
    is_gamma = t >= n_el
    o = np.where(is_gamma, t - n_el, t)
    half = n_el // 2
    is_pos = ~is_gamma & (o >= half)
    in_kind = np.where(is_gamma, o, np.where(is_pos, o - half, o))
    kind_size = np.where(is_gamma, n_gm, np.where(is_pos, n_el - half, half))
    imc = ((in_kind * num_couples) // np.maximum(kind_size, 1)).astype(np.int32)
    el = _new_electrons(n_el)
    e = np.flatnonzero(~is_gamma)
    el.ekin_logekin[:, 0] = ekin[e]
    el.dirx_diry[:, 0] = sint[e] * cphi[e]
    el.dirx_diry[:, 1] = sint[e] * sphi[e]
    el.dirz_safety[:, 0] = cost[e]
    el.dirz_safety[:, 1] = safety[e]
    el.meta[:, 0] = imc[e]
    el.meta[:, 1] = _capi.F_MSC_FIRST_STEP | np.where(is_pos[e], _capi.F_POSITRON, 0)
    el.meta[:, 2] = ids[e]
    gm = _new_gammas(n_gm)
    g = np.flatnonzero(is_gamma)
    gm.ekin_logekin[:, 0] = ekin[g]
    gm.dirx_diry[:, 0] = sint[g] * cphi[g]
    gm.dirx_diry[:, 1] = sint[g] * sphi[g]
    gm.dirz_nia0[:, 0] = cost[g]
    gm.meta[:, 0] = imc[g]
    gm.meta[:, 2] = ids[g]
    return el, gm


def run_mixed(reference, num_electrons, num_gammas, num_steps, seed, num_couples, emin=1.0e-3, emax=1.0e5, threads=4):
    """CPU driver of BASELINE configs[3] around the reference's managers: (total deposit, stats) like g4hepem_b200.shower.run_mixed."""
    el, gm = mixed_population(reference, seed, num_electrons, num_gammas, num_couples, emin, emax)
    hist, stats = run(reference, MixedCell(), num_electrons + num_gammas, 0.0, seed, max_steps=num_steps, threads=threads, mixed=(el, gm))
    return float(hist.sum()), stats
