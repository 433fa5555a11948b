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


def run(reference, calo, num_primaries, primary_ekin, seed, kind=_capi.SEC_ELECTRON, first_track_id=0, max_steps=0, threads=4):
    """Returns (edep[num_layers, num_absorbers], stats) like g4hepem_b200.shower.run."""
    slab = Slab(calo)
    hist = np.zeros(slab.nl * slab.na)
    stats = dict(num_steps=0, electron_track_steps=0, gamma_track_steps=0, secondaries=0, peak_electrons=0, peak_gammas=0,
                 leak_electron=0.0, leak_gamma=0.0)
    ids = first_track_id + np.arange(num_primaries, dtype=np.int32)
    if kind == _capi.SEC_GAMMA:
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
    el_pos = np.zeros((el.n, 3)); el_pos[:, 0] = slab.xfront
    gm_pos = np.zeros((gm.n, 3)); gm_pos[:, 0] = slab.xfront
    el_vol = np.zeros(el.n, dtype=np.int32)
    gm_vol = np.zeros(gm.n, dtype=np.int32)

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
        imc = slab.couple[vol % slab.na]
        ie = np.flatnonzero(knd != _capi.SEC_GAMMA)
        ig = np.flatnonzero(knd == _capi.SEC_GAMMA)
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
        next_el, next_gm = [], []
        # ---- e-/e+ ------------------------------------------------------------------------------------------------
        if el.n > 0:
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
            surv.dirz_safety[:, 1] = np.where(onb[alive], 0.0, slab.safety(svol, spos))
            surv.edep_dispx[...] = 0.0
            surv.winner[...] = -1
            next_el += [(surv, spos, svol), (ce, cepos, cevol)]
            next_gm += [(cg, cgpos, cgvol)]
        # ---- gamma ------------------------------------------------------------------------------------------------------
        if gm.n > 0:
            reference.gamma_howfar(gm, seed, threads)
            dirs = np.stack([gm.dirx_diry[:, 0], gm.dirx_diry[:, 1], gm.dirz_nia0[:, 0]], axis=1)
            dist, nv = slab.distance(gm_vol, gm_pos, dirs)
            onb = dist < gm.gstep_mfp0[:, 0]
            step = np.where(onb, dist, gm.gstep_mfp0[:, 0])
            gm_pos = gm_pos + step[:, None] * dirs
            gm.gstep_mfp0[:, 0] = step
            gm.meta[:, 1] = np.where(onb, gm.meta[:, 1] | _capi.F_ON_BOUNDARY, gm.meta[:, 1] & ~_capi.F_ON_BOUNDARY)
            sec = batches.SecondaryHostQueue(2 * gm.n)
            reference.gamma_perform(gm, sec, seed, threads)
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
        el = _concat([p[0] for p in next_el], batches.ElectronHostBatch)
        el_pos = np.concatenate([p[1] for p in next_el], axis=0) if next_el else np.zeros((0, 3))
        el_vol = np.concatenate([p[2] for p in next_el]).astype(np.int32) if next_el else np.zeros(0, dtype=np.int32)
        gm = _concat([p[0] for p in next_gm], batches.GammaHostBatch)
        gm_pos = np.concatenate([p[1] for p in next_gm], axis=0) if next_gm else np.zeros((0, 3))
        gm_vol = np.concatenate([p[2] for p in next_gm]).astype(np.int32) if next_gm else np.zeros(0, dtype=np.int32)
    return hist.reshape(slab.nl, slab.na), stats
