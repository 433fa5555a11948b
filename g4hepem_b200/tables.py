"""Offline table loader: G4HepEmDataJsonIO state JSON -> flat G4HB200Tables descriptor.

Reads the schema written by the reference's G4HepEmStateToJson
(G4HepEm/G4HepEmDataJsonIO/src/G4HepEmDataJsonIOImpl.hh:127-977; keys listed in SURVEY.md App. D)
and lays the arrays out as include/g4hepem_b200.h describes.  The C++ twin of this function is
g4hepem_b200/host/G4HepEmB200Flatten.hh (same descriptor from in-memory G4HepEmData); the tests check
that both produce identical descriptors.
"""
import ctypes as C
import json

import numpy as np

from . import _capi

_F8 = np.float64
_I4 = np.int32


def _d(a):
    return np.ascontiguousarray(np.asarray([] if a is None else a, dtype=_F8))


def _i(a):
    return np.ascontiguousarray(np.asarray([] if a is None else a, dtype=_I4))


class FlatTables:
    """Owns the numpy arrays the ctypes descriptor points into."""

    def __init__(self, state):
        self._keep = []
        self.desc = _capi.Tables()
        self.state = state
        self._fill(state)

    # -- helpers -------------------------------------------------------------------------------
    def _pd(self, arr):
        arr = _d(arr)
        self._keep.append(arr)
        return arr.ctypes.data_as(_capi.c_dp) if arr.size else C.cast(None, _capi.c_dp)

    def _pi(self, arr):
        arr = _i(arr)
        self._keep.append(arr)
        return arr.ctypes.data_as(_capi.c_ip) if arr.size else C.cast(None, _capi.c_ip)

    def _electron(self, ed, t):
        grid = _d(ed["fELossEnergyGrid"])
        t.num_loss = grid.size
        t.loss_log_min_ekin = ed["fELossLogMinEkin"]
        t.loss_eil_delta = ed["fELossEILDelta"]
        t.loss_egrid = self._pd(grid)
        t.loss_data = self._pd(ed["fELossData"])
        t.resmx_start = self._pi(ed["fResMacXSecStartIndexPerMatCut"])
        res = _d(ed["fResMacXSecData"])
        t.resmx_data = self._pd(res)
        t.num_resmx = res.size
        t.enuc_log_min_ekin = ed["fENucLogMinEkin"]
        t.enuc_eil_delta = ed["fENucEILDelta"]
        t.enuc_egrid = self._pd(ed["fENucEnergyGrid"])
        t.enuc_data = self._pd(ed["fENucMacXsecData"])
        t.tr1_data = self._pd(ed["fTr1MacXSecData"])
        for short, key in (("ioni", "Ioni"), ("sb", "BremSB"), ("rb", "BremRB")):
            data = _d(ed[f"fElemSelector{key}Data"])
            setattr(t, f"sel_{short}_start", self._pi(ed[f"fElemSelector{key}StartIndexPerMatCut"]))
            setattr(t, f"sel_{short}_data", self._pd(data))
            setattr(t, f"num_sel_{short}", data.size)

    def _fill(self, state):
        t = self.desc
        p = state["fParameters"]
        d = state["fData"]
        t.electron_tracking_cut = p["fElectronTrackingCut"]
        t.gamma_tracking_cut = p["fGammaTrackingCut"]
        t.min_loss_table_energy = p["fMinLossTableEnergy"]
        t.electron_brem_model_lim = p["fElectronBremModelLim"]
        t.is_msc_positron_cor = int(bool(p["fIsMSCPositronCor"]))
        t.is_msc_displacement = int(bool(p["fIsMSCDisplacement"]))
        regs = p["fParametersPerRegion"] or []
        t.num_regions = len(regs)
        rp = []
        for r in regs:
            rp += [r["fFinalRange"], r["fDRoverRange"], r["fLinELossLimit"], r["fMSCRangeFactor"], r["fMSCSafetyFactor"],
                   float(bool(r["fIsMSCMinimalStepLimit"])), float(bool(r["fIsELossFluctuation"])),
                   float(bool(r["fIsMultipleStepsInMSCTrans"])) + 2.0 * float(bool(r.get("fIsApplyCuts", True)))]
        t.region_pars = self._pd(rp)
        # couples
        mcs = d["fTheMatCutData"]["fMatCutData"]
        t.num_matcut = len(mcs)
        t.mc_cuts = self._pd([v for m in mcs for v in (m["fSecElProdCutE"], m["fSecPosProdCutE"], m["fSecGamProdCutE"], m["fLogSecGamCutE"])])
        t.mc_imat = self._pi([m["fHepEmMatIndex"] for m in mcs])
        t.mc_ireg = self._pi([m["fG4RegionIndex"] for m in mcs])
        # materials + Sandia pools
        mats = d["fTheMaterialData"]["fMaterialData"]
        t.num_mat = len(mats)
        nelem, estart, ez, enat, mpars, snum, sstart, sen, scof = [], [], [], [], [], [], [], [], []
        for m in mats:
            zs = m["fElementVect"] or []
            nelem.append(len(zs))
            estart.append(len(ez))
            ez += list(zs)
            enat += list(m["fNumOfAtomsPerVolumeVect"] or [])
            mpars += [m["fDensityCorfactor"], m["fElectronDensity"], m["fRadiationLength"], m["fMeanExEnergy"], m["fZeff"],
                      m["fZeff23"], m["fZeffSqrt"], m["fUMSCPar"], *m["fUMSCStepMinPars"], *m["fUMSCTailCoeff"],
                      *m["fUMSCThetaCoeff"]]
            se = m["fSandiaEnergies"] or []
            snum.append(len(se))
            sstart.append(len(sen))
            sen += list(se)
            scof += list(m["fSandiaCoefficients"] or [])
        epars = np.zeros(12 * 121)
        esn = np.zeros(121, dtype=_I4)
        ess = np.zeros(121, dtype=_I4)
        for e in sorted(d["fTheElementData"] or [], key=lambda e: e["fZet"]):
            z = int(e["fZet"])
            epars[12 * z: 12 * z + 11] = [e["fZet"], e["fZet13"], e["fZet23"], e["fCoulomb"], e["fLogZ"], e["fZFactor1"],
                                          e["fDeltaMaxLow"], e["fDeltaMaxHigh"], e["fILVarS1"], e["fILVarS1Cond"],
                                          e["fKShellBindingEnergy"]]
            se = e["fSandiaEnergies"] or []
            esn[z] = len(se)
            ess[z] = len(sen)
            sen += list(se)
            scof += list(e["fSandiaCoefficients"] or [])
        t.mat_num_elem = self._pi(nelem)
        t.mat_elem_start = self._pi(estart)
        t.mat_elem_z = self._pi(ez)
        t.mat_elem_natoms = self._pd(enat)
        t.mat_pars = self._pd(mpars)
        t.mat_sandia_num = self._pi(snum)
        t.mat_sandia_start = self._pi(sstart)
        t.elem_pars = self._pd(epars)
        t.elem_sandia_num = self._pi(esn)
        t.elem_sandia_start = self._pi(ess)
        t.num_sandia = len(sen)
        t.sandia_energies = self._pd(sen)
        t.sandia_cof = self._pd(scof)
        self._electron(d["fTheElectronData"], t.electron)
        self._electron(d["fThePositronData"], t.positron)
        sb = d["fTheSBTableData"]
        t.sb_log_min_el_energy = sb["fLogMinElEnergy"]
        t.sb_il_delta_el_energy = sb["fILDeltaElEnergy"]
        t.sb_el_energy = self._pd(sb["fElEnergyVect"])
        t.sb_lel_energy = self._pd(sb["fLElEnergyVect"])
        t.sb_lkappa = self._pd(sb["fLKappaVect"])
        t.sb_gcut_start = self._pi(sb["fGammaCutIndxStartIndexPerMC"])
        gci = _i(sb["fGammaCutIndices"])
        t.sb_gcut_indices = self._pi(gci)
        t.num_sb_gcut = gci.size
        t.sb_start_per_z = self._pi(sb["fSBStartTablesStartPerZ"])
        sbd = _d(sb["fSBTableData"])
        t.sb_data = self._pd(sbd)
        t.num_sb_data = sbd.size
        gm = d["fTheGammaData"]
        t.gm_data_per_mat = gm["fDataPerMat"]
        t.gm_num_data0 = gm["fNumData0"]
        t.gm_num_data1 = gm["fNumData1"]
        t.gm_emax0 = gm["fEMax0"]
        t.gm_log_emin0 = gm["fLogEMin0"]
        t.gm_eil_delta0 = gm["fEILDelta0"]
        t.gm_emax1 = gm["fEMax1"]
        t.gm_log_emin1 = gm["fLogEMin1"]
        t.gm_eil_delta1 = gm["fEILDelta1"]
        t.gm_log_emin2 = gm["fLogEMin2"]
        t.gm_eil_delta2 = gm["fEILDelta2"]
        t.gm_mxsec = self._pd(gm["fMacXsecData"])
        cg = _d(gm["fElemSelectorConvEgrid"])
        t.gm_conv_egrid_size = cg.size
        t.gm_conv_log_min_ekin = gm["fElemSelectorConvLogMinEkin"]
        t.gm_conv_eil_delta = gm["fElemSelectorConvEILDelta"]
        t.gm_conv_start = self._pi(gm["fElemSelectorConvStartIndexPerMat"])
        t.gm_conv_egrid = self._pd(cg)
        cd = _d(gm["fElemSelectorConvData"])
        t.gm_conv_data = self._pd(cd)
        t.num_gm_conv = cd.size

    # -- convenience views used by the batch generators -----------------------------------------
    @property
    def num_matcut(self):
        return int(self.desc.num_matcut)

    @property
    def num_mat(self):
        return int(self.desc.num_mat)

    def couple_cuts(self):
        n = self.num_matcut
        return np.ctypeslib.as_array(self.desc.mc_cuts, shape=(n, 4)).copy()

    def couple_material(self):
        return np.ctypeslib.as_array(self.desc.mc_imat, shape=(self.num_matcut,)).copy()

    def couple_region(self):
        return np.ctypeslib.as_array(self.desc.mc_ireg, shape=(self.num_matcut,)).copy()

    def region_pars(self):
        """(num_regions, 8): final_range, dr_over_range, lin_eloss_limit, msc_range_factor, msc_safety_factor,
        is_msc_minimal_step_limit, is_eloss_fluctuation, caller flags (multiple MSC steps + 2 * apply cuts)"""
        return np.ctypeslib.as_array(self.desc.region_pars, shape=(int(self.desc.num_regions), 8)).copy()


def load_state_json(path):
    with open(path) as f:
        return FlatTables(json.load(f))


# array-valued fields of the descriptor and how to compute their length, for comparisons / dumps
def descriptor_arrays(t):
    """Yield (name, numpy array) for every array the descriptor points to."""
    def arr(ptr, n):
        n = int(n)
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n > 0 and ptr else np.zeros(0)

    nmc, nmat = t.num_matcut, t.num_mat
    nel_tot = int(sum(arr(t.mat_num_elem, nmat))) if nmat else 0
    yield "region_pars", arr(t.region_pars, 8 * t.num_regions)
    yield "mc_cuts", arr(t.mc_cuts, 4 * nmc)
    yield "mc_imat", arr(t.mc_imat, nmc)
    yield "mc_ireg", arr(t.mc_ireg, nmc)
    yield "mat_num_elem", arr(t.mat_num_elem, nmat)
    yield "mat_elem_start", arr(t.mat_elem_start, nmat)
    yield "mat_elem_z", arr(t.mat_elem_z, nel_tot)
    yield "mat_elem_natoms", arr(t.mat_elem_natoms, nel_tot)
    yield "mat_pars", arr(t.mat_pars, 16 * nmat)
    yield "mat_sandia_num", arr(t.mat_sandia_num, nmat)
    yield "mat_sandia_start", arr(t.mat_sandia_start, nmat)
    yield "elem_pars", arr(t.elem_pars, 12 * 121)
    yield "elem_sandia_num", arr(t.elem_sandia_num, 121)
    yield "elem_sandia_start", arr(t.elem_sandia_start, 121)
    yield "sandia_energies", arr(t.sandia_energies, t.num_sandia)
    yield "sandia_cof", arr(t.sandia_cof, 4 * t.num_sandia)
    for name in ("electron", "positron"):
        e = getattr(t, name)
        yield f"{name}.loss_egrid", arr(e.loss_egrid, e.num_loss)
        yield f"{name}.loss_data", arr(e.loss_data, 5 * e.num_loss * nmc)
        yield f"{name}.resmx_start", arr(e.resmx_start, nmc)
        yield f"{name}.resmx_data", arr(e.resmx_data, e.num_resmx)
        yield f"{name}.enuc_egrid", arr(e.enuc_egrid, 128)
        yield f"{name}.enuc_data", arr(e.enuc_data, 2 * 128 * nmat)
        yield f"{name}.tr1_data", arr(e.tr1_data, 2 * e.num_loss * nmat)
        for s in ("ioni", "sb", "rb"):
            yield f"{name}.sel_{s}_start", arr(getattr(e, f"sel_{s}_start"), nmc)
            yield f"{name}.sel_{s}_data", arr(getattr(e, f"sel_{s}_data"), getattr(e, f"num_sel_{s}"))
    yield "sb_el_energy", arr(t.sb_el_energy, 65)
    yield "sb_lel_energy", arr(t.sb_lel_energy, 65)
    yield "sb_lkappa", arr(t.sb_lkappa, 54)
    yield "sb_gcut_start", arr(t.sb_gcut_start, nmc)
    yield "sb_gcut_indices", arr(t.sb_gcut_indices, t.num_sb_gcut)
    yield "sb_start_per_z", arr(t.sb_start_per_z, 121)
    yield "sb_data", arr(t.sb_data, t.num_sb_data)
    yield "gm_mxsec", arr(t.gm_mxsec, nmat * t.gm_data_per_mat)
    yield "gm_conv_start", arr(t.gm_conv_start, nmat)
    yield "gm_conv_egrid", arr(t.gm_conv_egrid, t.gm_conv_egrid_size)
    yield "gm_conv_data", arr(t.gm_conv_data, t.num_gm_conv)


SCALAR_FIELDS = [n for n, ty in _capi.Tables._fields_ if ty in (C.c_double, C.c_int32)]
ELECTRON_SCALAR_FIELDS = [n for n, ty in _capi.ElectronTables._fields_ if ty in (C.c_double, C.c_int32)]
