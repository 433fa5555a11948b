// G4HepEmB200Flatten.hh -- host side (C++) of the drop-in boundary: G4HepEmData/G4HepEmParameters
// --> flat G4HB200Tables descriptor (include/g4hepem_b200.h).
//
// Header-only and compiled in the *application's* translation unit, against the application's own
// G4HepEm headers (G4HepEmData.hh, G4HepEmParameters.hh, ... must be included before this file).
// It replaces the AoS-of-pointers deep copies of CopyG4HepEmDataToGPU
// (G4HepEm/G4HepEmData/src/G4HepEmData.cc:78-101 and the per-struct Copy*ToGPU/Device functions):
// instead of one cudaMalloc+cudaMemcpy per array, the arrays are described by plain pointers and
// g4hb200_create() packs them into one device arena.
//
// Ownership: the descriptor borrows the pointers of G4HepEmData for the big tables and owns the
// small re-packed arrays (region parameters, per-material / per-element scalars, Sandia pools).
#ifndef G4HEPEMB200_FLATTEN_HH
#define G4HEPEMB200_FLATTEN_HH

#include <cstdint>
#include <vector>

#include "g4hepem_b200.h"

struct G4HepEmB200FlatTables {
  G4HB200Tables desc;
  std::vector<double> regionPars, mcCuts, matPars, matElemNatoms, elemPars, sandiaEnergies, sandiaCof;
  std::vector<int32_t> mcImat, mcIreg, matNumElem, matElemStart, matElemZ, matSandiaNum, matSandiaStart,
      elemSandiaNum, elemSandiaStart;
};

namespace g4hepemb200 {

inline void FlattenElectronData(const G4HepEmElectronData* d, G4HB200ElectronTables& t) {
  t.num_loss           = d->fELossEnergyGridSize;
  t.loss_log_min_ekin  = d->fELossLogMinEkin;
  t.loss_eil_delta     = d->fELossEILDelta;
  t.loss_egrid         = d->fELossEnergyGrid;
  t.loss_data          = d->fELossData;
  t.resmx_start        = d->fResMacXSecStartIndexPerMatCut;
  t.resmx_data         = d->fResMacXSecData;
  t.num_resmx          = d->fResMacXSecNumData;
  t.enuc_log_min_ekin  = d->fENucLogMinEkin;
  t.enuc_eil_delta     = d->fENucEILDelta;
  t.enuc_egrid         = d->fENucEnergyGrid;
  t.enuc_data          = d->fENucMacXsecData;
  t.tr1_data           = d->fTr1MacXSecData;
  t.sel_ioni_start     = d->fElemSelectorIoniStartIndexPerMatCut;
  t.sel_ioni_data      = d->fElemSelectorIoniData;
  t.num_sel_ioni       = d->fElemSelectorIoniNumData;
  t.sel_sb_start       = d->fElemSelectorBremSBStartIndexPerMatCut;
  t.sel_sb_data        = d->fElemSelectorBremSBData;
  t.num_sel_sb         = d->fElemSelectorBremSBNumData;
  t.sel_rb_start       = d->fElemSelectorBremRBStartIndexPerMatCut;
  t.sel_rb_data        = d->fElemSelectorBremRBData;
  t.num_sel_rb         = d->fElemSelectorBremRBNumData;
}

}  // namespace g4hepemb200

// Fill `out` from the reference's host structures.  `out` must outlive every use of out.desc.
inline void G4HepEmB200Flatten(const G4HepEmData* data, const G4HepEmParameters* pars, G4HepEmB200FlatTables& out) {
  G4HB200Tables& t = out.desc;
  // --- parameters
  t.electron_tracking_cut   = pars->fElectronTrackingCut;
  t.gamma_tracking_cut      = pars->fGammaTrackingCut;
  t.min_loss_table_energy   = pars->fMinLossTableEnergy;
  t.electron_brem_model_lim = pars->fElectronBremModelLim;
  t.is_msc_positron_cor     = pars->fIsMSCPositronCor ? 1 : 0;
  t.is_msc_displacement     = pars->fIsMSCDisplacement ? 1 : 0;
  t.num_regions             = pars->fNumRegions;
  out.regionPars.clear();
  for (int i = 0; i < pars->fNumRegions; ++i) {
    const G4HepEmRegionParmeters& r = pars->fParametersPerRegion[i];
    const double v[8] = {r.fFinalRange, r.fDRoverRange, r.fLinELossLimit, r.fMSCRangeFactor, r.fMSCSafetyFactor,
                         r.fIsMSCMinimalStepLimit ? 1.0 : 0.0, r.fIsELossFluctuation ? 1.0 : 0.0,
                         (r.fIsMultipleStepsInMSCTrans ? 1.0 : 0.0) + (r.fIsApplyCuts ? 2.0 : 0.0)};
    out.regionPars.insert(out.regionPars.end(), v, v + 8);
  }
  t.region_pars = out.regionPars.data();
  // --- material-cuts couples
  const G4HepEmMatCutData* mc = data->fTheMatCutData;
  t.num_matcut = mc->fNumMatCutData;
  out.mcCuts.clear(); out.mcImat.clear(); out.mcIreg.clear();
  for (int i = 0; i < mc->fNumMatCutData; ++i) {
    const G4HepEmMCCData& c = mc->fMatCutData[i];
    const double v[4] = {c.fSecElProdCutE, c.fSecPosProdCutE, c.fSecGamProdCutE, c.fLogSecGamCutE};
    out.mcCuts.insert(out.mcCuts.end(), v, v + 4);
    out.mcImat.push_back(c.fHepEmMatIndex);
    out.mcIreg.push_back(c.fG4RegionIndex);
  }
  t.mc_cuts = out.mcCuts.data();
  t.mc_imat = out.mcImat.data();
  t.mc_ireg = out.mcIreg.data();
  // --- materials
  const G4HepEmMaterialData* md = data->fTheMaterialData;
  t.num_mat = md->fNumMaterialData;
  out.matNumElem.clear(); out.matElemStart.clear(); out.matElemZ.clear(); out.matElemNatoms.clear();
  out.matPars.clear(); out.matSandiaNum.clear(); out.matSandiaStart.clear();
  out.sandiaEnergies.clear(); out.sandiaCof.clear();
  for (int i = 0; i < md->fNumMaterialData; ++i) {
    const G4HepEmMatData& m = md->fMaterialData[i];
    out.matNumElem.push_back(m.fNumOfElement);
    out.matElemStart.push_back(static_cast<int32_t>(out.matElemZ.size()));
    for (int e = 0; e < m.fNumOfElement; ++e) {
      out.matElemZ.push_back(m.fElementVect[e]);
      out.matElemNatoms.push_back(m.fNumOfAtomsPerVolumeVect[e]);
    }
    const double v[16] = {m.fDensityCorFactor, m.fElectronDensity, m.fRadiationLength, m.fMeanExEnergy, m.fZeff,
                          m.fZeff23, m.fZeffSqrt, m.fUMSCPar, m.fUMSCStepMinPars[0], m.fUMSCStepMinPars[1],
                          m.fUMSCTailCoeff[0], m.fUMSCTailCoeff[1], m.fUMSCTailCoeff[2], m.fUMSCTailCoeff[3],
                          m.fUMSCThetaCoeff[0], m.fUMSCThetaCoeff[1]};
    out.matPars.insert(out.matPars.end(), v, v + 16);
    out.matSandiaNum.push_back(m.fNumOfSandiaIntervals);
    out.matSandiaStart.push_back(static_cast<int32_t>(out.sandiaEnergies.size()));
    for (int s = 0; s < m.fNumOfSandiaIntervals; ++s) {
      out.sandiaEnergies.push_back(m.fSandiaEnergies[s]);
      for (int k = 0; k < 4; ++k) out.sandiaCof.push_back(m.fSandiaCoefficients[4 * s + k]);
    }
  }
  // --- elements (indexed by Z)
  const G4HepEmElementData* ed = data->fTheElementData;
  out.elemPars.assign(12 * 121, 0.0);
  out.elemSandiaNum.assign(121, 0);
  out.elemSandiaStart.assign(121, 0);
  for (int z = 0; z <= ed->fMaxZet && z < 121; ++z) {
    const G4HepEmElemData& e = ed->fElementData[z];
    if (e.fZet <= 0.0) continue;
    const double v[12] = {e.fZet, e.fZet13, e.fZet23, e.fCoulomb, e.fLogZ, e.fZFactor1, e.fDeltaMaxLow,
                          e.fDeltaMaxHigh, e.fILVarS1, e.fILVarS1Cond, e.fKShellBindingEnergy, 0.0};
    for (int k = 0; k < 12; ++k) out.elemPars[12 * z + k] = v[k];
    out.elemSandiaNum[z]   = e.fNumOfSandiaIntervals;
    out.elemSandiaStart[z] = static_cast<int32_t>(out.sandiaEnergies.size());
    for (int s = 0; s < e.fNumOfSandiaIntervals; ++s) {
      out.sandiaEnergies.push_back(e.fSandiaEnergies[s]);
      for (int k = 0; k < 4; ++k) out.sandiaCof.push_back(e.fSandiaCoefficients[4 * s + k]);
    }
  }
  t.mat_num_elem      = out.matNumElem.data();
  t.mat_elem_start    = out.matElemStart.data();
  t.mat_elem_z        = out.matElemZ.data();
  t.mat_elem_natoms   = out.matElemNatoms.data();
  t.mat_pars          = out.matPars.data();
  t.mat_sandia_num    = out.matSandiaNum.data();
  t.mat_sandia_start  = out.matSandiaStart.data();
  t.elem_pars         = out.elemPars.data();
  t.elem_sandia_num   = out.elemSandiaNum.data();
  t.elem_sandia_start = out.elemSandiaStart.data();
  t.num_sandia        = static_cast<int32_t>(out.sandiaEnergies.size());
  t.sandia_energies   = out.sandiaEnergies.data();
  t.sandia_cof        = out.sandiaCof.data();
  // --- e-/e+
  g4hepemb200::FlattenElectronData(data->fTheElectronData, t.electron);
  g4hepemb200::FlattenElectronData(data->fThePositronData, t.positron);
  // --- Seltzer-Berger tables
  const G4HepEmSBTableData* sb = data->fTheSBTableData;
  t.sb_log_min_el_energy  = sb->fLogMinElEnergy;
  t.sb_il_delta_el_energy = sb->fILDeltaElEnergy;
  t.sb_el_energy          = sb->fElEnergyVect;
  t.sb_lel_energy         = sb->fLElEnergyVect;
  t.sb_lkappa             = sb->fLKappaVect;
  t.sb_gcut_start         = sb->fGammaCutIndxStartIndexPerMC;
  t.sb_gcut_indices       = sb->fGammaCutIndices;
  t.num_sb_gcut           = sb->fNumElemsInMatCuts;
  t.sb_start_per_z        = sb->fSBTablesStartPerZ;
  t.sb_data               = sb->fSBTableData;
  t.num_sb_data           = sb->fNumSBTableData;
  // --- gamma
  const G4HepEmGammaData* gm = data->fTheGammaData;
  t.gm_data_per_mat      = gm->fDataPerMat;
  t.gm_num_data0         = gm->fNumData0;
  t.gm_num_data1         = gm->fNumData1;
  t.gm_emax0             = gm->fEMax0;
  t.gm_log_emin0         = gm->fLogEMin0;
  t.gm_eil_delta0        = gm->fEILDelta0;
  t.gm_emax1             = gm->fEMax1;
  t.gm_log_emin1         = gm->fLogEMin1;
  t.gm_eil_delta1        = gm->fEILDelta1;
  t.gm_log_emin2         = gm->fLogEMin2;
  t.gm_eil_delta2        = gm->fEILDelta2;
  t.gm_mxsec             = gm->fMacXsecData;
  t.gm_conv_egrid_size   = gm->fElemSelectorConvEgridSize;
  t.gm_conv_log_min_ekin = gm->fElemSelectorConvLogMinEkin;
  t.gm_conv_eil_delta    = gm->fElemSelectorConvEILDelta;
  t.gm_conv_start        = gm->fElemSelectorConvStartIndexPerMat;
  t.gm_conv_egrid        = gm->fElemSelectorConvEgrid;
  t.gm_conv_data         = gm->fElemSelectorConvData;
  t.num_gm_conv          = gm->fElemSelectorConvNumData;
}

#endif  // G4HEPEMB200_FLATTEN_HH
