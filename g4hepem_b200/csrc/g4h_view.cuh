// g4h_view.cuh -- TablesView from a flat descriptor whose pointers are valid in the address space the
// view will be used in (device pointers for the kernels).
#ifndef G4H_VIEW_CUH
#define G4H_VIEW_CUH

#include "../../include/g4hepem_b200.h"
#include "g4h_tables.cuh"

namespace g4h {

inline void MakeElectronView(const G4HB200ElectronTables& t, ElectronTablesView& v) {
  v.numLoss = t.num_loss;
  v.lossLogMinEkin = t.loss_log_min_ekin;
  v.lossEILDelta = t.loss_eil_delta;
  v.lossEGrid = t.loss_egrid;
  v.lossData = t.loss_data;
  v.resStart = t.resmx_start;
  v.resData = t.resmx_data;
  v.enucLogMinEkin = t.enuc_log_min_ekin;
  v.enucEILDelta = t.enuc_eil_delta;
  v.enucEGrid = t.enuc_egrid;
  v.enucData = t.enuc_data;
  v.tr1Data = t.tr1_data;
  v.selSBStart = t.sel_sb_start;
  v.selSBData = t.sel_sb_data;
  v.selRBStart = t.sel_rb_start;
  v.selRBData = t.sel_rb_data;
}

inline TablesView MakeView(const G4HB200Tables& t) {
  TablesView v;
  v.elTrackingCut = t.electron_tracking_cut;
  v.gammaTrackingCut = t.gamma_tracking_cut;
  v.minLossTableEnergy = t.min_loss_table_energy;
  v.bremModelLim = t.electron_brem_model_lim;
  v.isMSCPositronCor = t.is_msc_positron_cor;
  v.isMSCDisplacement = t.is_msc_displacement;
  v.numRegions = t.num_regions;
  v.numMatCut = t.num_matcut;
  v.numMat = t.num_mat;
  v.regionPars = t.region_pars;
  v.mcCuts = t.mc_cuts;
  v.mcImat = t.mc_imat;
  v.mcIreg = t.mc_ireg;
  v.matNumElem = t.mat_num_elem;
  v.matElemStart = t.mat_elem_start;
  v.matElemZ = t.mat_elem_z;
  v.matElemNatoms = t.mat_elem_natoms;
  v.matPars = t.mat_pars;
  v.matSandiaNum = t.mat_sandia_num;
  v.matSandiaStart = t.mat_sandia_start;
  v.elemPars = t.elem_pars;
  v.elemSandiaNum = t.elem_sandia_num;
  v.elemSandiaStart = t.elem_sandia_start;
  v.sandiaEnergies = t.sandia_energies;
  v.sandiaCof = t.sandia_cof;
  MakeElectronView(t.electron, v.el[0]);
  MakeElectronView(t.positron, v.el[1]);
  v.sbLogMinElEnergy = t.sb_log_min_el_energy;
  v.sbILDeltaElEnergy = t.sb_il_delta_el_energy;
  v.sbElEnergy = t.sb_el_energy;
  v.sbLElEnergy = t.sb_lel_energy;
  v.sbLKappa = t.sb_lkappa;
  v.sbGCutStart = t.sb_gcut_start;
  v.sbGCutIndices = t.sb_gcut_indices;
  v.sbStartPerZ = t.sb_start_per_z;
  v.sbData = t.sb_data;
  v.gmDataPerMat = t.gm_data_per_mat;
  v.gmNumData0 = t.gm_num_data0;
  v.gmNumData1 = t.gm_num_data1;
  v.gmEMax0 = t.gm_emax0;
  v.gmLogEMin0 = t.gm_log_emin0;
  v.gmEILDelta0 = t.gm_eil_delta0;
  v.gmEMax1 = t.gm_emax1;
  v.gmLogEMin1 = t.gm_log_emin1;
  v.gmEILDelta1 = t.gm_eil_delta1;
  v.gmLogEMin2 = t.gm_log_emin2;
  v.gmEILDelta2 = t.gm_eil_delta2;
  v.gmMXsec = t.gm_mxsec;
  v.gmConvEGridSize = t.gm_conv_egrid_size;
  v.gmConvLogMinEkin = t.gm_conv_log_min_ekin;
  v.gmConvEILDelta = t.gm_conv_eil_delta;
  v.gmConvStart = t.gm_conv_start;
  v.gmConvEGrid = t.gm_conv_egrid;
  v.gmConvData = t.gm_conv_data;
  return v;
}

}  // namespace g4h
#endif
