/* g4hepem_b200.h -- C-ABI of the B200-native G4HepEm stepping path.
 *
 * This is the drop-in boundary: a host program (the G4HepEm C++ managers, a tracking
 * manager, a Python harness over ctypes) calls these entry points with plain pointers and
 * sizes.  Each entry point names the reference interface it replaces (paths relative to
 * G4HepEm/ in mnovak42/g4hepem).
 *
 * Units follow Geant4 internal units (MeV, mm).  All functions return 0 on success, a
 * negative G4HB200_E* code otherwise; they never call exit() (the reference's gpuErrchk does,
 * G4HepEmData/include/G4HepEmCuUtils.hh:19-26).
 *
 * There is no CPU fallback: every compute entry point launches sm_100a kernels and fails with
 * G4HB200_ENODEVICE when no CUDA device is usable.
 */
#ifndef G4HEPEM_B200_H
#define G4HEPEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G4HB200_OK 0
#define G4HB200_EINVAL (-1)    /* bad argument / inconsistent table sizes */
#define G4HB200_ENODEVICE (-2) /* no usable CUDA device */
#define G4HB200_ECUDA (-3)     /* CUDA runtime error (see g4hb200_last_error) */
#define G4HB200_ENOMEM (-4)
#define G4HB200_ECAPACITY (-5) /* batch / secondary queue capacity exceeded */

/* ------------------------------------------------------------------------------------------
 * Flat table descriptor: the content of G4HepEmData + G4HepEmParameters as plain arrays.
 * Replaces CopyG4HepEmDataToGPU / CopyG4HepEmParametersToGPU (G4HepEmData/src/G4HepEmData.cc:78-101,
 * G4HepEmData/src/G4HepEmParameters.cc) -- the AoS-of-pointers deep copy becomes one contiguous
 * device arena.  Array layouts are the reference's (G4HepEmData/include/G4HepEm*Data.hh).
 * ------------------------------------------------------------------------------------------ */
typedef struct G4HB200ElectronTables { /* G4HepEmElectronData (G4HepEmElectronData.hh:81-354) */
  int32_t num_loss;                    /* fELossEnergyGridSize */
  double loss_log_min_ekin;            /* fELossLogMinEkin */
  double loss_eil_delta;               /* fELossEILDelta */
  const double* loss_egrid;            /* [num_loss] */
  const double* loss_data;             /* [5*num_loss*num_matcut] range,sd | dedx,sd | inv-range sd */
  const int32_t* resmx_start;          /* [num_matcut] fResMacXSecStartIndexPerMatCut */
  const double* resmx_data;            /* [num_resmx] fResMacXSecData */
  int32_t num_resmx;
  double enuc_log_min_ekin;            /* fENucLogMinEkin */
  double enuc_eil_delta;               /* fENucEILDelta */
  const double* enuc_egrid;            /* [128] */
  const double* enuc_data;             /* [2*128*num_mat] */
  const double* tr1_data;              /* [2*num_loss*num_mat] */
  const int32_t* sel_ioni_start;       /* [num_matcut] (-1: single element) */
  const double* sel_ioni_data;
  int32_t num_sel_ioni;
  const int32_t* sel_sb_start;
  const double* sel_sb_data;
  int32_t num_sel_sb;
  const int32_t* sel_rb_start;
  const double* sel_rb_data;
  int32_t num_sel_rb;
} G4HB200ElectronTables;

typedef struct G4HB200Tables {
  /* G4HepEmParameters (G4HepEmParameters.hh:51-92) */
  double electron_tracking_cut;
  double gamma_tracking_cut;
  double min_loss_table_energy;
  double electron_brem_model_lim;
  int32_t is_msc_positron_cor;
  int32_t is_msc_displacement;
  int32_t num_regions;
  /* per region, G4HepEmRegionParmeters (G4HepEmParameters.hh:24-48): 8 doubles each:
   * final_range, dr_over_range, lin_eloss_limit, msc_range_factor, msc_safety_factor,
   * is_msc_minimal_step_limit, is_eloss_fluctuation (flags as 0/1), caller flags = is_multiple_steps_in_msc_trans
   * + 2 * is_apply_cuts (the two region parameters only the stepping loop around the managers reads,
   * G4HepEmTrackingManager.cc:426-427) */
  const double* region_pars;
  /* G4HepEmMatCutData (G4HepEmMatCutData.hh:41-68) */
  int32_t num_matcut;
  const double* mc_cuts;  /* [4*num_matcut] el cut, pos cut, gamma cut, log gamma cut */
  const int32_t* mc_imat; /* [num_matcut] fHepEmMatIndex */
  const int32_t* mc_ireg; /* [num_matcut] fG4RegionIndex */
  /* G4HepEmMaterialData (G4HepEmMaterialData.hh:41-90) */
  int32_t num_mat;
  const int32_t* mat_num_elem;   /* [num_mat] */
  const int32_t* mat_elem_start; /* [num_mat] offset into mat_elem_z / mat_elem_natoms */
  const int32_t* mat_elem_z;     /* [sum num_elem] */
  const double* mat_elem_natoms; /* [sum num_elem] fNumOfAtomsPerVolumeVect */
  /* [16*num_mat]: density_cor_factor, electron_density, radiation_length, mean_exc_energy, zeff,
   * zeff23, zeff_sqrt, umsc_par, umsc_stepmin[2], umsc_tail[4], umsc_theta[2] */
  const double* mat_pars;
  const int32_t* mat_sandia_num;   /* [num_mat] */
  const int32_t* mat_sandia_start; /* [num_mat] offset into sandia_energies (x4 into sandia_cof) */
  /* G4HepEmElementData (G4HepEmElementData.hh:36-85), indexed by Z in [0,120] */
  /* [12*121]: zet, zet13, zet23, coulomb, logz, zfactor1, delta_max_low, delta_max_high, il_var_s1,
   * il_var_s1_cond, kshell_binding, unused */
  const double* elem_pars;
  const int32_t* elem_sandia_num;   /* [121] */
  const int32_t* elem_sandia_start; /* [121] */
  int32_t num_sandia;               /* total Sandia intervals (materials + elements) */
  const double* sandia_energies;    /* [num_sandia] */
  const double* sandia_cof;         /* [4*num_sandia] */
  /* e- / e+ */
  G4HB200ElectronTables electron;
  G4HB200ElectronTables positron;
  /* G4HepEmSBTableData (G4HepEmSBTableData.hh:8-42) */
  double sb_log_min_el_energy;
  double sb_il_delta_el_energy;
  const double* sb_el_energy;       /* [65] */
  const double* sb_lel_energy;      /* [65] */
  const double* sb_lkappa;          /* [54] */
  const int32_t* sb_gcut_start;     /* [num_matcut] fGammaCutIndxStartIndexPerMC */
  const int32_t* sb_gcut_indices;   /* [num_sb_gcut] fGammaCutIndices */
  int32_t num_sb_gcut;
  const int32_t* sb_start_per_z;    /* [121] */
  const double* sb_data;            /* [num_sb_data] */
  int32_t num_sb_data;
  /* G4HepEmGammaData (G4HepEmGammaData.hh:15-79) */
  int32_t gm_data_per_mat, gm_num_data0, gm_num_data1;
  double gm_emax0, gm_log_emin0, gm_eil_delta0;
  double gm_emax1, gm_log_emin1, gm_eil_delta1;
  double gm_log_emin2, gm_eil_delta2;
  const double* gm_mxsec;           /* [num_mat*gm_data_per_mat] */
  int32_t gm_conv_egrid_size;
  double gm_conv_log_min_ekin, gm_conv_eil_delta;
  const int32_t* gm_conv_start;     /* [num_mat] */
  const double* gm_conv_egrid;      /* [gm_conv_egrid_size] */
  const double* gm_conv_data;
  int32_t num_gm_conv;
} G4HB200Tables;

/* ------------------------------------------------------------------------------------------
 * Track batches: structure of (paired) arrays.  Every double group is an array of n
 * {a,b} pairs (16 B per track, one 128-bit load per thread, a warp reads 512 contiguous bytes).
 * The same struct describes a host staging batch (pointers into pinned host memory) and a
 * device batch; which one is meant is stated per entry point.  A NULL group is not touched.
 *
 * flags bits (G4HepEmTrack/G4HepEmMSCTrackData booleans): */
#define G4HB200_F_POSITRON 0x01u        /* charge > 0 (G4HepEmTrack::fCharge) */
#define G4HB200_F_ON_BOUNDARY 0x02u     /* G4HepEmTrack::fOnBoundary */
#define G4HB200_F_MSC_FIRST_STEP 0x04u  /* G4HepEmMSCTrackData::fIsFirstStep */
#define G4HB200_F_MSC_ACTIVE 0x08u      /* fIsActive */
#define G4HB200_F_MSC_DISPLACE 0x10u    /* fIsDisplace */
#define G4HB200_F_MSC_NO_SCATTER 0x20u  /* fIsNoScatteringInMSC */
#define G4HB200_F_GAUSS_CACHED 0x40u    /* G4HepEmRandomEngine::fIsGauss */
#define G4HB200_F_WDT_ON 0x80u          /* gamma, stepping loop: under Woodcock tracking (isWDTOn of TrackGamma) */
#define G4HB200_F_MSC_SUBSTEP 0x100u    /* e-/e+, stepping loop: between two MSC sub-steps of one step (continueStepping of
                                           TrackElectron, G4HepEmTrackingManager.cc:447-597) */

/* e-/e+ state: G4HepEmElectronTrack (G4HepEmRun/include/G4HepEmElectronTrack.hh:20-93) */
typedef struct G4HB200ElectronBatch {
  int64_t n;
  /* persistent between steps (read+written every step) */
  double* ekin_logekin;    /* {fEKin, fLogEKin (>99: not cached, G4HepEmTrack.hh:95-100)} */
  double* dirx_diry;       /* fDirection[0,1] */
  double* dirz_safety;     /* fDirection[2], fSafety */
  double* nia01;           /* fNumIALeft[0,1] (ioni, brem) */
  double* nia23;           /* fNumIALeft[2,3] (annihilation, lepto-nuclear) */
  double* msc_irange_dynrf;/* fInitialRange, fDynamicRangeFactor */
  double* msc_tlimmin_gauss;/* fTlimitMin, cached Gaussian variate (G4HepEmRandomEngine::fGauss) */
  int32_t* meta;           /* int4 per track: {fMCIndex, flags, fID, rng draw counter} */
  /* step results */
  double* gstep_pstep;     /* fGStepLength, fPStepLength */
  double* edep_dispx;      /* fEDeposit, MSC fDisplacement[0] */
  double* dispy_dispz;     /* MSC fDisplacement[1,2] */
  int32_t* winner;         /* fPIndxWon */
  /* HowFar -> Perform hand-over (only used when the two run as separate launches) */
  double* mfp01;           /* fMFPs[0,1] */
  double* mfp23;           /* fMFPs[2,3] */
  double* range_lambtr1;   /* fRange, MSC fLambtr1 */
  double* tstep_zpath;     /* MSC fTrueStepLength, fZPathLength */
  double* par12;           /* MSC fPar1, fPar2 */
  double* par3_pad;        /* MSC fPar3, unused */
  /* G4HepEmElectronTrack::fPreStepEKin, fPreStepLogEKin (G4HepEmElectronTrack.hh:57-72): state between the track-level
   * calls only (g4hb200_electron_track_op); may be NULL for the HowFar / Perform / step entry points, which keep
   * the pre-step energy in their own workspace */
  double* prestep;
} G4HB200ElectronBatch;

/* gamma state: G4HepEmGammaTrack (G4HepEmRun/include/G4HepEmGammaTrack.hh:17-46) */
typedef struct G4HB200GammaBatch {
  int64_t n;
  double* ekin_logekin; /* {fEKin, fLogEKin} */
  double* dirx_diry;
  double* dirz_nia0;    /* fDirection[2], fNumIALeft[0] */
  int32_t* meta;        /* int4: {fMCIndex, flags, fID, rng draw counter} */
  double* gstep_mfp0;   /* fGStepLength, fMFPs[0] */
  double* edep_pemxsec; /* fEDeposit, fPEmxSec */
  int32_t* winner;      /* fPIndxWon */
} G4HB200GammaBatch;

/* secondaries: what G4HepEmTLData::AddSecondary{Electron,Gamma}Track hands back
 * (G4HepEmRun/include/G4HepEmTLData.hh:52-82), as an append-only queue. */
#define G4HB200_SEC_ELECTRON 0
#define G4HB200_SEC_POSITRON 1
#define G4HB200_SEC_GAMMA 2
typedef struct G4HB200SecondaryQueue {
  int64_t capacity;
  double* dirx_diry;  /* [capacity] pairs */
  double* dirz_ekin;  /* [capacity] pairs */
  int32_t* parent_kind; /* int2: {parent fID, G4HB200_SEC_*}; slot order within a parent is preserved */
  int32_t* parent_slot; /* int2: {index of the parent track in its batch, 0/1 = first/second secondary} */
  int32_t* count;     /* [1] number of valid entries */
  int32_t parent_base; /* added to the parent index written by the kernels (a batch processed in chunks); normally 0 */
} G4HB200SecondaryQueue;

typedef struct G4HB200 G4HB200; /* opaque handle: device arena with the flattened tables */

/* ---- life cycle ------------------------------------------------------------------------- */
/* Flatten + upload the tables to `device`.  Replaces InitG4HepEmData/CopyG4HepEmDataToGPU. */
int g4hb200_create(const G4HB200Tables* tables, int device, G4HB200** out);
int g4hb200_destroy(G4HB200* h);
const char* g4hb200_last_error(void);
int g4hb200_device_count(void);

/* Device batch management (library-owned device memory). */
int g4hb200_electron_batch_alloc(G4HB200* h, int64_t capacity, G4HB200ElectronBatch* out_dev);
int g4hb200_electron_batch_free(G4HB200* h, G4HB200ElectronBatch* dev);
int g4hb200_gamma_batch_alloc(G4HB200* h, int64_t capacity, G4HB200GammaBatch* out_dev);
int g4hb200_gamma_batch_free(G4HB200* h, G4HB200GammaBatch* dev);
int g4hb200_secondary_queue_alloc(G4HB200* h, int64_t capacity, G4HB200SecondaryQueue* out_dev);
int g4hb200_secondary_queue_free(G4HB200* h, G4HB200SecondaryQueue* dev);
int g4hb200_secondary_queue_reset(G4HB200* h, G4HB200SecondaryQueue* dev, void* stream);
/* host <-> device copies of the non-NULL groups (n tracks), asynchronous on `stream` */
int g4hb200_electron_batch_upload(G4HB200* h, const G4HB200ElectronBatch* host, G4HB200ElectronBatch* dev, void* stream);
int g4hb200_electron_batch_download(G4HB200* h, const G4HB200ElectronBatch* dev, G4HB200ElectronBatch* host, void* stream);
int g4hb200_gamma_batch_upload(G4HB200* h, const G4HB200GammaBatch* host, G4HB200GammaBatch* dev, void* stream);
int g4hb200_gamma_batch_download(G4HB200* h, const G4HB200GammaBatch* dev, G4HB200GammaBatch* host, void* stream);
int g4hb200_secondary_queue_download(G4HB200* h, const G4HB200SecondaryQueue* dev, G4HB200SecondaryQueue* host, void* stream);
int g4hb200_sync(G4HB200* h, void* stream);
/* plain device scratch for the callers of the track-level entry points (flag / uniform arrays): no CUDA header needed on
 * the host side.  to_device: 1 = host -> device, 0 = device -> host; asynchronous on `stream` */
int g4hb200_device_alloc(G4HB200* h, size_t bytes, void** out);
int g4hb200_device_free(G4HB200* h, void* p);
int g4hb200_memcpy(G4HB200* h, void* dst, const void* src, size_t bytes, int to_device, void* stream);

/* ---- table look-ups (BASELINE config 1) ---------------------------------------------------
 * One launch evaluates, per track, G4HepEmElectronManager::GetRestRange, GetRestDEDX,
 * GetInvRange(range), GetRestMacXSec(ioni), GetRestMacXSec(brem), GetMacXSecNuclear,
 * GetTransportMFP (G4HepEmElectronManager.icc:486-582).  All pointers are device pointers;
 * out is 7 arrays of n doubles: out[k*n + i]. is_electron selects e- / e+ tables. */
int g4hb200_electron_lookups(G4HB200* h, int64_t n, const int32_t* imc, const double* ekin, const double* logekin,
                             int is_electron, double* out, void* stream);
/* The same seven look-ups in single precision on a float copy of the tables (float device arrays in and out):
 * |f32 - f64| <= 2e-5 |f64| + 1e-6 max|f64| per output over a batch (measured on the configs[0] inputs: 4e-6 relative;
 * 2.5e-8 of the maximum right above a production threshold) for couples whose table entries and their squares fit in
 * single precision (everything but the vacuum couple, whose ranges exceed 1e19 mm); 1.9x the FP64 rate
 * (3.9e10 look-up sets/s on a B200) -- a stated bound, not bit identity; the FP64
 * entry point above is the drop-in.
 * (The reference's authors note that parts of this path "could probably be computed in float",
 * G4HepEmElectronInteractionUMSC.icc:296,312.) */
int g4hb200_electron_lookups_f32(G4HB200* h, int64_t n, const int32_t* imc, const float* ekin, const float* logekin,
                                 int is_electron, float* out, void* stream);
/* G4HepEmElectronManager::GetRestMacXSecForStepping (ioni, brem), GetMacXSecNuclearForStepping,
 * ComputeMacXsecAnnihilationForStepping (.icc:544-599): out[k*n+i], k = 0..3 */
int g4hb200_electron_stepping_xsecs(G4HB200* h, int64_t n, const int32_t* imc, const double* ekin, const double* logekin,
                                    int is_electron, double* out, void* stream);
/* G4HepEmGammaManager::GetTotalMacXSec + SampleInteraction with a caller supplied uniform
 * (G4HepEmGammaManager.icc:108-152,173-219): out_mxsec[n], out_pid[n] */
int g4hb200_gamma_lookups(G4HB200* h, int64_t n, const int32_t* imc, const double* ekin, const double* logekin,
                          const double* urnd, double* out_mxsec, int32_t* out_pid, void* stream);
/* Target element selectors with a caller supplied uniform: kind 0 = brem SB, 1 = brem RB
 * (G4HepEmElectronInteractionBrem.icc:266-296), 2 = conversion (G4HepEmGammaInteractionConversion.icc:151-176;
 * imc is then the material index).  out_elem[n]. */
int g4hb200_select_target_element(G4HB200* h, int kind, int is_electron, int64_t n, const int32_t* imc,
                                  const double* ekin, const double* logekin, const double* urnd,
                                  int32_t* out_elem, void* stream);
/* device VDT log / exp (G4HepEmLog.hh:228-263, G4HepEmExp.hh:182-223), for bit-exactness tests */
int g4hb200_vdt_log_exp(G4HB200* h, int64_t n, const double* x, double* out_log, double* out_exp, void* stream);
/* the counter based uniform stream: out[i*ndraw + j] = draw j of track id[i] */
int g4hb200_rng_uniforms(G4HB200* h, uint64_t seed, int64_t n, const int32_t* track_id, int32_t ndraw,
                         double* out, void* stream);

/* ---- stepping entry points (device batches) ------------------------------------------------
 * seed: global seed of the counter based stream; the per-track key is (seed, fID) and the
 * draw counter lives in meta[3], so results do not depend on batch order or launch shape. */
/* G4HepEmElectronManager::HowFar(data, pars, tlData) (.icc:35-45): resample fNumIALeft,
 * HowFarToDiscreteInteraction (.icc:48-101) and HowFarToMSC (.icc:103-164). */
int g4hb200_electron_howfar(G4HB200* h, G4HB200ElectronBatch* dev, uint64_t seed, void* stream);
/* G4HepEmElectronManager::Perform(data, pars, tlData) (.icc:461-483): continuous part, then e+
 * annihilation at rest or the discrete interaction; secondaries are appended to `sec`. */
int g4hb200_electron_perform(G4HB200* h, G4HB200ElectronBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream);
/* HowFar + (geometry accepts the proposed step, fOnBoundary unchanged) + Perform in one pass:
 * the hand-over groups stay in registers. */
int g4hb200_electron_step(G4HB200* h, G4HB200ElectronBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream);
/* G4HepEmGammaManager::HowFar (.icc:27-48) */
int g4hb200_gamma_howfar(G4HB200* h, G4HB200GammaBatch* dev, uint64_t seed, void* stream);
/* G4HepEmGammaManager::SelectInteraction (.icc:173-219; skipped when on boundary, as the
 * callers do: G4HepEmTrackingManager.cc:1092-1108) followed by Perform (.icc:54-94). */
int g4hb200_gamma_perform(G4HB200* h, G4HB200GammaBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream);
int g4hb200_gamma_step(G4HB200* h, G4HB200GammaBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream);

/* ---- track-level entry points: the pieces of a step, one launch each -------------------------------------------------
 * The production caller of the reference does not use the two-call protocol: G4HepEmTrackingManager::TrackElectron
 * (G4HepEm/G4HepEm/src/G4HepEmTrackingManager.cc:428-665) calls the static pieces of G4HepEmElectronManager
 * (G4HepEmRun/include/G4HepEmElectronManager.hh:90-206) one by one with the geometry step and the MSC sub-step loop in
 * between, TrackGamma (.cc:985-1140) those of G4HepEmGammaManager (G4HepEmGammaManager.hh:32-55).  One call applies
 * ONE of them, in place, to every track of the device batch -- all groups of the batch, `prestep` included, are the
 * state between calls (what the reference keeps in the G4HepEmElectronTrack / G4HepEmGammaTrack object).
 * out_flag (device, int32[n], may be NULL): the bool the reference's function returns, for the ops that return one. */
#define G4HB200_OP_HOWFAR_DISCRETE 0    /* HowFarToDiscreteInteraction(data, pars, elTrack)          .hh:90   .icc:48-101  */
#define G4HB200_OP_HOWFAR_MSC 1         /* HowFarToMSC(data, pars, elTrack, rng)                     .hh:108  .icc:103-164 */
#define G4HB200_OP_UPDATE_PSTEP 2       /* UpdatePStepLength(elTrack)                                .hh:133  .icc:171-203 */
#define G4HB200_OP_UPDATE_NIA 3         /* UpdateNumIALeft(elTrack)                                  .hh:140  .icc:205-214 */
#define G4HB200_OP_MEAN_ELOSS 4         /* ApplyMeanEnergyLoss(data, pars, elTrack) -> stopped       .hh:149  .icc:216-259 */
#define G4HB200_OP_SAMPLE_MSC 5         /* SampleMSC(data, pars, elTrack, rng)                       .hh:158  .icc:261-322 */
#define G4HB200_OP_LOSS_FLUCT 6         /* SampleLossFluctuations(data, pars, elTrack, rng) -> stopped .hh:167 .icc:324-368 */
#define G4HB200_OP_DISCRETE 7           /* PerformDiscrete(data, pars, tlData)                       .hh:206  .icc:425-459 */
#define G4HB200_OP_ANNIHILATE_AT_REST 8 /* G4HepEmPositronInteractionAnnihilation::Perform(tlData, true)  (...Annihilation.icc:15-50) */
#define G4HB200_OP_PERFORM_CONTINUOUS 9 /* PerformContinuous(data, pars, elTrack, rng) -> stopped    .hh:184  .icc:375-405 */
#define G4HB200_OP_RESAMPLE_NIA 10      /* fNumIALeft[ip] = -log(u) where <= 0: the loop of HowFar (.icc:39-43) and of
                                           TrackElectron (G4HepEmTrackingManager.cc:430-434) */
/* sec: required for the ops that create secondaries (DISCRETE, ANNIHILATE_AT_REST), may be NULL otherwise */
int g4hb200_electron_track_op(G4HB200* h, int op, G4HB200ElectronBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed,
                              int32_t* out_flag, void* stream);
/* CheckDelta(data, track, rand) with caller supplied uniforms (.hh:195, .icc:408-423): out_flag[i] = 1 for a delta interaction */
int g4hb200_electron_check_delta(G4HB200* h, G4HB200ElectronBatch* dev, const double* urnd, int32_t* out_flag, void* stream);
#define G4HB200_GOP_HOWFAR_TRACK 0       /* HowFar(data, pars, gammaTrack): no resampling of fNumIALeft  GammaManager.hh:34 .icc:38-48 */
#define G4HB200_GOP_UPDATE_NIA 1         /* UpdateNumIALeft(track)                                        .hh:41 .icc:97-105  */
#define G4HB200_GOP_SELECT_INTERACTION 2 /* SelectInteraction(data, tlData)                               .hh:51 .icc:173-177 */
#define G4HB200_GOP_PERFORM_SELECTED 3   /* Perform(data, pars, tlData) for an interaction selected before .hh:39 .icc:54-94   */
int g4hb200_gamma_track_op(G4HB200* h, int op, G4HB200GammaBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream);

/* ---- host-buffer entry points (the call a host application makes) ---------------------------
 * Upload the persistent groups of `host`, run the fused step, download state + results and the
 * secondaries; everything asynchronous on the handle's internal stream, then synchronised. */
int g4hb200_electron_step_host(G4HB200* h, G4HB200ElectronBatch* host, G4HB200SecondaryQueue* host_sec, uint64_t seed);
int g4hb200_gamma_step_host(G4HB200* h, G4HB200GammaBatch* host, G4HB200SecondaryQueue* host_sec, uint64_t seed);

/* number of kernels launched through this handle so far (for the bench's gpu_launches) */
int64_t g4hb200_launch_count(const G4HB200* h);

/* ---- per-kernel device timing of the e-/e+ step pipeline -------------------------------------------
 * With timing enabled every pipelined g4hb200_electron_step / _perform call records CUDA events on its
 * stream around each of its kernels.  g4hb200_kernel_times synchronises those events and returns, per
 * pipeline stage k < G4HB200_NUM_STAGES, the summed device time in ms (ms_sum[k]), the number of launches
 * (launches[k]) and the summed number of tracks the stage processed (items[k], from the queue counters);
 * the sums restart at every call.  Stage names: g4hb200_stage_name(k). */
#define G4HB200_NUM_STAGES 20 /* capacity; unused slots have an empty name and zero launches */
int g4hb200_set_kernel_timing(G4HB200* h, int enable);
int g4hb200_kernel_times(G4HB200* h, double* ms_sum, int64_t* launches, int64_t* items);
const char* g4hb200_stage_name(int k);

/* ---- offered variant: multiple scattering in single precision ----------------------------------------------------
 * SURVEY.md 8(f) rank 4; the reference's authors: "all these could probably be computed in float"
 * (G4HepEm/G4HepEmRun/include/G4HepEmElectronInteractionUMSC.icc:296,312).  bits = 32: G4HepEmElectronManager::SampleMSC
 * (G4HepEmElectronManager.icc:261-322, UMSC.icc:129-357) inside g4hb200_electron_step / _perform and the stepping loops
 * computes the Urban model parameters and samples the polar angle in float (csrc/g4h_msc_f32.cuh); bits = 64 (the
 * default): the drop-in, bit-exact path.  With 32 everything but the post-step direction and the MSC displacement is
 * still identical to the reference for all but a few tracks in a million (those whose model regime is decided within
 * float rounding of a threshold consume a different number of uniforms); direction and displacement agree with the FP64
 * path within the bound stated and tested in tests/test_msc_f32.py.  Returns G4HB200_EINVAL for any other value. */
int g4hb200_set_msc_precision(G4HB200* h, int bits);

/* ---- stepping loop over a slab calorimeter (BASELINE configs[4]) -----------------------------------------------
 * The loop the reference's callers run around the managers -- G4HepEmTrackingManager::TrackElectron / TrackGamma
 * (G4HepEm/G4HepEm/src/G4HepEmTrackingManager.cc:428-705,985-1140) inside the TestEm3 sampling calorimeter
 * (apps/examples/TestEm3/src/DetectorConstruction.cc:281-384: num_layers x {absorbers} stacked along x, square
 * cross section) with the per-(layer, absorber) energy-deposit score of apps/examples/TestEm3/src/SteppingAction.cc:83-84 --
 * for num_primaries showers at once, breadth first, entirely on the device: HowFar, geometry step, Perform, MSC
 * displacement, scoring, relocation, secondaries -> new tracks, until no track is left (or max_steps).
 * Primaries start on the front face (x = -thickness/2) along +x.  A track leaving the calorimeter is dropped
 * and its kinetic energy booked as leakage.  Per-track uniform streams do not depend on the batching (a secondary's
 * stream is derived from its parent's), so edep_out is the same for any sharding of the primaries over GPUs up to
 * floating point summation order; ranks sum their histograms with one allreduce (g4hepem_b200/sharding.py). */
typedef struct G4HB200SlabGeometry {
  int32_t num_layers;           /* TestEm3 fNbOfLayers (50) */
  int32_t num_absorbers;        /* fNbOfAbsor (2), at most 4 */
  double absorber_thickness[4]; /* mm: 2.3 (Pb), 5.7 (lAr) */
  int32_t absorber_couple[4];   /* material-cuts couple index of each absorber */
  double half_yz;               /* half of fCalorSizeYZ (200 mm) */
  /* Woodcock tracking of gammas with the calorimeter as the tracking region (G4HepEmWoodcockHelper,
   * G4HepEm/G4HepEm/src/G4HepEmWoodcockHelper.cc:105-300; TestEm3 turns it on for its calorimeter region,
   * apps/examples/TestEm3/src/PhysListHepEmTracking.cc:42): a gamma above woodcock_ekin_min steps with the
   * cross section of woodcock_couple (the couple of the densest absorber) across the slab boundaries and
   * interacts where it stops with probability sigma(local material) / sigma(woodcock material). */
  int32_t woodcock_on;          /* 0: every gamma step ends on the slab boundaries (no Woodcock tracking) */
  int32_t woodcock_couple;      /* G4HepEmWoodcockHelper::fWDTHepEmIMC */
  double woodcock_ekin_min;     /* MeV; G4HepEmConfig::fWDTEnergyLimit (0.2) */
} G4HB200SlabGeometry;

typedef struct G4HB200ShowerStats {
  int64_t num_steps;            /* iterations of the loop */
  int64_t electron_track_steps; /* sum over iterations of the e-/e+ population */
  int64_t gamma_track_steps;
  int64_t secondaries;          /* tracks created */
  int64_t peak_electrons, peak_gammas;
  double leak_electron, leak_gamma; /* kinetic energy [MeV] that left the calorimeter */
  double device_ms;             /* CUDA-event time of the whole loop on the handle's stream */
  int64_t kernel_launches;
  /* what was still alive when the loop stopped on max_steps (all zero for a loop that ran to the end): the deposits and
   * the leakage are then incomplete by that much kinetic energy */
  int64_t remaining_electrons, remaining_gammas;
  double remaining_ekin;        /* MeV */
} G4HB200ShowerStats;

/* primary_kind: G4HB200_SEC_ELECTRON / _POSITRON / _GAMMA.  first_track_id: id of the first primary (ids are consecutive:
 * give every rank its own range).  capacity: tracks per store (e-/e+ and gamma each); G4HB200_ECAPACITY if exceeded.
 * max_steps: 0 = until no track is left (at most 1 000 000 iterations: G4HB200_ECUDA-free safety net against a track that
 * never ends); > 0: stop after that many iterations and report what is left in stats->remaining_*.
 * edep_out: host array [num_layers * num_absorbers], MeV.
 * Streams of secondaries: a secondary's (track id, first draw) is a Philox hash of (parent id, parent draw counter, slot):
 * 32 + 29 random bits.  Ids are NOT guaranteed unique: among N tracks about N^2 / 2^33 pairs share an id, and two
 * tracks with the same id consume overlapping stretches of one stream only if their first draws lie within their
 * lifetimes' draws of each other (~1e3 of 2^29) -- for the 1e8 tracks of 4096 x 10 GeV showers that is a few pairs of
 * tracks in 1e8, each sharing a few hundred uniforms.  The deposits do not depend on batching or sharding either way. */
int g4hb200_shower_run(G4HB200* h, const G4HB200SlabGeometry* geom, int64_t num_primaries, int32_t primary_kind,
                       double primary_ekin, uint64_t seed, int32_t first_track_id, int64_t capacity, int32_t max_steps,
                       double* edep_out, G4HB200ShowerStats* stats);

/* ---- mixed stepping batches without geometry (BASELINE configs[3]) -------------------------------------------------
 * num_electrons e-/e+ (half each) and num_gammas gammas, generated on the device in queue order (particle, then
 * couple): E log-uniform in [emin, emax] MeV, couples uniform over the table set, isotropic.  num_steps consecutive
 * fused steps (g4hb200_electron_step / g4hb200_gamma_step); after every step the survivors are compacted and the
 * secondaries become tracks of the next step (a secondary keeps its parent's couple).  edep_total: one double, the
 * energy deposited over all steps [MeV].  stats->device_ms times the loop without the generation of the population. */
int g4hb200_mixed_run(G4HB200* h, int64_t num_electrons, int64_t num_gammas, double emin, double emax, uint64_t seed,
                      int64_t capacity, int32_t num_steps, double* edep_total, G4HB200ShowerStats* stats);

#ifdef __cplusplus
}
#endif
#endif /* G4HEPEM_B200_H */
