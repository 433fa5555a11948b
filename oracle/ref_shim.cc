// ref_shim.cc -- TEST INFRASTRUCTURE: batch driver around the UNMODIFIED reference (mnovak42/g4hepem).
//
// Compiled by oracle/Makefile together with the reference's own sources *where they lie* under
// /root/reference (G4HepEmRun/include/*.icc, G4HepEmData/src/*.cc, G4HepEmDataJsonIO/src/*.cc) into
// oracle/_ref/libg4hepem_ref.so.  Nothing of the reference is copied into this repository.
//
// What this file adds on top of the reference:
//   * G4HepEmRandomEngine::flat()/flatArray(), which the reference leaves to the consumer
//     (G4HepEmRun/include/G4HepEmRandomEngine.hh:21-28) -> the counter based stream of g4h_rng_host.h
//   * extern "C" batch entry points with the same batch structs as the product's C-ABI
//     (include/g4hepem_b200.h): every track is unpacked into a real G4HepEmElectronTrack /
//     G4HepEmGammaTrack, pushed through the reference's G4HepEmElectronManager / G4HepEmGammaManager
//     static functions, and packed back.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference use it.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <thread>
#include <vector>

#include "G4HepEmData.hh"
#include "G4HepEmElectronData.hh"
#include "G4HepEmElementData.hh"
#include "G4HepEmGammaData.hh"
#include "G4HepEmMatCutData.hh"
#include "G4HepEmMaterialData.hh"
#include "G4HepEmParameters.hh"
#include "G4HepEmSBTableData.hh"
#include "G4HepEmState.hh"
#include "G4HepEmDataJsonIO.hh"

// the engine's Gauss cache (fIsGauss/fGauss) is private; the batch driver has to persist it per track
#define private public
#include "G4HepEmRandomEngine.hh"
#undef private

#include "G4HepEmTLData.hh"
#include "G4HepEmElectronManager.hh"
#include "G4HepEmGammaManager.hh"
#include "G4HepEmElectronInteractionBrem.hh"
#include "G4HepEmPositronInteractionAnnihilation.hh"
#include "G4HepEmGammaInteractionConversion.hh"
#include "G4HepEmRunUtils.hh"

#include "g4hepem_b200.h"
#include "../g4hepem_b200/host/G4HepEmB200Flatten.hh"
#include "g4h_rng_host.h"

// ---- the injected random stream -------------------------------------------------------------------
double G4HepEmRandomEngine::flat() { return g4h_stream_next(static_cast<G4HStream*>(fObject)); }
void G4HepEmRandomEngine::flatArray(const int size, double* vect) {
  G4HStream* s = static_cast<G4HStream*>(fObject);
  for (int i = 0; i < size; ++i) vect[i] = g4h_stream_next(s);
}

namespace {

struct RefState {
  G4HepEmState* state = nullptr;
  G4HepEmB200FlatTables flat;
};

inline bool Has(uint32_t flags, uint32_t bit) { return (flags & bit) != 0u; }

void UnpackElectron(const G4HB200ElectronBatch* b, int64_t i, G4HepEmElectronTrack& et, G4HepEmRandomEngine& eng,
                    G4HStream& stream, uint64_t seed, bool withHandOver) {
  et.ReSet();
  G4HepEmTrack* t = et.GetTrack();
  const int32_t* meta = b->meta + 4 * i;
  const uint32_t flags = static_cast<uint32_t>(meta[1]);
  t->SetEKin(b->ekin_logekin[2 * i], b->ekin_logekin[2 * i + 1]);
  t->SetDirection(b->dirx_diry[2 * i], b->dirx_diry[2 * i + 1], b->dirz_safety[2 * i]);
  t->SetSafety(b->dirz_safety[2 * i + 1]);
  t->SetNumIALeft(b->nia01[2 * i], 0);
  t->SetNumIALeft(b->nia01[2 * i + 1], 1);
  t->SetNumIALeft(b->nia23[2 * i], 2);
  t->SetNumIALeft(b->nia23[2 * i + 1], 3);
  t->SetMCIndex(meta[0]);
  t->SetCharge(Has(flags, G4HB200_F_POSITRON) ? 1.0 : -1.0);
  t->SetOnBoundary(Has(flags, G4HB200_F_ON_BOUNDARY));
  t->SetID(meta[2]);
  G4HepEmMSCTrackData* msc = et.GetMSCTrackData();
  msc->fInitialRange        = b->msc_irange_dynrf[2 * i];
  msc->fDynamicRangeFactor  = b->msc_irange_dynrf[2 * i + 1];
  msc->fTlimitMin           = b->msc_tlimmin_gauss[2 * i];
  msc->fIsFirstStep         = Has(flags, G4HB200_F_MSC_FIRST_STEP);
  msc->fIsActive            = Has(flags, G4HB200_F_MSC_ACTIVE);
  msc->fIsDisplace          = Has(flags, G4HB200_F_MSC_DISPLACE);
  msc->fIsNoScatteringInMSC = Has(flags, G4HB200_F_MSC_NO_SCATTER);
  stream.seed     = seed;
  stream.track_id = static_cast<uint32_t>(meta[2]);
  stream.draw     = static_cast<uint32_t>(meta[3]);
  eng.fIsGauss = Has(flags, G4HB200_F_GAUSS_CACHED);
  eng.fGauss   = b->msc_tlimmin_gauss[2 * i + 1];
  // a track object persists between the calls: HowFar does not touch the deposit of the previous step
  t->SetEnergyDeposit(b->edep_dispx[2 * i]);
  if (withHandOver) {
    t->SetGStepLength(b->gstep_pstep[2 * i]);
    et.SetPStepLength(b->gstep_pstep[2 * i + 1]);
    msc->SetDisplacement(b->edep_dispx[2 * i + 1], b->dispy_dispz[2 * i], b->dispy_dispz[2 * i + 1]);
    t->SetWinnerProcessIndex(b->winner[i]);
    t->SetMFP(b->mfp01[2 * i], 0);
    t->SetMFP(b->mfp01[2 * i + 1], 1);
    t->SetMFP(b->mfp23[2 * i], 2);
    t->SetMFP(b->mfp23[2 * i + 1], 3);
    et.SetRange(b->range_lambtr1[2 * i]);
    msc->fLambtr1        = b->range_lambtr1[2 * i + 1];
    msc->fTrueStepLength = b->tstep_zpath[2 * i];
    msc->fZPathLength    = b->tstep_zpath[2 * i + 1];
    msc->fPar1           = b->par12[2 * i];
    msc->fPar2           = b->par12[2 * i + 1];
    msc->fPar3           = b->par3_pad[2 * i];
    if (b->prestep != nullptr) et.SetPreStepEKin(b->prestep[2 * i], b->prestep[2 * i + 1]);
  }
}

// fLogEKin is private with a lazy getter; read it without triggering the lazy evaluation
double PeekLogEKin(G4HepEmTrack* t) {
  // layout: fPosition[3], fDirection[3], fEKin, fLogEKin, ... (G4HepEmTrack.hh:244-251)
  double raw[8];
  std::memcpy(raw, static_cast<void*>(t), sizeof(raw));
  return raw[7];
}

void PackElectron(G4HB200ElectronBatch* b, int64_t i, G4HepEmElectronTrack& et, const G4HepEmRandomEngine& eng,
                  const G4HStream& stream) {
  G4HepEmTrack* t = et.GetTrack();
  G4HepEmMSCTrackData* msc = et.GetMSCTrackData();
  int32_t* meta = b->meta + 4 * i;
  b->ekin_logekin[2 * i]     = t->GetEKin();
  b->ekin_logekin[2 * i + 1] = PeekLogEKin(t);
  const double* dir = t->GetDirection();
  b->dirx_diry[2 * i]       = dir[0];
  b->dirx_diry[2 * i + 1]   = dir[1];
  b->dirz_safety[2 * i]     = dir[2];
  b->dirz_safety[2 * i + 1] = t->GetSafety();
  b->nia01[2 * i]     = t->GetNumIALeft(0);
  b->nia01[2 * i + 1] = t->GetNumIALeft(1);
  b->nia23[2 * i]     = t->GetNumIALeft(2);
  b->nia23[2 * i + 1] = t->GetNumIALeft(3);
  b->msc_irange_dynrf[2 * i]      = msc->fInitialRange;
  b->msc_irange_dynrf[2 * i + 1]  = msc->fDynamicRangeFactor;
  b->msc_tlimmin_gauss[2 * i]     = msc->fTlimitMin;
  b->msc_tlimmin_gauss[2 * i + 1] = eng.fGauss;
  uint32_t flags = 0;
  if (t->GetCharge() > 0.0) flags |= G4HB200_F_POSITRON;
  if (t->GetOnBoundary()) flags |= G4HB200_F_ON_BOUNDARY;
  if (msc->fIsFirstStep) flags |= G4HB200_F_MSC_FIRST_STEP;
  if (msc->fIsActive) flags |= G4HB200_F_MSC_ACTIVE;
  if (msc->fIsDisplace) flags |= G4HB200_F_MSC_DISPLACE;
  if (msc->fIsNoScatteringInMSC) flags |= G4HB200_F_MSC_NO_SCATTER;
  if (eng.fIsGauss) flags |= G4HB200_F_GAUSS_CACHED;
  meta[1] = static_cast<int32_t>(flags);
  meta[3] = static_cast<int32_t>(stream.draw);
  b->gstep_pstep[2 * i]     = t->GetGStepLength();
  b->gstep_pstep[2 * i + 1] = et.GetPStepLength();
  const double* disp = msc->GetDisplacement();
  b->edep_dispx[2 * i]      = t->GetEnergyDeposit();
  b->edep_dispx[2 * i + 1]  = disp[0];
  b->dispy_dispz[2 * i]     = disp[1];
  b->dispy_dispz[2 * i + 1] = disp[2];
  b->winner[i] = t->GetWinnerProcessIndex();
  if (b->mfp01 != nullptr) {
    b->mfp01[2 * i]             = t->GetMFP(0);
    b->mfp01[2 * i + 1]         = t->GetMFP(1);
    b->mfp23[2 * i]             = t->GetMFP(2);
    b->mfp23[2 * i + 1]         = t->GetMFP(3);
    b->range_lambtr1[2 * i]     = et.GetRange();
    b->range_lambtr1[2 * i + 1] = msc->fLambtr1;
    b->tstep_zpath[2 * i]       = msc->fTrueStepLength;
    b->tstep_zpath[2 * i + 1]   = msc->fZPathLength;
    b->par12[2 * i]             = msc->fPar1;
    b->par12[2 * i + 1]         = msc->fPar2;
    b->par3_pad[2 * i]          = msc->fPar3;
    b->par3_pad[2 * i + 1]      = 0.0;
    if (b->prestep != nullptr) {
      b->prestep[2 * i]     = et.GetPreStepEKin();
      b->prestep[2 * i + 1] = et.GetPreStepLogEKin();
    }
  }
}

void UnpackGamma(const G4HB200GammaBatch* b, int64_t i, G4HepEmGammaTrack& gt, G4HStream& stream, uint64_t seed,
                 bool withHandOver) {
  gt.ReSet();
  G4HepEmTrack* t = gt.GetTrack();
  const int32_t* meta = b->meta + 4 * i;
  const uint32_t flags = static_cast<uint32_t>(meta[1]);
  t->SetEKin(b->ekin_logekin[2 * i], b->ekin_logekin[2 * i + 1]);
  t->SetDirection(b->dirx_diry[2 * i], b->dirx_diry[2 * i + 1], b->dirz_nia0[2 * i]);
  t->SetNumIALeft(b->dirz_nia0[2 * i + 1], 0);
  t->SetMCIndex(meta[0]);
  t->SetOnBoundary(Has(flags, G4HB200_F_ON_BOUNDARY));
  t->SetID(meta[2]);
  stream.seed     = seed;
  stream.track_id = static_cast<uint32_t>(meta[2]);
  stream.draw     = static_cast<uint32_t>(meta[3]);
  // persistent between the calls (a stale winner index survives a boundary-limited step in the reference)
  t->SetEnergyDeposit(b->edep_pemxsec[2 * i]);
  gt.SetPEmxSec(b->edep_pemxsec[2 * i + 1]);
  t->SetWinnerProcessIndex(b->winner[i]);
  if (withHandOver) {
    t->SetGStepLength(b->gstep_mfp0[2 * i]);
    t->SetMFP(b->gstep_mfp0[2 * i + 1], 0);
  }
}

void PackGamma(G4HB200GammaBatch* b, int64_t i, G4HepEmGammaTrack& gt, const G4HStream& stream) {
  G4HepEmTrack* t = gt.GetTrack();
  int32_t* meta = b->meta + 4 * i;
  b->ekin_logekin[2 * i]     = t->GetEKin();
  b->ekin_logekin[2 * i + 1] = PeekLogEKin(t);
  const double* dir = t->GetDirection();
  b->dirx_diry[2 * i]     = dir[0];
  b->dirx_diry[2 * i + 1] = dir[1];
  b->dirz_nia0[2 * i]     = dir[2];
  b->dirz_nia0[2 * i + 1] = t->GetNumIALeft(0);
  meta[3] = static_cast<int32_t>(stream.draw);
  b->gstep_mfp0[2 * i]       = t->GetGStepLength();
  b->gstep_mfp0[2 * i + 1]   = t->GetMFP(0);
  b->edep_pemxsec[2 * i]     = t->GetEnergyDeposit();
  b->edep_pemxsec[2 * i + 1] = gt.GetPEmxSec();
  b->winner[i] = t->GetWinnerProcessIndex();
}

struct SecRecord {
  double dir[3];
  double ekin;
  int32_t parentId, kind, parentIndex, slot;
};

void CollectSecondaries(G4HepEmTLData& tl, int64_t parentIndex, std::vector<SecRecord>& out) {
  int slot = 0;
  const int ne = static_cast<int>(tl.GetNumSecondaryElectronTrack());
  for (int k = 0; k < ne; ++k) {
    G4HepEmTrack* s = tl.GetSecondaryElectronTrack(k)->GetTrack();
    SecRecord r;
    std::memcpy(r.dir, s->GetDirection(), 3 * sizeof(double));
    r.ekin = s->GetEKin();
    r.parentId = s->GetParentID();
    r.kind = s->GetCharge() > 0.0 ? G4HB200_SEC_POSITRON : G4HB200_SEC_ELECTRON;
    r.parentIndex = static_cast<int32_t>(parentIndex);
    r.slot = slot++;
    out.push_back(r);
  }
  const int ng = static_cast<int>(tl.GetNumSecondaryGammaTrack());
  for (int k = 0; k < ng; ++k) {
    G4HepEmTrack* s = tl.GetSecondaryGammaTrack(k)->GetTrack();
    SecRecord r;
    std::memcpy(r.dir, s->GetDirection(), 3 * sizeof(double));
    r.ekin = s->GetEKin();
    r.parentId = s->GetParentID();
    r.kind = G4HB200_SEC_GAMMA;
    r.parentIndex = static_cast<int32_t>(parentIndex);
    r.slot = slot++;
    out.push_back(r);
  }
  tl.ResetNumSecondaryElectronTrack();
  tl.ResetNumSecondaryGammaTrack();
}

int AppendSecondaries(G4HB200SecondaryQueue* q, const std::vector<SecRecord>& recs) {
  if (q == nullptr) return 0;
  int64_t n = q->count[0];
  for (const SecRecord& r : recs) {
    if (n >= q->capacity) return G4HB200_ECAPACITY;
    q->dirx_diry[2 * n]       = r.dir[0];
    q->dirx_diry[2 * n + 1]   = r.dir[1];
    q->dirz_ekin[2 * n]       = r.dir[2];
    q->dirz_ekin[2 * n + 1]   = r.ekin;
    q->parent_kind[2 * n]     = r.parentId;
    q->parent_kind[2 * n + 1] = r.kind;
    q->parent_slot[2 * n]     = r.parentIndex;
    q->parent_slot[2 * n + 1] = r.slot;
    ++n;
  }
  q->count[0] = static_cast<int32_t>(n);
  return 0;
}

enum class Mode { kHowFar, kPerform, kStep };

void ElectronRange(RefState* rs, G4HB200ElectronBatch* b, uint64_t seed, Mode mode, int64_t lo, int64_t hi,
                   std::vector<SecRecord>* secs) {
  G4HepEmData* data       = rs->state->fData;
  G4HepEmParameters* pars = rs->state->fParameters;
  G4HepEmTLData tl;
  G4HStream stream{0, 0, 0};
  G4HepEmRandomEngine eng(&stream);
  tl.SetRandomEngine(&eng);
  G4HepEmElectronTrack* et = tl.GetPrimaryElectronTrack();
  for (int64_t i = lo; i < hi; ++i) {
    UnpackElectron(b, i, *et, eng, stream, seed, mode == Mode::kPerform);
    if (mode != Mode::kPerform) {
      G4HepEmElectronManager::HowFar(data, pars, &tl);
    }
    if (mode != Mode::kHowFar) {
      // the geometry stub of the fused step: the proposed geometrical step is accepted as it is
      G4HepEmElectronManager::Perform(data, pars, &tl);
      CollectSecondaries(tl, i, *secs);
    }
    PackElectron(b, i, *et, eng, stream);
  }
}

void GammaRange(RefState* rs, G4HB200GammaBatch* b, uint64_t seed, Mode mode, int64_t lo, int64_t hi,
                std::vector<SecRecord>* secs) {
  G4HepEmData* data       = rs->state->fData;
  G4HepEmParameters* pars = rs->state->fParameters;
  G4HepEmTLData tl;
  G4HStream stream{0, 0, 0};
  G4HepEmRandomEngine eng(&stream);
  tl.SetRandomEngine(&eng);
  G4HepEmGammaTrack* gt = tl.GetPrimaryGammaTrack();
  for (int64_t i = lo; i < hi; ++i) {
    UnpackGamma(b, i, *gt, stream, seed, mode == Mode::kPerform);
    if (mode != Mode::kPerform) {
      G4HepEmGammaManager::HowFar(data, pars, &tl);
    }
    if (mode != Mode::kHowFar) {
      // callers select the interaction only when the step was not limited by a boundary
      // (apps/examples/TestEm3/src/G4HepEmProcess.cc:166-181, G4HepEmTrackingManager.cc:1092-1108)
      if (!gt->GetTrack()->GetOnBoundary()) {
        G4HepEmGammaManager::SelectInteraction(data, &tl);
      }
      G4HepEmGammaManager::Perform(data, pars, &tl);
      CollectSecondaries(tl, i, *secs);
    }
    PackGamma(b, i, *gt, stream);
  }
}

// ---- the track-level statics, one at a time (the pieces G4HepEmTrackingManager::TrackElectron / TrackGamma call) --------------
// op codes: include/g4hepem_b200.h (G4HB200_OP_*, G4HB200_GOP_*)
void ElectronOpRange(RefState* rs, int op, G4HB200ElectronBatch* b, uint64_t seed, int32_t* flags, std::vector<SecRecord>* secs) {
  G4HepEmData* data       = rs->state->fData;
  G4HepEmParameters* pars = rs->state->fParameters;
  G4HepEmTLData tl;
  G4HStream stream{0, 0, 0};
  G4HepEmRandomEngine eng(&stream);
  tl.SetRandomEngine(&eng);
  G4HepEmElectronTrack* et = tl.GetPrimaryElectronTrack();
  for (int64_t i = 0; i < b->n; ++i) {
    UnpackElectron(b, i, *et, eng, stream, seed, true);
    G4HepEmTrack* t = et->GetTrack();
    bool result = false;
    switch (op) {
      case G4HB200_OP_HOWFAR_DISCRETE: G4HepEmElectronManager::HowFarToDiscreteInteraction(data, pars, et); break;
      case G4HB200_OP_HOWFAR_MSC: G4HepEmElectronManager::HowFarToMSC(data, pars, et, &eng); break;
      case G4HB200_OP_UPDATE_PSTEP: G4HepEmElectronManager::UpdatePStepLength(et); break;
      case G4HB200_OP_UPDATE_NIA: G4HepEmElectronManager::UpdateNumIALeft(et); break;
      case G4HB200_OP_MEAN_ELOSS: result = G4HepEmElectronManager::ApplyMeanEnergyLoss(data, pars, et); break;
      case G4HB200_OP_SAMPLE_MSC: G4HepEmElectronManager::SampleMSC(data, pars, et, &eng); break;
      case G4HB200_OP_LOSS_FLUCT: result = G4HepEmElectronManager::SampleLossFluctuations(data, pars, et, &eng); break;
      case G4HB200_OP_DISCRETE: G4HepEmElectronManager::PerformDiscrete(data, pars, &tl); break;
      case G4HB200_OP_ANNIHILATE_AT_REST: G4HepEmPositronInteractionAnnihilation::Perform(&tl, true); break;
      case G4HB200_OP_PERFORM_CONTINUOUS: result = G4HepEmElectronManager::PerformContinuous(data, pars, et, &eng); break;
      case G4HB200_OP_RESAMPLE_NIA:
        // the loop at the top of HowFar (G4HepEmElectronManager.icc:39-43) and of TrackElectron (G4HepEmTrackingManager.cc:430-434)
        for (int ip = 0; ip < 4; ++ip) {
          if (t->GetNumIALeft(ip) <= 0.) t->SetNumIALeft(-G4HepEmLog(eng.flat()), ip);
        }
        break;
      default: break;
    }
    if (flags != nullptr && (op == G4HB200_OP_MEAN_ELOSS || op == G4HB200_OP_LOSS_FLUCT || op == G4HB200_OP_PERFORM_CONTINUOUS))
      flags[i] = result ? 1 : 0;
    if (secs != nullptr) CollectSecondaries(tl, i, *secs);
    PackElectron(b, i, *et, eng, stream);
  }
}

void GammaOpRange(RefState* rs, int op, G4HB200GammaBatch* b, uint64_t seed, std::vector<SecRecord>* secs) {
  G4HepEmData* data       = rs->state->fData;
  G4HepEmParameters* pars = rs->state->fParameters;
  G4HepEmTLData tl;
  G4HStream stream{0, 0, 0};
  G4HepEmRandomEngine eng(&stream);
  tl.SetRandomEngine(&eng);
  G4HepEmGammaTrack* gt = tl.GetPrimaryGammaTrack();
  for (int64_t i = 0; i < b->n; ++i) {
    UnpackGamma(b, i, *gt, stream, seed, true);
    switch (op) {
      case G4HB200_GOP_HOWFAR_TRACK: G4HepEmGammaManager::HowFar(data, pars, gt); break;
      case G4HB200_GOP_UPDATE_NIA: G4HepEmGammaManager::UpdateNumIALeft(gt->GetTrack()); break;
      case G4HB200_GOP_SELECT_INTERACTION: G4HepEmGammaManager::SelectInteraction(data, &tl); break;
      case G4HB200_GOP_PERFORM_SELECTED: G4HepEmGammaManager::Perform(data, pars, &tl); break;
      default: break;
    }
    if (secs != nullptr) CollectSecondaries(tl, i, *secs);
    PackGamma(b, i, *gt, stream);
  }
}

template <class Batch, class Fn>
int RunThreaded(RefState* rs, Batch* b, G4HB200SecondaryQueue* sec, uint64_t seed, Mode mode, int nthreads, Fn fn) {
  const int64_t n = b->n;
  if (nthreads < 1) nthreads = 1;
  if (n < nthreads) nthreads = n > 0 ? static_cast<int>(n) : 1;
  std::vector<std::vector<SecRecord>> secs(nthreads);
  if (nthreads == 1) {
    fn(rs, b, seed, mode, 0, n, &secs[0]);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
      const int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
      pool.emplace_back([=, &secs]() { fn(rs, b, seed, mode, lo, hi, &secs[t]); });
    }
    for (auto& th : pool) th.join();
  }
  for (auto& s : secs) {
    const int rc = AppendSecondaries(sec, s);
    if (rc != 0) return rc;
  }
  return 0;
}

}  // namespace

extern "C" {

void* g4href_load_state(const char* jsonPath) {
  std::ifstream in(jsonPath);
  if (!in.good()) return nullptr;
  RefState* rs = new RefState;
  rs->state    = G4HepEmStateFromJson(in);
  if (rs->state == nullptr || rs->state->fData == nullptr || rs->state->fParameters == nullptr) {
    delete rs;
    return nullptr;
  }
  G4HepEmB200Flatten(rs->state->fData, rs->state->fParameters, rs->flat);
  return rs;
}

void g4href_free_state(void* p) {
  RefState* rs = static_cast<RefState*>(p);
  if (rs == nullptr) return;
  FreeG4HepEmData(rs->state->fData);
  FreeG4HepEmParameters(rs->state->fParameters);
  delete rs->state->fData;
  delete rs->state->fParameters;
  delete rs->state;
  delete rs;
}

// JSON round trip through the reference's own writer (G4HepEmStateToJson)
int g4href_save_state(void* p, const char* jsonPath) {
  RefState* rs = static_cast<RefState*>(p);
  std::ofstream out(jsonPath);
  if (!out.good()) return -1;
  return G4HepEmStateToJson(out, rs->state) ? 0 : -1;
}

// the flat descriptor produced by the C++ adapter from the reference's structs
const G4HB200Tables* g4href_flat_tables(void* p) { return &static_cast<RefState*>(p)->flat.desc; }

int g4href_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }

void g4href_electron_lookups(void* p, int64_t n, const int32_t* imc, const double* ekin, const double* lekin,
                             int isElectron, double* out) {
  RefState* rs = static_cast<RefState*>(p);
  const G4HepEmData* data = rs->state->fData;
  const G4HepEmElectronData* ed = isElectron ? data->fTheElectronData : data->fThePositronData;
  for (int64_t i = 0; i < n; ++i) {
    const int imat = data->fTheMatCutData->fMatCutData[imc[i]].fHepEmMatIndex;
    const double range = G4HepEmElectronManager::GetRestRange(ed, imc[i], ekin[i], lekin[i]);
    out[0 * n + i] = range;
    out[1 * n + i] = G4HepEmElectronManager::GetRestDEDX(ed, imc[i], ekin[i], lekin[i]);
    out[2 * n + i] = G4HepEmElectronManager::GetInvRange(ed, imc[i], range);
    out[3 * n + i] = G4HepEmElectronManager::GetRestMacXSec(ed, imc[i], ekin[i], lekin[i], true);
    out[4 * n + i] = G4HepEmElectronManager::GetRestMacXSec(ed, imc[i], ekin[i], lekin[i], false);
    out[5 * n + i] = G4HepEmElectronManager::GetMacXSecNuclear(ed, imat, ekin[i], lekin[i]);
    out[6 * n + i] = G4HepEmElectronManager::GetTransportMFP(ed, imat, ekin[i], lekin[i]);
  }
}

void g4href_electron_stepping_xsecs(void* p, int64_t n, const int32_t* imc, const double* ekin, const double* lekin,
                                    int isElectron, double* out) {
  RefState* rs = static_cast<RefState*>(p);
  const G4HepEmData* data = rs->state->fData;
  const G4HepEmElectronData* ed = isElectron ? data->fTheElectronData : data->fThePositronData;
  for (int64_t i = 0; i < n; ++i) {
    const int imat = data->fTheMatCutData->fMatCutData[imc[i]].fHepEmMatIndex;
    out[0 * n + i] = G4HepEmElectronManager::GetRestMacXSecForStepping(ed, imc[i], ekin[i], lekin[i], true);
    out[1 * n + i] = G4HepEmElectronManager::GetRestMacXSecForStepping(ed, imc[i], ekin[i], lekin[i], false);
    out[2 * n + i] = G4HepEmElectronManager::GetMacXSecNuclearForStepping(ed, imat, ekin[i], lekin[i]);
    out[3 * n + i] = G4HepEmElectronManager::ComputeMacXsecAnnihilationForStepping(
        ekin[i], data->fTheMaterialData->fMaterialData[imat].fElectronDensity);
  }
}

void g4href_gamma_lookups(void* p, int64_t n, const int32_t* imc, const double* ekin, const double* lekin,
                          const double* urnd, double* outMxsec, int32_t* outPid) {
  RefState* rs = static_cast<RefState*>(p);
  G4HepEmData* data = rs->state->fData;
  G4HepEmGammaTrack gt;
  for (int64_t i = 0; i < n; ++i) {
    gt.ReSet();
    G4HepEmTrack* t = gt.GetTrack();
    t->SetEKin(ekin[i], lekin[i]);
    t->SetMCIndex(imc[i]);
    const double mx = G4HepEmGammaManager::GetTotalMacXSec(data, &gt);
    t->SetMFP(mx > 0.0 ? 1.0 / mx : 1.0e20, 0);
    G4HepEmGammaManager::SampleInteraction(data, &gt, urnd[i]);
    outMxsec[i] = mx;
    outPid[i]   = t->GetWinnerProcessIndex();
  }
}

void g4href_select_target_element(void* p, int kind, int isElectron, int64_t n, const int32_t* imc,
                                  const double* ekin, const double* lekin, const double* urnd, int32_t* outElem) {
  RefState* rs = static_cast<RefState*>(p);
  const G4HepEmData* data = rs->state->fData;
  const G4HepEmElectronData* ed = isElectron ? data->fTheElectronData : data->fThePositronData;
  for (int64_t i = 0; i < n; ++i) {
    if (kind == 2) {
      outElem[i] = G4HepEmGammaInteractionConversion::SelectTargetAtom(data->fTheGammaData, imc[i], ekin[i], lekin[i], urnd[i]);
    } else {
      outElem[i] = G4HepEmElectronInteractionBrem::SelectTargetAtom(ed, imc[i], ekin[i], lekin[i], urnd[i], kind == 0);
    }
  }
}

void g4href_vdt_log_exp(int64_t n, const double* x, double* outLog, double* outExp) {
  for (int64_t i = 0; i < n; ++i) {
    outLog[i] = G4HepEmLog(x[i]);
    outExp[i] = G4HepEmExp(x[i]);
  }
}

void g4href_rng_uniforms(uint64_t seed, int64_t n, const int32_t* trackId, int32_t ndraw, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    G4HStream s{seed, static_cast<uint32_t>(trackId[i]), 0};
    G4HepEmRandomEngine eng(&s);
    for (int j = 0; j < ndraw; ++j) out[i * ndraw + j] = eng.flat();
  }
}

int g4href_electron_track_op(void* p, int op, G4HB200ElectronBatch* b, G4HB200SecondaryQueue* sec, uint64_t seed, int32_t* flags) {
  std::vector<SecRecord> secs;
  ElectronOpRange(static_cast<RefState*>(p), op, b, seed, flags, sec != nullptr ? &secs : nullptr);
  return AppendSecondaries(sec, secs);
}
int g4href_electron_check_delta(void* p, G4HB200ElectronBatch* b, const double* urnd, int32_t* flags) {
  RefState* rs = static_cast<RefState*>(p);
  G4HepEmTLData tl;
  G4HStream stream{0, 0, 0};
  G4HepEmRandomEngine eng(&stream);
  G4HepEmElectronTrack* et = tl.GetPrimaryElectronTrack();
  for (int64_t i = 0; i < b->n; ++i) {
    UnpackElectron(b, i, *et, eng, stream, 0, true);
    flags[i] = G4HepEmElectronManager::CheckDelta(rs->state->fData, et->GetTrack(), urnd[i]) ? 1 : 0;
    b->ekin_logekin[2 * i + 1] = PeekLogEKin(et->GetTrack());
  }
  return 0;
}
int g4href_gamma_track_op(void* p, int op, G4HB200GammaBatch* b, G4HB200SecondaryQueue* sec, uint64_t seed) {
  std::vector<SecRecord> secs;
  GammaOpRange(static_cast<RefState*>(p), op, b, seed, sec != nullptr ? &secs : nullptr);
  return AppendSecondaries(sec, secs);
}
int g4href_electron_howfar(void* p, G4HB200ElectronBatch* b, uint64_t seed, int nthreads) {
  return RunThreaded(static_cast<RefState*>(p), b, nullptr, seed, Mode::kHowFar, nthreads, ElectronRange);
}
int g4href_electron_perform(void* p, G4HB200ElectronBatch* b, G4HB200SecondaryQueue* sec, uint64_t seed, int nthreads) {
  return RunThreaded(static_cast<RefState*>(p), b, sec, seed, Mode::kPerform, nthreads, ElectronRange);
}
int g4href_electron_step(void* p, G4HB200ElectronBatch* b, G4HB200SecondaryQueue* sec, uint64_t seed, int nthreads) {
  return RunThreaded(static_cast<RefState*>(p), b, sec, seed, Mode::kStep, nthreads, ElectronRange);
}
int g4href_gamma_howfar(void* p, G4HB200GammaBatch* b, uint64_t seed, int nthreads) {
  return RunThreaded(static_cast<RefState*>(p), b, nullptr, seed, Mode::kHowFar, nthreads, GammaRange);
}
int g4href_gamma_perform(void* p, G4HB200GammaBatch* b, G4HB200SecondaryQueue* sec, uint64_t seed, int nthreads) {
  return RunThreaded(static_cast<RefState*>(p), b, sec, seed, Mode::kPerform, nthreads, GammaRange);
}
int g4href_gamma_step(void* p, G4HB200GammaBatch* b, G4HB200SecondaryQueue* sec, uint64_t seed, int nthreads) {
  return RunThreaded(static_cast<RefState*>(p), b, sec, seed, Mode::kStep, nthreads, GammaRange);
}

}  // extern "C"
