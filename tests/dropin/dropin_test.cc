// dropin_test.cc -- TEST INFRASTRUCTURE: drives the same G4HepEmElectronTrack / G4HepEmGammaTrack objects through
//   (A) the reference's own managers (G4HepEmElectronManager::HowFar/Perform, G4HepEmGammaManager::HowFar/
//       SelectInteraction/Perform over a G4HepEmTLData, with the injected counter based engine of oracle/ref_shim.cc), and
//   (B) the product's C++ host layer (g4hepem_b200/host/G4HepEmB200Managers.hh -> C-ABI -> CUDA kernels)
// following the caller protocol of apps/examples/TestEm3/src/G4HepEmProcess.cc:106-217, for several steps with a
// geometry stub between HowFar and Perform, and counts the tracks whose public state differs.
// g4hdropin_track_level does the same for the manager classes of G4HepEmB200DropIn.hh (the reference's static signatures,
// one track per call) in the piece-by-piece order of G4HepEmTrackingManager::TrackElectron / TrackGamma.
// Built by oracle/Makefile (target dropin) against the reference headers into oracle/_ref/libg4hepem_dropin.so.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <vector>

#include "G4HepEmData.hh"
#include "G4HepEmElectronData.hh"
#include "G4HepEmElementData.hh"
#include "G4HepEmGammaData.hh"
#include "G4HepEmMatCutData.hh"
#include "G4HepEmMaterialData.hh"
#include "G4HepEmSBTableData.hh"
#include "G4HepEmParameters.hh"
#include "G4HepEmState.hh"
#include "G4HepEmDataJsonIO.hh"
#include "G4HepEmRandomEngine.hh"
#include "G4HepEmTLData.hh"
#include "G4HepEmElectronTrack.hh"
#include "G4HepEmGammaTrack.hh"
#include "G4HepEmElectronManager.hh"
#include "G4HepEmGammaManager.hh"

#include "G4HepEmPositronInteractionAnnihilation.hh"

#include "../../g4hepem_b200/host/G4HepEmB200Managers.hh"
#include "../../g4hepem_b200/host/G4HepEmB200DropIn.hh"
#include "../../oracle/g4h_rng_host.h"

namespace {

const double kRel = 1.0e-12;

bool Close(double a, double b) { return a == b || std::fabs(a - b) <= kRel * std::max(std::fabs(a), std::fabs(b)); }
bool CloseAbs(double a, double b, double scale) { return a == b || std::fabs(a - b) <= kRel * scale + 1.0e-12; }

struct Sec {
  int parent, slot, kind;
  double ekin, dir[3];
};
bool SecLess(const Sec& a, const Sec& b) { return a.parent != b.parent ? a.parent < b.parent : a.slot < b.slot; }

void Collect(G4HepEmTLData& tl, int parent, std::vector<Sec>& out) {
  int slot = 0;
  for (int k = 0; k < static_cast<int>(tl.GetNumSecondaryElectronTrack()); ++k) {
    G4HepEmTrack* s = tl.GetSecondaryElectronTrack(k)->GetTrack();
    out.push_back(Sec{parent, slot++, s->GetCharge() > 0.0 ? G4HB200_SEC_POSITRON : G4HB200_SEC_ELECTRON, s->GetEKin(),
                      {s->GetDirection()[0], s->GetDirection()[1], s->GetDirection()[2]}});
  }
  for (int k = 0; k < static_cast<int>(tl.GetNumSecondaryGammaTrack()); ++k) {
    G4HepEmTrack* s = tl.GetSecondaryGammaTrack(k)->GetTrack();
    out.push_back(Sec{parent, slot++, G4HB200_SEC_GAMMA, s->GetEKin(), {s->GetDirection()[0], s->GetDirection()[1], s->GetDirection()[2]}});
  }
  tl.ResetNumSecondaryElectronTrack();
  tl.ResetNumSecondaryGammaTrack();
}

int CompareSecondaries(std::vector<Sec>& a, const std::vector<G4HepEmB200Secondary>& bIn) {
  std::vector<Sec> b;
  for (const auto& s : bIn) b.push_back(Sec{s.fParentIndex, s.fSlot, s.fKind, s.fEKin, {s.fDirection[0], s.fDirection[1], s.fDirection[2]}});
  std::sort(a.begin(), a.end(), SecLess);
  std::sort(b.begin(), b.end(), SecLess);
  if (a.size() != b.size()) return static_cast<int>(std::max(a.size(), b.size()));
  int bad = 0;
  for (size_t k = 0; k < a.size(); ++k) {
    bool ok = a[k].parent == b[k].parent && a[k].slot == b[k].slot && a[k].kind == b[k].kind && Close(a[k].ekin, b[k].ekin);
    for (int d = 0; d < 3; ++d) ok = ok && CloseAbs(a[k].dir[d], b[k].dir[d], 1.0);
    bad += ok ? 0 : 1;
  }
  return bad;
}

bool SameTrack(G4HepEmTrack* a, G4HepEmTrack* b) {
  bool ok = Close(a->GetEKin(), b->GetEKin()) && Close(a->GetEnergyDeposit(), b->GetEnergyDeposit()) && Close(a->GetMFP(0), b->GetMFP(0)) &&
            Close(a->GetGStepLength(), b->GetGStepLength()) && a->GetWinnerProcessIndex() == b->GetWinnerProcessIndex();
  for (int d = 0; d < 3; ++d) ok = ok && CloseAbs(a->GetDirection()[d], b->GetDirection()[d], 1.0);
  for (int p = 0; p < 4; ++p) ok = ok && Close(a->GetNumIALeft(p), b->GetNumIALeft(p));
  return ok;
}


// ---- the production caller's order, written once against "a manager class" -------------------------------------------------
// G4HepEmTrackingManager::TrackElectron (G4HepEm/G4HepEm/src/G4HepEmTrackingManager.cc:408-665) without Geant4: the geometry
// is a deterministic stub (`cutEvery`: which sub-steps are cut short by a boundary), everything else is the reference's
// control flow, statement for statement: interaction lengths, HowFarToDiscreteInteraction, the MSC sub-step loop
// { HowFarToMSC, geometry, UpdatePStepLength, UpdateNumIALeft, ApplyMeanEnergyLoss, SampleMSC }, SampleLossFluctuations,
// annihilation at rest or PerformDiscrete.
struct RefAtRest {
  static void Do(G4HepEmTLData* tl) { G4HepEmPositronInteractionAnnihilation::Perform(tl, true); }
};
struct B200AtRest {
  static void Do(G4HepEmTLData* tl) { G4HepEmB200ElectronManager::AnnihilateAtRest(tl); }
};

template <class Manager, class AtRest>
int TrackElectronStep(G4HepEmData* data, G4HepEmParameters* pars, G4HepEmTLData* tl, bool preOnBoundary, double preSafety, int cutKey) {
  G4HepEmElectronTrack* theElTrack = tl->GetPrimaryElectronTrack();
  G4HepEmTrack* thePrimaryTrack    = theElTrack->GetTrack();
  G4HepEmRandomEngine* rnge        = tl->GetRNGEngine();
  const bool isElectron = thePrimaryTrack->GetCharge() < 0.0;
  const double preStepEkin    = thePrimaryTrack->GetEKin();
  const double preStepLogEkin = thePrimaryTrack->GetLogEKin();
  thePrimaryTrack->SetEKin(preStepEkin, preStepLogEkin);
  thePrimaryTrack->SetOnBoundary(preOnBoundary);
  thePrimaryTrack->SetSafety(preOnBoundary ? 0.0 : preSafety);
  const int indxRegion  = data->fTheMatCutData->fMatCutData[thePrimaryTrack->GetMCIndex()].fG4RegionIndex;
  bool continueStepping = pars->fParametersPerRegion[indxRegion].fIsMultipleStepsInMSCTrans;
  for (int ip = 0; ip < 4; ++ip) {
    if (thePrimaryTrack->GetNumIALeft(ip) <= 0.) thePrimaryTrack->SetNumIALeft(-G4HepEmLog(rnge->flat()), ip);
  }
  Manager::HowFarToDiscreteInteraction(data, pars, theElTrack);
  const int iDProc = thePrimaryTrack->GetWinnerProcessIndex();
  double stepLimitLeft = theElTrack->GetPStepLength();
  double totalEloss = 0;
  bool stopped = false;
  int subSteps = 0;
  theElTrack->SavePreStepEKin();
  do {
    Manager::HowFarToMSC(data, pars, theElTrack, rnge);
    if (thePrimaryTrack->GetWinnerProcessIndex() != -2) continueStepping = false;
    const double physicalStep = thePrimaryTrack->GetGStepLength();
    // geometry stub: some (sub-)steps end on a boundary half way
    const bool geometryLimitedStep = ((cutKey + 3 * subSteps) % 5) == 0;
    const double finalStep = geometryLimitedStep ? 0.5 * physicalStep : physicalStep;
    if (geometryLimitedStep) continueStepping = false;
    const bool postStepOnBoundary = geometryLimitedStep;
    thePrimaryTrack->SetGStepLength(finalStep);
    thePrimaryTrack->SetOnBoundary(postStepOnBoundary);
    ++subSteps;
    if (finalStep > 0) {
      do {
        Manager::UpdatePStepLength(theElTrack);
        const double pStepLength = theElTrack->GetPStepLength();
        if (pStepLength <= 0.0) break;
        Manager::UpdateNumIALeft(theElTrack);
        stopped = Manager::ApplyMeanEnergyLoss(data, pars, theElTrack);
        totalEloss += thePrimaryTrack->GetEnergyDeposit();
        if (stopped) {
          continueStepping = false;
          break;
        }
        Manager::SampleMSC(data, pars, theElTrack, rnge);
      } while (0);
      if (continueStepping) {
        thePrimaryTrack->SetEnergyDeposit(0);
        thePrimaryTrack->SetWinnerProcessIndex(iDProc);
        theElTrack->SavePreStepEKin();
        const double pStepLength = theElTrack->GetPStepLength();
        stepLimitLeft -= pStepLength;
        theElTrack->SetPStepLength(stepLimitLeft);
        thePrimaryTrack->SetGStepLength(stepLimitLeft);
        theElTrack->SetRange(theElTrack->GetRange() - pStepLength);
      }
    }
    if (subSteps > 64) continueStepping = false;  // the stub has no world to leave
  } while (continueStepping);
  thePrimaryTrack->SetEnergyDeposit(totalEloss);
  if (!stopped) {
    theElTrack->SetPreStepEKin(preStepEkin, preStepLogEkin);
    stopped = Manager::SampleLossFluctuations(data, pars, theElTrack, rnge);
  }
  if (stopped) {
    if (!isElectron) AtRest::Do(tl);
  } else if (!thePrimaryTrack->GetOnBoundary() && thePrimaryTrack->GetWinnerProcessIndex() != 3) {
    Manager::PerformDiscrete(data, pars, tl);
  } else if (!thePrimaryTrack->GetOnBoundary()) {
    // lepto-nuclear: the caller clears the interaction length and asks CheckDelta (.cc:655-665); no Geant4 process here
    thePrimaryTrack->SetNumIALeft(-1.0, 3);
    (void)Manager::CheckDelta(data, thePrimaryTrack, rnge->flat());
  }
  return subSteps;
}

// G4HepEmTrackingManager::TrackGamma (.cc:985-1140) without Woodcock tracking and without Geant4
template <class Manager>
void TrackGammaStep(G4HepEmData* data, G4HepEmParameters* pars, G4HepEmTLData* tl, int cutKey) {
  G4HepEmTrack* thePrimaryTrack = tl->GetPrimaryGammaTrack()->GetTrack();
  Manager::HowFar(data, pars, tl);
  const double physicalStep = thePrimaryTrack->GetGStepLength();
  const bool onBoundary = (cutKey % 4) == 0;
  thePrimaryTrack->SetGStepLength(onBoundary ? 0.5 * physicalStep : physicalStep);
  thePrimaryTrack->SetOnBoundary(onBoundary);
  if (onBoundary) {
    Manager::UpdateNumIALeft(thePrimaryTrack);
  } else {
    Manager::SelectInteraction(data, tl);
    if (thePrimaryTrack->GetWinnerProcessIndex() != 3) {
      Manager::Perform(data, pars, tl);
    } else {
      thePrimaryTrack->SetEnergyDeposit(0.0);
    }
  }
}

bool SameElectronTrack(G4HepEmElectronTrack& a, G4HepEmElectronTrack& b) {
  bool ok = SameTrack(a.GetTrack(), b.GetTrack()) && Close(a.GetPStepLength(), b.GetPStepLength()) && Close(a.GetRange(), b.GetRange());
  for (int p = 0; p < 4; ++p) ok = ok && Close(a.GetTrack()->GetMFP(p), b.GetTrack()->GetMFP(p));
  G4HepEmMSCTrackData* ma = a.GetMSCTrackData();
  G4HepEmMSCTrackData* mb = b.GetMSCTrackData();
  ok = ok && Close(ma->fTrueStepLength, mb->fTrueStepLength) && Close(ma->fZPathLength, mb->fZPathLength) &&
       Close(ma->fInitialRange, mb->fInitialRange) && Close(ma->fDynamicRangeFactor, mb->fDynamicRangeFactor) &&
       Close(ma->fTlimitMin, mb->fTlimitMin) && ma->fIsActive == mb->fIsActive && ma->fIsFirstStep == mb->fIsFirstStep &&
       ma->fIsDisplace == mb->fIsDisplace && ma->fIsNoScatteringInMSC == mb->fIsNoScatteringInMSC;
  const double* da = ma->GetDisplacement();
  const double* db = mb->GetDisplacement();
  const double norm = std::sqrt(da[0] * da[0] + da[1] * da[1] + da[2] * da[2]);
  for (int d = 0; d < 3; ++d) ok = ok && CloseAbs(da[d], db[d], norm);
  return ok;
}

void CollectTL(G4HepEmTLData& tl, int parent, std::vector<Sec>& out) { Collect(tl, parent, out); }

int CompareSecLists(std::vector<Sec>& a, std::vector<Sec>& b) {
  if (a.size() != b.size()) return static_cast<int>(std::max(a.size(), b.size()));
  int bad = 0;
  for (size_t k = 0; k < a.size(); ++k) {
    bool ok = a[k].parent == b[k].parent && a[k].slot == b[k].slot && a[k].kind == b[k].kind && Close(a[k].ekin, b[k].ekin);
    for (int d = 0; d < 3; ++d) ok = ok && CloseAbs(a[k].dir[d], b[k].dir[d], 1.0);
    bad += ok ? 0 : 1;
  }
  return bad;
}

}  // namespace

extern "C" {

// The manager classes of G4HepEmB200DropIn.hh against the reference's, in the production caller's order, one track per call.
// report: [0] e-/e+ tracks differing after a step, [1] e-/e+ secondaries differing, [2] steps with more than one MSC sub-step,
//         [3] gamma tracks differing after a step, [4] gamma secondaries differing, [5] total secondaries, [6] draw counters
//         differing, [7] status of the drop-in (0 = ok)
int g4hdropin_track_level(const char* jsonPath, int64_t n, uint64_t seed, int nsteps, const double* ekin, const int32_t* imc,
                          const int32_t* isPositron, const double* dir, const double* safety, int64_t* report) {
  for (int k = 0; k < 8; ++k) report[k] = 0;
  std::ifstream in(jsonPath);
  if (!in.good()) return -100;
  G4HepEmState* state = G4HepEmStateFromJson(in);
  if (state == nullptr) return -101;
  G4HepEmData* data = state->fData;
  G4HepEmParameters* pars = state->fParameters;
  int rc = G4HepEmB200DropIn::Attach(data, pars, 0);
  if (rc != 0) { report[7] = rc; return rc; }
  G4HepEmTLData tlA, tlB;
  for (int64_t i = 0; i < n; ++i) {
    G4HStream streamA{seed, static_cast<uint32_t>(i + 1), 0};
    G4HepEmB200Stream streamB{seed, static_cast<uint32_t>(i + 1), 0, 0, 0.0};
    G4HepEmRandomEngine engA(&streamA), engB(&streamB);
    G4HepEmB200DropIn::BindEngine(&engB, &streamB);
    tlA.SetRandomEngine(&engA);
    tlB.SetRandomEngine(&engB);
    // ---- e-/e+
    for (G4HepEmTLData* tl : {&tlA, &tlB}) {
      G4HepEmElectronTrack* et = tl->GetPrimaryElectronTrack();
      et->ReSet();
      G4HepEmTrack* t = et->GetTrack();
      t->SetCharge(isPositron[i] ? 1.0 : -1.0);
      t->SetEKin(ekin[i]);
      t->SetMCIndex(imc[i]);
      t->SetID(static_cast<int>(i + 1));
      t->SetDirection(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
      et->SetPreStepEKin(0.0, 0.0);
    }
    for (int step = 0; step < nsteps; ++step) {
      const bool onb = ((i + step) % 7) == 0;
      const int cutKey = static_cast<int>(i + 2 * step);
      const int subA = TrackElectronStep<G4HepEmElectronManager, RefAtRest>(data, pars, &tlA, onb, safety[i], cutKey);
      const int subB = TrackElectronStep<G4HepEmB200ElectronManager, B200AtRest>(data, pars, &tlB, onb, safety[i], cutKey);
      if (G4HepEmB200DropIn::LastStatus() != 0) { report[7] = G4HepEmB200DropIn::LastStatus(); return static_cast<int>(report[7]); }
      report[2] += subA > 1 ? 1 : 0;
      const bool ok = subA == subB && SameElectronTrack(*tlA.GetPrimaryElectronTrack(), *tlB.GetPrimaryElectronTrack());
      report[0] += ok ? 0 : 1;
      report[6] += streamA.draw == streamB.draw ? 0 : 1;
      std::vector<Sec> secA, secB;
      CollectTL(tlA, static_cast<int>(i), secA);
      CollectTL(tlB, static_cast<int>(i), secB);
      report[5] += static_cast<int64_t>(secA.size());
      report[1] += CompareSecLists(secA, secB);
      if (tlA.GetPrimaryElectronTrack()->GetTrack()->GetEKin() <= 0.0 || tlB.GetPrimaryElectronTrack()->GetTrack()->GetEKin() <= 0.0) {
        for (G4HepEmTLData* tl : {&tlA, &tlB}) {
          G4HepEmElectronTrack* et = tl->GetPrimaryElectronTrack();
          const double charge = et->GetTrack()->GetCharge();
          et->ReSet();
          et->GetTrack()->SetCharge(charge);
          et->GetTrack()->SetEKin(ekin[i] * 0.7);
          et->GetTrack()->SetMCIndex(imc[i]);
          et->GetTrack()->SetID(static_cast<int>(i + 1));
          et->GetTrack()->SetDirection(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
          et->SetPreStepEKin(0.0, 0.0);
        }
        engA.DiscardGauss();
        streamB.is_gauss = 0;
      }
    }
    // ---- gamma (its own stream: ids n+1 ...)
    streamA = G4HStream{seed, static_cast<uint32_t>(n + i + 1), 0};
    streamB = G4HepEmB200Stream{seed, static_cast<uint32_t>(n + i + 1), 0, 0, 0.0};
    engA.DiscardGauss();
    for (G4HepEmTLData* tl : {&tlA, &tlB}) {
      G4HepEmGammaTrack* gt = tl->GetPrimaryGammaTrack();
      gt->ReSet();
      G4HepEmTrack* t = gt->GetTrack();
      t->SetEKin(ekin[i]);
      t->SetMCIndex(imc[i]);
      t->SetID(static_cast<int>(n + i + 1));
      t->SetDirection(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
    }
    for (int step = 0; step < nsteps; ++step) {
      const int cutKey = static_cast<int>(i + step);
      TrackGammaStep<G4HepEmGammaManager>(data, pars, &tlA, cutKey);
      TrackGammaStep<G4HepEmB200GammaManager>(data, pars, &tlB, cutKey);
      if (G4HepEmB200DropIn::LastStatus() != 0) { report[7] = G4HepEmB200DropIn::LastStatus(); return static_cast<int>(report[7]); }
      G4HepEmTrack* a = tlA.GetPrimaryGammaTrack()->GetTrack();
      G4HepEmTrack* b = tlB.GetPrimaryGammaTrack()->GetTrack();
      report[3] += SameTrack(a, b) ? 0 : 1;
      report[6] += streamA.draw == streamB.draw ? 0 : 1;
      std::vector<Sec> secA, secB;
      CollectTL(tlA, static_cast<int>(i), secA);
      CollectTL(tlB, static_cast<int>(i), secB);
      report[5] += static_cast<int64_t>(secA.size());
      report[4] += CompareSecLists(secA, secB);
      if (a->GetEKin() <= 0.0 || b->GetEKin() <= 0.0) {
        for (G4HepEmTLData* tl : {&tlA, &tlB}) {
          G4HepEmGammaTrack* gt = tl->GetPrimaryGammaTrack();
          gt->ReSet();
          gt->GetTrack()->SetEKin(ekin[i] * 0.7);
          gt->GetTrack()->SetMCIndex(imc[i]);
          gt->GetTrack()->SetID(static_cast<int>(n + i + 1));
          gt->GetTrack()->SetDirection(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        }
      }
    }
    G4HepEmB200DropIn::UnbindEngine(&engB);
  }
  G4HepEmB200DropIn::Detach();
  return 0;
}

// report: [0] e-/e+ tracks differing after HowFar, [1] after Perform, [2] e-/e+ secondaries differing,
//         [3] gamma tracks differing after HowFar, [4] after Perform, [5] gamma secondaries differing,
//         [6] total secondaries seen, [7] return code of the session (0 = ok)
int g4hdropin_run(const char* jsonPath, int64_t n, uint64_t seed, int nsteps, const double* ekin, const int32_t* imc,
                  const int32_t* isPositron, const double* dir, const double* safety, int64_t* report) {
  for (int k = 0; k < 8; ++k) report[k] = 0;
  std::ifstream in(jsonPath);
  if (!in.good()) return -100;
  G4HepEmState* state = G4HepEmStateFromJson(in);
  if (state == nullptr) return -101;
  G4HepEmData* data = state->fData;
  G4HepEmParameters* pars = state->fParameters;
  G4HepEmB200Session session;
  int rc = session.Open(data, pars, 0, seed);
  if (rc != 0) { report[7] = rc; return rc; }

  // ---- e-/e+ ------------------------------------------------------------------------------------------------
  std::vector<G4HepEmElectronTrack> A(n), B(n);
  std::vector<G4HepEmB200TrackAux> aux(n);
  std::vector<G4HStream> streams(n);
  std::vector<G4HepEmRandomEngine> engines;
  engines.reserve(n);
  for (int64_t i = 0; i < n; ++i) {
    streams[i] = G4HStream{seed, static_cast<uint32_t>(i + 1), 0};
    engines.emplace_back(&streams[i]);
    for (G4HepEmElectronTrack* et : {&A[i], &B[i]}) {
      et->ReSet();
      G4HepEmTrack* t = et->GetTrack();
      t->SetCharge(isPositron[i] ? 1.0 : -1.0);
      t->SetEKin(ekin[i]);
      t->SetMCIndex(imc[i]);
      t->SetID(static_cast<int>(i + 1));
      t->SetDirection(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
    }
  }
  G4HepEmTLData tl;
  for (int step = 0; step < nsteps; ++step) {
    // pre-step point: boundary flag and safety from "geometry"
    for (int64_t i = 0; i < n; ++i) {
      const bool onb = ((i + step) % 7) == 0;
      for (G4HepEmElectronTrack* et : {&A[i], &B[i]}) {
        et->GetTrack()->SetOnBoundary(onb);
        et->GetTrack()->SetSafety(onb ? 0.0 : safety[i]);
      }
    }
    for (int64_t i = 0; i < n; ++i) {
      tl.SetRandomEngine(&engines[i]);
      *tl.GetPrimaryElectronTrack() = A[i];
      G4HepEmElectronManager::HowFar(data, pars, &tl);
      A[i] = *tl.GetPrimaryElectronTrack();
    }
    if ((rc = session.ElectronHowFar(B.data(), aux.data(), n)) != 0) { report[7] = rc; return rc; }
    for (int64_t i = 0; i < n; ++i) {
      const bool ok = SameTrack(A[i].GetTrack(), B[i].GetTrack()) && Close(A[i].GetPStepLength(), B[i].GetPStepLength());
      report[0] += ok ? 0 : 1;
    }
    // geometry stub: every 5th step is cut short and ends on a boundary
    for (int64_t i = 0; i < n; ++i) {
      const bool cut = ((i + 3 * step) % 5) == 0;
      for (G4HepEmElectronTrack* et : {&A[i], &B[i]}) {
        G4HepEmTrack* t = et->GetTrack();
        if (cut) t->SetGStepLength(0.5 * t->GetGStepLength());
        t->SetOnBoundary(cut);
      }
    }
    std::vector<Sec> secA;
    for (int64_t i = 0; i < n; ++i) {
      tl.SetRandomEngine(&engines[i]);
      *tl.GetPrimaryElectronTrack() = A[i];
      G4HepEmElectronManager::Perform(data, pars, &tl);
      A[i] = *tl.GetPrimaryElectronTrack();
      Collect(tl, static_cast<int>(i), secA);
    }
    std::vector<G4HepEmB200Secondary> secB;
    if ((rc = session.ElectronPerform(B.data(), aux.data(), n, &secB)) != 0) { report[7] = rc; return rc; }
    for (int64_t i = 0; i < n; ++i) {
      bool ok = SameTrack(A[i].GetTrack(), B[i].GetTrack()) && Close(A[i].GetPStepLength(), B[i].GetPStepLength());
      const double* da = A[i].GetMSCTrackData()->GetDisplacement();
      const double* db = B[i].GetMSCTrackData()->GetDisplacement();
      const double norm = std::sqrt(da[0] * da[0] + da[1] * da[1] + da[2] * da[2]);
      for (int d = 0; d < 3; ++d) ok = ok && CloseAbs(da[d], db[d], norm);
      report[1] += ok ? 0 : 1;
    }
    report[6] += static_cast<int64_t>(secA.size());
    report[2] += CompareSecondaries(secA, secB);
    // stopped tracks are re-born with a new energy, as new tracks would be (ReSet + DiscardGauss of StartTracking)
    for (int64_t i = 0; i < n; ++i) {
      if (A[i].GetTrack()->GetEKin() <= 0.0 || B[i].GetTrack()->GetEKin() <= 0.0) {
        for (G4HepEmElectronTrack* et : {&A[i], &B[i]}) {
          const double charge = et->GetTrack()->GetCharge();
          const int mc = et->GetTrack()->GetMCIndex();
          et->ReSet();
          et->GetTrack()->SetCharge(charge);
          et->GetTrack()->SetEKin(ekin[i] * 0.7);
          et->GetTrack()->SetMCIndex(mc);
          et->GetTrack()->SetID(static_cast<int>(i + 1));
          et->GetTrack()->SetDirection(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        }
        engines[i].DiscardGauss();
        aux[i].fIsGauss = false;
      }
    }
  }

  // ---- gamma -----------------------------------------------------------------------------------------------------
  std::vector<G4HepEmGammaTrack> GA(n), GB(n);
  std::vector<G4HepEmB200TrackAux> gaux(n);
  for (int64_t i = 0; i < n; ++i) {
    streams[i] = G4HStream{seed, static_cast<uint32_t>(n + i + 1), 0};
    for (G4HepEmGammaTrack* gt : {&GA[i], &GB[i]}) {
      gt->ReSet();
      G4HepEmTrack* t = gt->GetTrack();
      t->SetEKin(ekin[i]);
      t->SetMCIndex(imc[i]);
      t->SetID(static_cast<int>(n + i + 1));
      t->SetDirection(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
    }
  }
  for (int step = 0; step < nsteps; ++step) {
    for (int64_t i = 0; i < n; ++i) {
      tl.SetRandomEngine(&engines[i]);
      *tl.GetPrimaryGammaTrack() = GA[i];
      G4HepEmGammaManager::HowFar(data, pars, &tl);
      GA[i] = *tl.GetPrimaryGammaTrack();
    }
    if ((rc = session.GammaHowFar(GB.data(), gaux.data(), n)) != 0) { report[7] = rc; return rc; }
    for (int64_t i = 0; i < n; ++i) report[3] += SameTrack(GA[i].GetTrack(), GB[i].GetTrack()) ? 0 : 1;
    for (int64_t i = 0; i < n; ++i) {
      const bool cut = ((i + step) % 4) == 0;
      for (G4HepEmGammaTrack* gt : {&GA[i], &GB[i]}) {
        G4HepEmTrack* t = gt->GetTrack();
        if (cut) t->SetGStepLength(0.5 * t->GetGStepLength());
        t->SetOnBoundary(cut);
      }
    }
    std::vector<Sec> secA;
    for (int64_t i = 0; i < n; ++i) {
      tl.SetRandomEngine(&engines[i]);
      *tl.GetPrimaryGammaTrack() = GA[i];
      if (!GA[i].GetTrack()->GetOnBoundary()) G4HepEmGammaManager::SelectInteraction(data, &tl);
      G4HepEmGammaManager::Perform(data, pars, &tl);
      GA[i] = *tl.GetPrimaryGammaTrack();
      Collect(tl, static_cast<int>(i), secA);
    }
    std::vector<G4HepEmB200Secondary> secB;
    if ((rc = session.GammaPerform(GB.data(), gaux.data(), n, &secB)) != 0) { report[7] = rc; return rc; }
    for (int64_t i = 0; i < n; ++i) report[4] += SameTrack(GA[i].GetTrack(), GB[i].GetTrack()) ? 0 : 1;
    report[6] += static_cast<int64_t>(secA.size());
    report[5] += CompareSecondaries(secA, secB);
    for (int64_t i = 0; i < n; ++i) {
      if (GA[i].GetTrack()->GetEKin() <= 0.0 || GB[i].GetTrack()->GetEKin() <= 0.0) {
        for (G4HepEmGammaTrack* gt : {&GA[i], &GB[i]}) {
          const int mc = gt->GetTrack()->GetMCIndex();
          gt->ReSet();
          gt->GetTrack()->SetEKin(ekin[i] * 0.7);
          gt->GetTrack()->SetMCIndex(mc);
          gt->GetTrack()->SetID(static_cast<int>(n + i + 1));
          gt->GetTrack()->SetDirection(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        }
      }
    }
  }
  session.Close();
  return 0;
}

}  // extern "C"
