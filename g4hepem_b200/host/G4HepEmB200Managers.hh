// G4HepEmB200Managers.hh -- C++ host side of the drop-in boundary: the reference's track objects in, the
// reference's track objects out, the physics on the B200 through the C-ABI (include/g4hepem_b200.h).
//
// Header-only; compiled in the application's translation unit against the application's own G4HepEm headers
// (G4HepEmData.hh, G4HepEmParameters.hh, G4HepEmElectronTrack.hh, G4HepEmGammaTrack.hh must be included first),
// linked with libg4hepem_b200.so.  No CUDA header is needed here.
//
// What it mirrors (G4HepEm/G4HepEmRun/include/):
//   G4HepEmElectronManager::HowFar / Perform   (G4HepEmElectronManager.hh:71,222)  -> ElectronHowFar / ElectronPerform / ElectronStep
//   G4HepEmGammaManager::HowFar / SelectInteraction + Perform (G4HepEmGammaManager.hh:32-51) -> GammaHowFar / GammaPerform / GammaStep
// but over an ARRAY of primary tracks per call (the tracks of n independent workers) instead of the single
// primary of one G4HepEmTLData.  The caller protocol is the reference's (apps/examples/TestEm3/src/
// G4HepEmProcess.cc:106-217): set charge, SetEKin, SetMCIndex, SetOnBoundary, SetSafety before HowFar; set the
// direction, the final geometrical step and the post-step boundary flag before Perform; read ekin, edep,
// direction, pStep, displacement, winner and the secondaries after it.
//
// Random numbers: the reference draws from the worker's G4HepEmRandomEngine; here every track owns a counter based
// stream keyed (seed, track ID) -- G4HepEmTrack::fID must be unique per live track -- and the engine state the
// reference keeps per worker (next draw, cached Gauss variate) travels per track in G4HepEmB200TrackAux.
#ifndef G4HEPEMB200_MANAGERS_HH
#define G4HEPEMB200_MANAGERS_HH

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "G4HepEmB200Flatten.hh"
#include "g4hepem_b200.h"

struct G4HepEmB200TrackAux {
  int32_t fNextDraw = 0;   // index of the next uniform of the track's stream
  bool fIsGauss     = false;  // G4HepEmRandomEngine::fIsGauss
  double fGauss     = 0.0;    // G4HepEmRandomEngine::fGauss
};

struct G4HepEmB200Secondary {
  double fDirection[3];
  double fEKin;
  int32_t fParentID;     // G4HepEmTrack::fID of the parent
  int32_t fKind;         // G4HB200_SEC_ELECTRON / _POSITRON / _GAMMA
  int32_t fParentIndex;  // index of the parent in the array handed to Perform
  int32_t fSlot;         // 0 / 1: order in which the reference would have added it
};

class G4HepEmB200Session {
 public:
  G4HepEmB200Session() { std::memset(&fElDev, 0, sizeof(fElDev)); std::memset(&fGmDev, 0, sizeof(fGmDev)); std::memset(&fSecDev, 0, sizeof(fSecDev)); }
  ~G4HepEmB200Session() { Close(); }
  G4HepEmB200Session(const G4HepEmB200Session&) = delete;
  G4HepEmB200Session& operator=(const G4HepEmB200Session&) = delete;

  // replaces CopyG4HepEmDataToGPU (G4HepEmData/src/G4HepEmData.cc:78-101)
  int Open(const G4HepEmData* data, const G4HepEmParameters* pars, int device, uint64_t seed) {
    Close();
    G4HepEmB200Flatten(data, pars, fFlat);
    fSeed = seed;
    return g4hb200_create(&fFlat.desc, device, &fHandle);
  }
  void Close() {
    if (fHandle == nullptr) return;
    if (fElCap > 0) g4hb200_electron_batch_free(fHandle, &fElDev);
    if (fGmCap > 0) g4hb200_gamma_batch_free(fHandle, &fGmDev);
    if (fSecCap > 0) g4hb200_secondary_queue_free(fHandle, &fSecDev);
    if (fFlagDev != nullptr) g4hb200_device_free(fHandle, fFlagDev);
    g4hb200_destroy(fHandle);
    fHandle = nullptr;
    fFlagDev = nullptr;
    fElCap = fGmCap = fSecCap = fFlagCap = 0;
  }
  bool IsOpen() const { return fHandle != nullptr; }
  G4HB200* Handle() const { return fHandle; }
  uint64_t Seed() const { return fSeed; }
  void SetSeed(uint64_t seed) { fSeed = seed; }
  const char* LastError() const { return g4hb200_last_error(); }

  int ElectronHowFar(G4HepEmElectronTrack* tracks, G4HepEmB200TrackAux* aux, int64_t n) { return RunElectron(tracks, aux, n, 0, nullptr); }
  int ElectronPerform(G4HepEmElectronTrack* tracks, G4HepEmB200TrackAux* aux, int64_t n, std::vector<G4HepEmB200Secondary>* sec) {
    return RunElectron(tracks, aux, n, 1, sec);
  }
  // HowFar + Perform with the proposed geometrical step accepted (no geometry in between)
  int ElectronStep(G4HepEmElectronTrack* tracks, G4HepEmB200TrackAux* aux, int64_t n, std::vector<G4HepEmB200Secondary>* sec) {
    return RunElectron(tracks, aux, n, 2, sec);
  }
  int GammaHowFar(G4HepEmGammaTrack* tracks, G4HepEmB200TrackAux* aux, int64_t n) { return RunGamma(tracks, aux, n, 0, nullptr); }
  // SelectInteraction (when the step did not end on a boundary) + Perform, as the reference's callers sequence them
  int GammaPerform(G4HepEmGammaTrack* tracks, G4HepEmB200TrackAux* aux, int64_t n, std::vector<G4HepEmB200Secondary>* sec) {
    return RunGamma(tracks, aux, n, 1, sec);
  }
  int GammaStep(G4HepEmGammaTrack* tracks, G4HepEmB200TrackAux* aux, int64_t n, std::vector<G4HepEmB200Secondary>* sec) {
    return RunGamma(tracks, aux, n, 2, sec);
  }

  // ---- the track-level statics of the managers, one at a time over an array of tracks (G4HB200_OP_* / G4HB200_GOP_* of
  // include/g4hepem_b200.h; G4HepEmElectronManager.hh:90-206, G4HepEmGammaManager.hh:34-51).  flags[i] (may be null): the bool
  // the reference's function returns.  sec (may be null): the secondaries of the ops that create some.
  int ElectronTrackOp(int op, G4HepEmElectronTrack* tracks, G4HepEmB200TrackAux* aux, int64_t n, std::vector<G4HepEmB200Secondary>* sec,
                      int32_t* flags) {
    return RunElectron(tracks, aux, n, 3 + op, sec, flags);
  }
  int GammaTrackOp(int op, G4HepEmGammaTrack* tracks, G4HepEmB200TrackAux* aux, int64_t n, std::vector<G4HepEmB200Secondary>* sec) {
    return RunGamma(tracks, aux, n, 3 + op, sec);
  }
  // CheckDelta(data, track, rand) (G4HepEmElectronManager.hh:195): isDelta[i] for uniform urnd[i]
  int ElectronCheckDelta(G4HepEmElectronTrack* tracks, G4HepEmB200TrackAux* aux, int64_t n, const double* urnd, int32_t* isDelta) {
    return RunElectron(tracks, aux, n, -1, nullptr, isDelta, urnd);
  }

 private:
  struct HostElectron {
    std::vector<double> g[17];
    std::vector<int32_t> meta, winner;
    G4HB200ElectronBatch view;
    void Resize(int64_t n) {
      for (auto& v : g) v.assign(static_cast<size_t>(2 * n), 0.0);
      meta.assign(static_cast<size_t>(4 * n), 0);
      winner.assign(static_cast<size_t>(n), -1);
      view.n = n;
      double** p[17] = {&view.ekin_logekin, &view.dirx_diry, &view.dirz_safety, &view.nia01, &view.nia23, &view.msc_irange_dynrf,
                        &view.msc_tlimmin_gauss, &view.gstep_pstep, &view.edep_dispx, &view.dispy_dispz, &view.mfp01, &view.mfp23,
                        &view.range_lambtr1, &view.tstep_zpath, &view.par12, &view.par3_pad, &view.prestep};
      for (int k = 0; k < 17; ++k) *p[k] = g[k].data();
      view.meta = meta.data();
      view.winner = winner.data();
    }
  };
  struct HostGamma {
    std::vector<double> g[5];
    std::vector<int32_t> meta, winner;
    G4HB200GammaBatch view;
    void Resize(int64_t n) {
      for (auto& v : g) v.assign(static_cast<size_t>(2 * n), 0.0);
      meta.assign(static_cast<size_t>(4 * n), 0);
      winner.assign(static_cast<size_t>(n), -1);
      view.n = n;
      double** p[5] = {&view.ekin_logekin, &view.dirx_diry, &view.dirz_nia0, &view.gstep_mfp0, &view.edep_pemxsec};
      for (int k = 0; k < 5; ++k) *p[k] = g[k].data();
      view.meta = meta.data();
      view.winner = winner.data();
    }
  };
  struct HostSecondaries {
    std::vector<double> dxy, dze;
    std::vector<int32_t> pk, ps;
    int32_t count[4];
    G4HB200SecondaryQueue view;
    void Resize(int64_t cap) {
      dxy.assign(static_cast<size_t>(2 * cap), 0.0);
      dze.assign(static_cast<size_t>(2 * cap), 0.0);
      pk.assign(static_cast<size_t>(2 * cap), 0);
      ps.assign(static_cast<size_t>(2 * cap), 0);
      count[0] = 0;
      view.capacity = cap;
      view.dirx_diry = dxy.data();
      view.dirz_ekin = dze.data();
      view.parent_kind = pk.data();
      view.parent_slot = ps.data();
      view.count = count;
      view.parent_base = 0;
    }
  };

  // G4HepEmElectronTrack -> row i of the staging batch
  static void Pack(G4HepEmElectronTrack& et, const G4HepEmB200TrackAux& a, G4HB200ElectronBatch& b, int64_t i) {
    G4HepEmTrack* t = et.GetTrack();
    G4HepEmMSCTrackData* msc = et.GetMSCTrackData();
    b.ekin_logekin[2 * i]     = t->GetEKin();
    b.ekin_logekin[2 * i + 1] = t->GetLogEKin();  // evaluates the reference's lazy cache (same VDT log as the kernels)
    const double* dir = t->GetDirection();
    b.dirx_diry[2 * i]       = dir[0];
    b.dirx_diry[2 * i + 1]   = dir[1];
    b.dirz_safety[2 * i]     = dir[2];
    b.dirz_safety[2 * i + 1] = t->GetSafety();
    b.nia01[2 * i]     = t->GetNumIALeft(0);
    b.nia01[2 * i + 1] = t->GetNumIALeft(1);
    b.nia23[2 * i]     = t->GetNumIALeft(2);
    b.nia23[2 * i + 1] = t->GetNumIALeft(3);
    b.msc_irange_dynrf[2 * i]      = msc->fInitialRange;
    b.msc_irange_dynrf[2 * i + 1]  = msc->fDynamicRangeFactor;
    b.msc_tlimmin_gauss[2 * i]     = msc->fTlimitMin;
    b.msc_tlimmin_gauss[2 * i + 1] = a.fGauss;
    uint32_t flags = 0;
    if (t->GetCharge() > 0.0) flags |= G4HB200_F_POSITRON;
    if (t->GetOnBoundary()) flags |= G4HB200_F_ON_BOUNDARY;
    if (msc->fIsFirstStep) flags |= G4HB200_F_MSC_FIRST_STEP;
    if (msc->fIsActive) flags |= G4HB200_F_MSC_ACTIVE;
    if (msc->fIsDisplace) flags |= G4HB200_F_MSC_DISPLACE;
    if (msc->fIsNoScatteringInMSC) flags |= G4HB200_F_MSC_NO_SCATTER;
    if (a.fIsGauss) flags |= G4HB200_F_GAUSS_CACHED;
    int32_t* meta = b.meta + 4 * i;
    meta[0] = t->GetMCIndex();
    meta[1] = static_cast<int32_t>(flags);
    meta[2] = t->GetID();
    meta[3] = a.fNextDraw;
    // what HowFar left in the track for Perform
    b.gstep_pstep[2 * i]       = t->GetGStepLength();
    b.gstep_pstep[2 * i + 1]   = et.GetPStepLength();
    const double* disp = msc->GetDisplacement();
    b.edep_dispx[2 * i]        = t->GetEnergyDeposit();
    b.edep_dispx[2 * i + 1]    = disp[0];
    b.dispy_dispz[2 * i]       = disp[1];
    b.dispy_dispz[2 * i + 1]   = disp[2];
    b.winner[i]                = t->GetWinnerProcessIndex();
    b.mfp01[2 * i]             = t->GetMFP(0);
    b.mfp01[2 * i + 1]         = t->GetMFP(1);
    b.mfp23[2 * i]             = t->GetMFP(2);
    b.mfp23[2 * i + 1]         = t->GetMFP(3);
    b.range_lambtr1[2 * i]     = et.GetRange();
    b.range_lambtr1[2 * i + 1] = msc->fLambtr1;
    b.tstep_zpath[2 * i]       = msc->fTrueStepLength;
    b.tstep_zpath[2 * i + 1]   = msc->fZPathLength;
    b.par12[2 * i]             = msc->fPar1;
    b.par12[2 * i + 1]         = msc->fPar2;
    b.par3_pad[2 * i]          = msc->fPar3;
    b.prestep[2 * i]           = et.GetPreStepEKin();
    b.prestep[2 * i + 1]       = et.GetPreStepLogEKin();
  }

  // row i -> G4HepEmElectronTrack (in place, like the reference's managers)
  static void Unpack(const G4HB200ElectronBatch& b, int64_t i, G4HepEmElectronTrack& et, G4HepEmB200TrackAux& a) {
    G4HepEmTrack* t = et.GetTrack();
    G4HepEmMSCTrackData* msc = et.GetMSCTrackData();
    const double le = b.ekin_logekin[2 * i + 1];
    if (le > 99.0) t->SetEKin(b.ekin_logekin[2 * i]); else t->SetEKin(b.ekin_logekin[2 * i], le);
    t->SetDirection(b.dirx_diry[2 * i], b.dirx_diry[2 * i + 1], b.dirz_safety[2 * i]);
    t->SetNumIALeft(b.nia01[2 * i], 0);
    t->SetNumIALeft(b.nia01[2 * i + 1], 1);
    t->SetNumIALeft(b.nia23[2 * i], 2);
    t->SetNumIALeft(b.nia23[2 * i + 1], 3);
    const uint32_t flags = static_cast<uint32_t>(b.meta[4 * i + 1]);
    msc->fInitialRange        = b.msc_irange_dynrf[2 * i];
    msc->fDynamicRangeFactor  = b.msc_irange_dynrf[2 * i + 1];
    msc->fTlimitMin           = b.msc_tlimmin_gauss[2 * i];
    msc->fIsFirstStep         = (flags & G4HB200_F_MSC_FIRST_STEP) != 0u;
    msc->fIsActive            = (flags & G4HB200_F_MSC_ACTIVE) != 0u;
    msc->fIsDisplace          = (flags & G4HB200_F_MSC_DISPLACE) != 0u;
    msc->fIsNoScatteringInMSC = (flags & G4HB200_F_MSC_NO_SCATTER) != 0u;
    a.fIsGauss  = (flags & G4HB200_F_GAUSS_CACHED) != 0u;
    a.fGauss    = b.msc_tlimmin_gauss[2 * i + 1];
    a.fNextDraw = b.meta[4 * i + 3];
    t->SetGStepLength(b.gstep_pstep[2 * i]);
    et.SetPStepLength(b.gstep_pstep[2 * i + 1]);
    t->SetEnergyDeposit(b.edep_dispx[2 * i]);
    msc->SetDisplacement(b.edep_dispx[2 * i + 1], b.dispy_dispz[2 * i], b.dispy_dispz[2 * i + 1]);
    t->SetWinnerProcessIndex(b.winner[i]);
    t->SetMFP(b.mfp01[2 * i], 0);
    t->SetMFP(b.mfp01[2 * i + 1], 1);
    t->SetMFP(b.mfp23[2 * i], 2);
    t->SetMFP(b.mfp23[2 * i + 1], 3);
    et.SetRange(b.range_lambtr1[2 * i]);
    msc->fLambtr1        = b.range_lambtr1[2 * i + 1];
    msc->fTrueStepLength = b.tstep_zpath[2 * i];
    msc->fZPathLength    = b.tstep_zpath[2 * i + 1];
    msc->fPar1           = b.par12[2 * i];
    msc->fPar2           = b.par12[2 * i + 1];
    msc->fPar3           = b.par3_pad[2 * i];
    et.SetPreStepEKin(b.prestep[2 * i], b.prestep[2 * i + 1]);
  }

  static void Pack(G4HepEmGammaTrack& gt, const G4HepEmB200TrackAux& a, G4HB200GammaBatch& b, int64_t i) {
    G4HepEmTrack* t = gt.GetTrack();
    b.ekin_logekin[2 * i]     = t->GetEKin();
    b.ekin_logekin[2 * i + 1] = t->GetLogEKin();
    const double* dir = t->GetDirection();
    b.dirx_diry[2 * i]     = dir[0];
    b.dirx_diry[2 * i + 1] = dir[1];
    b.dirz_nia0[2 * i]     = dir[2];
    b.dirz_nia0[2 * i + 1] = t->GetNumIALeft(0);
    int32_t* meta = b.meta + 4 * i;
    meta[0] = t->GetMCIndex();
    meta[1] = t->GetOnBoundary() ? static_cast<int32_t>(G4HB200_F_ON_BOUNDARY) : 0;
    meta[2] = t->GetID();
    meta[3] = a.fNextDraw;
    b.gstep_mfp0[2 * i]       = t->GetGStepLength();
    b.gstep_mfp0[2 * i + 1]   = t->GetMFP(0);
    b.edep_pemxsec[2 * i]     = t->GetEnergyDeposit();
    b.edep_pemxsec[2 * i + 1] = gt.GetPEmxSec();
    b.winner[i]               = t->GetWinnerProcessIndex();
  }

  static void Unpack(const G4HB200GammaBatch& b, int64_t i, G4HepEmGammaTrack& gt, G4HepEmB200TrackAux& a) {
    G4HepEmTrack* t = gt.GetTrack();
    const double le = b.ekin_logekin[2 * i + 1];
    if (le > 99.0) t->SetEKin(b.ekin_logekin[2 * i]); else t->SetEKin(b.ekin_logekin[2 * i], le);
    t->SetDirection(b.dirx_diry[2 * i], b.dirx_diry[2 * i + 1], b.dirz_nia0[2 * i]);
    t->SetNumIALeft(b.dirz_nia0[2 * i + 1], 0);
    a.fNextDraw = b.meta[4 * i + 3];
    t->SetGStepLength(b.gstep_mfp0[2 * i]);
    t->SetMFP(b.gstep_mfp0[2 * i + 1], 0);
    t->SetEnergyDeposit(b.edep_pemxsec[2 * i]);
    gt.SetPEmxSec(b.edep_pemxsec[2 * i + 1]);
    t->SetWinnerProcessIndex(b.winner[i]);
  }

  int EnsureSecondaries(int64_t n) {
    const int64_t cap = 2 * n + 2;
    if (cap > fSecCap) {
      if (fSecCap > 0) g4hb200_secondary_queue_free(fHandle, &fSecDev);
      fSecCap = 0;
      const int rc = g4hb200_secondary_queue_alloc(fHandle, cap, &fSecDev);
      if (rc != 0) return rc;
      fSecCap = cap;
    }
    fSecHost.Resize(cap);
    return g4hb200_secondary_queue_reset(fHandle, &fSecDev, nullptr);
  }

  int FetchSecondaries(std::vector<G4HepEmB200Secondary>* out) {
    int rc = g4hb200_secondary_queue_download(fHandle, &fSecDev, &fSecHost.view, nullptr);
    if (rc != 0) return rc;
    if ((rc = g4hb200_sync(fHandle, nullptr)) != 0) return rc;
    if (out == nullptr) return 0;
    const int32_t n = fSecHost.count[0];
    out->clear();
    out->reserve(static_cast<size_t>(n));
    for (int32_t k = 0; k < n; ++k) {
      G4HepEmB200Secondary s;
      s.fDirection[0] = fSecHost.dxy[2 * k];
      s.fDirection[1] = fSecHost.dxy[2 * k + 1];
      s.fDirection[2] = fSecHost.dze[2 * k];
      s.fEKin         = fSecHost.dze[2 * k + 1];
      s.fParentID     = fSecHost.pk[2 * k];
      s.fKind         = fSecHost.pk[2 * k + 1];
      s.fParentIndex  = fSecHost.ps[2 * k];
      s.fSlot         = fSecHost.ps[2 * k + 1];
      out->push_back(s);
    }
    return 0;
  }

  int EnsureFlags(int64_t n) {
    if (n <= fFlagCap) return 0;
    if (fFlagDev != nullptr) g4hb200_device_free(fHandle, fFlagDev);
    fFlagDev = nullptr;
    fFlagCap = 0;
    void* p = nullptr;
    const int rc = g4hb200_device_alloc(fHandle, static_cast<size_t>(n) * 16, &p);
    if (rc != 0) return rc;
    fFlagDev = p;
    fFlagCap = n;
    return 0;
  }

  // mode 0: HowFar, 1: Perform, 2: fused step, 3 + op: one track-level op, -1: CheckDelta with the uniforms urnd
  int RunElectron(G4HepEmElectronTrack* tracks, G4HepEmB200TrackAux* aux, int64_t n, int mode, std::vector<G4HepEmB200Secondary>* sec,
                  int32_t* flags = nullptr, const double* urnd = nullptr) {
    if (fHandle == nullptr) return G4HB200_EINVAL;
    if (n <= 0) return 0;
    int rc = 0;
    if (mode >= 3 || mode < 0) {
      if ((rc = EnsureFlags(n)) != 0) return rc;
    }
    int32_t* flagDev = static_cast<int32_t*>(fFlagDev);
    double* urndDev  = reinterpret_cast<double*>(static_cast<char*>(fFlagDev) + 8 * (fFlagCap > 0 ? fFlagCap : 0));
    if (n > fElCap) {
      if (fElCap > 0) g4hb200_electron_batch_free(fHandle, &fElDev);
      fElCap = 0;
      if ((rc = g4hb200_electron_batch_alloc(fHandle, n, &fElDev)) != 0) return rc;
      fElCap = n;
    }
    fElHost.Resize(n);
    for (int64_t i = 0; i < n; ++i) Pack(tracks[i], aux[i], fElHost.view, i);
    if ((rc = g4hb200_electron_batch_upload(fHandle, &fElHost.view, &fElDev, nullptr)) != 0) return rc;
    if (mode != 0 && (rc = EnsureSecondaries(n)) != 0) return rc;
    if (mode == 0) rc = g4hb200_electron_howfar(fHandle, &fElDev, fSeed, nullptr);
    if (mode == 1) rc = g4hb200_electron_perform(fHandle, &fElDev, &fSecDev, fSeed, nullptr);
    if (mode == 2) rc = g4hb200_electron_step(fHandle, &fElDev, &fSecDev, fSeed, nullptr);
    if (mode >= 3) rc = g4hb200_electron_track_op(fHandle, mode - 3, &fElDev, &fSecDev, fSeed, flagDev, nullptr);
    if (mode < 0) {
      if ((rc = g4hb200_memcpy(fHandle, urndDev, urnd, static_cast<size_t>(n) * 8, 1, nullptr)) != 0) return rc;
      rc = g4hb200_electron_check_delta(fHandle, &fElDev, urndDev, flagDev, nullptr);
    }
    if (rc != 0) return rc;
    if ((rc = g4hb200_electron_batch_download(fHandle, &fElDev, &fElHost.view, nullptr)) != 0) return rc;
    if (flags != nullptr && (mode >= 3 || mode < 0)) {
      if ((rc = g4hb200_memcpy(fHandle, flags, flagDev, static_cast<size_t>(n) * 4, 0, nullptr)) != 0) return rc;
    }
    if ((rc = g4hb200_sync(fHandle, nullptr)) != 0) return rc;
    for (int64_t i = 0; i < n; ++i) Unpack(fElHost.view, i, tracks[i], aux[i]);
    return mode != 0 ? FetchSecondaries(sec) : 0;
  }

  int RunGamma(G4HepEmGammaTrack* tracks, G4HepEmB200TrackAux* aux, int64_t n, int mode, std::vector<G4HepEmB200Secondary>* sec) {
    if (fHandle == nullptr) return G4HB200_EINVAL;
    if (n <= 0) return 0;
    int rc = 0;
    if (n > fGmCap) {
      if (fGmCap > 0) g4hb200_gamma_batch_free(fHandle, &fGmDev);
      fGmCap = 0;
      if ((rc = g4hb200_gamma_batch_alloc(fHandle, n, &fGmDev)) != 0) return rc;
      fGmCap = n;
    }
    fGmHost.Resize(n);
    for (int64_t i = 0; i < n; ++i) Pack(tracks[i], aux[i], fGmHost.view, i);
    if ((rc = g4hb200_gamma_batch_upload(fHandle, &fGmHost.view, &fGmDev, nullptr)) != 0) return rc;
    if (mode != 0 && (rc = EnsureSecondaries(n)) != 0) return rc;
    if (mode == 0) rc = g4hb200_gamma_howfar(fHandle, &fGmDev, fSeed, nullptr);
    if (mode == 1) rc = g4hb200_gamma_perform(fHandle, &fGmDev, &fSecDev, fSeed, nullptr);
    if (mode == 2) rc = g4hb200_gamma_step(fHandle, &fGmDev, &fSecDev, fSeed, nullptr);
    if (mode >= 3) rc = g4hb200_gamma_track_op(fHandle, mode - 3, &fGmDev, &fSecDev, fSeed, nullptr);
    if (rc != 0) return rc;
    if ((rc = g4hb200_gamma_batch_download(fHandle, &fGmDev, &fGmHost.view, nullptr)) != 0) return rc;
    if ((rc = g4hb200_sync(fHandle, nullptr)) != 0) return rc;
    for (int64_t i = 0; i < n; ++i) Unpack(fGmHost.view, i, tracks[i], aux[i]);
    return mode != 0 ? FetchSecondaries(sec) : 0;
  }

  G4HB200* fHandle = nullptr;
  G4HepEmB200FlatTables fFlat;
  uint64_t fSeed = 0;
  G4HB200ElectronBatch fElDev;
  G4HB200GammaBatch fGmDev;
  G4HB200SecondaryQueue fSecDev;
  int64_t fElCap = 0, fGmCap = 0, fSecCap = 0, fFlagCap = 0;
  void* fFlagDev = nullptr;  // device scratch of the track-level calls: int32 flags [cap] + padding, then uniforms [cap]
  HostElectron fElHost;
  HostGamma fGmHost;
  HostSecondaries fSecHost;
};

#endif  // G4HEPEMB200_MANAGERS_HH
