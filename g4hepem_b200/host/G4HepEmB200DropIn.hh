// G4HepEmB200DropIn.hh -- the reference's manager classes, signature for signature, with the physics on the B200.
//
//   G4HepEmB200ElectronManager  <->  G4HepEmElectronManager  (G4HepEm/G4HepEmRun/include/G4HepEmElectronManager.hh:71-275)
//   G4HepEmB200GammaManager     <->  G4HepEmGammaManager     (G4HepEm/G4HepEmRun/include/G4HepEmGammaManager.hh:32-55)
//
// Every static below has the argument list of its namesake and works in place on the same objects: the primary track of
// the G4HepEmTLData (or the track handed in), the secondaries appended to the G4HepEmTLData with
// AddSecondaryElectronTrack() / AddSecondaryGammaTrack() exactly as the reference's models do
// (G4HepEmRun/include/G4HepEmTLData.hh:52-82).  A caller written against the reference -- the two-call protocol of
// apps/examples/TestEm3/src/G4HepEmProcess.cc:106-217 or the piece-by-piece order of G4HepEmTrackingManager::TrackElectron /
// TrackGamma (G4HepEm/G4HepEm/src/G4HepEmTrackingManager.cc:428-665, 985-1140) -- compiles against these classes unchanged
// (define G4HEPEMB200_REPLACE_MANAGERS before including this header to get the reference's class names as aliases).
//
// Each call forwards a ONE-track batch through the C-ABI (pack -> upload -> kernel -> download -> unpack): this is the
// functional drop-in, with the latency of a kernel launch per call.  Throughput comes from handing whole arrays of tracks
// to G4HepEmB200Session (G4HepEmB200Managers.hh) or to the C-ABI directly.
//
// Set-up:   G4HepEmB200DropIn::Attach(data, pars, device);            once, replaces CopyG4HepEmDataToGPU
// Random numbers: the kernels draw from the counter based stream of G4HepEmB200Stream.h.  The G4HepEmRandomEngine of a worker
// must be built on a G4HepEmB200Stream (engine.fObject = &stream, flat() = G4HepEmB200StreamNext) and registered with
// G4HepEmB200DropIn::BindEngine(&engine, &stream): fObject is private in the reference, the binding is how a static that
// receives `G4HepEmRandomEngine*` finds the stream.  The stream carries seed, track id (= G4HepEmTrack::fID), next draw and
// the Gauss cache; after each call it is where the reference's engine would be.
//
// Include after the reference's headers (G4HepEmData.hh, G4HepEmParameters.hh, G4HepEmTLData.hh, G4HepEmElectronTrack.hh,
// G4HepEmGammaTrack.hh, G4HepEmRandomEngine.hh); link with libg4hepem_b200.so.
#ifndef G4HEPEMB200_DROPIN_HH
#define G4HEPEMB200_DROPIN_HH

#include <map>
#include <vector>

#include "G4HepEmB200Managers.hh"
#include "G4HepEmB200Stream.h"

class G4HepEmB200DropIn {
 public:
  // flatten + upload the tables (replaces CopyG4HepEmDataToGPU, G4HepEmData/src/G4HepEmData.cc:78-101); 0 or a G4HB200_E* code
  static int Attach(const G4HepEmData* data, const G4HepEmParameters* pars, int device = 0) {
    State& st = Get();
    st.data = data;
    return st.session.Open(data, pars, device, 0);
  }
  static void Detach() {
    State& st = Get();
    st.session.Close();
    st.engines.clear();
    st.data = nullptr;
  }
  static void BindEngine(G4HepEmRandomEngine* engine, G4HepEmB200Stream* stream) { Get().engines[engine] = stream; }
  static void UnbindEngine(G4HepEmRandomEngine* engine) { Get().engines.erase(engine); }
  static const char* LastError() { return g4hb200_last_error(); }
  // status of the last forwarded call (the reference's statics return void / bool: errors cannot travel in the signature)
  static int LastStatus() { return Get().status; }

  // ---- used by the manager classes below
  static G4HepEmB200Stream* StreamOf(G4HepEmRandomEngine* engine) {
    State& st = Get();
    auto it = st.engines.find(engine);
    return it != st.engines.end() ? it->second : nullptr;
  }
  static G4HepEmB200TrackAux ToAux(const G4HepEmB200Stream* s) {
    G4HepEmB200TrackAux a;
    if (s != nullptr) {
      a.fNextDraw = static_cast<int32_t>(s->draw);
      a.fIsGauss  = s->is_gauss != 0;
      a.fGauss    = s->gauss;
    }
    return a;
  }
  static void FromAux(const G4HepEmB200TrackAux& a, G4HepEmB200Stream* s) {
    if (s == nullptr) return;
    s->draw     = static_cast<uint32_t>(a.fNextDraw);
    s->is_gauss = a.fIsGauss ? 1 : 0;
    s->gauss    = a.fGauss;
  }
  // one electron track through one track-level op (G4HB200_OP_*); returns the bool of the reference's function
  static bool ElectronOp(int op, G4HepEmElectronTrack* track, G4HepEmRandomEngine* engine, G4HepEmTLData* tlData = nullptr) {
    State& st = Get();
    G4HepEmB200Stream* stream = StreamOf(engine);
    G4HepEmB200TrackAux aux = ToAux(stream);
    if (stream != nullptr) {
      st.session.SetSeed(stream->seed);
      // the stream is keyed by the track id the kernels see
      stream->track_id = static_cast<uint32_t>(track->GetTrack()->GetID());
    }
    int32_t flag = 0;
    std::vector<G4HepEmB200Secondary> sec;
    st.status = st.session.ElectronTrackOp(op, track, &aux, 1, tlData != nullptr ? &sec : nullptr, &flag);
    FromAux(aux, stream);
    if (tlData != nullptr) PushSecondaries(sec, track->GetTrack()->GetID(), tlData);
    return flag != 0;
  }
  static void GammaOp(int op, G4HepEmGammaTrack* track, G4HepEmRandomEngine* engine, G4HepEmTLData* tlData = nullptr) {
    State& st = Get();
    G4HepEmB200Stream* stream = StreamOf(engine);
    G4HepEmB200TrackAux aux = ToAux(stream);
    if (stream != nullptr) {
      st.session.SetSeed(stream->seed);
      stream->track_id = static_cast<uint32_t>(track->GetTrack()->GetID());
    }
    std::vector<G4HepEmB200Secondary> sec;
    st.status = st.session.GammaTrackOp(op, track, &aux, 1, tlData != nullptr ? &sec : nullptr);
    FromAux(aux, stream);
    if (tlData != nullptr) PushSecondaries(sec, track->GetTrack()->GetID(), tlData);
  }
  // the two-call entry points on the primary of a G4HepEmTLData: mode 0 HowFar, 1 Perform (as G4HepEmB200Session numbers them)
  static void ElectronCall(int mode, G4HepEmTLData* tlData) {
    State& st = Get();
    G4HepEmElectronTrack* track = tlData->GetPrimaryElectronTrack();
    G4HepEmB200Stream* stream   = StreamOf(tlData->GetRNGEngine());
    G4HepEmB200TrackAux aux     = ToAux(stream);
    if (stream != nullptr) {
      st.session.SetSeed(stream->seed);
      stream->track_id = static_cast<uint32_t>(track->GetTrack()->GetID());
    }
    std::vector<G4HepEmB200Secondary> sec;
    st.status = mode == 0 ? st.session.ElectronHowFar(track, &aux, 1) : st.session.ElectronPerform(track, &aux, 1, &sec);
    FromAux(aux, stream);
    PushSecondaries(sec, track->GetTrack()->GetID(), tlData);
  }
  static void GammaCall(int mode, G4HepEmTLData* tlData) {
    State& st = Get();
    G4HepEmGammaTrack* track  = tlData->GetPrimaryGammaTrack();
    G4HepEmB200Stream* stream = StreamOf(tlData->GetRNGEngine());
    G4HepEmB200TrackAux aux   = ToAux(stream);
    if (stream != nullptr) {
      st.session.SetSeed(stream->seed);
      stream->track_id = static_cast<uint32_t>(track->GetTrack()->GetID());
    }
    st.status = st.session.GammaHowFar(track, &aux, 1);
    (void)mode;
    FromAux(aux, stream);
  }
  static bool CheckDelta(G4HepEmTrack* theTrack, double rand) {
    // CheckDelta reads the track only (energy, couple, charge, winner, its mean free path): a scratch electron track carries it
    State& st = Get();
    G4HepEmElectronTrack scratch;
    *scratch.GetTrack() = *theTrack;
    G4HepEmB200TrackAux aux;
    int32_t flag = 0;
    st.status = st.session.ElectronCheckDelta(&scratch, &aux, 1, &rand, &flag);
    // CheckDelta caches the logarithm of the energy in the track (GetLogEKin)
    theTrack->SetEKin(scratch.GetTrack()->GetEKin(), scratch.GetTrack()->GetLogEKin());
    return flag != 0;
  }

 private:
  struct State {
    G4HepEmB200Session session;
    const G4HepEmData* data = nullptr;
    std::map<G4HepEmRandomEngine*, G4HepEmB200Stream*> engines;
    int status = 0;
  };
  static State& Get() {
    static State st;
    return st;
  }
  // secondaries -> G4HepEmTLData, the way the reference's models hand them back (e.g. G4HepEmElectronInteractionIoni.icc:35-46):
  // e-/e+ into the electron buffer, gammas into the gamma buffer, in the order they were created
  static void PushSecondaries(const std::vector<G4HepEmB200Secondary>& sec, int parentID, G4HepEmTLData* tlData) {
    for (int slot = 0; slot < 2; ++slot) {
      for (const G4HepEmB200Secondary& s : sec) {
        if (s.fSlot != slot) continue;
        G4HepEmTrack* t = nullptr;
        if (s.fKind == G4HB200_SEC_GAMMA) {
          t = tlData->AddSecondaryGammaTrack()->GetTrack();
        } else {
          G4HepEmElectronTrack* et = tlData->AddSecondaryElectronTrack();
          t = et->GetTrack();
          t->ReSet();
          t->SetCharge(s.fKind == G4HB200_SEC_POSITRON ? +1.0 : -1.0);
        }
        if (s.fKind == G4HB200_SEC_GAMMA) t->ReSet();
        t->SetDirection(s.fDirection[0], s.fDirection[1], s.fDirection[2]);
        t->SetEKin(s.fEKin);
        t->SetParentID(parentID);
      }
    }
  }
};

class G4HepEmB200ElectronManager {
 private:
  G4HepEmB200ElectronManager() = delete;

 public:
  // G4HepEmElectronManager.hh:71
  static void HowFar(struct G4HepEmData* /*hepEmData*/, struct G4HepEmParameters* /*hepEmPars*/, G4HepEmTLData* tlData) {
    G4HepEmB200DropIn::ElectronCall(0, tlData);
  }
  // .hh:90
  static void HowFarToDiscreteInteraction(struct G4HepEmData*, struct G4HepEmParameters*, G4HepEmElectronTrack* theElTrack) {
    G4HepEmB200DropIn::ElectronOp(G4HB200_OP_HOWFAR_DISCRETE, theElTrack, nullptr);
  }
  // .hh:108
  static void HowFarToMSC(struct G4HepEmData*, struct G4HepEmParameters*, G4HepEmElectronTrack* theElTrack, G4HepEmRandomEngine* rnge) {
    G4HepEmB200DropIn::ElectronOp(G4HB200_OP_HOWFAR_MSC, theElTrack, rnge);
  }
  // .hh:126: HowFarToDiscreteInteraction + HowFarToMSC (.icc:165-168)
  static void HowFar(struct G4HepEmData* d, struct G4HepEmParameters* p, G4HepEmElectronTrack* theElTrack, G4HepEmRandomEngine* rnge) {
    HowFarToDiscreteInteraction(d, p, theElTrack);
    HowFarToMSC(d, p, theElTrack, rnge);
  }
  // .hh:133
  static void UpdatePStepLength(G4HepEmElectronTrack* theElTrack) { G4HepEmB200DropIn::ElectronOp(G4HB200_OP_UPDATE_PSTEP, theElTrack, nullptr); }
  // .hh:140
  static void UpdateNumIALeft(G4HepEmElectronTrack* theElTrack) { G4HepEmB200DropIn::ElectronOp(G4HB200_OP_UPDATE_NIA, theElTrack, nullptr); }
  // .hh:149
  static bool ApplyMeanEnergyLoss(struct G4HepEmData*, struct G4HepEmParameters*, G4HepEmElectronTrack* theElTrack) {
    return G4HepEmB200DropIn::ElectronOp(G4HB200_OP_MEAN_ELOSS, theElTrack, nullptr);
  }
  // .hh:158
  static void SampleMSC(struct G4HepEmData*, struct G4HepEmParameters*, G4HepEmElectronTrack* theElTrack, G4HepEmRandomEngine* rnge) {
    G4HepEmB200DropIn::ElectronOp(G4HB200_OP_SAMPLE_MSC, theElTrack, rnge);
  }
  // .hh:167
  static bool SampleLossFluctuations(struct G4HepEmData*, struct G4HepEmParameters*, G4HepEmElectronTrack* theElTrack,
                                     G4HepEmRandomEngine* rnge) {
    return G4HepEmB200DropIn::ElectronOp(G4HB200_OP_LOSS_FLUCT, theElTrack, rnge);
  }
  // .hh:184
  static bool PerformContinuous(struct G4HepEmData*, struct G4HepEmParameters*, G4HepEmElectronTrack* theElTrack, G4HepEmRandomEngine* rnge) {
    return G4HepEmB200DropIn::ElectronOp(G4HB200_OP_PERFORM_CONTINUOUS, theElTrack, rnge);
  }
  // .hh:195
  static bool CheckDelta(struct G4HepEmData*, G4HepEmTrack* theTrack, double rand) { return G4HepEmB200DropIn::CheckDelta(theTrack, rand); }
  // .hh:206
  static void PerformDiscrete(struct G4HepEmData*, struct G4HepEmParameters*, G4HepEmTLData* tlData) {
    G4HepEmB200DropIn::ElectronOp(G4HB200_OP_DISCRETE, tlData->GetPrimaryElectronTrack(), tlData->GetRNGEngine(), tlData);
  }
  // .hh:222
  static void Perform(struct G4HepEmData*, struct G4HepEmParameters*, G4HepEmTLData* tlData) { G4HepEmB200DropIn::ElectronCall(1, tlData); }
  // G4HepEmPositronInteractionAnnihilation::Perform(tlData, isatrest = true) (G4HepEmPositronInteractionAnnihilation.hh:21), which
  // TrackElectron calls itself for a stopped e+ (G4HepEmTrackingManager.cc:621)
  static void AnnihilateAtRest(G4HepEmTLData* tlData) {
    G4HepEmB200DropIn::ElectronOp(G4HB200_OP_ANNIHILATE_AT_REST, tlData->GetPrimaryElectronTrack(), tlData->GetRNGEngine(), tlData);
  }
};

class G4HepEmB200GammaManager {
 private:
  G4HepEmB200GammaManager() = delete;

 public:
  // G4HepEmGammaManager.hh:32
  static void HowFar(struct G4HepEmData*, struct G4HepEmParameters*, G4HepEmTLData* tlData) { G4HepEmB200DropIn::GammaCall(0, tlData); }
  // .hh:34
  static void HowFar(struct G4HepEmData*, struct G4HepEmParameters*, G4HepEmGammaTrack* theGammaTrack) {
    G4HepEmB200DropIn::GammaOp(G4HB200_GOP_HOWFAR_TRACK, theGammaTrack, nullptr);
  }
  // .hh:39: the interaction must have been selected (SelectInteraction) unless the step ended on a boundary
  static void Perform(struct G4HepEmData*, struct G4HepEmParameters*, G4HepEmTLData* tlData) {
    G4HepEmB200DropIn::GammaOp(G4HB200_GOP_PERFORM_SELECTED, tlData->GetPrimaryGammaTrack(), tlData->GetRNGEngine(), tlData);
  }
  // .hh:41 (takes the G4HepEmTrack of a gamma track: the batch is built around a scratch gamma track)
  static void UpdateNumIALeft(G4HepEmTrack* theTrack) {
    G4HepEmGammaTrack scratch;
    *scratch.GetTrack() = *theTrack;
    G4HepEmB200DropIn::GammaOp(G4HB200_GOP_UPDATE_NIA, &scratch, nullptr);
    theTrack->SetNumIALeft(scratch.GetTrack()->GetNumIALeft(0), 0);
  }
  // .hh:51
  static void SelectInteraction(const struct G4HepEmData*, G4HepEmTLData* tlData) {
    G4HepEmB200DropIn::GammaOp(G4HB200_GOP_SELECT_INTERACTION, tlData->GetPrimaryGammaTrack(), tlData->GetRNGEngine());
  }
};

#ifdef G4HEPEMB200_REPLACE_MANAGERS
#define G4HepEmElectronManager G4HepEmB200ElectronManager
#define G4HepEmGammaManager G4HepEmB200GammaManager
#endif

#endif  // G4HEPEMB200_DROPIN_HH
