// g4h_interactions.cuh -- the discrete interactions of one track: dispatch and the samplers without a rejection
// loop worth cutting up.
//
// The samplers with a rejection loop (Moller / Bhabha, Seltzer-Berger / relativistic brem, Compton, conversion,
// photoelectric) live in g4h_samplers.cuh as Setup / Trial / Finish pieces; Perform*() below run them one track at a
// time (RunSampler), the queue kernels run them a warp at a time with lane refill (g4h_refill.cuh).  Here:
//   e+ annihilation            G4HepEmPositronInteractionAnnihilation.icc:15-119
//   PerformDiscrete / Perform  G4HepEmElectronManager.icc:425-483
//   gamma HowFar / Perform     G4HepEmGammaManager.icc:27-105,173-219
// Secondaries are returned in registers (at most two per interaction); the kernels append them to
// the secondary queue.
#ifndef G4H_INTERACTIONS_CUH
#define G4H_INTERACTIONS_CUH

#include "g4h_samplers.cuh"

namespace g4h {

// Ioni::Perform (Ioni.icc:19-47)
G4H_FN void PerformIoni(const TablesView& tv, ElectronState& s, Rng& rng, Secondaries& sec) {
  if (s.isPositron) {
    RunSampler<BhabhaSampler>(tv, s, rng, sec);
  } else {
    RunSampler<MollerSampler>(tv, s, rng, sec);
  }
}

// Brem::Perform (Brem.icc:35-70)
G4H_FN void PerformBrem(const TablesView& tv, ElectronState& s, Rng& rng, Secondaries& sec, bool isSBmodel) {
  if (isSBmodel) {
    RunSampler<SBSampler>(tv, s, rng, sec);
  } else {
    RunSampler<RBSampler>(tv, s, rng, sec);
  }
}

// ---- e+ annihilation ----------------------------------------------------------------------------------
// AnnihilateAtRest (Annihilation.icc:23-50)
G4H_FN void AnnihilateAtRest(Rng& rng, Secondaries& sec) {
  const double cost = 2. * rng.Flat() - 1.;
  const double sint = sqrt((1. - cost) * (1. + cost));
  const double phi  = k2Pi * rng.Flat();
  double sphi, cphi;
  SinCos(phi, sphi, cphi);
  Secondary& g1 = sec.s[sec.n++];
  Secondary& g2 = sec.s[sec.n++];
  g1.dir[0] = sint * cphi;
  g1.dir[1] = sint * sphi;
  g1.dir[2] = cost;
  g2.dir[0] = -g1.dir[0];
  g2.dir[1] = -g1.dir[1];
  g2.dir[2] = -g1.dir[2];
  g1.ekin = kElectronMassC2;
  g2.ekin = kElectronMassC2;
  g1.kind = kSecGamma;
  g2.kind = kSecGamma;
}

// SampleEnergyAndDirectionsInFlight + AnnihilateInFlight (Annihilation.icc:52-119)
G4H_FN void AnnihilateInFlight(ElectronState& s, Rng& rng, Secondaries& sec) {
  const double thePrimEkin = s.ekin;
  const double tau     = thePrimEkin * kInvElectronMassC2;
  const double gam     = tau + 1.0;
  const double tau2    = tau + 2.0;
  const double sqgrate = sqrt(tau / tau2) * 0.5;
  const double epsmin  = 0.5 - sqgrate;
  const double epsmax  = 0.5 + sqgrate;
  const double epsqot  = epsmax / epsmin;
  const double tau4    = tau2 * tau2;
  double eps   = 0.0;
  double rfunc = 0.0;
  double r1;
  do {
    const double r0 = rng.Flat();
    r1 = rng.Flat();
    eps   = epsmin * Exp(Log(epsqot) * r0);
    rfunc = 1. - eps + (2. * gam * eps - 1.) / (eps * tau4);
  } while (rfunc < r1);
  const double sqg2m1 = sqrt(tau * tau2);
  const double cost   = Min(1., Max(-1., (eps * tau2 - 1.) / (eps * sqg2m1)));
  const double sint   = sqrt((1. + cost) * (1. - cost));
  const double phi    = k2Pi * rng.Flat();
  double sphi, cphi;
  SinCos(phi, sphi, cphi);
  const double initEt = thePrimEkin + 2. * kElectronMassC2;
  const double ekinG1 = eps * initEt;
  Secondary& g1 = sec.s[sec.n++];
  Secondary& g2 = sec.s[sec.n++];
  g1.ekin   = ekinG1;
  g1.dir[0] = sint * cphi;
  g1.dir[1] = sint * sphi;
  g1.dir[2] = cost;
  RotateToReferenceFrame(g1.dir, s.dir);
  g2.ekin = initEt - ekinG1;
  const double initPt = sqrt(thePrimEkin * (thePrimEkin + 2 * kElectronMassC2));
  const double px = initPt * s.dir[0] - g1.dir[0] * ekinG1;
  const double py = initPt * s.dir[1] - g1.dir[1] * ekinG1;
  const double pz = initPt * s.dir[2] - g1.dir[2] * ekinG1;
  const double norm = 1.0 / sqrt(px * px + py * py + pz * pz);
  g2.dir[0] = px * norm;
  g2.dir[1] = py * norm;
  g2.dir[2] = pz * norm;
  g1.kind = kSecGamma;
  g2.kind = kSecGamma;
  SetEKin(s, 0.0);
}

// PerformDiscrete (G4HepEmElectronManager.icc:425-459)
G4H_FN void PerformDiscrete(const TablesView& tv, ElectronState& s, Rng& rng, Secondaries& sec) {
  const int iDProc = s.winner;
  if (iDProc < 0 || s.onBoundary) {
    return;
  }
  s.nIA[iDProc] = -1.0;
  if (CheckDelta(tv, s, rng.Flat())) {
    return;
  }
  const double theEkin = s.ekin;
  switch (iDProc) {
    case 0:
      PerformIoni(tv, s, rng, sec);
      break;
    case 1:
      PerformBrem(tv, s, rng, sec, theEkin < tv.bremModelLim);
      break;
    case 2:
      AnnihilateInFlight(s, rng, sec);
      break;
    case 3:
      break;
  }
}

// G4HepEmElectronManager::Perform (.icc:461-483)
G4H_FN void ElectronPerform(const TablesView& tv, ElectronState& s, Rng& rng, Secondaries& sec) {
  s.edep  = 0;
  s.pStep = s.gStep;
  if (s.gStep <= 0.) return;
  const bool stopped = PerformContinuous(tv, s, rng);
  if (stopped) {
    if (s.isPositron) {
      AnnihilateAtRest(rng, sec);
    }
    return;
  }
  PerformDiscrete(tv, s, rng, sec);
}

// ---- gamma ----------------------------------------------------------------------------------------------
// G4HepEmGammaManager::HowFar (G4HepEmGammaManager.icc:27-48)
G4H_FN void GammaHowFar(const TablesView& tv, GammaState& s, Rng& rng) {
  if (s.nIA0 <= 0.0) {
    s.nIA0 = -Log(rng.Flat());
  }
  const double lekin   = GetLogEKin(s);
  const int imat       = G4H_LD(tv.mcImat + s.imc);
  const double totMXSec = GammaTotalMacXSec(tv, imat, s.ekin, lekin, s.peMXsec);
  const double totalMFP = (totMXSec > 0.) ? 1. / totMXSec : kALargeValue;
  s.mfp0  = totalMFP;
  s.gStep = totalMFP * s.nIA0;
}


G4H_FN void PerformCompton(const TablesView& tv, GammaState& s, Rng& rng, Secondaries& sec) {
  RunSampler<ComptonSampler>(tv, s, rng, sec);
}
G4H_FN void PerformConversion(const TablesView& tv, GammaState& s, Rng& rng, Secondaries& sec) {
  RunSampler<ConversionSampler>(tv, s, rng, sec);
}
G4H_FN void PerformPhotoelectric(const TablesView& tv, GammaState& s, Rng& rng, Secondaries& sec) {
  RunSampler<PhotoelectricSampler>(tv, s, rng, sec);
}

// G4HepEmGammaManager::SelectInteraction (if not on boundary) + Perform (G4HepEmGammaManager.icc:54-105,173-219)
G4H_FN void GammaPerform(const TablesView& tv, GammaState& s, Rng& rng, Secondaries& sec) {
  if (!s.onBoundary) {
    // SelectInteraction -> SampleInteraction
    const double urnd = rng.Flat();
    s.nIA0 = -1.0;
    const double lekin = (s.ekin > tv.gmEMax1) ? GetLogEKin(s) : 0.0;
    s.winner = GammaSampleInteraction(tv, G4H_LD(tv.mcImat + s.imc), s.ekin, lekin, s.mfp0, urnd, s.peMXsec);
  }
  // UpdateNumIALeft
  s.nIA0 -= s.gStep / s.mfp0;
  s.edep = 0.0;
  if (s.onBoundary) {
    return;
  }
  const int iDProc = s.winner;
  // SetNumIALeft(-1.0, iDProc): only slot 0 (the total) is live state for a gamma; slots 1-3 stay at -1
  if (iDProc == 0) s.nIA0 = -1.0;
  switch (iDProc) {
    case 0:
      PerformConversion(tv, s, rng, sec);
      break;
    case 1:
      PerformCompton(tv, s, rng, sec);
      break;
    case 2:
      PerformPhotoelectric(tv, s, rng, sec);
      break;
    case 3:
      break;
  }
  const double finalEkin = s.ekin;
  if (finalEkin > 0.0 && finalEkin <= tv.gammaTrackingCut) {
    SetEKin(s, 0.0);
    s.edep += finalEkin;
  }
}

}  // namespace g4h
#endif