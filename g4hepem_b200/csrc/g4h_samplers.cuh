// g4h_samplers.cuh -- the final-state samplers with a rejection loop, cut into three pieces each so that a warp can
// run the pieces at full width (g4h_refill.cuh):
//
//   Setup   everything in front of the rejection loop: cuts, target element, table row, the constants of the trial;
//           tells whether there is nothing to do (kSamplerDone), no loop (kSamplerFinish) or a loop (kSamplerLoop)
//   Trial   ONE pass through the body of the rejection loop with the uniforms of that pass; true = accepted
//   Finish  everything behind the loop: kinematics, directions, secondaries, the primary's new state
//
// Pars is what travels between the pieces: plain doubles only (integers are stored as doubles), so that the refill
// executor can park it in shared memory word by word.  RunSampler() is the per-track form (Setup, loop, Finish) the
// track-level ops, the fused kernels and the host pre-flight harness use: one source for both.
//
// Arithmetic: the reference's, operation by operation (same order, same uniform consumption):
//   Moller / Bhabha            G4HepEmElectronInteractionIoni.icc:19-138
//   Seltzer-Berger / rel. brem G4HepEmElectronInteractionBrem.icc:35-344 (+ LPM functions, Tsai angles G4HepEmInteractionUtils.icc:11-76)
//   Klein-Nishina Compton      G4HepEmGammaInteractionCompton.icc:17-103
//   Bethe-Heitler conversion   G4HepEmGammaInteractionConversion.icc:19-273
//   photoelectric              G4HepEmGammaInteractionPhotoelectric.icc:13-121
// Values that the reference recomputes inside its loop from loop-invariant inputs (e.g. the e+ factor of the
// Seltzer-Berger suppression) are computed once in Setup: same operands, same operations, same bits.
#ifndef G4H_SAMPLERS_CUH
#define G4H_SAMPLERS_CUH

#include "g4h_electron.cuh"

namespace g4h {

struct Secondary {
  double dir[3];
  double ekin;
  int kind;  // G4HB200_SEC_ELECTRON / _POSITRON / _GAMMA
};

struct Secondaries {
  int n;
  Secondary s[2];
};

constexpr int kSecElectron = 0, kSecPositron = 1, kSecGamma = 2;

enum SamplerNext { kSamplerDone = 0, kSamplerFinish = 1, kSamplerLoop = 2 };

struct GammaState {
  double ekin, logEkin;
  double dir[3];
  double nIA0, mfp0, gStep, edep, peMXsec;
  int imc, id, winner;
  bool onBoundary;
};

G4H_FN double GetLogEKin(GammaState& s) {
  if (s.logEkin > 99.0) {
    s.logEkin = (s.ekin > 0.) ? Log(s.ekin) : -30;
  }
  return s.logEkin;
}
G4H_FN void SetEKin(GammaState& s, double ekin) {
  s.ekin    = ekin;
  s.logEkin = 100.0;
}

// LPM G(s), Phi(s) on s in [0,2], ds = 0.05 (G4HepEmInteractionUtils.hh:21-36)
#if defined(__CUDACC__)
__device__ __constant__
#else
static const
#endif
double kFuncLPM[82] = {
  0.0000E+00, 0.0000E+00, 6.9163E-02, 2.5747E-01, 2.0597E-01, 4.4573E-01, 3.5098E-01, 5.8373E-01, 4.8095E-01, 6.8530E-01,
  5.8926E-01, 7.6040E-01, 6.7626E-01, 8.1626E-01, 7.4479E-01, 8.5805E-01, 7.9826E-01, 8.8952E-01, 8.4003E-01, 9.1338E-01,
  8.7258E-01, 9.3159E-01, 8.9794E-01, 9.4558E-01, 9.1776E-01, 9.5640E-01, 9.3332E-01, 9.6483E-01, 9.4560E-01, 9.7143E-01,
  9.5535E-01, 9.7664E-01, 9.6313E-01, 9.8078E-01, 9.6939E-01, 9.8408E-01, 9.7444E-01, 9.8673E-01, 9.7855E-01, 9.8888E-01,
  9.8191E-01, 9.9062E-01, 9.8467E-01, 9.9204E-01, 9.8695E-01, 9.9321E-01, 9.8884E-01, 9.9417E-01, 9.9042E-01, 9.9497E-01,
  9.9174E-01, 9.9564E-01, 9.9285E-01, 9.9619E-01, 9.9379E-01, 9.9666E-01, 9.9458E-01, 9.9706E-01, 9.9526E-01, 9.9739E-01,
  9.9583E-01, 9.9768E-01, 9.9632E-01, 9.9794E-01, 9.9674E-01, 9.9818E-01, 9.9710E-01, 9.9839E-01, 9.9741E-01, 9.9857E-01,
  9.9767E-01, 9.9873E-01, 9.9790E-01, 9.9887E-01, 9.9809E-01, 9.9898E-01, 9.9826E-01, 9.9909E-01, 9.9840E-01, 9.9918E-01,
  9.9856E-01, 9.9926E-01};

// ---- pieces several samplers share ------------------------------------------------------------------------------
// SampleCostModifiedTsai (InteractionUtils.icc:11-23): 3 uniforms per pass, nearly always one pass (the tail beyond
// uMax >= 2 holds 3 % of the weight at 1 MeV): stays a loop inside Finish
G4H_FN double SampleCostModifiedTsai(double thePrimEkin, Rng& rng) {
  const double uMax = 2.0 * (1.0 + thePrimEkin * kInvElectronMassC2);
  double u;
  do {
    const double r0 = rng.Flat();
    const double r1 = rng.Flat();
    const double r2 = rng.Flat();
    const double uu = -Log(r0 * r1);
    u = (0.25 > r2) ? uu * 1.6 : uu * 0.533333333;
  } while (u > uMax);
  return 1.0 - 2.0 * u * u / (uMax * uMax);
}

// EvaluateLPMFunctions (InteractionUtils.icc:28-76)
G4H_FN void EvaluateLPMFunctions(double& funcXiS, double& funcGS, double& funcPhiS, double egamma, double etotal,
                                 double elpm, double z23, double ilVarS1, double ilVarS1Cond, double densityCor, double times) {
  const double sqrt2     = 1.414213562373095;
  const double redegamma = egamma / etotal;
  const double varSprime = sqrt(0.125 * redegamma * elpm / (times * (1.0 - redegamma) * etotal));
  const double varS1     = z23 / (184.15 * 184.15);
  const double condition = sqrt2 * varS1;
  double funcXiSprime = 2.0;
  if (varSprime > 1.0) {
    funcXiSprime = 1.0;
  } else if (varSprime > condition) {
    const double funcHSprime = Log(varSprime) * ilVarS1Cond;
    funcXiSprime = 1.0 + funcHSprime - 0.08 * (1.0 - funcHSprime) * funcHSprime * (2.0 - funcHSprime) * ilVarS1Cond;
  }
  funcXiS = funcXiSprime;
  const double varS = varSprime / sqrt(funcXiSprime);
  double varShat = varS;
  if (densityCor != 0.0) {
    varShat *= (1.0 + densityCor / (egamma * egamma));
    funcXiS = 2.0;
    if (varShat > 1.0) {
      funcXiS = 1.0;
    } else if (varShat > varS1) {
      funcXiS = 1.0 + Log(varShat) * ilVarS1;
    }
  }
  const double lpmSLimit = 2.0;
  const double lpmISDelt = 20.0;
  if (varShat < lpmSLimit) {
    double val = varShat * lpmISDelt;
    int ilow   = static_cast<int>(val);
    val -= ilow;
    ilow *= 2;
    funcGS   = (kFuncLPM[ilow + 2] - kFuncLPM[ilow]) * val + kFuncLPM[ilow];
    funcPhiS = (kFuncLPM[ilow + 3] - kFuncLPM[ilow + 1]) * val + kFuncLPM[ilow + 1];
  } else {
    double ss = 1.0 / (varShat * varShat);
    ss *= ss;
    funcGS   = 1.0 - 0.0230655 * ss;
    funcPhiS = 1.0 - 0.01190476 * ss;
  }
  if (funcXiS * funcPhiS > 1.0 || varShat > 0.57) {
    funcXiS = 1.0 / funcPhiS;
  }
}

// the element parameter row of atomic number Z (G4HepEmElementData is indexed by Z, 121 slots)
G4H_FN const double* ElemParsOfZ(const TablesView& tv, int iZet) { return tv.elemPars + 12 * (iZet < 120 ? iZet : 120); }

// ---- Moller / Bhabha (Ioni.icc:19-138) --------------------------------------------------------------------------
// Ioni::SampleDirections (Ioni.icc:111-138): two-body kinematics, one uniform for the azimuth
G4H_FN void IoniSampleDirections(double thePrimEkin, double deltaEkin, double* theSecElecDir, double* thePrimElecDir, Rng& rng) {
  const double elInitETot = thePrimEkin + kElectronMassC2;
  const double elInitPTot = sqrt(thePrimEkin * (elInitETot + kElectronMassC2));
  const double deltaPTot  = sqrt(deltaEkin * (deltaEkin + 2.0 * kElectronMassC2));
  const double cost       = deltaEkin * (elInitETot + kElectronMassC2) / (deltaPTot * elInitPTot);
  const double cosTheta   = Max(-1.0, Min(cost, 1.0));
  const double sinTheta   = sqrt((1.0 - cosTheta) * (1.0 + cosTheta));
  const double phi        = k2Pi * rng.Flat();
  double sphi, cphi;
  SinCos(phi, sphi, cphi);
  theSecElecDir[0] = sinTheta * cphi;
  theSecElecDir[1] = sinTheta * sphi;
  theSecElecDir[2] = cosTheta;
  RotateToReferenceFrame(theSecElecDir, thePrimElecDir);
  thePrimElecDir[0] = elInitPTot * thePrimElecDir[0] - deltaPTot * theSecElecDir[0];
  thePrimElecDir[1] = elInitPTot * thePrimElecDir[1] - deltaPTot * theSecElecDir[1];
  thePrimElecDir[2] = elInitPTot * thePrimElecDir[2] - deltaPTot * theSecElecDir[2];
  const double norm = 1.0 / sqrt(thePrimElecDir[0] * thePrimElecDir[0] + thePrimElecDir[1] * thePrimElecDir[1] +
                                 thePrimElecDir[2] * thePrimElecDir[2]);
  thePrimElecDir[0] *= norm;
  thePrimElecDir[1] *= norm;
  thePrimElecDir[2] *= norm;
}

// what Ioni::Perform does with the sampled transfer (Ioni.icc:28-46)
G4H_FN void IoniFinish(ElectronState& s, double thePrimEkin, double deltaEkin, Rng& rng, Secondaries& sec) {
  Secondary& sc = sec.s[sec.n++];
  IoniSampleDirections(thePrimEkin, deltaEkin, sc.dir, s.dir, rng);
  SetEKin(s, thePrimEkin - deltaEkin);
  sc.ekin = deltaEkin;
  sc.kind = kSecElectron;
}

// e- e- -> e- e-: SampleETransferMoller (Ioni.icc:50-74); the transfer is sampled as a fraction of the primary's energy
struct MollerSampler {
  static constexpr int kDraws = 2;
  static constexpr int kNumResults = 1;  // trailing fields of Pars a trial writes
  struct Pars {
    double primEkin, xmin, xmax, xminmax, gg, gf;
    double delta;  // the accepted transfer (fraction)
  };
  static G4H_MFN int Setup(const TablesView& tv, ElectronState& s, Rng&, Pars& p) {
    const double primEkin = s.ekin;
    const double elCut    = G4H_LD(tv.mcCuts + 4 * s.imc + kCElCut);
    if (0.5 * primEkin <= elCut) return kSamplerDone;  // Ioni::Perform (Ioni.icc:24-27)
    const double tmax   = 0.5 * primEkin;
    const double xmin   = elCut / primEkin;
    const double xmax   = tmax / primEkin;
    const double gamma  = primEkin * kInvElectronMassC2 + 1.0;
    const double gamma2 = gamma * gamma;
    const double gg     = (2.0 * gamma - 1.0) / gamma2;
    const double y      = 1. - xmax;
    p.primEkin = primEkin;
    p.xmin     = xmin;
    p.xmax     = xmax;
    p.xminmax  = xmin * xmax;
    p.gg       = gg;
    p.gf       = 1.0 - gg * xmax + xmax * xmax * (1.0 - gg + (1.0 - gg * y) / (y * y));
    p.delta    = 0.0;
    return kSamplerLoop;
  }
  static G4H_MFN bool Trial(const TablesView&, Pars& p, const double* u) {
    const double gg = p.gg;
    const double deltaEkin = p.xminmax / (p.xmin * (1.0 - u[0]) + p.xmax * u[0]);
    const double xx  = 1.0 - deltaEkin;
    const double dum = 1.0 - gg * deltaEkin + deltaEkin * deltaEkin * (1.0 - gg + (1.0 - gg * xx) / (xx * xx));
    p.delta = deltaEkin;
    return !(p.gf * u[1] > dum);
  }
  static G4H_MFN void Finish(const TablesView&, ElectronState& s, const Pars& p, Rng& rng, Secondaries& sec) {
    IoniFinish(s, p.primEkin, p.delta * p.primEkin, rng, sec);
  }
};

// e+ e- -> e+ e-: SampleETransferBhabha (Ioni.icc:76-108)
struct BhabhaSampler {
  static constexpr int kDraws = 2;
  static constexpr int kNumResults = 1;  // trailing fields of Pars a trial writes
  struct Pars {
    double primEkin, xmin, xmax, xminmax, b1, b2, b3, b4, beta2, gf;
    double delta;
  };
  static G4H_MFN int Setup(const TablesView& tv, ElectronState& s, Rng&, Pars& p) {
    const double primEkin = s.ekin;
    const double elCut    = G4H_LD(tv.mcCuts + 4 * s.imc + kCElCut);
    if (primEkin <= elCut) return kSamplerDone;
    const double xmin   = elCut / primEkin;
    const double xmax   = primEkin / primEkin;
    const double gamma  = primEkin * kInvElectronMassC2 + 1.0;
    const double gamma2 = gamma * gamma;
    const double beta2  = 1. - 1. / gamma2;
    const double y      = 1.0 / (1.0 + gamma);
    const double y2     = y * y;
    const double y12    = 1.0 - 2.0 * y;
    const double b1     = 2.0 - y2;
    const double b2     = y12 * (3.0 + y2);
    const double y122   = y12 * y12;
    const double b4     = y122 * y12;
    const double b3     = b4 + y122;
    const double xmax2  = xmax * xmax;
    p.primEkin = primEkin;
    p.xmin     = xmin;
    p.xmax     = xmax;
    p.xminmax  = xmin * xmax;
    p.b1 = b1; p.b2 = b2; p.b3 = b3; p.b4 = b4;
    p.beta2 = beta2;
    p.gf    = 1.0 + (xmax2 * b4 - xmin * xmin * xmin * b3 + xmax2 * b2 - xmin * b1) * beta2;
    p.delta = 0.0;
    return kSamplerLoop;
  }
  static G4H_MFN bool Trial(const TablesView&, Pars& p, const double* u) {
    const double deltaEkin = p.xminmax / (p.xmin * (1.0 - u[0]) + p.xmax * u[0]);
    const double xx  = deltaEkin * deltaEkin;
    const double dum = 1.0 + (xx * xx * p.b4 - deltaEkin * xx * p.b3 + xx * p.b2 - deltaEkin * p.b1) * p.beta2;
    p.delta = deltaEkin;
    return !(p.gf * u[1] > dum);
  }
  static G4H_MFN void Finish(const TablesView&, ElectronState& s, const Pars& p, Rng& rng, Secondaries& sec) {
    IoniFinish(s, p.primEkin, p.delta * p.primEkin, rng, sec);
  }
};

// ---- bremsstrahlung (Brem.icc:35-344) -----------------------------------------------------------------------------
// Brem::SampleDirections (Brem.icc:299-322)
G4H_FN void BremSampleDirections(double thePrimEkin, double theSecGammaEkin, double* theSecGammaDir, double* thePrimElecDir, Rng& rng) {
  const double cost = SampleCostModifiedTsai(thePrimEkin, rng);
  const double sint = sqrt((1.0 - cost) * (1.0 + cost));
  const double phi  = k2Pi * rng.Flat();
  double sphi, cphi;
  SinCos(phi, sphi, cphi);
  theSecGammaDir[0] = sint * cphi;
  theSecGammaDir[1] = sint * sphi;
  theSecGammaDir[2] = cost;
  RotateToReferenceFrame(theSecGammaDir, thePrimElecDir);
  const double primETot = thePrimEkin + kElectronMassC2;
  const double primPTot = sqrt(thePrimEkin * (primETot + kElectronMassC2));
  thePrimElecDir[0] = primPTot * thePrimElecDir[0] - theSecGammaEkin * theSecGammaDir[0];
  thePrimElecDir[1] = primPTot * thePrimElecDir[1] - theSecGammaEkin * theSecGammaDir[1];
  thePrimElecDir[2] = primPTot * thePrimElecDir[2] - theSecGammaEkin * theSecGammaDir[2];
  const double norm = 1.0 / sqrt(thePrimElecDir[0] * thePrimElecDir[0] + thePrimElecDir[1] * thePrimElecDir[1] +
                                 thePrimElecDir[2] * thePrimElecDir[2]);
  thePrimElecDir[0] *= norm;
  thePrimElecDir[1] *= norm;
  thePrimElecDir[2] *= norm;
}

// what Brem::Perform does with the sampled photon energy (Brem.icc:52-69)
G4H_FN void BremFinish(ElectronState& s, double thePrimEkin, double eGamma, Rng& rng, Secondaries& sec) {
  Secondary& sc = sec.s[sec.n++];
  BremSampleDirections(thePrimEkin, eGamma, sc.dir, s.dir, rng);
  SetEKin(s, thePrimEkin - eGamma);
  sc.ekin = eGamma;
  sc.kind = kSecGamma;
}

// Brem::LinSearch (Brem.icc:328-344): first index (stride 3) whose cumulative exceeds val.  The reference scans
// the 54 kappa points linearly (up to 54 dependent loads); the cumulative is non-decreasing by construction
// (Init/src/G4HepEmElectronTableBuilder.cc:685-834), so the upper bound found by bisection (6 loads) is the same index.
G4H_FN int SBLinSearch(const double* vect, int size, double val) {
  int lo  = 0;
  int len = size;
  while (len > 0) {
    const int half = len >> 1;
    if (G4H_LD(vect + 3 * (lo + half)) > val) {
      len = half;
    } else {
      lo += half + 1;
      len -= half + 1;
    }
  }
  return 3 * lo;
}

// Seltzer-Berger: SampleETransferSB (Brem.icc:73-182).  Setup selects the target atom and the row of the sampling
// table (up to two uniforms); a trial inverts the rational-function CDF in kappa and tests the dielectric suppression
// (times the e+ factor)
struct SBSampler {
  static constexpr int kDraws = 2;
  static constexpr int kNumResults = 1;  // trailing fields of Pars a trial writes
  struct Pars {
    double primEkin, gamCut, minV, lKTrans, dielSupConst;
    double posFactor;  // alpha 2 pi Z for e+ (0: e-), iBeta1: the e+ correction of the suppression (Brem.icc:168-176)
    double iBeta1;
    double stOffset;   // index of the first (cumulative, a, b) triplet of the selected row in the SB table
    double isSimply;
    double eGamma;     // the accepted photon energy
  };
  static G4H_MFN int Setup(const TablesView& tv, ElectronState& s, Rng& rng, Pars& p) {
    const double thePrimEkin = s.ekin;
    const double theLogEkin  = GetLogEKin(s);
    const int theMCIndx      = s.imc;
    const bool iselectron    = !s.isPositron;
    const double theGamCut    = G4H_LD(tv.mcCuts + 4 * theMCIndx + kCGamCut);
    if (thePrimEkin <= theGamCut) return kSamplerDone;  // Brem::Perform (Brem.icc:43-45)
    const double theLogGamCut = G4H_LD(tv.mcCuts + 4 * theMCIndx + kCLogGamCut);
    const int imat            = G4H_LD(tv.mcImat + theMCIndx);
    const ElectronTablesView& ed = tv.el[iselectron ? 0 : 1];
    const int numElem  = G4H_LD(tv.matNumElem + imat);
    const int elemIndx = (numElem > 1) ? SelectTargetAtomBrem(ed, theMCIndx, thePrimEkin, theLogEkin, rng.Flat(), true) : 0;
    const int iZet     = G4H_LD(tv.matElemZ + G4H_LD(tv.matElemStart + imat) + elemIndx);
    const double dZet  = static_cast<double>(iZet);
    const int iStart   = G4H_LD(tv.sbStartPerZ + iZet);
    const int iGamCut  = G4H_LD(tv.sbGCutIndices + G4H_LD(tv.sbGCutStart + theMCIndx) + elemIndx);
    bool isCorner = false;
    bool isSimply = false;
    int elEnergyIndx = static_cast<int>(G4H_LD(tv.sbData + iStart + 2));
    if (thePrimEkin < G4H_LD(tv.sbElEnergy + elEnergyIndx)) {
      const double val = (theLogEkin - tv.sbLogMinElEnergy) * tv.sbILDeltaElEnergy;
      elEnergyIndx  = static_cast<int>(val);
      double pIndxH = val - elEnergyIndx;
      if (G4H_LD(tv.sbElEnergy + elEnergyIndx) <= theGamCut) {
        pIndxH   = (theLogEkin - theLogGamCut) / (G4H_LD(tv.sbLElEnergy + elEnergyIndx + 1) - theLogGamCut);
        isCorner = true;
      }
      if (rng.Flat() < pIndxH) {
        ++elEnergyIndx;
      } else if (isCorner) {
        isSimply = true;
      }
    }
    const int numKappa   = 54;
    const int minEIndx   = static_cast<int>(G4H_LD(tv.sbData + iStart + 1));
    const int numGamCuts = static_cast<int>(G4H_LD(tv.sbData + iStart + 3));
    const int sizeOneE   = static_cast<int>(numGamCuts + 3 * numKappa);
    const int iSTStart   = iStart + 4 + (elEnergyIndx - minEIndx) * sizeOneE;
    const double primETot = thePrimEkin + kElectronMassC2;
    p.primEkin     = thePrimEkin;
    p.gamCut       = theGamCut;
    p.minV         = G4H_LD(tv.sbData + iSTStart + iGamCut);
    p.lKTrans      = (theLogGamCut - theLogEkin) / (theLogGamCut - G4H_LD(tv.sbLElEnergy + elEnergyIndx));
    p.dielSupConst = G4H_LD(tv.matPars + 16 * imat + kMDensityCorFactor) * primETot * primETot;
    p.posFactor    = 0.0;
    p.iBeta1       = 0.0;
    if (!iselectron) {
      const double e1 = thePrimEkin - theGamCut;
      p.iBeta1    = (e1 + kElectronMassC2) / sqrt(e1 * (e1 + 2.0 * kElectronMassC2));
      p.posFactor = kAlpha * k2Pi * dZet;
    }
    p.stOffset = static_cast<double>(iSTStart + numGamCuts);
    p.isSimply = isSimply ? 1.0 : 0.0;
    p.eGamma   = 0.0;
    return kSamplerLoop;
  }
  static G4H_MFN bool Trial(const TablesView& tv, Pars& p, const double* u) {
    const int numKappa = 54;
    const double thePrimEkin = p.primEkin;
    const double r0 = u[0];
    double kappa = 1.0;
    if (p.isSimply == 0.0) {
      const double* stData = tv.sbData + static_cast<int>(p.stOffset);
      const double minV    = p.minV;
      const double cumRV   = r0 * (1.0 - minV) + minV;
      const int cumLIndx3  = SBLinSearch(stData, numKappa, cumRV) - 3;
      const int cumLIndx   = cumLIndx3 / 3;
      const double cumL = G4H_LD(stData + cumLIndx3);
      const double pA   = G4H_LD(stData + cumLIndx3 + 1);
      const double pB   = G4H_LD(stData + cumLIndx3 + 2);
      const double cumH = G4H_LD(stData + cumLIndx3 + 3);
      const double lKL  = G4H_LD(tv.sbLKappa + cumLIndx);
      const double lKH  = G4H_LD(tv.sbLKappa + cumLIndx + 1);
      const double dm1  = (cumRV - cumL) / (cumH - cumL);
      const double dm2  = (1.0 + pA + pB) * dm1;
      const double dm3  = 1.0 + dm1 * (pA + pB * dm1);
      const double lKappa = lKL + dm2 / dm3 * (lKH - lKL);
      kappa = Exp(lKappa * p.lKTrans);
    } else {
      kappa = 1.0 - r0 * (1.0 - p.gamCut / thePrimEkin);
    }
    const double eGamma    = kappa * thePrimEkin;
    const double invEGamma = 1.0 / eGamma;
    double suppression = 1.0 / (1.0 + p.dielSupConst * invEGamma * invEGamma);
    if (p.posFactor != 0.0) {
      const double e2     = thePrimEkin - eGamma;
      const double iBeta2 = (e2 + kElectronMassC2) / sqrt(e2 * (e2 + 2.0 * kElectronMassC2));
      const double dum    = p.posFactor * (p.iBeta1 - iBeta2);
      suppression = (dum > -12.) ? suppression * Exp(dum) : 0.;
    }
    p.eGamma = eGamma;
    return !(u[1] > suppression);
  }
  static G4H_MFN void Finish(const TablesView&, ElectronState& s, const Pars& p, Rng& rng, Secondaries& sec) {
    BremFinish(s, p.primEkin, p.eGamma, rng, sec);
  }
};

// relativistic brem with LPM: SampleETransferRB (Brem.icc:184-262)
struct RBSampler {
  static constexpr int kDraws = 2;
  static constexpr int kNumResults = 1;  // trailing fields of Pars a trial writes
  struct Pars {
    double primEkin, densityCorr, xmin, xrange, rejFuncMax, zFactor1, zFactor2, lpmEnergy;
    double zet;        // atomic number of the target
    double isLPMActive;
    double eGamma;
  };
  static G4H_MFN int Setup(const TablesView& tv, ElectronState& s, Rng& rng, Pars& p) {
    const double thePrimEkin = s.ekin;
    const double theLogEkin  = GetLogEKin(s);
    const int theMCIndx      = s.imc;
    const bool iselectron    = !s.isPositron;
    const double theGamCut = G4H_LD(tv.mcCuts + 4 * theMCIndx + kCGamCut);
    if (thePrimEkin <= theGamCut) return kSamplerDone;
    const int imat         = G4H_LD(tv.mcImat + theMCIndx);
    const double* mp       = tv.matPars + 16 * imat;
    const ElectronTablesView& ed = tv.el[iselectron ? 0 : 1];
    const int numElem  = G4H_LD(tv.matNumElem + imat);
    const int elemIndx = (numElem > 1) ? SelectTargetAtomBrem(ed, theMCIndx, thePrimEkin, theLogEkin, rng.Flat(), false) : 0;
    const int iZet     = G4H_LD(tv.matElemZ + G4H_LD(tv.matElemStart + imat) + elemIndx);
    const double dZet  = static_cast<double>(iZet);
    const double* ep   = ElemParsOfZ(tv, iZet);
    const double densityFactor = kMigdalConst * G4H_LD(mp + kMElectronDensity);
    const double lpmEnergy     = kLPMconstant * G4H_LD(mp + kMRadLength);
    const double lpmEnergyLim  = sqrt(densityFactor) * lpmEnergy;
    const double thePrimTotalE = thePrimEkin + kElectronMassC2;
    const double densityCorr   = densityFactor * thePrimTotalE * thePrimTotalE;
    const double zFactor1 = G4H_LD(ep + kEZFactor1);
    const double zFactor2 = (1. + 1. / dZet) / 12.;
    const double xmin     = Log(theGamCut * theGamCut + densityCorr);
    p.primEkin    = thePrimEkin;
    p.densityCorr = densityCorr;
    p.xmin        = xmin;
    p.xrange      = Log(thePrimEkin * thePrimEkin + densityCorr) - xmin;
    p.rejFuncMax  = zFactor1 + zFactor2;
    p.zFactor1    = zFactor1;
    p.zFactor2    = zFactor2;
    p.lpmEnergy   = lpmEnergy;
    p.zet         = dZet;
    p.isLPMActive = (thePrimTotalE > lpmEnergyLim) ? 1.0 : 0.0;
    p.eGamma      = 0.0;
    return kSamplerLoop;
  }
  static G4H_MFN bool Trial(const TablesView& tv, Pars& p, const double* u) {
    const double thePrimTotalE = p.primEkin + kElectronMassC2;
    const double densityCorr   = p.densityCorr;
    const double zFactor1 = p.zFactor1, zFactor2 = p.zFactor2;
    const int iZet     = static_cast<int>(p.zet);
    const double dZet  = p.zet;
    const double* ep   = ElemParsOfZ(tv, iZet);
    const double eGamma = sqrt(Max(Exp(p.xmin + u[0] * p.xrange) - densityCorr, 0.0));
    const double y     = eGamma / thePrimTotalE;
    const double onemy = 1. - y;
    const double dum0  = 0.25 * y * y;
    double funcVal;
    if (p.isLPMActive != 0.0) {
      double funcGS, funcPhiS, funcXiS;
      EvaluateLPMFunctions(funcXiS, funcGS, funcPhiS, eGamma, thePrimTotalE, p.lpmEnergy, G4H_LD(ep + kEZet23),
                           G4H_LD(ep + kEILVarS1), G4H_LD(ep + kEILVarS1Cond), densityCorr, 1.0);
      const double term1 = funcXiS * (dum0 * funcGS + (onemy + 2.0 * dum0) * funcPhiS);
      funcVal = term1 * zFactor1 + onemy * zFactor2;
    } else {
      const double dum1 = onemy + 3. * dum0;
      if (iZet < 5) {
        funcVal = dum1 * zFactor1 + onemy * zFactor2;
      } else {
        const double zet13 = G4H_LD(ep + kEZet13);
        const double dum2 = y / (thePrimTotalE - eGamma);
        const double gam  = dum2 * 100. * kElectronMassC2 / zet13;
        const double eps  = gam / zet13;
        const double gam2 = gam * gam;
        const double phi1 = 16.863 - 2.0 * Log(1.0 + 0.311877 * gam2) + 2.4 * Exp(-0.9 * gam) + 1.6 * Exp(-1.5 * gam);
        const double phi2 = 2.0 / (3.0 + 19.5 * gam + 18.0 * gam2);
        const double eps2 = eps * eps;
        const double psi1 = 24.34 - 2.0 * Log(1.0 + 13.111641 * eps2) + 2.8 * Exp(-8.0 * eps) + 1.2 * Exp(-29.2 * eps);
        const double psi2 = 2.0 / (3.0 + 120.0 * eps + 1200.0 * eps2);
        const double logZ = G4H_LD(ep + kELogZ);
        const double Fz   = logZ / 3. + G4H_LD(ep + kECoulomb);
        const double invZ = 1. / dZet;
        funcVal = dum1 * ((0.25 * phi1 - Fz) + (0.25 * psi1 - 2. * logZ / 3.) * invZ) + 0.125 * onemy * (phi2 + psi2 * invZ);
      }
    }
    funcVal  = Max(0.0, funcVal);
    p.eGamma = eGamma;
    return !(funcVal < p.rejFuncMax * u[1]);
  }
  static G4H_MFN void Finish(const TablesView&, ElectronState& s, const Pars& p, Rng& rng, Secondaries& sec) {
    BremFinish(s, p.primEkin, p.eGamma, rng, sec);
  }
};

// ---- Klein-Nishina Compton (Compton.icc:17-103) ---------------------------------------------------------------------
struct ComptonSampler {
  static constexpr int kDraws = 3;
  static constexpr int kNumResults = 3;  // trailing fields of Pars a trial writes
  struct Pars {
    double primEkin, kappa, eps02, al1, al2;
    double eps, oneMinusCost, sint2;  // the accepted kinematics
  };
  static G4H_MFN int Setup(const TablesView&, GammaState& s, Rng&, Pars& p) {
    const double thePrimGmE = s.ekin;
    const double theLowEnergyThreshold = 0.0001;
    if (thePrimGmE < theLowEnergyThreshold) return kSamplerDone;
    const double kappa = thePrimGmE * kInvElectronMassC2;
    const double eps0  = 1. / (1. + 2. * kappa);
    const double eps02 = eps0 * eps0;
    const double al1   = -Log(eps0);
    p.primEkin = thePrimGmE;
    p.kappa    = kappa;
    p.eps02    = eps02;
    p.al1      = al1;
    p.al2      = al1 + 0.5 * (1. - eps02);
    p.eps = 0.0; p.oneMinusCost = 0.0; p.sint2 = 0.0;
    return kSamplerLoop;
  }
  // both branches of the reference (eps from the exponential or from the square root) are evaluated and one is
  // selected, so that the lanes of a warp stay together
  static G4H_MFN bool Trial(const TablesView&, Pars& p, const double* u) {
    const double al1 = p.al1, eps02 = p.eps02;
    const bool expBranch = al1 > p.al2 * u[0];
    const double epsE  = Exp(expBranch ? -al1 * u[1] : 0.0);
    const double eps2S = eps02 + (1. - eps02) * u[1];
    const double epsS  = sqrt(eps2S);
    const double eps   = expBranch ? epsE : epsS;
    const double eps2  = expBranch ? epsE * epsE : eps2S;
    const double oneMinusCost = (1. - eps) / (eps * p.kappa);
    const double sint2 = oneMinusCost * (2. - oneMinusCost);
    const double gf    = 1. - eps * sint2 / (1. + eps2);
    p.eps = eps;
    p.oneMinusCost = oneMinusCost;
    p.sint2 = sint2;
    return !(gf < u[2]);
  }
  static G4H_MFN void Finish(const TablesView&, GammaState& s, const Pars& p, Rng& rng, Secondaries& sec) {
    const double theLowEnergyThreshold = 0.0001;
    const double thePrimGmE = p.primEkin;
    const double theOrgGmDir[3] = {s.dir[0], s.dir[1], s.dir[2]};
    const double cost = 1.0 - p.oneMinusCost;
    const double sint = sqrt(Max(0., p.sint2));
    const double phi  = k2Pi * rng.Flat();
    double sphi, cphi;
    SinCos(phi, sphi, cphi);
    s.dir[0] = sint * cphi;
    s.dir[1] = sint * sphi;
    s.dir[2] = cost;
    RotateToReferenceFrame(s.dir, theOrgGmDir);
    const double thePostGmE = thePrimGmE * p.eps;
    const double theSecElE  = thePrimGmE - thePostGmE;
    double theEnergyDeposit = 0.0;
    if (theSecElE > theLowEnergyThreshold) {
      Secondary& sc = sec.s[sec.n++];
      sc.dir[0] = thePrimGmE * theOrgGmDir[0] - thePostGmE * s.dir[0];
      sc.dir[1] = thePrimGmE * theOrgGmDir[1] - thePostGmE * s.dir[1];
      sc.dir[2] = thePrimGmE * theOrgGmDir[2] - thePostGmE * s.dir[2];
      const double norm = 1.0 / sqrt(sc.dir[0] * sc.dir[0] + sc.dir[1] * sc.dir[1] + sc.dir[2] * sc.dir[2]);
      sc.dir[0] *= norm;
      sc.dir[1] *= norm;
      sc.dir[2] *= norm;
      sc.ekin = theSecElE;
      sc.kind = kSecElectron;
    } else {
      theEnergyDeposit += theSecElE;
    }
    if (thePostGmE > theLowEnergyThreshold) {
      SetEKin(s, thePostGmE);
    } else {
      theEnergyDeposit += thePostGmE;
      SetEKin(s, 0.0);
    }
    s.edep = theEnergyDeposit;
  }
};

// ---- Bethe-Heitler conversion (Conversion.icc:19-273) -----------------------------------------------------------------
// std::pow(x, 1./3.) of Conversion.icc:188,216 for x in (0, 1): the cube root (libdevice's cbrt is a quarter of the
// instructions of its pow; the two differ from the host's pow by an ulp either way, well inside the 1e-12 of energies)
G4H_FN double CubeRoot(double x) {
#if defined(__CUDA_ARCH__)
  return cbrt(x);
#else
  return pow(x, 1. / 3.);
#endif
}

struct ConversionSampler {
  static constexpr int kDraws = 3;
  static constexpr int kNumResults = 1;  // trailing fields of Pars a trial writes
  struct Pars {
    double primEkin, deltaFactor, epsMin, epsRange, FZ, normCond, invF10, invF20, lpmEnr;
    double zet;  // atomic number of the target
    double eps;  // the accepted (or, below 2 MeV, directly sampled) energy share
  };
  // Conversion::Perform head (Conversion.icc:19-27) + SampleKinEnergies up to the loop (:56-100, :180-195)
  static G4H_MFN int Setup(const TablesView& tv, GammaState& s, Rng& rng, Pars& p) {
    const double thePrimEkin = s.ekin;
    if (thePrimEkin < 2. * kElectronMassC2) return kSamplerDone;
    const double theLogEkin = GetLogEKin(s);
    const int matIndx  = G4H_LD(tv.mcImat + s.imc);
    const int numElem  = G4H_LD(tv.matNumElem + matIndx);
    const int elemIndx = (numElem > 1) ? SelectTargetAtomConversion(tv, matIndx, thePrimEkin, theLogEkin, rng.Flat()) : 0;
    const int iZet     = G4H_LD(tv.matElemZ + G4H_LD(tv.matElemStart + matIndx) + elemIndx);
    const double* ep   = ElemParsOfZ(tv, iZet);
    const double eps0  = kElectronMassC2 / thePrimEkin;
    p.primEkin = thePrimEkin;
    p.zet      = static_cast<double>(iZet);
    p.lpmEnr   = kLPMconstant * G4H_LD(tv.matPars + 16 * matIndx + kMRadLength);
    if (thePrimEkin < 2.0) {
      p.eps = eps0 + (0.5 - eps0) * rng.Flat();
      p.deltaFactor = p.epsMin = p.epsRange = p.FZ = p.normCond = p.invF10 = p.invF20 = 0.0;
      return kSamplerFinish;
    }
    const double deltaFactor = eps0 * 136. / G4H_LD(ep + kEZet13);
    const double deltaMin    = 4. * deltaFactor;
    const double deltaMax    = (thePrimEkin < 50.0) ? G4H_LD(ep + kEDeltaMaxLow) : G4H_LD(ep + kEDeltaMaxHigh);
    const double logZ13      = 0.333333 * G4H_LD(ep + kELogZ);
    const double FZ          = (thePrimEkin < 50.0) ? 8. * logZ13 : 8. * (logZ13 + G4H_LD(ep + kECoulomb));
    const double epsp     = 0.5 - 0.5 * sqrt(1. - deltaMin / deltaMax);
    const double epsMin   = Max(eps0, epsp);
    const double epsRange = 0.5 - epsMin;
    double F10, F20;
    // ScreenFunction12 (Conversion.icc:264-273)
    if (deltaMin > 1.4) {
      F10 = 42.038 - 8.29 * Log(deltaMin + 0.958);
      F20 = F10;
    } else {
      F10 = 42.184 - deltaMin * (7.444 - 1.623 * deltaMin);
      F20 = 41.326 - deltaMin * (5.848 - 0.902 * deltaMin);
    }
    F10 -= FZ;
    F20 -= FZ;
    const double NormF1 = Max(F10 * epsRange * epsRange, 0.);
    const double NormF2 = Max(1.5 * F20, 0.);
    p.deltaFactor = deltaFactor;
    p.epsMin      = epsMin;
    p.epsRange    = epsRange;
    p.FZ          = FZ;
    p.normCond    = NormF1 / (NormF1 + NormF2);
    p.invF10      = 1. / F10;
    p.invF20      = 1. / F20;
    p.eps         = 0.0;
    return kSamplerLoop;
  }
  // SampleEnergyRateNoLPM / WithLPM (Conversion.icc:180-234), one pass: one call site for what the four branches of
  // the reference share -- the screening variable, its logarithm (ScreenFunction1/2 and ComputePhi12 take the same
  // Log(delta + 0.958) above 1.4) and the LPM functions; the branches then only combine them
  static G4H_MFN bool Trial(const TablesView& tv, Pars& p, const double* u) {
    const double thePrimEkin = p.primEkin;
    const double epsRange = p.epsRange, FZ = p.FZ;
    const bool withLPM = !(thePrimEkin < 100000.0);
    const bool first   = p.normCond > u[0];
    const double eps   = first ? 0.5 - epsRange * CubeRoot(u[1]) : p.epsMin + epsRange * u[1];
    const double delta    = p.deltaFactor / (eps * (1. - eps));
    const bool highDelta  = delta > 1.4;
    const double logDelta = Log(highDelta ? delta + 0.958 : 1.0);
    double greject;
    if (!withLPM) {
      // ScreenFunction1 / ScreenFunction2 (Conversion.icc:237-248)
      const double screen = highDelta ? 42.038 - 8.29 * logDelta
                                      : (first ? 42.184 - delta * (7.444 - 1.623 * delta) : 41.326 - delta * (5.848 - 0.902 * delta));
      greject = (screen - FZ) * (first ? p.invF10 : p.invF20);
    } else {
      // ComputePhi12 (Conversion.icc:250-262)
      const double* ep  = ElemParsOfZ(tv, static_cast<int>(p.zet));
      const double phi1 = highDelta ? 21.0190 - 4.145 * logDelta : 20.806 - delta * (3.190 - 0.5710 * delta);
      const double phi2 = highDelta ? phi1 : 20.234 - delta * (2.126 - 0.0903 * delta);
      double funcXiS, funcGS, funcPhiS;
      EvaluateLPMFunctions(funcXiS, funcGS, funcPhiS, thePrimEkin, eps * thePrimEkin, p.lpmEnr, G4H_LD(ep + kEZet23),
                           G4H_LD(ep + kEILVarS1), G4H_LD(ep + kEILVarS1Cond), 0.0, -1.0);
      greject = first ? funcXiS * ((2. * funcPhiS + funcGS) * phi1 - funcGS * phi2 - funcPhiS * FZ) * p.invF10
                      : funcXiS * ((funcPhiS + 0.5 * funcGS) * phi1 + 0.5 * funcGS * phi2 - 0.5 * (funcGS + funcPhiS) * FZ) * p.invF20;
    }
    p.eps = eps;
    return !(greject < u[2]);
  }
  // the rest of SampleKinEnergies (Conversion.icc:102-120) + SampleDirections + Perform (Conversion.icc:28-53, 123-146)
  static G4H_MFN void Finish(const TablesView&, GammaState& s, const Pars& p, Rng& rng, Secondaries& sec) {
    const double thePrimEkin = p.primEkin;
    const double eps = p.eps;
    double eTotEnergy, pTotEnergy;
    if (rng.Flat() > 0.5) {
      eTotEnergy = (1. - eps) * thePrimEkin;
      pTotEnergy = eps * thePrimEkin;
    } else {
      pTotEnergy = (1. - eps) * thePrimEkin;
      eTotEnergy = eps * thePrimEkin;
    }
    const double elKinEnergy  = Max(0., eTotEnergy - kElectronMassC2);
    const double posKinEnergy = Max(0., pTotEnergy - kElectronMassC2);
    Secondary& el  = sec.s[sec.n++];
    Secondary& pos = sec.s[sec.n++];
    const double phi = k2Pi * rng.Flat();
    double sinPhi, cosPhi;
    SinCos(phi, sinPhi, cosPhi);
    const double costEl = SampleCostModifiedTsai(elKinEnergy, rng);
    const double sintEl = sqrt((1.0 - costEl) * (1.0 + costEl));
    el.dir[0] = sintEl * cosPhi;
    el.dir[1] = sintEl * sinPhi;
    el.dir[2] = costEl;
    RotateToReferenceFrame(el.dir, s.dir);
    const double costPos = SampleCostModifiedTsai(posKinEnergy, rng);
    const double sintPos = sqrt((1.0 - costPos) * (1.0 + costPos));
    pos.dir[0] = -sintPos * cosPhi;
    pos.dir[1] = -sintPos * sinPhi;
    pos.dir[2] = costPos;
    RotateToReferenceFrame(pos.dir, s.dir);
    el.ekin  = elKinEnergy;
    el.kind  = kSecElectron;
    pos.ekin = posKinEnergy;
    pos.kind = kSecPositron;
    SetEKin(s, 0.0);
  }
};

// ---- photoelectric (Photoelectric.icc:13-121) ------------------------------------------------------------------------
// Setup picks the target atom (SelectElementBindingEnergy :41-83) and ends the photon when no electron comes out --
// which is what happens to most photons below the K edge; only photo-electrons go through the Sauter-Gavrila loop
struct PhotoelectricSampler {
  static constexpr int kDraws = 2;
  static constexpr int kNumResults = 1;  // trailing fields of Pars a trial writes
  struct Pars {
    double photoElecE, bindingEnergy, ac, a1, a2, gtmax;
    double tsam;  // the accepted 1 - cos(theta)
  };
  static G4H_MFN int Setup(const TablesView& tv, GammaState& s, Rng& rng, Pars& p) {
    const double theGammaE = s.ekin;
    const double mxsec     = s.peMXsec;
    const int theMatIndx   = G4H_LD(tv.mcImat + s.imc);
    const int numElem      = G4H_LD(tv.matNumElem + theMatIndx);
    const int elemStart    = G4H_LD(tv.matElemStart + theMatIndx);
    int ielem = 0;
    if (numElem > 1) {
      const double x = rng.Flat() * mxsec;
      double sum = 0;
      const double invE = 1 / theGammaE;
      for (int i = 0; i < numElem; i++) {
        const int z  = G4H_LD(tv.matElemZ + elemStart + i);
        const int st = G4H_LD(tv.elemSandiaStart + z);
        const double poly = SandiaPoly(tv.sandiaEnergies + st, tv.sandiaCof + 4 * st, G4H_LD(tv.elemSandiaNum + z), theGammaE, invE);
        sum += G4H_LD(tv.matElemNatoms + elemStart + i) * invE * poly;
        if (x <= sum) {
          ielem = i;
          break;
        }
      }
    }
    const double bindingEnergy = G4H_LD(tv.elemPars + 12 * G4H_LD(tv.matElemZ + elemStart + ielem) + kEKShell);
    const double theLowEnergyThreshold = 0.000001;
    const double photoElecE = theGammaE - bindingEnergy;
    if (!(photoElecE > theLowEnergyThreshold)) {
      s.edep = theGammaE;
      SetEKin(s, 0.0);
      return kSamplerDone;
    }
    // SamplePhotoElectronDirection (Sauter-Gavrila), up to the loop
    const double tau   = photoElecE * kInvElectronMassC2;
    const double gamma = 1.0 + tau;
    const double beta  = sqrt(tau * (tau + 2.0)) / gamma;
    const double ac    = (1.0 - beta) / beta;
    const double a1    = 0.5 * beta * gamma * tau * (gamma - 2.0);
    p.photoElecE    = photoElecE;
    p.bindingEnergy = bindingEnergy;
    p.ac    = ac;
    p.a1    = a1;
    p.a2    = ac + 2.0;
    p.gtmax = 2.0 * (a1 + 1.0 / ac);
    p.tsam  = 0.0;
    return kSamplerLoop;
  }
  static G4H_MFN bool Trial(const TablesView&, Pars& p, const double* u) {
    const double ac = p.ac, a2 = p.a2, r0 = u[0];
    const double tsam = 2.0 * ac * (2.0 * r0 + a2 * sqrt(r0)) / (a2 * a2 - 4.0 * r0);
    const double gtr  = (2.0 - tsam) * (p.a1 + 1.0 / (ac + tsam));
    p.tsam = tsam;
    return !(u[1] * p.gtmax > gtr);
  }
  static G4H_MFN void Finish(const TablesView&, GammaState& s, const Pars& p, Rng& rng, Secondaries& sec) {
    const double tsam = p.tsam;
    Secondary& sc = sec.s[sec.n++];
    const double costheta = 1.0 - tsam;
    const double sint = sqrt(tsam * (2.0 - tsam));
    const double phi  = k2Pi * rng.Flat();
    double sphi, cphi;
    SinCos(phi, sphi, cphi);
    sc.dir[0] = sint * cphi;
    sc.dir[1] = sint * sphi;
    sc.dir[2] = costheta;
    RotateToReferenceFrame(sc.dir, s.dir);
    sc.ekin = p.photoElecE;
    sc.kind = kSecElectron;
    s.edep  = p.bindingEnergy;
    SetEKin(s, 0.0);
  }
};

// ---- the per-track form: Setup, the rejection loop, Finish ------------------------------------------------------------
template <class S, class Track>
G4H_FN void RunSampler(const TablesView& tv, Track& s, Rng& rng, Secondaries& sec) {
  typename S::Pars p;
  const int next = S::Setup(tv, s, rng, p);
  if (next == kSamplerDone) return;
  if (next == kSamplerLoop) {
    double u[3];
    do {
      u[0] = rng.Flat();
      u[1] = rng.Flat();
      if (S::kDraws > 2) u[2] = rng.Flat();
    } while (!S::Trial(tv, p, u));
  }
  S::Finish(tv, s, p, rng, sec);
}

}  // namespace g4h
#endif
