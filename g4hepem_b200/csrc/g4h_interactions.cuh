// g4h_interactions.cuh -- final-state samplers of the discrete interactions, one track per thread.
//
// Restates (same operation order, same uniform consumption):
//   Moller / Bhabha            G4HepEmElectronInteractionIoni.icc:19-138
//   Seltzer-Berger / rel. brem G4HepEmElectronInteractionBrem.icc:35-344, LPM functions and modified
//                              Tsai angles G4HepEmInteractionUtils.icc:11-76 (+ table :21-36 of the .hh)
//   e+ annihilation            G4HepEmPositronInteractionAnnihilation.icc:15-119
//   Klein-Nishina Compton      G4HepEmGammaInteractionCompton.icc:17-103
//   Bethe-Heitler conversion   G4HepEmGammaInteractionConversion.icc:19-273
//   photoelectric              G4HepEmGammaInteractionPhotoelectric.icc:13-121
// Secondaries are returned in registers (at most two per interaction); the kernels append them to
// the secondary queue.
#ifndef G4H_INTERACTIONS_CUH
#define G4H_INTERACTIONS_CUH

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

// ---- ionisation ---------------------------------------------------------------------------------------
// SampleETransferMoller (Ioni.icc:50-74)
G4H_FN double SampleETransferMoller(double elCut, double primEkin, Rng& rng) {
  const double tmin    = elCut;
  const double tmax    = 0.5 * primEkin;
  const double xmin    = tmin / primEkin;
  const double xmax    = tmax / primEkin;
  const double gamma   = primEkin * kInvElectronMassC2 + 1.0;
  const double gamma2  = gamma * gamma;
  const double xminmax = xmin * xmax;
  const double gg = (2.0 * gamma - 1.0) / gamma2;
  const double y  = 1. - xmax;
  const double gf = 1.0 - gg * xmax + xmax * xmax * (1.0 - gg + (1.0 - gg * y) / (y * y));
  double dum;
  double deltaEkin = 0.;
  double r1;
  do {
    const double r0 = rng.Flat();
    r1 = rng.Flat();
    deltaEkin       = xminmax / (xmin * (1.0 - r0) + xmax * r0);
    const double xx = 1.0 - deltaEkin;
    dum = 1.0 - gg * deltaEkin + deltaEkin * deltaEkin * (1.0 - gg + (1.0 - gg * xx) / (xx * xx));
  } while (gf * r1 > dum);
  return deltaEkin * primEkin;
}

// SampleETransferBhabha (Ioni.icc:76-108)
G4H_FN double SampleETransferBhabha(double elCut, double primEkin, Rng& rng) {
  const double tmin    = elCut;
  const double tmax    = primEkin;
  const double xmin    = tmin / primEkin;
  const double xmax    = tmax / primEkin;
  const double gamma   = primEkin * kInvElectronMassC2 + 1.0;
  const double gamma2  = gamma * gamma;
  const double beta2   = 1. - 1. / gamma2;
  const double xminmax = xmin * xmax;
  const double y    = 1.0 / (1.0 + gamma);
  const double y2   = y * y;
  const double y12  = 1.0 - 2.0 * y;
  const double b1   = 2.0 - y2;
  const double b2   = y12 * (3.0 + y2);
  const double y122 = y12 * y12;
  const double b4   = y122 * y12;
  const double b3   = b4 + y122;
  const double xmax2 = xmax * xmax;
  const double gf = 1.0 + (xmax2 * b4 - xmin * xmin * xmin * b3 + xmax2 * b2 - xmin * b1) * beta2;
  double dum;
  double deltaEkin = 0.;
  double r1;
  do {
    const double r0 = rng.Flat();
    r1 = rng.Flat();
    deltaEkin       = xminmax / (xmin * (1.0 - r0) + xmax * r0);
    const double xx = deltaEkin * deltaEkin;
    dum = 1.0 + (xx * xx * b4 - deltaEkin * xx * b3 + xx * b2 - deltaEkin * b1) * beta2;
  } while (gf * r1 > dum);
  return deltaEkin * primEkin;
}

// Ioni::SampleDirections (Ioni.icc:111-138)
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

// Ioni::Perform (Ioni.icc:19-47)
G4H_FN void PerformIoni(const TablesView& tv, ElectronState& s, Rng& rng, Secondaries& sec) {
  const bool iselectron = !s.isPositron;
  const double thePrimEkin = s.ekin;
  const double theElCut    = G4H_LD(tv.mcCuts + 4 * s.imc + kCElCut);
  const double maxETransfer = iselectron ? 0.5 * thePrimEkin : thePrimEkin;
  if (maxETransfer <= theElCut) return;
  const double deltaEkin = iselectron ? SampleETransferMoller(theElCut, thePrimEkin, rng)
                                      : SampleETransferBhabha(theElCut, thePrimEkin, rng);
  Secondary& sc = sec.s[sec.n++];
  IoniSampleDirections(thePrimEkin, deltaEkin, sc.dir, s.dir, rng);
  SetEKin(s, thePrimEkin - deltaEkin);
  sc.ekin = deltaEkin;
  sc.kind = kSecElectron;
}

// ---- bremsstrahlung -----------------------------------------------------------------------------------
// SampleCostModifiedTsai (InteractionUtils.icc:11-23)
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

// SampleETransferSB (Brem.icc:73-182)
G4H_FN double SampleETransferSB(const TablesView& tv, double thePrimEkin, double theLogEkin, int theMCIndx, Rng& rng,
                                bool iselectron) {
  const double theGamCut    = G4H_LD(tv.mcCuts + 4 * theMCIndx + kCGamCut);
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
  const double minV    = G4H_LD(tv.sbData + iSTStart + iGamCut);
  const double* stData = tv.sbData + iSTStart + numGamCuts;
  const double lKTrans = (theLogGamCut - theLogEkin) / (theLogGamCut - G4H_LD(tv.sbLElEnergy + elEnergyIndx));
  const double primETot     = thePrimEkin + kElectronMassC2;
  const double dielSupConst = G4H_LD(tv.matPars + 16 * imat + kMDensityCorFactor) * primETot * primETot;
  double suppression = 1.0;
  double eGamma = 0.0;
  double r1;
  do {
    const double r0 = rng.Flat();
    r1 = rng.Flat();
    double kappa = 1.0;
    if (!isSimply) {
      const double cumRV  = r0 * (1.0 - minV) + minV;
      const int cumLIndx3 = SBLinSearch(stData, numKappa, cumRV) - 3;
      const int cumLIndx  = cumLIndx3 / 3;
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
      kappa = Exp(lKappa * lKTrans);
    } else {
      kappa = 1.0 - r0 * (1.0 - theGamCut / thePrimEkin);
    }
    eGamma = kappa * thePrimEkin;
    const double invEGamma = 1.0 / eGamma;
    suppression = 1.0 / (1.0 + dielSupConst * invEGamma * invEGamma);
    if (!iselectron) {
      const double e1     = thePrimEkin - theGamCut;
      const double iBeta1 = (e1 + kElectronMassC2) / sqrt(e1 * (e1 + 2.0 * kElectronMassC2));
      const double e2     = thePrimEkin - eGamma;
      const double iBeta2 = (e2 + kElectronMassC2) / sqrt(e2 * (e2 + 2.0 * kElectronMassC2));
      const double dum    = kAlpha * k2Pi * dZet * (iBeta1 - iBeta2);
      suppression = (dum > -12.) ? suppression * Exp(dum) : 0.;
    }
  } while (r1 > suppression);
  return eGamma;
}

// SampleETransferRB (Brem.icc:184-262)
G4H_FN double SampleETransferRB(const TablesView& tv, double thePrimEkin, double theLogEkin, int theMCIndx, Rng& rng,
                                bool iselectron) {
  const double theGamCut = G4H_LD(tv.mcCuts + 4 * theMCIndx + kCGamCut);
  const int imat         = G4H_LD(tv.mcImat + theMCIndx);
  const double* mp       = tv.matPars + 16 * imat;
  const ElectronTablesView& ed = tv.el[iselectron ? 0 : 1];
  const int numElem  = G4H_LD(tv.matNumElem + imat);
  const int elemIndx = (numElem > 1) ? SelectTargetAtomBrem(ed, theMCIndx, thePrimEkin, theLogEkin, rng.Flat(), false) : 0;
  const int iZet     = G4H_LD(tv.matElemZ + G4H_LD(tv.matElemStart + imat) + elemIndx);
  const double dZet  = static_cast<double>(iZet);
  const double* ep   = tv.elemPars + 12 * (iZet < 120 ? iZet : 120);
  const double densityFactor = kMigdalConst * G4H_LD(mp + kMElectronDensity);
  const double lpmEnergy     = kLPMconstant * G4H_LD(mp + kMRadLength);
  const double lpmEnergyLim  = sqrt(densityFactor) * lpmEnergy;
  const double thePrimTotalE = thePrimEkin + kElectronMassC2;
  const double densityCorr   = densityFactor * thePrimTotalE * thePrimTotalE;
  const bool isLPMActive     = (thePrimTotalE > lpmEnergyLim);
  const double zFactor1   = G4H_LD(ep + kEZFactor1);
  const double zFactor2   = (1. + 1. / dZet) / 12.;
  const double rejFuncMax = zFactor1 + zFactor2;
  const double xmin   = Log(theGamCut * theGamCut + densityCorr);
  const double xrange = Log(thePrimEkin * thePrimEkin + densityCorr) - xmin;
  const double zet13 = G4H_LD(ep + kEZet13);
  double eGamma, funcVal, r1;
  do {
    const double r0 = rng.Flat();
    r1 = rng.Flat();
    eGamma = sqrt(Max(Exp(xmin + r0 * xrange) - densityCorr, 0.0));
    const double y     = eGamma / thePrimTotalE;
    const double onemy = 1. - y;
    const double dum0  = 0.25 * y * y;
    if (isLPMActive) {
      double funcGS, funcPhiS, funcXiS;
      EvaluateLPMFunctions(funcXiS, funcGS, funcPhiS, eGamma, thePrimTotalE, lpmEnergy, G4H_LD(ep + kEZet23),
                           G4H_LD(ep + kEILVarS1), G4H_LD(ep + kEILVarS1Cond), densityCorr, 1.0);
      const double term1 = funcXiS * (dum0 * funcGS + (onemy + 2.0 * dum0) * funcPhiS);
      funcVal = term1 * zFactor1 + onemy * zFactor2;
    } else {
      const double dum1 = onemy + 3. * dum0;
      if (iZet < 5) {
        funcVal = dum1 * zFactor1 + onemy * zFactor2;
      } else {
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
    funcVal = Max(0.0, funcVal);
  } while (funcVal < rejFuncMax * r1);
  return eGamma;
}

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

// Brem::Perform (Brem.icc:35-70)
G4H_FN void PerformBrem(const TablesView& tv, ElectronState& s, Rng& rng, Secondaries& sec, bool isSBmodel) {
  const bool iselectron    = !s.isPositron;
  const double thePrimEkin = s.ekin;
  const double theLogEkin  = GetLogEKin(s);
  const double theGamCut   = G4H_LD(tv.mcCuts + 4 * s.imc + kCGamCut);
  if (thePrimEkin <= theGamCut) return;
  const double eGamma = isSBmodel ? SampleETransferSB(tv, thePrimEkin, theLogEkin, s.imc, rng, iselectron)
                                  : SampleETransferRB(tv, thePrimEkin, theLogEkin, s.imc, rng, iselectron);
  Secondary& sc = sec.s[sec.n++];
  BremSampleDirections(thePrimEkin, eGamma, sc.dir, s.dir, rng);
  SetEKin(s, thePrimEkin - eGamma);
  sc.ekin = eGamma;
  sc.kind = kSecGamma;
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

// Compton: SamplePhotonEnergyAndDirection + Perform (Compton.icc:17-103)
G4H_FN void PerformCompton(GammaState& s, Rng& rng, Secondaries& sec) {
  const double thePrimGmE = s.ekin;
  const double theLowEnergyThreshold = 0.0001;
  if (thePrimGmE < theLowEnergyThreshold) {
    return;
  }
  const double theOrgGmDir[3] = {s.dir[0], s.dir[1], s.dir[2]};
  const double kappa = thePrimGmE * kInvElectronMassC2;
  const double eps0  = 1. / (1. + 2. * kappa);
  const double eps02 = eps0 * eps0;
  const double al1   = -Log(eps0);
  const double al2   = al1 + 0.5 * (1. - eps02);
  double eps, eps2, gf;
  double oneMinusCost, sint2;
  double r2;
  do {
    const double r0 = rng.Flat();
    const double r1 = rng.Flat();
    r2 = rng.Flat();
    if (al1 > al2 * r0) {
      eps  = Exp(-al1 * r1);
      eps2 = eps * eps;
    } else {
      eps2 = eps02 + (1. - eps02) * r1;
      eps  = sqrt(eps2);
    }
    oneMinusCost = (1. - eps) / (eps * kappa);
    sint2 = oneMinusCost * (2. - oneMinusCost);
    gf    = 1. - eps * sint2 / (1. + eps2);
  } while (gf < r2);
  const double cost = 1.0 - oneMinusCost;
  const double sint = sqrt(Max(0., sint2));
  const double phi  = k2Pi * rng.Flat();
  double sphi, cphi;
  SinCos(phi, sphi, cphi);
  s.dir[0] = sint * cphi;
  s.dir[1] = sint * sphi;
  s.dir[2] = cost;
  RotateToReferenceFrame(s.dir, theOrgGmDir);
  const double thePostGmE = thePrimGmE * eps;
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

// std::pow(x, 1./3.) of Conversion.icc:188,216 for x in (0, 1): the cube root (libdevice's cbrt is a quarter of the
// instructions of its pow; the two differ from the host's pow by an ulp either way, well inside the 1e-12 of energies)
G4H_FN double CubeRoot(double x) {
#if defined(__CUDA_ARCH__)
  return cbrt(x);
#else
  return pow(x, 1. / 3.);
#endif
}

// Conversion screening functions (Conversion.icc:237-273)
G4H_FN double ScreenFunction1(double delta) {
  return (delta > 1.4) ? 42.038 - 8.29 * Log(delta + 0.958) : 42.184 - delta * (7.444 - 1.623 * delta);
}
G4H_FN double ScreenFunction2(double delta) {
  return (delta > 1.4) ? 42.038 - 8.29 * Log(delta + 0.958) : 41.326 - delta * (5.848 - 0.902 * delta);
}
G4H_FN void ComputePhi12(double delta, double& phi1, double& phi2) {
  if (delta > 1.4) {
    phi1 = 21.0190 - 4.145 * Log(delta + 0.958);
    phi2 = phi1;
  } else {
    phi1 = 20.806 - delta * (3.190 - 0.5710 * delta);
    phi2 = 20.234 - delta * (2.126 - 0.0903 * delta);
  }
}

// Conversion::SampleKinEnergies (Conversion.icc:56-120) with SampleEnergyRateNoLPM / WithLPM (:180-234)
G4H_FN void ConversionSampleKinEnergies(const TablesView& tv, double thePrimEkin, double theLogEkin, int theMCIndx,
                                        double& eKinEnergy, double& pKinEnergy, Rng& rng) {
  const int matIndx  = G4H_LD(tv.mcImat + theMCIndx);
  const int numElem  = G4H_LD(tv.matNumElem + matIndx);
  const int elemIndx = (numElem > 1) ? SelectTargetAtomConversion(tv, matIndx, thePrimEkin, theLogEkin, rng.Flat()) : 0;
  const int iZet     = G4H_LD(tv.matElemZ + G4H_LD(tv.matElemStart + matIndx) + elemIndx);
  const double lpmEnr = kLPMconstant * G4H_LD(tv.matPars + 16 * matIndx + kMRadLength);
  const double* ep   = tv.elemPars + 12 * (iZet < 120 ? iZet : 120);
  const double eps0 = kElectronMassC2 / thePrimEkin;
  double eps = 0.0;
  if (thePrimEkin < 2.0) {
    eps = eps0 + (0.5 - eps0) * rng.Flat();
  } else {
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
    const double NormF1   = Max(F10 * epsRange * epsRange, 0.);
    const double NormF2   = Max(1.5 * F20, 0.);
    const double NormCond = NormF1 / (NormF1 + NormF2);
    const double invF10 = 1. / F10;
    const double invF20 = 1. / F20;
    const bool withLPM  = !(thePrimEkin < 100000.0);
    const double z23 = G4H_LD(ep + kEZet23), ilVarS1 = G4H_LD(ep + kEILVarS1), ilVarS1Cond = G4H_LD(ep + kEILVarS1Cond);
    double greject = 0.;
    double r2;
    do {
      const double r0 = rng.Flat();
      const double r1 = rng.Flat();
      r2 = rng.Flat();
      // one call site for what the four branches of the reference share (Conversion.icc:196-231): the screening
      // variable, its logarithm (ScreenFunction1/2 and ComputePhi12 take the same Log(delta + 0.958) above 1.4) and
      // the LPM functions; the branches then only combine them.  (Called from the four branches the logarithm ran
      // at 6 of 32 lanes and was a sixth of the kernel's instructions.)
      const bool first = NormCond > r0;
      eps = first ? 0.5 - epsRange * CubeRoot(r1) : epsMin + epsRange * r1;
      const double delta    = deltaFactor / (eps * (1. - eps));
      const bool highDelta  = delta > 1.4;
      const double logDelta = Log(highDelta ? delta + 0.958 : 1.0);
      if (!withLPM) {
        // ScreenFunction1 / ScreenFunction2 (Conversion.icc:237-248)
        const double screen = highDelta ? 42.038 - 8.29 * logDelta
                                        : (first ? 42.184 - delta * (7.444 - 1.623 * delta) : 41.326 - delta * (5.848 - 0.902 * delta));
        greject = (screen - FZ) * (first ? invF10 : invF20);
      } else {
        // ComputePhi12 (Conversion.icc:250-262)
        const double phi1 = highDelta ? 21.0190 - 4.145 * logDelta : 20.806 - delta * (3.190 - 0.5710 * delta);
        const double phi2 = highDelta ? phi1 : 20.234 - delta * (2.126 - 0.0903 * delta);
        double funcXiS, funcGS, funcPhiS;
        EvaluateLPMFunctions(funcXiS, funcGS, funcPhiS, thePrimEkin, eps * thePrimEkin, lpmEnr, z23, ilVarS1, ilVarS1Cond, 0.0, -1.0);
        greject = first ? funcXiS * ((2. * funcPhiS + funcGS) * phi1 - funcGS * phi2 - funcPhiS * FZ) * invF10
                        : funcXiS * ((funcPhiS + 0.5 * funcGS) * phi1 + 0.5 * funcGS * phi2 - 0.5 * (funcGS + funcPhiS) * FZ) * invF20;
      }
    } while (greject < r2);
  }
  double eTotEnergy, pTotEnergy;
  if (rng.Flat() > 0.5) {
    eTotEnergy = (1. - eps) * thePrimEkin;
    pTotEnergy = eps * thePrimEkin;
  } else {
    pTotEnergy = (1. - eps) * thePrimEkin;
    eTotEnergy = eps * thePrimEkin;
  }
  eKinEnergy = Max(0., eTotEnergy - kElectronMassC2);
  pKinEnergy = Max(0., pTotEnergy - kElectronMassC2);
}

// Conversion::Perform + SampleDirections (Conversion.icc:19-53, 123-146)
G4H_FN void PerformConversion(const TablesView& tv, GammaState& s, Rng& rng, Secondaries& sec) {
  const double thePrimGmE = s.ekin;
  if (thePrimGmE < 2. * kElectronMassC2) {
    return;
  }
  const double theLogPrimGmE = GetLogEKin(s);
  double elKinEnergy, posKinEnergy;
  ConversionSampleKinEnergies(tv, thePrimGmE, theLogPrimGmE, s.imc, elKinEnergy, posKinEnergy, rng);
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

// Photoelectric::Perform, SelectElementBindingEnergy, SamplePhotoElectronDirection (Photoelectric.icc:13-121)
G4H_FN void PerformPhotoelectric(const TablesView& tv, GammaState& s, Rng& rng, Secondaries& sec) {
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
      const int z = G4H_LD(tv.matElemZ + elemStart + i);
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
  if (photoElecE > theLowEnergyThreshold) {
    Secondary& sc = sec.s[sec.n++];
    // SamplePhotoElectronDirection (Sauter-Gavrila)
    const double tau   = photoElecE * kInvElectronMassC2;
    const double gamma = 1.0 + tau;
    const double beta  = sqrt(tau * (tau + 2.0)) / gamma;
    const double ac = (1.0 - beta) / beta;
    const double a1 = 0.5 * beta * gamma * tau * (gamma - 2.0);
    const double a2 = ac + 2.0;
    const double gtmax = 2.0 * (a1 + 1.0 / ac);
    double tsam = 0.0;
    double gtr  = 0.0;
    double r1;
    do {
      const double r0 = rng.Flat();
      r1 = rng.Flat();
      tsam = 2.0 * ac * (2.0 * r0 + a2 * sqrt(r0)) / (a2 * a2 - 4.0 * r0);
      gtr  = (2.0 - tsam) * (a1 + 1.0 / (ac + tsam));
    } while (r1 * gtmax > gtr);
    const double costheta = 1.0 - tsam;
    const double sint = sqrt(tsam * (2.0 - tsam));
    const double phi  = k2Pi * rng.Flat();
    double sphi, cphi;
    SinCos(phi, sphi, cphi);
    sc.dir[0] = sint * cphi;
    sc.dir[1] = sint * sphi;
    sc.dir[2] = costheta;
    RotateToReferenceFrame(sc.dir, s.dir);
    sc.ekin = photoElecE;
    sc.kind = kSecElectron;
    s.edep  = bindingEnergy;
  } else {
    s.edep = theGammaE;
  }
  SetEKin(s, 0.0);
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
      PerformCompton(s, rng, sec);
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
