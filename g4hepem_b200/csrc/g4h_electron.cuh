// g4h_electron.cuh -- e-/e+ step limit and along-step (continuous) physics, one track per thread.
//
// Restates, with identical floating point operation order and random number consumption:
//   G4HepEmElectronManager::HowFar / HowFarToDiscreteInteraction / HowFarToMSC / UpdatePStepLength /
//     UpdateNumIALeft / ApplyMeanEnergyLoss / SampleMSC / SampleLossFluctuations / PerformContinuous /
//     CheckDelta / ConvertTrueToGeometricLength / ConvertGeometricToTrueLength
//     (G4HepEmRun/include/G4HepEmElectronManager.icc:35-483, 602-691)
//   G4HepEmElectronInteractionUMSC (G4HepEmElectronInteractionUMSC.icc:19-357), G4VERSION_NUM >= 1100 branches
//   G4HepEmElectronEnergyLossFluctuation (G4HepEmElectronEnergyLossFluctuation.icc:11-111)
// The track lives in registers (ElectronState) for the whole step.
#ifndef G4H_ELECTRON_CUH
#define G4H_ELECTRON_CUH

#include "g4h_math.cuh"
#include "g4h_rng.cuh"
#include "g4h_tables.cuh"

namespace g4h {

struct ElectronState {
  // G4HepEmTrack
  double ekin, logEkin;  // logEkin > 99: not cached (G4HepEmTrack.hh:95-100)
  double dir[3];
  double safety;
  double nIA[4];
  double mfp[4];
  double gStep, edep;
  int imc, id, winner;
  bool isPositron, onBoundary;
  // G4HepEmElectronTrack
  double range, pStep, preStepEkin, preStepLogEkin;
  // G4HepEmMSCTrackData
  double lambtr1, trueStep, zPath;
  double disp[3];
  double initialRange, dynRangeFactor, tlimitMin;
  double par1, par2, par3;
  bool mscNoScatter, mscDisplace, mscFirstStep, mscActive;
};

// G4HepEmTrack::GetLogEKin (G4HepEmTrack.hh:95-100)
G4H_FN double GetLogEKin(ElectronState& s) {
  if (s.logEkin > 99.0) {
    s.logEkin = (s.ekin > 0.) ? Log(s.ekin) : -30;
  }
  return s.logEkin;
}
// G4HepEmTrack::SetEKin(ekin) (G4HepEmTrack.hh:71-74)
G4H_FN void SetEKin(ElectronState& s, double ekin) {
  s.ekin    = ekin;
  s.logEkin = 100.0;
}

// rotate (u,v,w) given in the scattering frame into the frame of refDir (G4HepEmRunUtils.icc:31-46)
G4H_FN void RotateToReferenceFrame(double* dir, const double* refDir) {
  double up = refDir[0] * refDir[0] + refDir[1] * refDir[1];
  if (up > 0.) {
    up = sqrt(up);
    const double px = dir[0];
    const double py = dir[1];
    const double pz = dir[2];
    dir[0] = (refDir[0] * refDir[2] * px - refDir[1] * py) / up + refDir[0] * pz;
    dir[1] = (refDir[1] * refDir[2] * px + refDir[0] * py) / up + refDir[1] * pz;
    dir[2] = -up * px + refDir[2] * pz;
  } else if (refDir[2] < 0.) {
    dir[0] = -dir[0];
    dir[2] = -dir[2];
  }
}

// ---- HowFar -----------------------------------------------------------------------------------------
// the resampling loop of G4HepEmElectronManager::HowFar(data, pars, tlData) (.icc:39-43)
G4H_FN void ResampleNumIALeft(ElectronState& s, Rng& rng) {
#pragma unroll
  for (int ip = 0; ip < 4; ++ip) {
    if (s.nIA[ip] <= 0.) {
      s.nIA[ip] = -Log(rng.Flat());
    }
  }
}

// HowFarToDiscreteInteraction (.icc:48-101)
G4H_FN void HowFarToDiscreteInteraction(const TablesView& tv, ElectronState& s) {
  int indxWinnerProcess = -1;
  const double theEkin  = s.ekin;
  const double theLEkin = GetLogEKin(s);
  const int theIMC      = s.imc;
  const bool isElectron = !s.isPositron;
  const ElectronTablesView& ed = tv.el[isElectron ? 0 : 1];
  const double range = RestRange(ed, theIMC, theEkin, theLEkin);
  s.range = range;
  const double* rp    = tv.regionPars + 8 * G4H_LD(tv.mcIreg + theIMC);
  const double frange = G4H_LD(rp + kRFinalRange);
  const double drange = G4H_LD(rp + kRDRoverRange);
  double pStepLength = (range > frange) ? range * drange + frange * (1.0 - drange) * (2.0 - frange / range) : range;
  const int theImat = G4H_LD(tv.mcImat + theIMC);
  double mxSecs[4];
  mxSecs[0] = RestMacXSecForStepping(ed, theIMC, theEkin, theLEkin, true);
  mxSecs[1] = RestMacXSecForStepping(ed, theIMC, theEkin, theLEkin, false);
  mxSecs[2] = isElectron ? 0.0 : MacXSecAnnihilation(0.8 * theEkin, G4H_LD(tv.matPars + 16 * theImat + kMElectronDensity));
  mxSecs[3] = MacXSecNuclear(ed, theImat, theEkin, theLEkin);
#pragma unroll
  for (int ip = 0; ip < 4; ++ip) {
    const double mxsec = mxSecs[ip];
    const double mfp   = (mxsec > 0.) ? 1. / mxsec : kALargeValue;
    s.mfp[ip] = mfp;
    const double dStepLimit = mfp * s.nIA[ip];
    if (dStepLimit < pStepLength) {
      pStepLength       = dStepLimit;
      indxWinnerProcess = ip;
    }
  }
  s.pStep  = pStepLength;
  s.winner = indxWinnerProcess;
  s.gStep  = pStepLength;
}

// G4HepEmElectronInteractionUMSC::StepLimit (G4HepEmElectronInteractionUMSC.icc:19-126)
G4H_FN void UMSCStepLimit(const TablesView& tv, ElectronState& s, double ekin, int imat, int iregion, double range,
                          double presafety, bool onBoundary, bool iselectron, Rng& rng) {
  s.mscNoScatter = false;
  s.mscDisplace  = true;
  const double kTLimitMinfix = 1.0E-8;
  const double* mp = tv.matPars + 16 * imat;
  if (s.trueStep < kTLimitMinfix || range * G4H_LD(mp + kMUMSCPar) < presafety) {
    s.mscDisplace = false;
    return;
  }
  const double* rp = tv.regionPars + 8 * iregion;
  const double mscRangeFactor  = G4H_LD(rp + kRMSCRangeFactor);
  const double mscSafetyFactor = G4H_LD(rp + kRMSCSafetyFactor);
  const bool mscIsUseSafety    = !(G4H_LD(rp + kRIsMSCMinimal) != 0.0);
  double tlimit = 0.0;
  if (mscIsUseSafety) {
    if (s.mscFirstStep || onBoundary) {
      const double lambdaTr1 = s.lambtr1;
      s.initialRange   = Max(range, lambdaTr1);
      s.dynRangeFactor = lambdaTr1 > 1.0 ? mscRangeFactor * (0.75 + 0.25 * lambdaTr1) : mscRangeFactor;
      const double stepMin = lambdaTr1 * 1.0E-3 / (2.0E-3 + ekin * (G4H_LD(mp + kMStepMin0) + ekin * G4H_LD(mp + kMStepMin1)));
      const double dum0 = iselectron ? 0.87 * G4H_LD(mp + kMZeff23) : 0.70 * G4H_LD(mp + kMZeffSqrt);
      const double dum1 = ekin > 5.0E-3 ? dum0 * stepMin : dum0 * stepMin * 0.5 * (1.0 + ekin * 200.0);
      s.tlimitMin    = Max(dum1, kTLimitMinfix);
      s.mscFirstStep = false;
    }
    const double tlimitmin = s.tlimitMin;
    tlimit = range > presafety ? Max(Max(s.initialRange * s.dynRangeFactor, mscSafetyFactor * presafety), tlimitmin)
                               : Max(range, tlimitmin);
  } else {
    if (onBoundary) {
      const double lambdaTr1 = s.lambtr1;
      const double tmpTlimit = range > lambdaTr1 ? mscRangeFactor * range : mscRangeFactor * lambdaTr1;
      s.initialRange = Max(tmpTlimit, 10 * kTLimitMinfix);
    }
    tlimit = s.initialRange;
  }
  const double tlimitmin = s.tlimitMin;
  if (tlimit < s.trueStep) {
    const double dum0 = tlimit > tlimitmin ? Max(rng.Gauss(tlimit, 0.1 * (tlimit - tlimitmin)), tlimitmin) : tlimitmin;
    s.trueStep = Min(dum0, s.trueStep);
  }
}

// ConvertTrueToGeometricLength (G4HepEmElectronManager.icc:602-650) in two pieces: everything but the regime that
// needs the inverse range table (Head; returns true when that regime applies and leaves zPath = trueStep), and that
// regime (RangeRegime).  The staged HowFar runs the second piece as its own kernel over a queue: one track in
// seven takes it and it is half the instructions of the conversion.
G4H_FN bool ConvertTrueToGeometricLengthHead(ElectronState& s, double ekin, double range) {
  s.par1 = -1.;
  s.par2 = 0.;
  s.par3 = 0.;
  s.trueStep = Min(s.trueStep, range);
  s.zPath    = s.trueStep;
  const double kTlimitMinfix2 = 1.0E-6;
  if (s.trueStep < kTlimitMinfix2) {
    return false;
  }
  const double kTauSmall = 1.0e-16;
  const double kDtrl     = 0.05;
  const double tau       = s.trueStep / s.lambtr1;
  if (tau < kTauSmall) {
    s.zPath = Min(s.trueStep, s.lambtr1);
  } else if (s.trueStep < range * kDtrl) {
    const double kTauLim = 1.0e-6;
    s.zPath = (tau < kTauLim) ? s.trueStep * (1. - 0.5 * tau) : s.lambtr1 * (1. - Exp(-tau));
  } else if (ekin < kElectronMassC2 || s.trueStep == range) {
    s.par1  = 1. / range;
    s.par2  = 1. / (s.par1 * s.lambtr1);
    s.par3  = 1. + s.par2;
    s.zPath = 1. / (s.par1 * s.par3);
    if (s.trueStep < range) {
      s.zPath *= (1. - Pow(1. - s.trueStep / range, s.par3));
    }
  } else {
    return true;
  }
  s.zPath = Min(s.zPath, s.lambtr1);
  return false;
}

G4H_FN void ConvertTrueToGeometricLengthRangeRegime(const TablesView& tv, ElectronState& s, double range, int imc,
                                                    bool iselectron) {
  const double rfin = Max(range - s.trueStep, 0.01 * range);
  const ElectronTablesView& ed = tv.el[iselectron ? 0 : 1];
  const double t1      = InvRange(ed, imc, rfin);
  const int imat       = G4H_LD(tv.mcImat + imc);
  const double lambda1 = TransportMFP(ed, imat, t1, Log(t1));
  s.par1  = (s.lambtr1 - lambda1) / (s.lambtr1 * s.trueStep);
  s.par2  = 1. / (s.par1 * s.lambtr1);
  s.par3  = 1. + s.par2;
  s.zPath = (1. - Pow(lambda1 / s.lambtr1, s.par3)) / (s.par1 * s.par3);
  s.zPath = Min(s.zPath, s.lambtr1);
}

G4H_FN void ConvertTrueToGeometricLength(const TablesView& tv, ElectronState& s, double ekin, double range, int imc,
                                         bool iselectron) {
  if (ConvertTrueToGeometricLengthHead(s, ekin, range)) {
    ConvertTrueToGeometricLengthRangeRegime(tv, s, range, imc, iselectron);
  }
}

// HowFarToMSC (.icc:103-164)
G4H_FN void HowFarToMSC(const TablesView& tv, ElectronState& s, Rng& rng) {
  double pStepLength    = s.pStep;
  const double range    = s.range;
  const double theEkin  = s.ekin;
  const double theLEkin = GetLogEKin(s);
  const int theIMC      = s.imc;
  const bool isElectron = !s.isPositron;
  const ElectronTablesView& ed = tv.el[isElectron ? 0 : 1];
  const int theImat = G4H_LD(tv.mcImat + theIMC);
  const int theIreg = G4H_LD(tv.mcIreg + theIMC);
  s.trueStep  = pStepLength;
  s.zPath     = pStepLength;
  s.mscActive = false;
  s.disp[0] = 0.;
  s.disp[1] = 0.;
  s.disp[2] = 0.;
  const double kGeomMinLength = 5.E-8;
  if (pStepLength > kGeomMinLength && theEkin > 1.0E-3) {
    s.mscActive = true;
    s.lambtr1   = TransportMFP(ed, theImat, theEkin, theLEkin);
    UMSCStepLimit(tv, s, theEkin, theImat, theIreg, range, s.safety, s.onBoundary, isElectron, rng);
    ConvertTrueToGeometricLength(tv, s, theEkin, range, theIMC, isElectron);
    const double mscTruStepLength = s.trueStep;
    if (mscTruStepLength < pStepLength) {
      s.winner    = -2;
      pStepLength = mscTruStepLength;
      s.pStep     = pStepLength;
    }
    s.gStep = Min(s.zPath, pStepLength);
  }
}

// ---- along step -----------------------------------------------------------------------------------------
// ConvertGeometricToTrueLength (.icc:653-691)
G4H_FN void ConvertGeometricToTrueLength(ElectronState& s, double range, double gStepToConvert) {
  s.zPath = gStepToConvert;
  const double kTLimitMinfix2 = 1.0E-6;
  if (gStepToConvert < kTLimitMinfix2) {
    s.trueStep = gStepToConvert;
  } else {
    const double kTauSmall = 1.0e-16;
    double tlength = gStepToConvert;
    if (gStepToConvert > s.lambtr1 * kTauSmall) {
      if (s.par1 < 0.) {
        tlength = -s.lambtr1 * Log(1. - gStepToConvert / s.lambtr1);
      } else {
        const double dum = s.par1 * s.par3 * gStepToConvert;
        if (dum < 1.) {
          tlength = (1. - Pow(1. - dum, 1. / s.par3)) / s.par1;
        } else {
          tlength = range;
        }
      }
      if (tlength < gStepToConvert) {
        tlength = gStepToConvert;
      }
    }
    s.trueStep = tlength;
  }
}

// UpdatePStepLength (.icc:171-203)
G4H_FN void UpdatePStepLength(ElectronState& s) {
  const double gStepLength = s.gStep;
  double pStepLength       = gStepLength;
  const double theRange    = s.range;
  if (s.mscActive) {
    pStepLength = s.trueStep;
    if (gStepLength < s.zPath) {
      ConvertGeometricToTrueLength(s, theRange, gStepLength);
      pStepLength = Min(pStepLength, s.trueStep);
      s.trueStep  = pStepLength;
    }
    const double kGeomMinLength = 5.E-8;
    if (pStepLength <= kGeomMinLength || theRange <= pStepLength) {
      s.mscActive = false;
    }
  }
  s.pStep = pStepLength;
}

// ApplyMeanEnergyLoss (.icc:216-259)
G4H_FN bool ApplyMeanEnergyLoss(const TablesView& tv, ElectronState& s) {
  const double pStepLength = s.pStep;
  const bool isElectron    = !s.isPositron;
  const double theEkin     = s.ekin;
  const double theRange    = s.range;
  if (pStepLength >= theRange || theEkin <= tv.minLossTableEnergy) {
    s.edep = theEkin;
    SetEKin(s, 0.0);
    return true;
  }
  const ElectronTablesView& ed = tv.el[isElectron ? 0 : 1];
  const int theIMC      = s.imc;
  const double theLEkin = GetLogEKin(s);
  double eloss = pStepLength * RestDEDX(ed, theIMC, theEkin, theLEkin);
  const double parLinELossLimit = G4H_LD(tv.regionPars + 8 * G4H_LD(tv.mcIreg + theIMC) + kRLinELossLimit);
  if (eloss > theEkin * parLinELossLimit) {
    const double postStepRange = theRange - pStepLength;
    eloss = theEkin - InvRange(ed, theIMC, postStepRange);
  }
  eloss = Max(eloss, 0.0);
  if (eloss >= theEkin) {
    eloss = theEkin;
    SetEKin(s, 0);
    s.edep = eloss;
    return true;
  }
  SetEKin(s, theEkin - eloss);
  s.edep = eloss;
  return false;
}

// G4HepEmElectronInteractionUMSC::Theta0PositronCorrection (UMSC.icc:309-335)
G4H_FN double Theta0PositronCorrection(double eekin, double zeff) {
  const double ff = 1. + zeff * (1.84035E-4 * zeff - 1.86427E-2) + 0.41125;
  const double a  = 0.994 - 4.08E-3 * zeff;
  const double b  = 7.16 + (52.6 + 365. / zeff) / zeff;
  const double tu = sqrt(eekin) * kInvElectronMassC2;
  const double x  = sqrt(tu * (tu + 2.) / ((tu + 1.) * (tu + 1.)));
  const double xl = 0.6;
  if (x < xl) {
    return ff * a * (1. - Exp(-b * x));
  }
  const double c  = 1.00 - 4.47E-3 * zeff;
  const double d  = 1.21E-3 * zeff;
  const double e  = 113.0;
  const double xh = 0.9;
  if (x > xh) {
    return ff * (c + d * Exp(e * (x - 1.)));
  }
  const double yl = a * (1. - Exp(-b * xl));
  const double yh = c + d * Exp(e * (xh - 1.));
  const double y0 = (yh - yl) / (xh - xl);
  const double y1 = yl - y0 * xl;
  return ff * (y0 * x + y1);
}

// ComputeTheta0 (UMSC.icc:293-305): Highland formula with correction
G4H_FN double ComputeTheta0(double stepInRadLength, double postStepEkin, double preStepEkin, double zeff,
                            double thetaCoeff0, double thetaCoeff1, bool isElectron, bool isPosCor) {
  const double kHighland     = 13.6;
  const double postInvBetaPc = (postStepEkin + kElectronMassC2) / (postStepEkin * (postStepEkin + 2. * kElectronMassC2));
  const double invBetaPc     = preStepEkin != postStepEkin
                                   ? sqrt(postInvBetaPc * (preStepEkin + kElectronMassC2) / (preStepEkin * (preStepEkin + 2. * kElectronMassC2)))
                                   : postInvBetaPc;
  const double y = (isElectron || !isPosCor) ? stepInRadLength
                                             : stepInRadLength * Theta0PositronCorrection(preStepEkin * postStepEkin, zeff);
  return kHighland * sqrt(y) * invBetaPc * (thetaCoeff0 + thetaCoeff1 * Log(y));
}

// SimpleScattering (UMSC.icc:274-289)
G4H_FN double SimpleScattering(double xmeanth, double x2meanth, Rng& rng) {
  const double dum0 = 3. * x2meanth - 1;
  const double dum1 = 2. * xmeanth - dum0;
  const double a    = 1. + 4. * dum0 / dum1;
  const double prob = (2. + a) * xmeanth / a;
  const double r0 = rng.Flat();
  const double r1 = rng.Flat();
  return (r0 < prob) ? -1. + 2. * Pow(r1, 1. / (1. + a)) : -1. + 2. * r1;
}

// SampleCosineTheta (UMSC.icc:153-271)
G4H_FN double SampleCosineTheta(double pStepLength, double preStepEkin, double preStepTr1mfp, double postStepEkin,
                                double postStepTr1mfp, double umscTlimitMin, const double* mp, bool isElectron,
                                bool isPosCor, Rng& rng) {
  const double radLength = G4H_LD(mp + kMRadLength);
  const double zeff      = G4H_LD(mp + kMZeff);
  const double iPreStepTr1mfp = 1.0 / preStepTr1mfp;
  const double deltaR1mfp     = preStepTr1mfp - postStepTr1mfp;
  const double tau = fabs(deltaR1mfp) > 0.01 * preStepTr1mfp ? pStepLength * Log(preStepTr1mfp / postStepTr1mfp) / deltaR1mfp
                                                              : pStepLength * iPreStepTr1mfp;
  const double kTauBig = 8.0;
  if (tau > kTauBig) {
    return 2.0 * rng.Flat() - 1.0;
  }
  const double kTauSmall = 1.0E-16;
  if (tau < kTauSmall) {
    return 1.0;
  }
  double xmeanth, x2meanth;
  if (tau < 0.01) {
    xmeanth  = 1.0 - tau * (1.0 - 0.5 * tau);
    x2meanth = 1.0 - tau * (5.0 - 6.25 * tau) * 0.333333;
  } else {
    xmeanth  = Exp(-tau);
    x2meanth = (1.0 + 2.0 * Exp(-2.5 * tau)) * 0.333333;
  }
  if (postStepEkin < 0.5 * preStepEkin) {
    return SimpleScattering(xmeanth, x2meanth, rng);
  }
  const double tsmall      = Min(umscTlimitMin, 1.0);
  const bool stpNotExSmall = pStepLength > tsmall;
  const double tc0 = G4H_LD(mp + kMTheta0), tc1 = G4H_LD(mp + kMTheta1);
  const double theta0 = stpNotExSmall
                            ? ComputeTheta0(pStepLength / radLength, postStepEkin, preStepEkin, zeff, tc0, tc1, isElectron, isPosCor)
                            : ComputeTheta0(tsmall / radLength, postStepEkin, preStepEkin, zeff, tc0, tc1, isElectron, isPosCor) *
                                  sqrt(pStepLength / tsmall);
  if (theta0 > kPi * 0.166666) {
    return SimpleScattering(xmeanth, x2meanth, rng);
  }
  const double theta2 = theta0 * theta0;
  if (theta2 < kTauSmall) {
    return 1.0;
  }
  const double dumtau = stpNotExSmall ? tau : tsmall * iPreStepTr1mfp;
  const double parU   = Pow(dumtau, 0.1666666);
  const double dumxsi = G4H_LD(mp + kMTail0) + parU * (G4H_LD(mp + kMTail1) + parU * G4H_LD(mp + kMTail2)) +
                        G4H_LD(mp + kMTail3) * Log(pStepLength / (tau * radLength));
  const double parXsi = Max(dumxsi, 1.9);
  const double parC   = fabs(parXsi - 3.) < 0.001 ? 3.001 : fabs(parXsi - 2.) < 0.001 ? 2.001 : parXsi;
  const double dumC1  = parC - 1.;
  const double dumEa  = Exp(-parXsi);
  const double dumEaa = 1. / (1. - dumEa);
  double thex = theta2 * (1.0 - theta2 * 0.0833333);
  if (theta2 > 0.01) {
    const double dum = 2.0 * Sin(0.5 * theta0);
    thex = dum * dum;
  }
  const double xmean1 = 1. - (1. - (1. + parXsi) * dumEa) * thex * dumEaa;
  if (xmean1 <= 0.999 * xmeanth) {
    return SimpleScattering(xmeanth, x2meanth, rng);
  }
  const double x0 = 1. - parXsi * thex;
  const double bx = parC * thex;
  const double b  = bx + x0;
  const double b1 = b + 1.;
  const double eb1 = Pow(b1, dumC1);
  const double ebx = Pow(bx, dumC1);
  const double d   = ebx / eb1;
  const double xmean2 = (x0 + d - (bx - b1 * d) / (parC - 2.)) / (1. - d);
  const double f1x0 = dumEa * dumEaa;
  const double f2x0 = dumC1 / (parC * (1. - d));
  const double prob = f2x0 / (f1x0 + f2x0);
  const double qprb = xmeanth / (prob * xmean1 + (1. - prob) * xmean2);
  const double r0 = rng.Flat();
  const double r1 = rng.Flat();
  const double r2 = rng.Flat();
  if (r0 < qprb) {
    if (r1 < prob) {
      return 1. + Log(dumEa + r2 / dumEaa) * thex;
    } else {
      const double var0 = (1.0 - d) * r2;
      if (var0 < 0.01 * d) {
        const double var = var0 / (d * dumC1);
        return -1.0 + var * (1.0 - var * 0.5 * parC) * b1;
      } else {
        return 1.0 + thex * (parC - parXsi - parC * Pow(var0 + d, -1. / dumC1));
      }
    }
  } else {
    return 2.0 * r1 - 1.0;
  }
}

// SampleMSC (.icc:261-322) with UMSC::SampleScattering (UMSC.icc:129-150) and SampleDisplacement (UMSC.icc:339-357)
G4H_FN void SampleMSC(const TablesView& tv, ElectronState& s, Rng& rng) {
  const double pStepLength = s.pStep;
  const bool isElectron    = !s.isPositron;
  const int theIMC         = s.imc;
  const double preStepEkin = s.preStepEkin;
  const double theRange    = s.range;
  const double kTLimitMinfix = 1.0E-8;
  const double kTauSmall     = 1.0e-16;
  if (s.mscActive && (pStepLength > Max(kTLimitMinfix, kTauSmall * s.lambtr1))) {
    double postStepEkin  = preStepEkin;
    double postStepLEkin = s.preStepLogEkin;
    if (pStepLength > theRange * 0.01) {  // G4VERSION_NUM >= 1100
      postStepEkin  = s.ekin;
      postStepLEkin = GetLogEKin(s);
    }
    const ElectronTablesView& ed = tv.el[isElectron ? 0 : 1];
    const int theImat = G4H_LD(tv.mcImat + theIMC);
    const double postStepTr1mfp = TransportMFP(ed, theImat, postStepEkin, postStepLEkin);
    const bool isPosCor   = tv.isMSCPositronCor != 0;
    const bool isDisplace = tv.isMSCDisplacement != 0;
    // --- SampleScattering
    const double* mp = tv.matPars + 16 * theImat;
    const double cost = SampleCosineTheta(pStepLength, preStepEkin, s.lambtr1, postStepEkin, postStepTr1mfp, s.tlimitMin, mp,
                                          isElectron, isPosCor, rng);
    if (fabs(cost) >= 1.0) {
      s.mscNoScatter = true;
      return;  // fIsNoScatteringInMSC: direction and displacement untouched
    }
    const double sth = sqrt((1.0 - cost) * (1.0 + cost));
    const double phi = k2Pi * rng.Flat();
    double sphi, cphi;
    SinCos(phi, sphi, cphi);
    double newDir[3] = {sth * cphi, sth * sphi, cost};
    s.mscDisplace = s.mscDisplace && isDisplace;
    if (s.mscDisplace && pStepLength > s.zPath) {
      // SampleDisplacement
      const double r = 0.73 * sqrt((pStepLength - s.zPath) * (pStepLength + s.zPath));
      const double cbeta  = 2.16;
      const double cbeta1 = 1. - Exp(-cbeta * kPi);
      const double r0 = rng.Flat();
      const double r1 = rng.Flat();
      const double psi  = -Log(1. - r0 * cbeta1) / cbeta;
      const double dphi = (r1 < 0.5) ? phi + psi : phi - psi;
      double sd, cd;
      SinCos(dphi, sd, cd);
      s.disp[0] = r * cd;
      s.disp[1] = r * sd;
      s.disp[2] = 0.0;
    }
    // rotate direction and displacement to the lab frame and update the direction
    RotateToReferenceFrame(newDir, s.dir);
    if (s.mscDisplace) {
      RotateToReferenceFrame(s.disp, s.dir);
    }
    s.dir[0] = newDir[0];
    s.dir[1] = newDir[1];
    s.dir[2] = newDir[2];
  }
}

// G4HepEmElectronEnergyLossFluctuation::SampleGaussianLoss (ELossFluctuation.icc:100-111)
G4H_FN double SampleGaussianLoss(double meane, double sig2e, Rng& rng) {
  const double twom = 2. * meane;
  if (meane * meane < 0.0625 * sig2e) {
    return twom * rng.Flat();
  }
  const double sig = sqrt(sig2e);
  double eloss;
  do {
    eloss = rng.Gauss(meane, sig);
  } while (eloss < 0. || eloss > twom);
  return eloss;
}

// G4HepEmElectronEnergyLossFluctuation::SampleEnergyLossFLuctuation (ELossFluctuation.icc:11-97)
G4H_FN double SampleEnergyLossFluctuation(double tcut, double excEner, double meanELoss, Rng& rng) {
  const double scaling  = Min(1. + 5.E-4 / tcut, 1.5);
  const double meanLoss = meanELoss / scaling;
  const double kFluctParRate     = 0.56;
  const double kFluctParE0       = 1.E-5;
  const double kFluctParNMaxCont = 8.;
  const double w1 = tcut / kFluctParE0;
  double a3 = meanLoss * (tcut - kFluctParE0) / (kFluctParE0 * tcut * Log(w1));
  double a1 = 0.;
  double e1 = excEner;
  double eloss = 0.0;
  if (tcut > excEner) {
    const double a1Tmp = meanLoss * (1. - kFluctParRate) / excEner;
    const double kFluctParA0 = 42.;
    const double kFluctParFw = 4.;
    const double dum0 = a1Tmp < kFluctParA0 ? .1 + (kFluctParFw - .1) * sqrt(a1Tmp / kFluctParA0) : kFluctParFw;
    a1 = a1Tmp / dum0;
    e1 *= dum0;
    a3 *= kFluctParRate;
    if (a1 > kFluctParNMaxCont) {
      const double emean = a1 * e1;
      const double sig2e = emean * e1;
      eloss = SampleGaussianLoss(emean, sig2e, rng);
    } else {
      const int p = rng.Poisson(a1);
      eloss = p > 0 ? ((p + 1) - 2. * rng.Flat()) * e1 : 0.;
    }
  }
  if (a3 > 0.) {
    double p3   = a3;
    double alfa = 1.;
    if (a3 > kFluctParNMaxCont) {
      alfa = w1 * (kFluctParNMaxCont + a3) / (w1 * kFluctParNMaxCont + a3);
      const double alfa1  = alfa * Log(alfa) / (alfa - 1.);
      const double namean = a3 * w1 * (alfa - 1.) / ((w1 - 1.) * alfa);
      const double emean  = namean * kFluctParE0 * alfa1;
      const double sig2e  = kFluctParE0 * kFluctParE0 * namean * (alfa - alfa1 * alfa1);
      eloss += SampleGaussianLoss(emean, sig2e, rng);
      p3 = a3 - namean;
    }
    const double w3 = alfa * kFluctParE0;
    if (tcut > w3) {
      const double w = (tcut - w3) / tcut;
      const int nnb  = rng.Poisson(p3);
      // the reference draws in blocks of 8 + a tail; the stream is consumed one uniform per term either way
      // 1 - w u is in [w3 / tcut, 1]: the division is in the safe range of FastDiv
      for (int i = 0; i < nnb; ++i) {
        eloss += FastDiv(w3, 1. - w * rng.Flat());
      }
    }
  }
  return eloss * scaling;
}

// SampleLossFluctuations (.icc:324-368) in three pieces, so that the pipelined Perform can run the sampling
// itself as its own kernel over the tracks that need it:
//   LossFluctuationIsSampled  the condition of .icc:337
//   LossFluctuationSample     .icc:338-350: the fluctuated loss and the energy after it
//   LossFluctuationFinish     .icc:352-367: tracking cut, final energy and deposit
G4H_FN bool LossFluctuationIsSampled(const TablesView& tv, const ElectronState& s) {
  const int iregion = G4H_LD(tv.mcIreg + s.imc);
  const bool isFluctuation = G4H_LD(tv.regionPars + 8 * iregion + kRIsFluct) != 0.0;
  const double kFluctParMinEnergy = 1.E-5;
  return isFluctuation && s.edep > kFluctParMinEnergy;
}

G4H_FN void LossFluctuationSample(const TablesView& tv, int imc, bool isElectron, double thePreStepEkin, double meanLoss,
                                  Rng& rng, double& finalEkin, double& eloss) {
  const double elCut   = G4H_LD(tv.mcCuts + 4 * imc + kCElCut);
  const int theImat    = G4H_LD(tv.mcImat + imc);
  const double meanExE = G4H_LD(tv.matPars + 16 * theImat + kMMeanExE);
  const double tmax = isElectron ? 0.5 * thePreStepEkin : thePreStepEkin;
  const double tcut = Min(elCut, tmax);
  eloss = SampleEnergyLossFluctuation(tcut, meanExE, meanLoss, rng);
  eloss = Max(eloss, 0.0);
  finalEkin = thePreStepEkin - eloss;
}

// returns true if the track was stopped; sets {ekin, logEkin = not cached} and edep
G4H_FN bool LossFluctuationFinish(const TablesView& tv, double thePreStepEkin, double finalEkin, double eloss, double& ekinOut,
                                  double& edepOut) {
  if (finalEkin <= tv.elTrackingCut) {
    ekinOut = 0.0;
    edepOut = thePreStepEkin;
    return true;
  }
  ekinOut = finalEkin;
  edepOut = eloss;
  return false;
}

G4H_FN bool SampleLossFluctuations(const TablesView& tv, ElectronState& s, Rng& rng) {
  double finalEkin = s.ekin;
  double eloss     = s.edep;
  if (LossFluctuationIsSampled(tv, s)) {
    LossFluctuationSample(tv, s.imc, !s.isPositron, s.preStepEkin, eloss, rng, finalEkin, eloss);
  }
  double ekin, edep;
  const bool stopped = LossFluctuationFinish(tv, s.preStepEkin, finalEkin, eloss, ekin, edep);
  SetEKin(s, ekin);
  s.edep = edep;
  return stopped;
}

// PerformContinuous (.icc:375-405)
G4H_FN bool PerformContinuous(const TablesView& tv, ElectronState& s, Rng& rng) {
  s.preStepEkin    = s.ekin;  // SavePreStepEKin (G4HepEmElectronTrack.hh:57-60)
  s.preStepLogEkin = GetLogEKin(s);
  UpdatePStepLength(s);
  const double pStepLength = s.pStep;
  if (pStepLength <= 0.0) {
    return false;
  }
  // UpdateNumIALeft (.icc:205-214)
  s.nIA[0] -= pStepLength / s.mfp[0];
  s.nIA[1] -= pStepLength / s.mfp[1];
  s.nIA[2] -= pStepLength / s.mfp[2];
  s.nIA[3] -= pStepLength / s.mfp[3];
  const bool stopped = ApplyMeanEnergyLoss(tv, s);
  if (stopped) {
    return true;
  }
  SampleMSC(tv, s, rng);
  return SampleLossFluctuations(tv, s, rng);
}

// CheckDelta (.icc:408-423); mfpWon is fMFPs[fPIndxWon]
G4H_FN bool CheckDeltaWith(const TablesView& tv, ElectronState& s, double mfpWon, double rand) {
  const bool isElectron = !s.isPositron;
  const ElectronTablesView& ed = tv.el[isElectron ? 0 : 1];
  const int iDProc      = s.winner;
  const int theIMC      = s.imc;
  const int theMatIndex = G4H_LD(tv.mcImat + theIMC);
  const double theEkin  = s.ekin;
  const double theLEkin = GetLogEKin(s);
  const double mxsec =
      (iDProc < 2 ? RestMacXSec(ed, theIMC, theEkin, theLEkin, iDProc == 0)
                  : (iDProc < 3 ? MacXSecAnnihilation(theEkin, G4H_LD(tv.matPars + 16 * theMatIndex + kMElectronDensity))
                                : MacXSecNuclear(ed, theMatIndex, theEkin, theLEkin)));
  return mxsec <= 0.0 || rand > mxsec * mfpWon;
}

G4H_FN bool CheckDelta(const TablesView& tv, ElectronState& s, double rand) {
  const int w = s.winner;
  const double mfpWon = w == 0 ? s.mfp[0] : w == 1 ? s.mfp[1] : w == 2 ? s.mfp[2] : s.mfp[3];
  return CheckDeltaWith(tv, s, mfpWon, rand);
}

}  // namespace g4h
#endif
