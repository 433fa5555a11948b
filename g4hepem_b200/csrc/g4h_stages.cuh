// g4h_stages.cuh -- the e-/e+ step cut into per-track stage functions, each with a small live state.
//
// Measured on the B200 (profiles/r01_*): the one-thread-per-track HowFar / continuous kernels issue one
// instruction per ~8 cycles and warp (the FP64 dependent-issue latency): 110-130 registers leave 4 warps per
// scheduler and the long dependency chains (VDT log/exp, splines) have no ILP because every leaf is an
// out-of-line call.  The stages below are what the kernels of g4h_pipeline.cuh run instead:
//
//   StageHowFarXS    HowFar, part 1: nIA resampling + HowFarToDiscreteInteraction (.icc:35-101).  Straight-line,
//                    (almost) branch free: the four -log(u), the range / ioni / brem / tr1 splines and the four
//                    1/sigma are independent chains the compiler interleaves.
//   StageHowFarMSC   HowFar, part 2: HowFarToMSC (.icc:103-164) given lambda_1 from part 1: Urban step limit
//                    and true -> geometrical conversion (branchy: four conversion regimes, Gaussian smearing).
//
// Each function loads what it needs from the batch groups and stores what it changed, so that the very same
// text runs per track in the host pre-flight harness (tests/hostsim).  Arithmetic is the reference's, operation
// by operation; only the order of *independent* evaluations differs.
#ifndef G4H_STAGES_CUH
#define G4H_STAGES_CUH

#include "g4h_batch_io.cuh"
#include "g4h_perform_stages.cuh"

namespace g4h {

// ---- HowFar, part 1 -----------------------------------------------------------------------------------------
// the resampling loop of HowFar (.icc:39-43); draws are consumed in process order, the logs are independent
G4H_FN void ResampleNumIALeftILP(double* nIA, Rng& rng) {
  double u[4];
  bool need[4];
  int count = 0;
#pragma unroll
  for (int ip = 0; ip < 4; ++ip) {
    need[ip] = nIA[ip] <= 0.;
    u[ip]    = 1.0;
    if (need[ip]) {
      u[ip] = rng.Flat();
      ++count;
    }
  }
  if (count > 1) {
    const double l0 = LogInl(u[0]);
    const double l1 = LogInl(u[1]);
    const double l2 = LogInl(u[2]);
    const double l3 = LogInl(u[3]);
    if (need[0]) nIA[0] = -l0;
    if (need[1]) nIA[1] = -l1;
    if (need[2]) nIA[2] = -l2;
    if (need[3]) nIA[3] = -l3;
  } else if (count == 1) {
    const double us = need[0] ? u[0] : need[1] ? u[1] : need[2] ? u[2] : u[3];
    const double l  = -Log(us);
    if (need[0]) nIA[0] = l;
    if (need[1]) nIA[1] = l;
    if (need[2]) nIA[2] = l;
    if (need[3]) nIA[3] = l;
  }
}

// GetRestMacXSecForStepping (.icc:544-568) without branches: the spline is always evaluated (indices are clamped
// into the table) and the three outcomes -- plateau value, zero below the table, interpolated value -- are selected
G4H_FN double RestMacXSecForSteppingSel(const double* d, double ekin, double lekin) {
  const double log08 = -0.22314355131420971;
  const int numData      = static_cast<int>(G4H_LD(d));
  const double mxsecMaxE = G4H_LD(d + 1);
  const double mxsecMaxV = G4H_LD(d + 2);
  const double logMinE   = G4H_LD(d + 3);
  const double invLD     = G4H_LD(d + 4);
  const double mxsecMinE = G4H_LD(d + 5);
  const bool above    = ekin > mxsecMaxE;
  const double er     = 0.8 * ekin;
  const bool plateau  = above && er < mxsecMaxE;
  const double e2     = above ? er : ekin;
  const double le2    = above ? lekin + log08 : lekin;
  const double mx     = Max(0.0, SplineLogXYSDInl(numData, d + 5, e2, le2, logMinE, invLD));
  return plateau ? Max(0.0, mxsecMaxV) : (e2 < mxsecMinE ? 0.0 : mx);
}

// HowFarToDiscreteInteraction (.icc:48-101) + the lambda_1 look-up of HowFarToMSC (.icc:134-135)
// returns lambda_1 (only meaningful when the MSC step limit will be evaluated)
G4H_FN double HowFarToDiscreteInteractionILP(const TablesView& tv, ElectronState& s) {
  const double theEkin  = s.ekin;
  const double theLEkin = GetLogEKin(s);
  const int theIMC      = s.imc;
  const bool isElectron = !s.isPositron;
  const ElectronTablesView& ed = tv.el[isElectron ? 0 : 1];
  const int theImat = G4H_LD(tv.mcImat + theIMC);
  const int n       = ed.numLoss;
  // range and lambda_1 sit on the same energy grid (same bin, same abscissas)
  const double range = Max(0.0, SplineLogYSDInl(n, ed.lossEGrid, ed.lossData + 5 * n * theIMC, theEkin, theLEkin,
                                                ed.lossLogMinEkin, ed.lossEILDelta));
  const double tr1 = Max(0.0, SplineLogYSDInl(n, ed.lossEGrid, ed.tr1Data + 2 * n * theImat, theEkin, theLEkin,
                                              ed.lossLogMinEkin, ed.lossEILDelta));
  s.range = range;
  const double* rp    = tv.regionPars + 8 * G4H_LD(tv.mcIreg + theIMC);
  const double frange = G4H_LD(rp + kRFinalRange);
  const double drange = G4H_LD(rp + kRDRoverRange);
  // FastDiv: the quotient is only used when range > frange > 0, sigma > 0, sigma_tr1 > 0 (selected below)
  double pStepLength = (range > frange) ? range * drange + frange * (1.0 - drange) * (2.0 - FastDiv(frange, range)) : range;
  // restricted ioni / brem: per couple header {numIoni, ...}; brem follows the 3*numIoni + 5 ioni entries
  const int iIoni   = G4H_LD(ed.resStart + theIMC);
  const int numIoni = static_cast<int>(G4H_LD(ed.resData + iIoni));
  double mxSecs[4];
  mxSecs[0] = RestMacXSecForSteppingSel(ed.resData + iIoni, theEkin, theLEkin);
  mxSecs[1] = RestMacXSecForSteppingSel(ed.resData + iIoni + 3 * numIoni + 5, theEkin, theLEkin);
  mxSecs[2] = 0.0;
  if (!isElectron) mxSecs[2] = MacXSecAnnihilation(0.8 * theEkin, G4H_LD(tv.matPars + 16 * theImat + kMElectronDensity));
  mxSecs[3] = 0.0;
  if (theEkin >= G4H_LD(ed.enucEGrid)) {
    mxSecs[3] = Max(0.0, SplineLogYSD(128, ed.enucEGrid, ed.enucData + theImat * 2 * 128, theEkin, theLEkin,
                                      ed.enucLogMinEkin, ed.enucEILDelta));
  }
  int indxWinnerProcess = -1;
#pragma unroll
  for (int ip = 0; ip < 4; ++ip) {
    const double mxsec = mxSecs[ip];
    const double mfp   = (mxsec > 0.) ? FastDiv(1., mxsec) : kALargeValue;
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
  return tr1 > 0. ? FastDiv(1., tr1) : kALargeValue;  // GetTransportMFP (.icc:576-582)
}

// the condition under which HowFarToMSC evaluates anything (.icc:131-132)
G4H_FN bool MSCStepLimitApplies(double pStepLength, double ekin) {
  const double kGeomMinLength = 5.E-8;
  return pStepLength > kGeomMinLength && ekin > 1.0E-3;
}

// kStoreResults: also reset the result groups HowFar defines (MSC displacement); the fused step skips that
// because its Perform stages overwrite them
G4H_FN void StageHowFarXS(const TablesView& tv, const G4HB200ElectronBatch& b, int64_t i, uint64_t seed) {
  const Meta m   = LoadMeta(b.meta, i);
  const Pair e   = LoadPair(b.ekin_logekin, i);
  const Pair n01 = LoadPair(b.nia01, i);
  const Pair n23 = LoadPair(b.nia23, i);
  ElectronState s;
  s.ekin = e.a; s.logEkin = e.b;
  s.imc = m.imc; s.id = m.id;
  s.isPositron = (static_cast<uint32_t>(m.flags) & G4HB200_F_POSITRON) != 0u;
  s.nIA[0] = n01.a; s.nIA[1] = n01.b; s.nIA[2] = n23.a; s.nIA[3] = n23.b;
  Rng rng;
  rng.Init(seed, static_cast<uint32_t>(m.id), static_cast<uint32_t>(m.draw), false, 0.0);  // no Gauss() in this stage
  ResampleNumIALeftILP(s.nIA, rng);
  const double lambtr1 = HowFarToDiscreteInteractionILP(tv, s);
  const bool msc = MSCStepLimitApplies(s.pStep, s.ekin);
  StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
  StorePair(b.nia01, i, s.nIA[0], s.nIA[1]);
  StorePair(b.nia23, i, s.nIA[2], s.nIA[3]);
  StorePair(b.mfp01, i, s.mfp[0], s.mfp[1]);
  StorePair(b.mfp23, i, s.mfp[2], s.mfp[3]);
  StorePair(b.range_lambtr1, i, s.range, msc ? lambtr1 : 0.0);
  StorePair(b.gstep_pstep, i, s.gStep, s.pStep);
  StoreMeta(b.meta, i, Meta{m.imc, m.flags, m.id, static_cast<int>(rng.draw)});
  b.winner[i] = s.winner;
}

// ---- HowFar, part 2 -----------------------------------------------------------------------------------------
// HowFarToMSC (.icc:103-164) on registers, in lock step: Urban StepLimit (UMSC.icc:19-126) and all four regimes of
// ConvertTrueToGeometricLength (.icc:602-650) through ONE sequence of Exp / inverse range / Log / spline / Log / Exp
// that every lane of the warp walks; a lane selects the result of the regime it is in.  (The branchy version ran
// at 12.5 of 32 lanes and sent the inverse-range regime, one track in seven, through a queue and a third kernel.)
//   in : s.ekin, logEkin, imc, isPositron, onBoundary, mscFirstStep, safety, range, lambtr1, pStep, winner,
//        initialRange, dynRangeFactor, tlimitMin; the Gauss cache; uA, uB = draws `draw`, `draw + 1` of the track
//   out: s.trueStep, zPath, par1-3, mscActive / mscDisplace / mscNoScatter / mscFirstStep, initialRange,
//        dynRangeFactor, tlimitMin, pStep, gStep, winner; returns the number of uniforms consumed
G4H_FN int HowFarMSCCore(const TablesView& tv, ElectronState& s, bool& hasGauss, double& gauss, double uA, double uB,
                         uint32_t k0, uint32_t k1, uint32_t draw) {
  const double pStepLength = s.pStep;
  s.trueStep = pStepLength;
  s.zPath    = pStepLength;
  s.par1 = -1.; s.par2 = 0.; s.par3 = 0.;
  s.mscActive = false;
  if (!MSCStepLimitApplies(pStepLength, s.ekin)) return 0;  // .icc:117-132; the MSC flags keep their values
  s.mscActive = true;
  const bool isElectron = !s.isPositron;
  const ElectronTablesView& et = tv.el[isElectron ? 0 : 1];
  const int imat = G4H_LD(tv.mcImat + s.imc);
  const double* mp = tv.matPars + 16 * imat;
  const double* rp = tv.regionPars + 8 * G4H_LD(tv.mcIreg + s.imc);
  const double range = s.range, presafety = s.safety, ekin = s.ekin, lam = s.lambtr1;
  // --- UMSC::StepLimit (UMSC.icc:19-126)
  const double kTLimitMinfix = 1.0E-8;
  s.mscNoScatter = false;
  s.mscDisplace  = true;
  const bool noLimit = s.trueStep < kTLimitMinfix || range * G4H_LD(mp + kMUMSCPar) < presafety;
  if (noLimit) s.mscDisplace = false;
  double tlimit = 0.0;
  if (!noLimit) {
    const double mscRangeFactor  = G4H_LD(rp + kRMSCRangeFactor);
    const double mscSafetyFactor = G4H_LD(rp + kRMSCSafetyFactor);
    const bool mscIsUseSafety    = !(G4H_LD(rp + kRIsMSCMinimal) != 0.0);
    if (mscIsUseSafety) {
      if (s.mscFirstStep || s.onBoundary) {
        s.initialRange   = Max(range, lam);
        s.dynRangeFactor = lam > 1.0 ? mscRangeFactor * (0.75 + 0.25 * lam) : mscRangeFactor;
        const double stepMin = FastDiv(lam * 1.0E-3, 2.0E-3 + ekin * (G4H_LD(mp + kMStepMin0) + ekin * G4H_LD(mp + kMStepMin1)));
        const double dum0 = isElectron ? 0.87 * G4H_LD(mp + kMZeff23) : 0.70 * G4H_LD(mp + kMZeffSqrt);
        const double dum1 = ekin > 5.0E-3 ? dum0 * stepMin : dum0 * stepMin * 0.5 * (1.0 + ekin * 200.0);
        s.tlimitMin    = Max(dum1, kTLimitMinfix);
        s.mscFirstStep = false;
      }
      tlimit = range > presafety ? Max(Max(s.initialRange * s.dynRangeFactor, mscSafetyFactor * presafety), s.tlimitMin)
                                 : Max(range, s.tlimitMin);
    } else {
      if (s.onBoundary) {
        const double tmpTlimit = range > lam ? mscRangeFactor * range : mscRangeFactor * lam;
        s.initialRange = Max(tmpTlimit, 10 * kTLimitMinfix);
      }
      tlimit = s.initialRange;
    }
  }
  // the Gaussian smearing of the limit (UMSC.icc:117-125) with G4HepEmRandomEngine::Gauss (RandomEngine.hh:50-67):
  // the first pair of uniforms comes from the caller, the logarithm is taken by every lane
  const double tlimitmin = s.tlimitMin;
  const bool smear     = !noLimit && tlimit < s.trueStep;
  const bool needGauss = smear && tlimit > tlimitmin;
  const bool newGauss  = needGauss && !hasGauss;
  int nDraw = 0;
  double v1 = 0.5, v2 = 0.5, r = 0.5;
  if (newGauss) {
    v1 = 2. * uA - 1.;
    v2 = 2. * uB - 1.;
    r  = v1 * v1 + v2 * v2;
    nDraw = 2;
    while (r > 1.) {  // 21 % of the pairs are rejected
      const uint32_t j = draw + static_cast<uint32_t>(nDraw);
      const Uniform2 p = UniformPair(k0, k1, s.id, j >> 1);
      const double w0  = (j & 1u) ? p.b : p.a;
      const double w1  = (j & 1u) ? UniformPair(k0, k1, s.id, (j >> 1) + 1u).a : p.b;
      v1 = 2. * w0 - 1.;
      v2 = 2. * w1 - 1.;
      r  = v1 * v1 + v2 * v2;
      nDraw += 2;
    }
  }
  const double fac = sqrt(-2. * LogInl(r) / r);
  if (needGauss) {
    const double stDev = 0.1 * (tlimit - tlimitmin);
    double g;
    if (hasGauss) {
      hasGauss = false;
      g = gauss * stDev + tlimit;
    } else {
      gauss    = v1 * fac;
      hasGauss = true;
      g = v2 * fac * stDev + tlimit;
    }
    s.trueStep = Min(Max(g, tlimitmin), s.trueStep);
  } else if (smear) {
    s.trueStep = Min(tlimitmin, s.trueStep);
  }
  // --- ConvertTrueToGeometricLength (.icc:602-650), the four regimes side by side
  s.trueStep = Min(s.trueStep, range);
  s.zPath    = s.trueStep;
  const double trueStep = s.trueStep;
  const bool conv = !(trueStep < 1.0E-6);
  const double tau = FastDiv(trueStep, lam);
  const bool reg1 = conv && tau < 1.0e-16;
  const bool reg2 = conv && !reg1 && trueStep < range * 0.05;
  const bool reg3 = conv && !reg1 && !reg2 && (ekin < kElectronMassC2 || trueStep == range);
  const bool reg4 = conv && !reg1 && !reg2 && !reg3;
  const double expTau = ExpInl(reg2 ? -tau : 0.0);
  // regime 4 (.icc:636-647): energy at the end of the step from the inverse range, lambda_1 there
  const double rfin = Max(range - trueStep, 0.01 * range);
  const double t1   = InvRangeInl(et, s.imc, reg4 ? rfin : range);
  const double lt1  = LogInl(t1);
  const int nl      = et.numLoss;
  const double tr1  = Max(0.0, SplineLogYSDInl(nl, et.lossEGrid, et.tr1Data + 2 * nl * imat, t1, lt1, et.lossLogMinEkin, et.lossEILDelta));
  const double lambda1 = tr1 > 0. ? FastDiv(1., tr1) : kALargeValue;
  const bool reg34 = reg3 || reg4;
  const double par1 = reg3 ? 1. / range : (reg4 ? (lam - lambda1) / (lam * trueStep) : 1.0);
  const double par2 = 1. / (par1 * lam);
  const double par3 = 1. + par2;
  const bool powNeeded = reg4 || (reg3 && trueStep < range);
  const double powArg  = reg3 ? 1. - trueStep / range : lambda1 / lam;
  const double logPow  = LogInl(powNeeded ? powArg : 1.0);
  const double powVal  = ExpInl(par3 * logPow);
  if (reg34) {
    s.par1 = par1;
    s.par2 = par2;
    s.par3 = par3;
  }
  double zPath = trueStep;
  if (reg1) zPath = Min(trueStep, lam);
  if (reg2) zPath = (tau < 1.0e-6) ? trueStep * (1. - 0.5 * tau) : lam * (1. - expTau);
  if (reg3) {
    zPath = 1. / (par1 * par3);
    if (trueStep < range) zPath *= (1. - powVal);
  }
  if (reg4) zPath = (1. - powVal) / (par1 * par3);
  if (conv) zPath = Min(zPath, lam);
  s.zPath = zPath;
  // --- the end of HowFarToMSC (.icc:155-163)
  if (trueStep < pStepLength) {
    s.winner = -2;
    s.pStep  = trueStep;
  }
  s.gStep = Min(zPath, s.pStep);
  return nDraw;
}

// uniforms `first` and `first + 1` of a track: two Philox blocks taken by every lane (one suffices when first is even)
G4H_FN Uniform2 DrawPairAt(uint32_t k0, uint32_t k1, uint32_t id, uint32_t first) {
  const Philox4 a = PhiloxBlockInl(k0, k1, id, first >> 1);
  const Philox4 b = PhiloxBlockInl(k0, k1, id, (first >> 1) + 1u);
  const double v0 = ToUniform(a.x, a.y), v1 = ToUniform(a.z, a.w), v2 = ToUniform(b.x, b.y);
  return (first & 1u) ? Uniform2{v1, v2} : Uniform2{v0, v1};
}

// kStoreResults: also reset the result groups HowFar defines (MSC displacement)
template <bool kStoreResults>
G4H_FN void StageHowFarMSC(const TablesView& tv, const G4HB200ElectronBatch& b, int64_t i, uint64_t seed) {
  const Meta m   = LoadMeta(b.meta, i);
  const Pair e   = LoadPair(b.ekin_logekin, i);
  const Pair gp  = LoadPair(b.gstep_pstep, i);
  const Pair rl  = LoadPair(b.range_lambtr1, i);
  const Pair dzs = LoadPair(b.dirz_safety, i);
  const Pair ir  = LoadPair(b.msc_irange_dynrf, i);
  const Pair tg  = LoadPair(b.msc_tlimmin_gauss, i);
  const int winner = b.winner[i];
  uint32_t f = static_cast<uint32_t>(m.flags);
  if (kStoreResults) {
    // fDisplacement = 0 (.icc:127-129); the energy deposit is not a HowFar field
    const Pair ed = LoadPair(b.edep_dispx, i);
    StorePair(b.edep_dispx, i, ed.a, 0.0);
    StorePair(b.dispy_dispz, i, 0.0, 0.0);
  }
  ElectronState s;
  s.ekin = e.a; s.logEkin = e.b;
  s.imc = m.imc; s.id = m.id;
  s.isPositron   = (f & G4HB200_F_POSITRON) != 0u;
  s.onBoundary   = (f & G4HB200_F_ON_BOUNDARY) != 0u;
  s.mscFirstStep = (f & G4HB200_F_MSC_FIRST_STEP) != 0u;
  s.mscDisplace  = (f & G4HB200_F_MSC_DISPLACE) != 0u;
  s.mscNoScatter = (f & G4HB200_F_MSC_NO_SCATTER) != 0u;
  s.safety = dzs.b;
  s.range = rl.a; s.lambtr1 = rl.b;
  s.initialRange = ir.a; s.dynRangeFactor = ir.b; s.tlimitMin = tg.a;
  s.pStep = gp.b; s.gStep = gp.a; s.winner = winner;
  bool hasGauss = (f & G4HB200_F_GAUSS_CACHED) != 0u;
  double gauss  = tg.b;
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  const Uniform2 u = DrawPairAt(k0, k1, static_cast<uint32_t>(m.id), static_cast<uint32_t>(m.draw));
  const int nDraw = HowFarMSCCore(tv, s, hasGauss, gauss, u.a, u.b, k0, k1, static_cast<uint32_t>(m.draw));
  f &= ~(G4HB200_F_MSC_FIRST_STEP | G4HB200_F_MSC_ACTIVE | G4HB200_F_MSC_DISPLACE | G4HB200_F_MSC_NO_SCATTER | G4HB200_F_GAUSS_CACHED);
  if (s.mscFirstStep) f |= G4HB200_F_MSC_FIRST_STEP;
  if (s.mscActive) f |= G4HB200_F_MSC_ACTIVE;
  if (s.mscDisplace) f |= G4HB200_F_MSC_DISPLACE;
  if (s.mscNoScatter) f |= G4HB200_F_MSC_NO_SCATTER;
  if (hasGauss) f |= G4HB200_F_GAUSS_CACHED;
  StorePair(b.msc_irange_dynrf, i, s.initialRange, s.dynRangeFactor);
  StorePair(b.msc_tlimmin_gauss, i, s.tlimitMin, gauss);
  StoreMeta(b.meta, i, Meta{m.imc, static_cast<int>(f), m.id, m.draw + nDraw});
  StorePair(b.tstep_zpath, i, s.trueStep, s.zPath);
  StorePair(b.par12, i, s.par1, s.par2);
  StorePair(b.par3_pad, i, s.par3, 0.0);
  StorePair(b.gstep_pstep, i, s.gStep, s.pStep);
  b.winner[i] = s.winner;
}

// ---- the fused step: HowFar and the along-step part of Perform in one pass over the track -------------------------------
// g4hb200_electron_step: geometry accepts the proposed step, so nothing happens between HowFar and Perform and the
// hand-over state (mean free paths, range, MSC step data) can stay in registers: HowFarXS + HowFarMSC + AlongStep
// as three kernels moved 670 MB per 1M tracks through HBM and were bound by memory latency, the fused stage reads
// the seven persistent groups and writes what the queue kernels behind it need.  par12 / par3_pad are not written.
// Draws come from one DrawWindow (at most four for the interaction lengths + two for the Gaussian).
G4H_FN int ResampleNumIALeftWindow(double* nIA, const DrawWindow& dw) {
  int count = 0;
  double u[4];
#pragma unroll
  for (int ip = 0; ip < 4; ++ip) {
    const bool need = nIA[ip] <= 0.;
    u[ip] = need ? (count == 0 ? dw.u[0] : count == 1 ? dw.u[1] : count == 2 ? dw.u[2] : dw.u[3]) : 1.0;
    count += need ? 1 : 0;
  }
  const double l0 = LogInl(u[0]);
  const double l1 = LogInl(u[1]);
  const double l2 = LogInl(u[2]);
  const double l3 = LogInl(u[3]);
  if (nIA[0] <= 0.) nIA[0] = -l0;
  if (nIA[1] <= 0.) nIA[1] = -l1;
  if (nIA[2] <= 0.) nIA[2] = -l2;
  if (nIA[3] <= 0.) nIA[3] = -l3;
  return count;
}

// What happens between HowFar and Perform: nothing (the proposed step is accepted) ...
struct NoGeometryStep {
  G4H_MFN void operator()(int64_t, ElectronState&, double) const {}
  // gamma: the step limit and what follows it up to SelectInteraction; flags: the track's flag bits (in / out)
  // returns false when the step ends without a boundary and without an interaction (a Woodcock pass that was cut)
  G4H_MFN bool GammaHowFarAndStep(const TablesView& tv, int64_t, GammaState& s, Rng& rng, int&) const {
    GammaHowFar(tv, s, rng);
    return true;
  }
};
// ... or a geometry step (g4h_shower.cuh: SlabGeometryStep) that shortens s.gStep and sets the post-step s.onBoundary.
// returns the queue the track goes to next (kQFluct, kQDiscrete, kQAtRest, kQMscEl, kQMscPos) or -1
template <class GeometryStep>
G4H_FN int StageStepHead(const TablesView& tv, const G4HB200ElectronBatch& b, double* prestep, int64_t i, uint64_t seed,
                         const GeometryStep& geometry) {
  const Meta m   = LoadMeta(b.meta, i);
  const Pair e   = LoadPair(b.ekin_logekin, i);
  const Pair n01 = LoadPair(b.nia01, i);
  const Pair n23 = LoadPair(b.nia23, i);
  const Pair dzs = LoadPair(b.dirz_safety, i);
  const Pair ir  = LoadPair(b.msc_irange_dynrf, i);
  const Pair tg  = LoadPair(b.msc_tlimmin_gauss, i);
  uint32_t f = static_cast<uint32_t>(m.flags);
  ElectronState s;
  s.ekin = e.a; s.logEkin = e.b;
  s.imc = m.imc; s.id = m.id;
  s.isPositron   = (f & G4HB200_F_POSITRON) != 0u;
  s.onBoundary   = (f & G4HB200_F_ON_BOUNDARY) != 0u;
  s.mscFirstStep = (f & G4HB200_F_MSC_FIRST_STEP) != 0u;
  s.mscDisplace  = (f & G4HB200_F_MSC_DISPLACE) != 0u;
  s.mscNoScatter = (f & G4HB200_F_MSC_NO_SCATTER) != 0u;
  s.safety = dzs.b;
  s.nIA[0] = n01.a; s.nIA[1] = n01.b; s.nIA[2] = n23.a; s.nIA[3] = n23.b;
  s.initialRange = ir.a; s.dynRangeFactor = ir.b; s.tlimitMin = tg.a;
  s.preStepEkin = 0.0; s.preStepLogEkin = 0.0;
  bool hasGauss = (f & G4HB200_F_GAUSS_CACHED) != 0u;
  double gauss  = tg.b;
  DrawWindow dw;
  dw.Init(seed, static_cast<uint32_t>(m.id), static_cast<uint32_t>(m.draw));
  // HowFar (.icc:35-45): interaction lengths, discrete step limit, MSC step limit
  const int nXS = ResampleNumIALeftWindow(s.nIA, dw);
  const double lam = HowFarToDiscreteInteractionILP(tv, s);
  s.lambtr1 = MSCStepLimitApplies(s.pStep, s.ekin) ? lam : 0.0;
  const double uA = nXS == 0 ? dw.u[0] : nXS == 1 ? dw.u[1] : nXS == 2 ? dw.u[2] : nXS == 3 ? dw.u[3] : dw.u[4];
  const double uB = nXS == 0 ? dw.u[1] : nXS == 1 ? dw.u[2] : nXS == 2 ? dw.u[3] : nXS == 3 ? dw.u[4] : dw.Sixth();
  const int nMSC = HowFarMSCCore(tv, s, hasGauss, gauss, uA, uB, dw.k0, dw.k1, static_cast<uint32_t>(m.draw + nXS));
  // the geometry step (or none: fGStepLength and fOnBoundary stay); Perform, along-step part
  geometry(i, s, dzs.a);
  const int route = AlongStepCore(tv, s);
  f &= ~(G4HB200_F_ON_BOUNDARY | G4HB200_F_MSC_FIRST_STEP | G4HB200_F_MSC_ACTIVE | G4HB200_F_MSC_DISPLACE | G4HB200_F_MSC_NO_SCATTER |
         G4HB200_F_GAUSS_CACHED);
  if (s.onBoundary) f |= G4HB200_F_ON_BOUNDARY;
  if (s.mscFirstStep) f |= G4HB200_F_MSC_FIRST_STEP;
  if (s.mscActive) f |= G4HB200_F_MSC_ACTIVE;
  if (s.mscDisplace) f |= G4HB200_F_MSC_DISPLACE;
  if (s.mscNoScatter) f |= G4HB200_F_MSC_NO_SCATTER;
  if (hasGauss) f |= G4HB200_F_GAUSS_CACHED;
  StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
  StorePair(b.nia01, i, s.nIA[0], s.nIA[1]);
  StorePair(b.nia23, i, s.nIA[2], s.nIA[3]);
  StorePair(b.msc_irange_dynrf, i, s.initialRange, s.dynRangeFactor);
  StorePair(b.msc_tlimmin_gauss, i, s.tlimitMin, gauss);
  StoreMeta(b.meta, i, Meta{m.imc, static_cast<int>(f), m.id, m.draw + nXS + nMSC});
  StorePair(b.gstep_pstep, i, s.gStep, s.pStep);
  StorePair(b.edep_dispx, i, s.edep, 0.0);  // fDisplacement = 0 (.icc:127-129)
  StorePair(b.dispy_dispz, i, 0.0, 0.0);
  b.winner[i] = s.winner;
  StorePair(b.mfp01, i, s.mfp[0], s.mfp[1]);
  StorePair(b.mfp23, i, s.mfp[2], s.mfp[3]);
  StorePair(b.range_lambtr1, i, s.range, s.lambtr1);
  StorePair(b.tstep_zpath, i, s.trueStep, s.zPath);
  StorePair(prestep, i, s.preStepEkin, s.preStepLogEkin);
  return route == -2 ? -1 : route;
}

// ---- gamma step in two stages ----------------------------------------------------------------------------------
//   StageGammaHead       HowFar (G4HepEmGammaManager.icc:27-48; kMode 2) + SelectInteraction (.icc:173-219) +
//                        UpdateNumIALeft and the head of Perform (.icc:54-76): returns the process that interacts
//                        (0 conversion, 1 Compton, 2 photoelectric) or -1
//   StageGammaInteract   the final state sampler of that process + the tracking cut (.icc:77-94), over a queue
// geometry (kMode 2 only): what happens between HowFar and Perform, see NoGeometryStep
template <int kMode, class GeometryStep>
G4H_FN int StageGammaHead(const TablesView& tv, const G4HB200GammaBatch& b, int64_t i, uint64_t seed, const GeometryStep& geometry) {
  GammaState s;
  Rng rng;
  LoadGamma(b, i, seed, s, rng);
  int flags = b.meta[4 * i + 1];
  if (kMode == 1) LoadGammaHandOver(b, i, s);
  bool interacts = true;
  if (kMode == 2) {
    interacts = geometry.GammaHowFarAndStep(tv, i, s, rng, flags);
    flags = s.onBoundary ? (flags | static_cast<int>(G4HB200_F_ON_BOUNDARY)) : (flags & ~static_cast<int>(G4HB200_F_ON_BOUNDARY));
  }
  int route = -1;
  if (!interacts) {
    s.edep = 0.0;
    StoreGamma(b, i, s, rng, flags);
    return -1;
  }
  if (!s.onBoundary) {
    const double urnd = rng.Flat();
    s.nIA0 = -1.0;
    const double lekin = (s.ekin > tv.gmEMax1) ? GetLogEKin(s) : 0.0;
    s.winner = GammaSampleInteraction(tv, G4H_LD(tv.mcImat + s.imc), s.ekin, lekin, s.mfp0, urnd, s.peMXsec);
  }
  s.nIA0 -= s.gStep / s.mfp0;
  s.edep = 0.0;
  if (!s.onBoundary) {
    if (s.winner == 0) s.nIA0 = -1.0;
    if (s.winner >= 0 && s.winner <= 2) {
      route = s.winner;
    } else if (s.ekin > 0.0 && s.ekin <= tv.gammaTrackingCut) {
      // no interaction (gamma-nuclear slot): only the tracking cut of .icc:88-93 applies
      s.edep += s.ekin;
      SetEKin(s, 0.0);
    }
  }
  StoreGamma(b, i, s, rng, flags);
  return route;
}

template <int kProc>
G4H_FN void StageGammaInteract(const TablesView& tv, const G4HB200GammaBatch& b, int64_t i, uint64_t seed, Secondaries& sec,
                               int& id, double* window = nullptr, uint32_t windowStride = 0u, uint32_t windowSlots = 0u) {
  const Meta m   = LoadMeta(b.meta, i);
  const Pair e   = LoadPair(b.ekin_logekin, i);
  const Pair dxy = LoadPair(b.dirx_diry, i);
  const Pair dzn = LoadPair(b.dirz_nia0, i);
  const Pair ep  = LoadPair(b.edep_pemxsec, i);
  GammaState s;
  s.ekin = e.a; s.logEkin = e.b;
  s.dir[0] = dxy.a; s.dir[1] = dxy.b; s.dir[2] = dzn.a;
  s.imc = m.imc; s.id = m.id;
  s.edep = ep.a; s.peMXsec = ep.b;
  id = m.id;
  Rng rng;
  rng.Init(seed, static_cast<uint32_t>(m.id), static_cast<uint32_t>(m.draw), false, 0.0);
  if (windowSlots != 0u) rng.FillWindow(window, windowStride, windowSlots);
  if (kProc == 0) PerformConversion(tv, s, rng, sec);
  if (kProc == 1) PerformCompton(tv, s, rng, sec);
  if (kProc == 2) PerformPhotoelectric(tv, s, rng, sec);
  const double finalEkin = s.ekin;
  if (finalEkin > 0.0 && finalEkin <= tv.gammaTrackingCut) {
    SetEKin(s, 0.0);
    s.edep += finalEkin;
  }
  StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
  StorePair(b.dirx_diry, i, s.dir[0], s.dir[1]);
  StorePair(b.dirz_nia0, i, s.dir[2], dzn.b);
  StorePair(b.edep_pemxsec, i, s.edep, s.peMXsec);
  StoreMeta(b.meta, i, Meta{m.imc, m.flags, m.id, static_cast<int>(rng.draw)});
}

}  // namespace g4h
#endif
