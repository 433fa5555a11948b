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
// the end of HowFarToMSC (.icc:155-163) once zPath is known
G4H_FN void FinishHowFarMSC(const G4HB200ElectronBatch& b, int64_t i, const ElectronState& s, double pStepLength, int winner) {
  if (s.trueStep < pStepLength) {
    winner      = -2;
    pStepLength = s.trueStep;
  }
  const double gStep = Min(s.zPath, pStepLength);
  StorePair(b.tstep_zpath, i, s.trueStep, s.zPath);
  StorePair(b.par12, i, s.par1, s.par2);
  StorePair(b.par3_pad, i, s.par3, 0.0);
  StorePair(b.gstep_pstep, i, gStep, pStepLength);
  b.winner[i] = winner;
}

// returns true when the true -> geometrical conversion needs the inverse range regime: the track then goes to the
// kQConvRange queue and StageHowFarMSCRange finishes it
template <bool kStoreResults>
G4H_FN bool StageHowFarMSC(const TablesView& tv, const G4HB200ElectronBatch& b, int64_t i, uint64_t seed) {
  const Meta m  = LoadMeta(b.meta, i);
  const Pair e  = LoadPair(b.ekin_logekin, i);
  const Pair gp = LoadPair(b.gstep_pstep, i);
  uint32_t f = static_cast<uint32_t>(m.flags);
  const double pStepLength = gp.b;
  if (kStoreResults) {
    // fDisplacement = 0 (.icc:127-129); the energy deposit is not a HowFar field
    const Pair ed = LoadPair(b.edep_dispx, i);
    StorePair(b.edep_dispx, i, ed.a, 0.0);
    StorePair(b.dispy_dispz, i, 0.0, 0.0);
  }
  if (!MSCStepLimitApplies(pStepLength, e.a)) {
    // HowFarToMSC (.icc:117-130): true = z = physical step, MSC inactive, no displacement; the rest keeps its
    // G4HepEmMSCTrackData::ReSet() value
    f &= ~G4HB200_F_MSC_ACTIVE;
    StorePair(b.tstep_zpath, i, pStepLength, pStepLength);
    StorePair(b.par12, i, -1.0, 0.0);
    StorePair(b.par3_pad, i, 0.0, 0.0);
    StoreMeta(b.meta, i, Meta{m.imc, static_cast<int>(f), m.id, m.draw});
    return false;
  }
  const Pair rl  = LoadPair(b.range_lambtr1, i);
  const Pair dzs = LoadPair(b.dirz_safety, i);
  const Pair ir  = LoadPair(b.msc_irange_dynrf, i);
  const Pair tg  = LoadPair(b.msc_tlimmin_gauss, i);
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
  s.pStep = pStepLength;
  Rng rng;
  rng.Init(seed, static_cast<uint32_t>(m.id), static_cast<uint32_t>(m.draw), (f & G4HB200_F_GAUSS_CACHED) != 0u, tg.b);
  const bool isElectron = !s.isPositron;
  const int theImat = G4H_LD(tv.mcImat + s.imc);
  const int theIreg = G4H_LD(tv.mcIreg + s.imc);
  s.trueStep  = pStepLength;
  s.zPath     = pStepLength;
  s.mscActive = true;
  UMSCStepLimit(tv, s, s.ekin, theImat, theIreg, s.range, s.safety, s.onBoundary, isElectron, rng);
  const bool rangeRegime = ConvertTrueToGeometricLengthHead(s, s.ekin, s.range);
  f &= ~(G4HB200_F_MSC_FIRST_STEP | G4HB200_F_MSC_ACTIVE | G4HB200_F_MSC_DISPLACE | G4HB200_F_MSC_NO_SCATTER |
         G4HB200_F_GAUSS_CACHED);
  if (s.mscFirstStep) f |= G4HB200_F_MSC_FIRST_STEP;
  f |= G4HB200_F_MSC_ACTIVE;
  if (s.mscDisplace) f |= G4HB200_F_MSC_DISPLACE;
  if (s.mscNoScatter) f |= G4HB200_F_MSC_NO_SCATTER;
  if (rng.hasGauss) f |= G4HB200_F_GAUSS_CACHED;
  StorePair(b.msc_irange_dynrf, i, s.initialRange, s.dynRangeFactor);
  StorePair(b.msc_tlimmin_gauss, i, s.tlimitMin, rng.gauss);
  StoreMeta(b.meta, i, Meta{m.imc, static_cast<int>(f), m.id, static_cast<int>(rng.draw)});
  if (rangeRegime) {
    // hand trueStep over (zPath = trueStep for now); gstep_pstep and the winner still hold the discrete limit
    StorePair(b.tstep_zpath, i, s.trueStep, s.zPath);
    return true;
  }
  FinishHowFarMSC(b, i, s, pStepLength, b.winner[i]);
  return false;
}

// the inverse-range regime of the true -> geometrical conversion (.icc:636-647) for the tracks queued by StageHowFarMSC
G4H_FN void StageHowFarMSCRange(const TablesView& tv, const G4HB200ElectronBatch& b, int64_t i) {
  const Meta m  = LoadMeta(b.meta, i);
  const Pair gp = LoadPair(b.gstep_pstep, i);
  const Pair rl = LoadPair(b.range_lambtr1, i);
  const Pair tz = LoadPair(b.tstep_zpath, i);
  ElectronState s;
  s.lambtr1  = rl.b;
  s.trueStep = tz.a;
  s.zPath    = tz.b;
  ConvertTrueToGeometricLengthRangeRegime(tv, s, rl.a, m.imc, (static_cast<uint32_t>(m.flags) & G4HB200_F_POSITRON) == 0u);
  FinishHowFarMSC(b, i, s, gp.b, b.winner[i]);
}

// ---- gamma step in two stages ----------------------------------------------------------------------------------
//   StageGammaHead       HowFar (G4HepEmGammaManager.icc:27-48; kMode 2) + SelectInteraction (.icc:173-219) +
//                        UpdateNumIALeft and the head of Perform (.icc:54-76): returns the process that interacts
//                        (0 conversion, 1 Compton, 2 photoelectric) or -1
//   StageGammaInteract   the final state sampler of that process + the tracking cut (.icc:77-94), over a queue
template <int kMode>
G4H_FN int StageGammaHead(const TablesView& tv, const G4HB200GammaBatch& b, int64_t i, uint64_t seed) {
  GammaState s;
  Rng rng;
  LoadGamma(b, i, seed, s, rng);
  const int flags = b.meta[4 * i + 1];
  if (kMode == 1) LoadGammaHandOver(b, i, s);
  if (kMode == 2) GammaHowFar(tv, s, rng);
  int route = -1;
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
                               int& id) {
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
  if (kProc == 0) PerformConversion(tv, s, rng, sec);
  if (kProc == 1) PerformCompton(s, rng, sec);
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
