// g4h_msc_f32.cuh -- SampleMSC in single precision: an OFFERED variant (g4hb200_set_msc_precision(h, 32)), not the drop-in.
//
// The reference's authors note that the Urban model parameters "could probably be computed in float"
// (G4HepEmElectronInteractionUMSC.icc:296,312).  StageMSCSampleF32 is StageMSCSample (g4h_perform_stages.cuh: SampleMSC
// .icc:261-322, UMSC.icc:129-357 in lock step) with the model parameters AND the sampling of the polar angle in float
// and the hardware's exp2 / log2 / sin / cos units behind __expf / __logf / __sincosf instead of the 60-instruction
// FP64 VDT chains.  Two things keep it usable at small angles: the angle is carried as 1 - cos(theta) (the sampling
// formulas all have the form 1 + small or -1 + large), and the new direction is assembled and rotated in double.
//
// What stays identical to the FP64 stage: which tracks are scattered at all, the uniform stream and -- for all but the
// few tracks in a million whose model regime is decided within float rounding of a threshold (tau > 8, theta0 > pi/6,
// xmean1 <= 0.999 xmeanth ...) -- the number of uniforms consumed, hence everything the later stages sample.  What
// differs: direction and displacement, within the bound tests/test_msc_f32.py states and checks against the FP64 stage.
#ifndef G4H_MSC_F32_CUH
#define G4H_MSC_F32_CUH

#include "g4h_perform_stages.cuh"

namespace g4h {

__device__ __forceinline__ float SplineLogYSDF32(int ndata, const double* xdata, const double* ydata, float x, float logx,
                                                 float logxmin, float invLDBin) {
  const float xlo = static_cast<float>(xdata[0]), xhi = static_cast<float>(xdata[ndata - 1]);
  const float xv  = fmaxf(xlo, fminf(xhi, x));
  const int idx   = static_cast<int>(fmaxf(0.f, fminf((logx - logxmin) * invLDBin, static_cast<float>(ndata) - 2.f)));
  const float x1 = static_cast<float>(xdata[idx]), x2 = static_cast<float>(xdata[idx + 1]);
  const float y1 = static_cast<float>(ydata[2 * idx]), y2 = static_cast<float>(ydata[2 * idx + 2]);
  const float s1 = static_cast<float>(ydata[2 * idx + 1]), s2 = static_cast<float>(ydata[2 * idx + 3]);
  const float dl = x2 - x1;
  const float b  = fmaxf(0.f, fminf(1.f, __fdividef(xv - x1, dl)));
  const float c0 = (2.0f - b) * s1;
  const float c1 = (1.0f + b) * s2;
  return y1 + b * (y2 - y1) + (b * (b - 1.0f)) * (c0 + c1) * (dl * dl * 0.166666666667f);
}

// returns the queue the track goes to next (kQFluct, kQDiscrete, kQAtRest) or -1; arguments as StageMSCSample
template <bool kPositron>
__device__ __forceinline__ int StageMSCSampleF32(const TablesView& tv, const G4HB200ElectronBatch& b, double* prestep, int64_t i,
                                                 uint64_t seed, const double* steppre = nullptr) {
  const Meta m   = LoadMeta(b.meta, i);
  const Pair pre = LoadPair(prestep, i);
  const Pair e   = LoadPair(b.ekin_logekin, i);
  const Pair dxy = LoadPair(b.dirx_diry, i);
  const Pair dzs = LoadPair(b.dirz_safety, i);
  const Pair gp  = LoadPair(b.gstep_pstep, i);
  const Pair rl  = LoadPair(b.range_lambtr1, i);
  const Pair tz  = LoadPair(b.tstep_zpath, i);
  const Pair tg  = LoadPair(b.msc_tlimmin_gauss, i);
  const Pair ed0 = LoadPair(b.edep_dispx, i);
  uint32_t f = static_cast<uint32_t>(m.flags);
  DrawWindow dw;
  dw.Init(seed, static_cast<uint32_t>(m.id), static_cast<uint32_t>(m.draw));
  const double pStepLengthD = gp.b;
  const double zPathD       = tz.b;
  const float pStepLength   = static_cast<float>(pStepLengthD);
  const float preStepEkin   = static_cast<float>(pre.a);
  const float preStepTr1mfp = static_cast<float>(rl.b);
  const ElectronTablesView& et = tv.el[kPositron ? 1 : 0];
  const int theImat = tv.mcImat[m.imc];
  const double* mp  = tv.matPars + 16 * theImat;
  // --- SampleMSC (.icc:272-285): transport mean free path at the post-step energy
  const bool usePost = pStepLengthD > rl.a * 0.01;  // G4VERSION_NUM >= 1100
  const float ekinNow = static_cast<float>(e.a);
  const float postStepEkin  = usePost ? ekinNow : preStepEkin;
  const float postStepLEkin = usePost ? (ekinNow > 0.f ? __logf(ekinNow) : -30.f) : static_cast<float>(pre.b);
  const int nl = et.numLoss;
  const float tr1 = fmaxf(0.f, SplineLogYSDF32(nl, et.lossEGrid, et.tr1Data + 2 * nl * theImat, postStepEkin, postStepLEkin,
                                               static_cast<float>(et.lossLogMinEkin), static_cast<float>(et.lossEILDelta)));
  const float postStepTr1mfp = tr1 > 0.f ? __fdividef(1.f, tr1) : 1.0e20f;
  // --- SampleCosineTheta (UMSC.icc:153-271)
  const float radLength = static_cast<float>(mp[kMRadLength]);
  const float zeff      = static_cast<float>(mp[kMZeff]);
  const float iPreStepTr1mfp = __fdividef(1.0f, preStepTr1mfp);
  const float deltaR1mfp     = preStepTr1mfp - postStepTr1mfp;
  const bool bigDelta = fabsf(deltaR1mfp) > 0.01f * preStepTr1mfp;
  const float lratio  = __logf(bigDelta ? __fdividef(preStepTr1mfp, postStepTr1mfp) : 1.0f);
  const float tau     = bigDelta ? __fdividef(pStepLength * lratio, deltaR1mfp) : pStepLength * iPreStepTr1mfp;
  const bool isIso  = tau > 8.0f;
  const bool isNone = !isIso && tau < 1.0E-16f;
  const bool tauSmall = tau < 0.01f;
  const float em1 = __expf(tauSmall ? 0.0f : -tau);
  const float em2 = __expf(tauSmall ? 0.0f : -2.5f * tau);
  const float xmeanth  = tauSmall ? 1.0f - tau * (1.0f - 0.5f * tau) : em1;
  const float x2meanth = tauSmall ? 1.0f - tau * (5.0f - 6.25f * tau) * 0.333333f : (1.0f + 2.0f * em2) * 0.333333f;
  const bool early = isIso || isNone;
  bool isSimple = !early && usePost && e.a < 0.5 * pre.a;  // :409-411 on the exact energies (post == pre when !usePost)
  // theta0 (ComputeTheta0, UMSC.icc:293-305)
  const float tsmall       = fminf(static_cast<float>(tg.a), 1.0f);
  const bool stpNotExSmall = pStepLength > tsmall;
  const float stepInRadLength = __fdividef(stpNotExSmall ? pStepLength : tsmall, radLength);
  const float kM = static_cast<float>(kElectronMassC2);
  const float postInvBetaPc = __fdividef(postStepEkin + kM, postStepEkin * (postStepEkin + 2.f * kM));
  const float invBetaPc     = preStepEkin != postStepEkin
                                  ? sqrtf(postInvBetaPc * __fdividef(preStepEkin + kM, preStepEkin * (preStepEkin + 2.f * kM)))
                                  : postInvBetaPc;
  float y = stepInRadLength;
  if (kPositron && tv.isMSCPositronCor != 0) {
    // Theta0PositronCorrection (UMSC.icc:309-335)
    const float eekin = preStepEkin * postStepEkin;
    const float ff = 1.f + zeff * (1.84035E-4f * zeff - 1.86427E-2f) + 0.41125f;
    const float a  = 0.994f - 4.08E-3f * zeff;
    const float bb = 7.16f + __fdividef(52.6f + __fdividef(365.f, zeff), zeff);
    const float tu = sqrtf(eekin) * static_cast<float>(kInvElectronMassC2);
    const float x  = sqrtf(__fdividef(tu * (tu + 2.f), (tu + 1.f) * (tu + 1.f)));
    const float xl = 0.6f, xh = 0.9f, ee = 113.0f;
    const float c  = 1.00f - 4.47E-3f * zeff;
    const float d  = 1.21E-3f * zeff;
    const float expLow  = __expf(x < xl ? -bb * x : -bb * xl);
    const float expHigh = __expf(x > xh ? ee * (x - 1.f) : ee * (xh - 1.f));
    const float yl = a * (1.f - expLow);
    const float yh = c + d * expHigh;
    const float y0 = __fdividef(yh - yl, xh - xl);
    const float y1 = yl - y0 * xl;
    const float corr = x < xl ? ff * a * (1.f - expLow) : (x > xh ? ff * (c + d * expHigh) : ff * (y0 * x + y1));
    y = stepInRadLength * corr;
  }
  const float logY = __logf(y);
  float theta0 = 13.6f * sqrtf(y) * invBetaPc * (static_cast<float>(mp[kMTheta0]) + static_cast<float>(mp[kMTheta1]) * logY);
  if (!stpNotExSmall) theta0 = theta0 * sqrtf(__fdividef(pStepLength, tsmall));
  isSimple = isSimple || (!early && theta0 > static_cast<float>(kPi * 0.166666));
  const float theta2 = theta0 * theta0;
  const bool isNone2 = !early && !isSimple && theta2 < 1.0E-16f;
  // the tail parameters (:426-434)
  const float dumtau    = stpNotExSmall ? tau : tsmall * iPreStepTr1mfp;
  const float logDumtau = __logf(dumtau);
  const float logTail   = __logf(__fdividef(pStepLength, tau * radLength));
  const float parU   = __expf(0.1666666f * logDumtau);
  const float dumxsi = static_cast<float>(mp[kMTail0]) + parU * (static_cast<float>(mp[kMTail1]) + parU * static_cast<float>(mp[kMTail2])) +
                       static_cast<float>(mp[kMTail3]) * logTail;
  const float parXsi = fmaxf(dumxsi, 1.9f);
  const float parC   = fabsf(parXsi - 3.f) < 0.001f ? 3.001f : fabsf(parXsi - 2.f) < 0.001f ? 2.001f : parXsi;
  const float dumC1  = parC - 1.f;
  const float dumEa  = __expf(-parXsi);
  const float dumEaa = __fdividef(1.f, 1.f - dumEa);
  const bool inFlow = !early && !isSimple && !isNone2;
  float thex = theta2 * (1.0f - theta2 * 0.0833333f);
  if (inFlow && theta2 > 0.01f) {
    const float dum = 2.0f * __sinf(0.5f * theta0);
    thex = dum * dum;
  }
  const float xmean1 = 1.f - (1.f - (1.f + parXsi) * dumEa) * thex * dumEaa;
  isSimple = isSimple || (inFlow && xmean1 <= 0.999f * xmeanth);
  const bool isMain = inFlow && !(xmean1 <= 0.999f * xmeanth);
  const float x0 = 1.f - parXsi * thex;
  const float bx = parC * thex;
  const float b1 = bx + x0 + 1.f;
  const float logB1 = __logf(isMain ? b1 : 1.0f);
  const float logBx = __logf(isMain ? bx : 1.0f);
  const float eb1 = __expf(dumC1 * logB1);
  const float ebx = __expf(dumC1 * logBx);
  const float d   = isMain ? __fdividef(ebx, eb1) : 0.5f;
  const float xmean2 = __fdividef(x0 + d - __fdividef(bx - b1 * d, parC - 2.f), 1.f - d);
  const float f1x0 = dumEa * dumEaa;
  const float f2x0 = __fdividef(dumC1, parC * (1.f - d));
  const float prob = __fdividef(f2x0, f1x0 + f2x0);
  const float qprb = __fdividef(xmeanth, prob * xmean1 + (1.f - prob) * xmean2);
  // SimpleScattering (UMSC.icc:274-289)
  const float sdum0 = 3.f * x2meanth - 1.f;
  const float sdum1 = 2.f * xmeanth - sdum0;
  const float sa    = 1.f + __fdividef(4.f * sdum0, sdum1);
  const float sprob = __fdividef((2.f + sa) * xmeanth, sa);
  const int nCost = (isNone || isNone2) ? 0 : (isIso ? 1 : (isSimple ? 2 : 3));
  const float r0 = static_cast<float>(dw.u[0]), r1 = static_cast<float>(dw.u[1]), r2 = static_cast<float>(dw.u[2]);
  const bool mainIn   = isMain && r0 < qprb;
  const bool mainExp  = mainIn && r1 < prob;
  const float var0    = (1.0f - d) * r2;
  const bool mainTail = mainIn && !mainExp;
  const bool tailSer  = mainTail && var0 < 0.01f * d;
  const bool tailPow  = mainTail && !tailSer;
  const bool simPow   = isSimple && r0 < sprob;
  // exponential part: log(dumEa + r2 / dumEaa) = log1p(-(1 - r2)(1 - dumEa)); the argument is close to 1 for most tracks and
  // its float rounding alone would be 1e-3 of the logarithm
  const float logExp = log1pf(-(static_cast<float>(1.0 - dw.u[2]) * (1.f - dumEa)));
  const float logArg = tailPow ? var0 + d : (simPow ? r1 : 1.0f);
  const float logFin = mainExp ? logExp : __logf(logArg);
  const float expFin = __expf(tailPow ? __fdividef(-1.f, dumC1) * logFin : (simPow ? __fdividef(1.f, 1.f + sa) * logFin : 0.0f));
  // 1 - cos(theta)
  float omc;
  if (isNone || isNone2) {
    omc = 0.0f;
  } else if (isIso) {
    omc = 2.0f - 2.0f * r0;
  } else if (isSimple) {
    omc = simPow ? 2.f - 2.f * expFin : 2.f - 2.f * r1;
  } else if (mainExp) {
    omc = -logFin * thex;
  } else if (tailSer) {
    const float var = __fdividef(var0, d * dumC1);
    omc = 2.0f - var * (1.0f - var * 0.5f * parC) * b1;
  } else if (tailPow) {
    omc = -thex * (parC - parXsi - parC * expFin);
  } else {
    omc = 2.0f - 2.0f * r1;
  }
  int nDraw = nCost;
  double dir[3]  = {dxy.a, dxy.b, dzs.a};
  double disp[3] = {ed0.b, 0.0, 0.0};
  bool storeDisp = false;
  // the reference's test is on cos(theta) as a double: an angle below 1.5e-8 rad does not scatter (UMSC.icc:137-140)
  const double cost = 1.0 - static_cast<double>(omc);
  if (fabs(cost) >= 1.0) {
    f |= G4HB200_F_MSC_NO_SCATTER;
  } else {
    const double sth  = static_cast<double>(sqrtf(omc * (2.0f - omc)));
    const double uPhi = nCost == 0 ? dw.u[0] : nCost == 1 ? dw.u[1] : nCost == 2 ? dw.u[2] : dw.u[3];
    const float phi = static_cast<float>(k2Pi) * static_cast<float>(uPhi);
    ++nDraw;
    const bool displace = (f & G4HB200_F_MSC_DISPLACE) != 0u && tv.isMSCDisplacement != 0;
    f = displace ? f : (f & ~G4HB200_F_MSC_DISPLACE);
    const bool sampleDisp = displace && pStepLengthD > zPathD;
    float uD0 = 0.0f, uD1 = 0.0f;
    if (sampleDisp) {
      uD0 = static_cast<float>(nCost == 0 ? dw.u[1] : nCost == 1 ? dw.u[2] : nCost == 2 ? dw.u[3] : dw.u[4]);
      uD1 = static_cast<float>(nCost == 0 ? dw.u[2] : nCost == 1 ? dw.u[3] : nCost == 2 ? dw.u[4] : dw.Sixth());
      nDraw += 2;
    }
    const float cbeta  = 2.16f;
    const float cbeta1 = 1.f - __expf(-cbeta * static_cast<float>(kPi));
    const float psi    = __fdividef(-__logf(sampleDisp ? 1.f - uD0 * cbeta1 : 1.0f), cbeta);
    const float dphi   = (uD1 < 0.5f) ? phi + psi : phi - psi;
    float sphi, cphi;
    __sincosf(phi, &sphi, &cphi);
    double newDir[3] = {sth * cphi, sth * sphi, cost};
    if (displace) {
      const Pair dyz = LoadPair(b.dispy_dispz, i);
      disp[1] = dyz.a;
      disp[2] = dyz.b;
      if (sampleDisp) {
        // the two factors separately: steps in near-vacuum reach 1e24 mm and their product leaves the float range
        const float r = 0.73f * sqrtf(static_cast<float>(pStepLengthD - zPathD)) * sqrtf(static_cast<float>(pStepLengthD + zPathD));
        float sd, cd;
        __sincosf(dphi, &sd, &cd);
        disp[0] = r * cd;
        disp[1] = r * sd;
        disp[2] = 0.0;
      }
      RotateToReferenceFrame(disp, dir);
      storeDisp = true;
    }
    RotateToReferenceFrame(newDir, dir);
    // the float sine / cosine leave the vector a few 1e-7 off unit length: renormalise (the reference's is exact to 1e-16)
    const double inv = rsqrt(newDir[0] * newDir[0] + newDir[1] * newDir[1] + newDir[2] * newDir[2]);
    StorePair(b.dirx_diry, i, newDir[0] * inv, newDir[1] * inv);
    StorePair(b.dirz_safety, i, newDir[2] * inv, dzs.b);
    if (storeDisp) StorePair(b.dispy_dispz, i, disp[1], disp[2]);
  }
  if ((f & G4HB200_F_MSC_SUBSTEP) != 0u) {
    if (storeDisp) StorePair(b.edep_dispx, i, ed0.a, disp[0]);
    StoreMeta(b.meta, i, Meta{m.imc, static_cast<int>(f), m.id, m.draw + nDraw});
    return -1;
  }
  double stepPreEkin = pre.a;
  if (steppre != nullptr) {
    const Pair sp = LoadPair(steppre, i);
    stepPreEkin = sp.a;
    StorePair(prestep, i, sp.a, sp.b);
  }
  // SampleLossFluctuations (.icc:324-368): sampled by StageFluctuation, or finished here -- as in the FP64 stage
  int route = -1;
  double ekin = e.a, edep = ed0.a;
  bool storeEkin = false;
  const bool isFluct = tv.regionPars[8 * tv.mcIreg[m.imc] + kRIsFluct] != 0.0;
  if (isFluct && edep > 1.E-5) {
    route = kQFluct;
  } else if (ekin <= tv.elTrackingCut) {
    ekin = 0.0;
    edep = stepPreEkin;
    storeEkin = true;
    if (kPositron) route = kQAtRest;
  } else if (b.winner[i] >= 0 && (f & G4HB200_F_ON_BOUNDARY) == 0u) {
    route = kQDiscrete;
  }
  if (storeEkin) StorePair(b.ekin_logekin, i, ekin, 100.0);
  if (storeEkin || storeDisp) StorePair(b.edep_dispx, i, edep, disp[0]);
  StoreMeta(b.meta, i, Meta{m.imc, static_cast<int>(f), m.id, m.draw + nDraw});
  return route;
}

}  // namespace g4h
#endif
