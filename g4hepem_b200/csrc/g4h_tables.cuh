// g4h_tables.cuh -- device view of the flattened tables and every table look-up of the stepping path.
//
// The arena is one contiguous device allocation (doubles first, then int32s); TablesView holds
// typed pointers into it and is passed to kernels by value (it lives in the constant bank).
// Look-ups restate G4HepEmRun/include/G4HepEmRunUtils.icc:49-237 (spline / linear interpolation,
// bin search) and the accessors of G4HepEmElectronManager.icc:486-599 / G4HepEmGammaManager.icc:108-219
// with identical floating point operation order (no FMA), so values agree bit for bit with the CPU.
#ifndef G4H_TABLES_CUH
#define G4H_TABLES_CUH

#include "g4h_math.cuh"

namespace g4h {

struct ElectronTablesView {
  int numLoss;
  double lossLogMinEkin, lossEILDelta;
  const double* lossEGrid;   // [numLoss]
  const double* lossData;    // [5*numLoss*numMatCut]
  const int* resStart;       // [numMatCut]
  const double* resData;
  double enucLogMinEkin, enucEILDelta;
  const double* enucEGrid;   // [128]
  const double* enucData;    // [2*128*numMat]
  const double* tr1Data;     // [2*numLoss*numMat]
  const int* selSBStart;
  const double* selSBData;
  const int* selRBStart;
  const double* selRBData;
};

struct TablesView {
  // parameters
  double elTrackingCut, gammaTrackingCut, minLossTableEnergy, bremModelLim;
  int isMSCPositronCor, isMSCDisplacement;
  int numRegions, numMatCut, numMat;
  const double* regionPars;  // [8*numRegions]
  const double* mcCuts;      // [4*numMatCut]
  const int* mcImat;
  const int* mcIreg;
  const int* matNumElem;
  const int* matElemStart;
  const int* matElemZ;
  const double* matElemNatoms;
  const double* matPars;     // [16*numMat]
  const int* matSandiaNum;
  const int* matSandiaStart;
  const double* elemPars;    // [12*121]
  const int* elemSandiaNum;
  const int* elemSandiaStart;
  const double* sandiaEnergies;
  const double* sandiaCof;
  ElectronTablesView el[2];  // [0] e-, [1] e+
  // Seltzer-Berger
  double sbLogMinElEnergy, sbILDeltaElEnergy;
  const double* sbElEnergy;
  const double* sbLElEnergy;
  const double* sbLKappa;
  const int* sbGCutStart;
  const int* sbGCutIndices;
  const int* sbStartPerZ;
  const double* sbData;
  // gamma
  int gmDataPerMat, gmNumData0, gmNumData1;
  double gmEMax0, gmLogEMin0, gmEILDelta0, gmEMax1, gmLogEMin1, gmEILDelta1, gmLogEMin2, gmEILDelta2;
  const double* gmMXsec;
  int gmConvEGridSize;
  double gmConvLogMinEkin, gmConvEILDelta;
  const int* gmConvStart;
  const double* gmConvEGrid;
  const double* gmConvData;
};

// indices into the packed parameter rows
enum RegionPar { kRFinalRange = 0, kRDRoverRange, kRLinELossLimit, kRMSCRangeFactor, kRMSCSafetyFactor, kRIsMSCMinimal, kRIsFluct, kRCallerFlags };  // kRCallerFlags: fIsMultipleStepsInMSCTrans + 2 * fIsApplyCuts
enum MatPar { kMDensityCorFactor = 0, kMElectronDensity, kMRadLength, kMMeanExE, kMZeff, kMZeff23, kMZeffSqrt, kMUMSCPar,
              kMStepMin0, kMStepMin1, kMTail0, kMTail1, kMTail2, kMTail3, kMTheta0, kMTheta1 };
enum ElemPar { kEZet = 0, kEZet13, kEZet23, kECoulomb, kELogZ, kEZFactor1, kEDeltaMaxLow, kEDeltaMaxHigh, kEILVarS1,
               kEILVarS1Cond, kEKShell };
enum CutPar { kCElCut = 0, kCPosCut, kCGamCut, kCLogGamCut };

#if defined(__CUDA_ARCH__)
// plain loads: the compiler emits LDG for pointers it can trace to a kernel parameter and LDS / generic loads for the
// shared-memory copies of the tables (ElectronLookupsSmemKernel); __ldg would be invalid for the latter
#define G4H_LD(p) (*(p))
#else
#define G4H_LD(p) (*(p))
#endif

// ---- G4HepEmRunUtils.icc:49-61 ---------------------------------------------------------------
G4H_FN double Spline(double x1, double x2, double y1, double y2, double sd1, double sd2, double x) {
  const double dl = x2 - x1;
  const double b  = Max(0., Min(1., FastDiv(x - x1, dl)));  // dl: spacing of a table grid
  const double os = 0.166666666667;
  const double c0 = (2.0 - b) * sd1;
  const double c1 = (1.0 + b) * sd2;
  return y1 + b * (y2 - y1) + (b * (b - 1.0)) * (c0 + c1) * (dl * dl * os);
}

// G4HepEmRunUtils.icc:63-71
G4H_FN double Linear(double x1, double x2, double y1, double y2, double x) {
  const double dl = x2 - x1;
  const double b  = Max(0., Min(1., (x - x1) / dl));
  return y1 + b * (y2 - y1);
}

// lower bin index of a logarithmic grid, G4HepEmRunUtils.icc:79,89,100 etc.
G4H_FN int LogBin(double logx, double logxmin, double invLDBin, int ndata) {
  return static_cast<int>(Max(0., Min((logx - logxmin) * invLDBin, ndata - 2.)));
}

// GetSplineLog, y and second derivative interleaved, separate x grid (G4HepEmRunUtils.icc:86-93)
G4H_FN double SplineLogYSDInl(int ndata, const double* xdata, const double* ydata, double x, double logx, double logxmin,
                              double invLDBin) {
  const double xv = Max(G4H_LD(xdata), Min(G4H_LD(xdata + ndata - 1), x));
  const int idx   = LogBin(logx, logxmin, invLDBin, ndata);
  const int idx2  = 2 * idx;
  return Spline(G4H_LD(xdata + idx), G4H_LD(xdata + idx + 1), G4H_LD(ydata + idx2), G4H_LD(ydata + idx2 + 2),
                G4H_LD(ydata + idx2 + 1), G4H_LD(ydata + idx2 + 3), xv);
}

G4H_LEAF double SplineLogYSD(int ndata, const double* xdata, const double* ydata, double x, double logx, double logxmin,
                             double invLDBin) {
  return SplineLogYSDInl(ndata, xdata, ydata, x, logx, logxmin, invLDBin);
}

// GetSplineLog, x, y and second derivative interleaved (G4HepEmRunUtils.icc:97-104)
G4H_FN double SplineLogXYSDInl(int ndata, const double* data, double x, double logx, double logxmin, double invLDBin) {
  const double xv = Max(G4H_LD(data), Min(G4H_LD(data + 3 * (ndata - 1)), x));
  const int idx   = LogBin(logx, logxmin, invLDBin, ndata);
  const int idx3  = 3 * idx;
  return Spline(G4H_LD(data + idx3), G4H_LD(data + idx3 + 3), G4H_LD(data + idx3 + 1), G4H_LD(data + idx3 + 4),
                G4H_LD(data + idx3 + 2), G4H_LD(data + idx3 + 5), xv);
}

G4H_LEAF double SplineLogXYSD(int ndata, const double* data, double x, double logx, double logxmin, double invLDBin) {
  return SplineLogXYSDInl(ndata, data, x, logx, logxmin, invLDBin);
}

// ---- e-/e+ accessors, G4HepEmElectronManager.icc:486-599 -----------------------------------------
// GetRestRange (.icc:486-492)
G4H_FN double RestRange(const ElectronTablesView& ed, int imc, double ekin, double lekin) {
  const int n = ed.numLoss;
  const double r = SplineLogYSD(n, ed.lossEGrid, ed.lossData + 5 * n * imc, ekin, lekin, ed.lossLogMinEkin, ed.lossEILDelta);
  return Max(0.0, r);
}

// GetRestDEDX (.icc:495-501)
G4H_FN double RestDEDX(const ElectronTablesView& ed, int imc, double ekin, double lekin) {
  const int n = ed.numLoss;
  const double d = SplineLogYSD(n, ed.lossEGrid, ed.lossData + n * (5 * imc + 2), ekin, lekin, ed.lossLogMinEkin, ed.lossEILDelta);
  return Max(0.0, d);
}

// GetInvRange (.icc:504-519) with FindLowerBinIndex (G4HepEmRunUtils.icc:227-237) and the strided
// GetSpline (G4HepEmRunUtils.icc:108-110)
G4H_LEAF double InvRange(const ElectronTablesView& ed, int imc, double range) {
  const int n = ed.numLoss;
  const double* rdata = ed.lossData + 5 * n * imc;
  const double minRange = G4H_LD(rdata);
  if (range < minRange) {
    const double dum = range / minRange;
    return Max(0.0, G4H_LD(ed.lossEGrid) * dum * dum);
  }
  int ml = -1;
  int mu = n - 1;
  while (mu - ml > 1) {
    const int mav = static_cast<int>(0.5 * (ml + mu));
    if (range < G4H_LD(rdata + 2 * mav)) {
      mu = mav;
    } else {
      ml = mav;
    }
  }
  const int i = mu - 1;
  const double* sd = rdata + 4 * n;
  const double e = Spline(G4H_LD(rdata + 2 * i), G4H_LD(rdata + 2 * (i + 1)), G4H_LD(ed.lossEGrid + i),
                          G4H_LD(ed.lossEGrid + i + 1), G4H_LD(sd + i), G4H_LD(sd + i + 1), range);
  return Max(0.0, e);
}

// GetRestMacXSec (.icc:522-532)
G4H_FN double RestMacXSec(const ElectronTablesView& ed, int imc, double ekin, double lekin, bool isIoni) {
  const int iIoni   = G4H_LD(ed.resStart + imc);
  const int numIoni = static_cast<int>(G4H_LD(ed.resData + iIoni));
  const int iStart  = isIoni ? iIoni : iIoni + 3 * numIoni + 5;
  const double* d   = ed.resData + iStart;
  const int numData = static_cast<int>(G4H_LD(d));
  if (ekin < G4H_LD(d + 5)) return 0.0;
  const double mx = SplineLogXYSD(numData, d + 5, ekin, lekin, G4H_LD(d + 3), G4H_LD(d + 4));
  return Max(0.0, mx);
}

// GetRestMacXSecForStepping (.icc:544-568)
G4H_FN double RestMacXSecForStepping(const ElectronTablesView& ed, int imc, double ekin, double lekin, bool isIoni) {
  const double log08 = -0.22314355131420971;
  const int iIoni   = G4H_LD(ed.resStart + imc);
  const int numIoni = static_cast<int>(G4H_LD(ed.resData + iIoni));
  const int iStart  = isIoni ? iIoni : iIoni + 3 * numIoni + 5;
  const double* d   = ed.resData + iStart;
  const int numData = static_cast<int>(G4H_LD(d));
  const double mxsecMinE = G4H_LD(d + 5);
  const double mxsecMaxE = G4H_LD(d + 1);
  const double mxsecMaxV = G4H_LD(d + 2);
  if (ekin > mxsecMaxE) {
    const double ekinReduced = 0.8 * ekin;
    if (ekinReduced < mxsecMaxE) {
      return Max(0.0, mxsecMaxV);
    } else {
      ekin = ekinReduced;
      lekin += log08;
    }
  }
  if (ekin < mxsecMinE) return 0.0;
  const double mx = SplineLogXYSD(numData, d + 5, ekin, lekin, G4H_LD(d + 3), G4H_LD(d + 4));
  return Max(0.0, mx);
}

// GetMacXSecNuclear (.icc:534-541); ...ForStepping (.icc:570-573) is the same function
G4H_FN double MacXSecNuclear(const ElectronTablesView& ed, int imat, double ekin, double lekin) {
  if (ekin < G4H_LD(ed.enucEGrid)) return 0.0;
  const double mx = SplineLogYSD(128, ed.enucEGrid, ed.enucData + imat * 2 * 128, ekin, lekin, ed.enucLogMinEkin, ed.enucEILDelta);
  return Max(0.0, mx);
}

// GetTransportMFP (.icc:576-582)
G4H_FN double TransportMFP(const ElectronTablesView& ed, int imat, double ekin, double lekin) {
  const int n = ed.numLoss;
  const double tr1 = Max(0.0, SplineLogYSD(n, ed.lossEGrid, ed.tr1Data + 2 * n * imat, ekin, lekin, ed.lossLogMinEkin, ed.lossEILDelta));
  return tr1 > 0. ? 1. / tr1 : kALargeValue;
}

// ComputeMacXsecAnnihilation (.icc:585-593): Heitler e+e- -> 2 gamma
G4H_LEAF double MacXSecAnnihilation(double ekin, double electronDensity) {
  const double tau  = ekin * kInvElectronMassC2;
  const double gam  = tau + 1.0;
  const double gam2 = gam * gam;
  const double bg2  = tau * (tau + 2.0);
  const double bg   = sqrt(bg2);
  return electronDensity * kPir02 * ((gam2 + 4. * gam + 1.) * Log(gam + bg) - (gam + 3.) * bg) / (bg2 * (gam + 1.));
}

// Brem::SelectTargetAtom (G4HepEmElectronInteractionBrem.icc:266-296)
G4H_FN int SelectTargetAtomBrem(const ElectronTablesView& ed, int imc, double ekin, double lekin, double urndn, bool isSB) {
  const int indxStart   = isSB ? G4H_LD(ed.selSBStart + imc) : G4H_LD(ed.selRBStart + imc);
  const double* theData = (isSB ? ed.selSBData : ed.selRBData) + indxStart;
  const int numData  = static_cast<int>(G4H_LD(theData));
  const int numElem  = static_cast<int>(G4H_LD(theData + 1));
  const double logE0 = G4H_LD(theData + 2);
  const double invLD = G4H_LD(theData + 3);
  const double* xdata = theData + 4;
  const double xv   = Max(G4H_LD(xdata), Min(G4H_LD(xdata + numElem * (numData - 1)), ekin));
  const int idxEkin = static_cast<int>(Max(0.0, Min((lekin - logE0) * invLD, numData - 2.0)));
  int indx0 = idxEkin * numElem;
  int indx1 = indx0 + numElem;
  const double x1 = G4H_LD(xdata + indx0++);
  const double x2 = G4H_LD(xdata + indx1++);
  const double dl = x2 - x1;
  const double b  = Max(0., Min(1., (xv - x1) / dl));
  int theElemIndex = 0;
  while (theElemIndex < numElem - 1 &&
         urndn > G4H_LD(xdata + indx0 + theElemIndex) + b * (G4H_LD(xdata + indx1 + theElemIndex) - G4H_LD(xdata + indx0 + theElemIndex))) {
    ++theElemIndex;
  }
  return theElemIndex;
}

// Conversion::SelectTargetAtom (G4HepEmGammaInteractionConversion.icc:151-176)
G4H_FN int SelectTargetAtomConversion(const TablesView& tv, int imat, double ekin, double lekin, double urndn) {
  const int indxStart   = G4H_LD(tv.gmConvStart + imat);
  const double* theData = tv.gmConvData + indxStart;
  const int numData  = tv.gmConvEGridSize;
  const int numElem  = static_cast<int>(G4H_LD(theData));
  const double* xdata = tv.gmConvEGrid;
  const double xv   = Max(G4H_LD(xdata), Min(G4H_LD(xdata + numData - 1), ekin));
  const int idxEkin = static_cast<int>(Max(0.0, Min((lekin - tv.gmConvLogMinEkin) * tv.gmConvEILDelta, numData - 2.0)));
  const double x1 = G4H_LD(xdata + idxEkin);
  const double x2 = G4H_LD(xdata + idxEkin + 1);
  const double dl = x2 - x1;
  const double b  = Max(0., Min(1., (xv - x1) / dl));
  const int indx0 = idxEkin * (numElem - 1) + 1;
  const int indx1 = indx0 + (numElem - 1);
  int theElemIndex = 0;
  while (theElemIndex < numElem - 1 &&
         urndn > G4H_LD(theData + indx0 + theElemIndex) + b * (G4H_LD(theData + indx1 + theElemIndex) - G4H_LD(theData + indx0 + theElemIndex))) {
    ++theElemIndex;
  }
  return theElemIndex;
}

// ---- gamma accessors, G4HepEmGammaManager.icc:108-170 -------------------------------------------
// Sandia interval search + polynomial (GetMacXSecPE .icc:155-170; the same form per atom in
// G4HepEmGammaInteractionPhotoelectric.icc:52-66)
// returns the bracketed polynomial c0 + inv*(c1 + inv*(c2 + inv*c3)) of the interval containing ekin; the
// callers multiply by inv (material) or by natoms*inv (per atom), in the reference's association order
G4H_FN double SandiaPoly(const double* energies, const double* cofs, int numIntervals, double ekin, double inv) {
  int interval = 0;
  if (ekin >= G4H_LD(energies)) {
    for (int i = numIntervals - 1; i >= 0; i--) {
      if (ekin >= G4H_LD(energies + i)) {
        interval = i;
        break;
      }
    }
  }
  const double* c = cofs + 4 * interval;
  return G4H_LD(c) + inv * (G4H_LD(c + 1) + inv * (G4H_LD(c + 2) + inv * G4H_LD(c + 3)));
}

G4H_FN double MacXSecPE(const TablesView& tv, int imat, double ekin) {
  const int s = G4H_LD(tv.matSandiaStart + imat);
  const double inv = 1 / ekin;
  return inv * SandiaPoly(tv.sandiaEnergies + s, tv.sandiaCof + 4 * s, G4H_LD(tv.matSandiaNum + imat), ekin, inv);
}

// GetSplineLog4 with one selected column (G4HepEmRunUtils.icc:178-187); iwhich = 1..4 (5 reads one slot
// past the PE column, exactly as the reference does when gamma-nuclear is probed)
G4H_LEAF double SplineLog4(const double* data, double x, double logx, double logxmin, double invLDBin, int iwhich) {
  const int idx    = LogBin(logx, logxmin, invLDBin, 256);
  const int idx9_0 = 9 * idx;
  const int idx9_1 = idx9_0 + 9;
  iwhich = (iwhich - 1) * 2 + 1;
  return Spline(G4H_LD(data + idx9_0), G4H_LD(data + idx9_1), G4H_LD(data + idx9_0 + iwhich), G4H_LD(data + idx9_1 + iwhich),
                G4H_LD(data + idx9_0 + iwhich + 1), G4H_LD(data + idx9_1 + iwhich + 1), x);
}

// The conversion, Compton and photoelectric columns (iwhich = 2, 3, 4) of one row pair in one go: SampleInteraction asks for
// them one after the other (.icc:196-208) and every call repeats the bin index, the abscissas and the interpolation weight.
// Same operations per column as Spline() above, hence the same bits.
G4H_FN void SplineLog4Three(const double* data, double x, double logx, double logxmin, double invLDBin, double* out) {
  const int idx    = LogBin(logx, logxmin, invLDBin, 256);
  const double* r0 = data + 9 * idx;
  const double* r1 = r0 + 9;
  const double x1 = G4H_LD(r0), x2 = G4H_LD(r1);
  const double dl = x2 - x1;
  const double b  = Max(0., Min(1., FastDiv(x - x1, dl)));
  const double os = 0.166666666667;
  const double bb = b * (b - 1.0);
  const double d2 = dl * dl * os;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double y1 = G4H_LD(r0 + 3 + 2 * k), sd1 = G4H_LD(r0 + 4 + 2 * k);
    const double y2 = G4H_LD(r1 + 3 + 2 * k), sd2 = G4H_LD(r1 + 4 + 2 * k);
    const double c0 = (2.0 - b) * sd1;
    const double c1 = (1.0 + b) * sd2;
    out[k] = y1 + b * (y2 - y1) + bb * (c0 + c1) * d2;
  }
}

// GetTotalMacXSec (.icc:108-152); peMXsec is G4HepEmGammaTrack::fPEmxSec (only written in windows 0/1)
G4H_FN double GammaTotalMacXSec(const TablesView& tv, int imat, double ekin, double lekin, double& peMXsec) {
  const double* matData = tv.gmMXsec + imat * tv.gmDataPerMat;
  if (ekin > tv.gmEMax1) {
    return SplineLog4(matData + tv.gmNumData0 + tv.gmNumData1, ekin, lekin, tv.gmLogEMin2, tv.gmEILDelta2, 1);
  }
  if (ekin > tv.gmEMax0) {
    // GetLinearLog2, both columns (G4HepEmRunUtils.icc:138-152)
    const double* data = matData + tv.gmNumData0;
    const int idx      = LogBin(lekin, tv.gmLogEMin1, tv.gmEILDelta1, 32);
    const int idx3_0   = 3 * idx;
    const int idx3_1   = idx3_0 + 3;
    const double x0    = G4H_LD(data + idx3_0);
    const double dl    = G4H_LD(data + idx3_1) - x0;
    const double b     = Max(0., Min(1., (ekin - x0) / dl));
    double res[2];
    for (int i = 1; i < 3; ++i) {
      const double y1 = G4H_LD(data + idx3_0 + i);
      const double y2 = G4H_LD(data + idx3_1 + i);
      res[i - 1] = Max(0.0, y1 + b * (y2 - y1));
    }
    peMXsec = res[1];
    return res[0];
  }
  // window 0: GetLinearLog (G4HepEmRunUtils.icc:116-123) + Sandia PE
  const int idx    = LogBin(lekin, tv.gmLogEMin0, tv.gmEILDelta0, 32);
  const int idx2_0 = 2 * idx;
  const int idx2_1 = idx2_0 + 2;
  const double comp = Linear(G4H_LD(matData + idx2_0), G4H_LD(matData + idx2_1), G4H_LD(matData + idx2_0 + 1),
                             G4H_LD(matData + idx2_1 + 1), ekin);
  const double pe = Max(0.0, MacXSecPE(tv, imat, ekin));
  peMXsec = pe;
  return comp + pe;
}

// SampleInteraction (.icc:181-219): returns the process index, updates peMXsec like the reference
G4H_FN int GammaSampleInteraction(const TablesView& tv, int imat, double ekin, double lekin, double totMFP, double urnd,
                                  double& peMXsec) {
  if (ekin > tv.gmEMax1) {
    const double* data = tv.gmMXsec + imat * tv.gmDataPerMat + tv.gmNumData0 + tv.gmNumData1;
    // The reference's do-while also probes a 4th column (pid = 5, gamma-nuclear) that does not exist: it
    // reads one slot past the PE column (out of bounds for the last bin of the last material) and the
    // value cannot change the outcome (winner = 3).  Only the three real columns are evaluated here;
    // fPEmxSec is "garbage otherwise" in the reference (.icc:182) and is set to 0 in that case.
    double mxSec[3];
    SplineLog4Three(data, ekin, lekin, tv.gmLogEMin2, tv.gmEILDelta2, mxSec);
    double cProb = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      cProb += mxSec[k] * totMFP;
      if (!(urnd > cProb)) {
        peMXsec = mxSec[k];
        return k;
      }
    }
    peMXsec = 0.0;
    return 3;
  }
  return (urnd > totMFP * peMXsec) ? 1 : 2;
}

}  // namespace g4h
#endif
