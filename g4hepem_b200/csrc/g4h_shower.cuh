// g4h_shower.cuh -- the stepping loop around HowFar / Perform for a TestEm3-style slab calorimeter, on the device.
//
// What the reference's callers do per track and step around the managers (G4HepEmTrackingManager::TrackElectron /
// TrackGamma, G4HepEm/G4HepEm/src/G4HepEmTrackingManager.cc:428-705,985-1140; apps/examples/TestEm3: geometry
// DetectorConstruction.cc:281-384, scoring SteppingAction.cc:83-84), restated for whole batches:
//
//   HowFar kernels          physics step limit                                           (existing pipeline)
//   ShowerGeomKernel        distance to the slab / side faces along the direction; the step becomes
//                           min(physics, geometry), the post-step boundary flag is set, the track is moved
//   Perform kernels         along-step + discrete physics, secondaries into the queue    (existing pipeline)
//   ShowerPostKernel        MSC displacement (limited by the safety like TrackingManager.cc:527-568), energy
//                           deposit into the per-(layer, absorber) histogram, relocation into the next volume or
//                           escape, safety for the next step, survivors compacted into the next-step store
//   ShowerSecondaryKernel   secondaries -> new tracks at the parent's post-step point, appended to the next-step stores
//
// Geometry: num_layers x num_absorbers slabs stacked along x, centred on the origin, half width half_yz in y and
// z; a track that leaves the calorimeter is dropped and its kinetic energy booked as leakage (the reference's world
// is vacuum around the calorimeter; nothing comes back from it).
// Track ids: a secondary's (id, first draw index) is a Philox hash of (parent id, parent draw counter after the
// step, slot), so the streams do not depend on the order in which tracks are created or on how they are batched --
// the loop gives the same shower breadth-first on the GPU as the depth-first CPU loop of tests/shower_oracle.py.
#ifndef G4H_SHOWER_CUH
#define G4H_SHOWER_CUH

#include "g4h_kernels.cuh"
#include "g4h_pipeline.cuh"
#include "g4h_fused.cuh"

namespace g4h {

constexpr int kMaxAbsorbers = 4;

struct SlabGeom {
  int numLayers, numAbsorbers;
  double thickness[kMaxAbsorbers];
  double absFront[kMaxAbsorbers + 1];  // x offset of the absorber front faces inside a layer; [numAbsorbers] = layer thickness
  int couple[kMaxAbsorbers];
  double halfYZ, xFront;               // xFront = -0.5 * numLayers * layer thickness
  int inheritCouple;                   // 1: no geometry (unbounded media): a secondary keeps its parent's couple
  int wdtOn, wdtCouple;                // Woodcock tracking of gammas: on, the couple whose cross section is used
  double wdtEkinMin;
};

// position / volume of the tracks of one store, next to its batch
struct TrackGeo {
  double* posx_posy;  // pairs
  double* posz_pad;   // pairs {z, unused}
  int32_t* vol;       // layer * numAbsorbers + absorber; -1: outside
  int32_t* nextVol;   // volume behind the face the step ended on (only meaningful while on a boundary)
  // e-/e+ stores: what TrackElectron keeps in local variables across the MSC sub-steps of one step (G4HepEmTrackingManager.cc:
  // 438-445, 574-596), meaningful while G4HB200_F_MSC_SUBSTEP is set; NULL in the gamma stores
  double* sub_left_eloss;  // pairs {stepLimitLeft, totalEloss}
  double* sub_pre;         // pairs {preStepEkin, preStepLogEkin} of the whole step
  double* sub_range_proc;  // pairs {range left after the sub-steps so far, iDProc (the discrete winner MSC replaced)}
  // number of live tracks of the store where the host does not know it (the graph-driven tail of the loop, capi_shower.inl);
  // NULL: the batch's n
  const int32_t* n_dev;
};

// device-side bookkeeping of the graph-driven tail of the loop: what the host keeps in local variables otherwise
struct ShowerCtrl {
  int32_t cur[2];  // populations of the running iteration {e-/e+, gamma}
  int32_t pad[2];
  long long steps, sumEl, sumGm, sumSec, peakEl, peakGm;
};

struct ShowerScore {
  double* hist;        // [numLayers * numAbsorbers] deposited energy
  double* leak;        // [2] kinetic energy that left the calorimeter {e-/e+, gamma}
  int32_t* nextCount;  // [2] number of tracks in the next-step stores {e-/e+, gamma}
  int32_t* overflow;   // [1] set when a store ran out of capacity (tracks were dropped: the run is invalid)
  int64_t capacity;    // of every store
};

constexpr double kGeoInfinity = 1.0e+30;

G4H_FN void SlabBounds(const SlabGeom& g, int vol, double& xlo, double& xhi) {
  const int layer = vol / g.numAbsorbers;
  const int iabs  = vol - layer * g.numAbsorbers;
  const double layerFront = g.xFront + layer * g.absFront[g.numAbsorbers];
  xlo = layerFront + g.absFront[iabs];
  xhi = layerFront + g.absFront[iabs + 1];
}

G4H_FN double DistanceAlong(double p, double d, double lo, double hi) {
  if (d > 0.) return Max(0.0, (hi - p) / d);
  if (d < 0.) return Max(0.0, (lo - p) / d);
  return kGeoInfinity;
}

// distance to the surface of the current slab along dir; nextVol = the volume behind that face (-1: outside)
G4H_FN double DistanceToBoundary(const SlabGeom& g, int vol, const double* pos, const double* dir, int& nextVol) {
  double xlo, xhi;
  SlabBounds(g, vol, xlo, xhi);
  const double dx = DistanceAlong(pos[0], dir[0], xlo, xhi);
  const double dy = DistanceAlong(pos[1], dir[1], -g.halfYZ, g.halfYZ);
  const double dz = DistanceAlong(pos[2], dir[2], -g.halfYZ, g.halfYZ);
  const int numVol = g.numLayers * g.numAbsorbers;
  int nv = dir[0] > 0. ? vol + 1 : vol - 1;
  if (nv >= numVol) nv = -1;
  double d = dx;
  if (dy < d) { d = dy; nv = -1; }
  if (dz < d) { d = dz; nv = -1; }
  nextVol = nv;
  return d;
}

// isotropic safety: distance to the nearest face of the current slab
G4H_FN double SlabSafety(const SlabGeom& g, int vol, const double* pos) {
  double xlo, xhi;
  SlabBounds(g, vol, xlo, xhi);
  double s = Min(pos[0] - xlo, xhi - pos[0]);
  s = Min(s, g.halfYZ - fabs(pos[1]));
  s = Min(s, g.halfYZ - fabs(pos[2]));
  return Max(0.0, s);
}

// the geometry step of one e-/e+ track between HowFar and Perform, on registers (for the fused head of the loop):
// distance to the faces of the current slab along the direction, step = min(physics, geometry), post-step boundary
// flag, move.  The same arithmetic as ShowerGeomKernel.
struct SlabGeometryStep {
  const SlabGeom& g;  // the kernel's __grid_constant__ parameter (indexed dynamically: a copy would live in local memory)
  const TrackGeo& geo;
  const double* dirx_diry;
  G4H_MFN void operator()(int64_t i, ElectronState& s, double dirz) const {
    const Pair dxy = LoadPair(dirx_diry, i);
    const Pair pxy = LoadPair(geo.posx_posy, i);
    const Pair pz  = LoadPair(geo.posz_pad, i);
    const int vol  = geo.vol[i];
    const double dir[3] = {dxy.a, dxy.b, dirz};
    double pos[3] = {pxy.a, pxy.b, pz.a};
    int nextVol;
    const double dist = DistanceToBoundary(g, vol, pos, dir, nextVol);
    double step = s.gStep;
    const bool onBoundary = dist < step;
    if (onBoundary) step = dist;
    pos[0] += step * dir[0];
    pos[1] += step * dir[1];
    pos[2] += step * dir[2];
    StorePair(geo.posx_posy, i, pos[0], pos[1]);
    StorePair(geo.posz_pad, i, pos[2], 0.0);
    geo.nextVol[i] = nextVol;
    s.gStep      = step;
    s.onBoundary = onBoundary;
  }
};

// distance from pos along dir to the surface of the whole calorimeter (the Woodcock tracking volume)
G4H_FN double DistanceToCalorimeterOut(const SlabGeom& g, const double* pos, const double* dir) {
  const double dx = DistanceAlong(pos[0], dir[0], g.xFront, -g.xFront);
  const double dy = DistanceAlong(pos[1], dir[1], -g.halfYZ, g.halfYZ);
  const double dz = DistanceAlong(pos[2], dir[2], -g.halfYZ, g.halfYZ);
  return Min(dx, Min(dy, dz));
}

#ifndef G4H_WDT_MAX_VIRTUAL_STEPS
#define G4H_WDT_MAX_VIRTUAL_STEPS 4
#endif
constexpr int kWdtMaxVirtualSteps = G4H_WDT_MAX_VIRTUAL_STEPS;

// the slab that holds the point x (inside the calorimeter): layer * numAbsorbers + absorber
G4H_FN int LocateSlab(const SlabGeom& g, double x) {
  const double layerT = g.absFront[g.numAbsorbers];
  const double t = x - g.xFront;
  int layer = static_cast<int>(t / layerT);
  layer = layer < 0 ? 0 : (layer > g.numLayers - 1 ? g.numLayers - 1 : layer);
  const double u = t - layer * layerT;
  int iabs = g.numAbsorbers - 1;
  for (int k = g.numAbsorbers - 1; k >= 1; --k) {
    if (u < g.absFront[k]) iabs = k - 1;
  }
  return layer * g.numAbsorbers + iabs;
}

// The step limit of a gamma track of the loop and the geometry step behind it, on registers:
//  - normally G4HepEmGammaManager::HowFar, then the distance to the faces of the current slab (ShowerGeomKernel);
//  - under Woodcock tracking (g.wdtOn, energy above the limit, not within 1e-3 mm of the calorimeter surface)
//    G4HepEmWoodcockHelper::KeepTracking (G4HepEmWoodcockHelper.cc:150-300) and the control flow around it in
//    G4HepEmTrackingManager::TrackGamma (G4HepEmTrackingManager.cc:955-1053): exponential steps with the cross
//    section of the Woodcock couple straight through the slab boundaries until a point is reached where the gamma
//    really interacts -- for sure in the Woodcock material, with probability sigma(local) / sigma(Woodcock) elsewhere --
//    or the surface of the calorimeter is 1e-3 mm away; then a zero step (interaction at that point) or a
//    10 mm step that ends on the surface.  The number of interaction lengths left is reset in both cases.
//    (A 10 mm step that does not end on a boundary -- `isWDTReachedBoundary && !geometryLimitedStep` -- cannot
//    happen here: the surface of the calorimeter is a face of the slab the point is in, 1e-3 mm ahead.)
struct SlabGammaGeometryStep {
  const SlabGeom& g;
  const TrackGeo& geo;
  G4H_MFN bool GammaHowFarAndStep(const TablesView& tv, int64_t i, GammaState& s, Rng& rng, int& flags) const {
    const Pair pxy = LoadPair(geo.posx_posy, i);
    const Pair pz  = LoadPair(geo.posz_pad, i);
    int vol        = geo.vol[i];
    double pos[3]  = {pxy.a, pxy.b, pz.a};
    double physicalStep;
    bool wdt = false;
    if (g.wdtOn != 0) {
      // FindWDTVolume (.cc:105-147) unless already under Woodcock tracking, and the energy limit (TrackingManager.cc:966-972)
      wdt = (static_cast<uint32_t>(flags) & G4HB200_F_WDT_ON) != 0u;
      const double distOut = Max(DistanceToCalorimeterOut(g, pos, s.dir) - 1.0E-3, 0.0);
      if (!wdt) wdt = !(s.ekin < g.wdtEkinMin) && !(distOut < 1.0E-6);
      wdt = wdt && s.ekin > g.wdtEkinMin;
      if (wdt) {
        // KeepTracking
        double distToBoundary = distOut;
        const double lekin  = GetLogEKin(s);
        const int wdtImat   = G4H_LD(tv.mcImat + g.wdtCouple);
        double wdtPE = s.peMXsec;
        const double wdtMXsec = GammaTotalMacXSec(tv, wdtImat, s.ekin, lekin, wdtPE);
        const double kDblMax  = 1.7976931348623157e+308;
        const double wdtMFP   = wdtMXsec > 0.0 ? 1.0 / wdtMXsec : kDblMax;
        s.peMXsec = wdtPE;
        double mxsec = 0.0;
        int prevIMC  = -1;
        bool doStop  = false;
        bool reached = false;
        double wdtStepLength = 0.0;
        // At most kWdtMaxVirtualSteps virtual steps per pass: a 0.2-1 MeV gamma crossing liquid argon with lead's cross
        // section takes dozens, and a warp waits for its longest lane (171 ms per 4096 showers without the cap, 128 ms
        // without Woodcock tracking at all).  A pass that is cut ends at a fictitious interaction point: the track is
        // moved there, nothing else happens, and the next pass carries on with the next uniform of its stream --
        // the exponential has no memory, so this is the uncut loop up to the rounding of the distance to the surface
        // (recomputed from the new position instead of decremented).
        int virtualSteps = 0;
        while (!doStop && virtualSteps < kWdtMaxVirtualSteps) {
          ++virtualSteps;
          const double pstep = wdtMFP < kDblMax ? -Log(rng.Flat()) * wdtMFP : kDblMax;
          if (distToBoundary < pstep) {
            wdtStepLength += distToBoundary;
            reached = true;
            doStop  = true;
          } else {
            wdtStepLength  += pstep;
            distToBoundary -= pstep;
            const int pvol = LocateSlab(g, pos[0] + wdtStepLength * s.dir[0]);
            const int pimc = g.couple[pvol % g.numAbsorbers];
            if (G4H_LD(tv.mcImat + pimc) != wdtImat) {
              if (pimc != prevIMC) {
                prevIMC = pimc;
                mxsec   = GammaTotalMacXSec(tv, G4H_LD(tv.mcImat + pimc), s.ekin, lekin, s.peMXsec);
              }
              doStop = mxsec * wdtMFP > rng.Flat();
              if (doStop) s.mfp0 = mxsec > 0.0 ? 1.0 / mxsec : kDblMax;
            } else {
              doStop    = true;
              s.mfp0    = wdtMFP;
              s.peMXsec = wdtPE;
            }
          }
        }
        pos[0] += wdtStepLength * s.dir[0];
        pos[1] += wdtStepLength * s.dir[1];
        pos[2] += wdtStepLength * s.dir[2];
        vol     = LocateSlab(g, pos[0]);
        s.imc   = g.couple[vol % g.numAbsorbers];
        s.nIA0  = -1.0;
        physicalStep = reached ? 10.0 : 0.0;
        flags = reached ? (flags & ~static_cast<int>(G4HB200_F_WDT_ON)) : (flags | static_cast<int>(G4HB200_F_WDT_ON));
        if (!doStop) {
          // the pass was cut: the track waits at the fictitious interaction point for the next pass
          StorePair(geo.posx_posy, i, pos[0], pos[1]);
          StorePair(geo.posz_pad, i, pos[2], 0.0);
          geo.vol[i]     = vol;
          geo.nextVol[i] = vol;
          s.gStep      = 0.0;
          s.onBoundary = false;
          return false;
        }
      }
    }
    if (!wdt) {
      flags &= ~static_cast<int>(G4HB200_F_WDT_ON);
      GammaHowFar(tv, s, rng);
      physicalStep = s.gStep;
    }
    int nextVol;
    const double dist = DistanceToBoundary(g, vol, pos, s.dir, nextVol);
    const bool onBoundary = dist < physicalStep;
    const double step = onBoundary ? dist : physicalStep;
    pos[0] += step * s.dir[0];
    pos[1] += step * s.dir[1];
    pos[2] += step * s.dir[2];
    StorePair(geo.posx_posy, i, pos[0], pos[1]);
    StorePair(geo.posz_pad, i, pos[2], 0.0);
    geo.vol[i]     = vol;
    geo.nextVol[i] = nextVol;
    // the track keeps the normal step length only: zero after Woodcock tracking (TrackingManager.cc:1038-1046)
    s.gStep      = wdt ? 0.0 : step;
    s.onBoundary = onBoundary;
    return true;
  }
};

#if defined(__CUDACC__)
// ---- the head of a gamma step of the loop: HowFar + geometry step + SelectInteraction + head of Perform -------------
__global__ void __launch_bounds__(kGammaThreads, 2 * G4H_MINB_QUEUE)
ShowerGammaHeadKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200GammaBatch b,
                      const __grid_constant__ ElectronWork w, uint64_t seed, const __grid_constant__ SlabGeom g,
                      const __grid_constant__ TrackGeo geo) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n      = geo.n_dev != nullptr ? *geo.n_dev : b.n;
  const int64_t nRound = RoundUpToCta(n);
  __shared__ CtaCounters<3> cc;
  cc.Init();
  const SlabGammaGeometryStep geometry{g, geo};
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    const int route = i < n ? StageGammaHead<2>(tv, b, i, seed, geometry) : -1;
    RouteToQueues<3>(cc, route, static_cast<int32_t>(i), w.queue, w.count);
  }
}

// ---- the head of an e-/e+ step of the loop: HowFar + geometry step + along-step part of Perform in one pass ---------
// (g4h_stages.cuh: StageStepHead; replaces ElHowFarXSKernel + ElHowFarMSCKernel + ShowerGeomKernel + ElAlongStepKernel)
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_HEAD)
ShowerElectronHeadKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                         const __grid_constant__ ElectronWork w, uint64_t seed, const __grid_constant__ SlabGeom g,
                         const __grid_constant__ TrackGeo geo) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n      = geo.n_dev != nullptr ? *geo.n_dev : b.n;
  const int64_t nRound = RoundUpToCta(n);
  __shared__ CtaCounters<5> cc;
  cc.Init();
  const SlabGeometryStep geometry{g, geo, b.dirx_diry};
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    const int route = i < n ? StageLoopHead(tv, b, w.prestep, w.steppre, i, seed, geometry) : -1;
    RouteToQueues<5>(cc, route, static_cast<int32_t>(i), w.queue, w.count);
  }
}

// ---- the head of an e-/e+ step of the loop ------------------------------------------------------------------------------
// StageStepHead (g4h_stages.cuh) + the geometry step + what G4HepEmTrackingManager::TrackElectron does around the pieces
// (G4HepEmTrackingManager.cc:428-613): with fIsMultipleStepsInMSCTrans of the region (the reference's default) a step whose
// length MSC limited goes on -- `continueStepping` -- with another HowFarToMSC / geometry / along-step / SampleMSC round over
// what is left of the step limit of the discrete interaction (stepLimitLeft), as long as MSC keeps limiting it, geometry
// does not and the track has energy; the mean losses of the rounds add up (totalEloss) and SampleLossFluctuations, the
// discrete interaction and the scoring see the whole step.  One round = one iteration of the device loop: a track that
// continues carries G4HB200_F_MSC_SUBSTEP and the caller's local variables (TrackGeo::sub_*) into the next iteration, where
// this head resumes it: no new interaction lengths, no HowFarToDiscreteInteraction, the mean free paths of the first round.
// prestep: energy at the beginning of this round (what SampleMSC reads); steppre: at the beginning of the step (what
// SampleLossFluctuations reads; StageMSCSample hands it on in `prestep`).
// A step of zero geometrical length follows G4HepEmElectronManager::Perform (.icc:461-470: nothing happens).
template <class GeometryStep>
G4H_FN int StageLoopHead(const TablesView& tv, const G4HB200ElectronBatch& b, double* prestep, double* steppre, int64_t i,
                         uint64_t seed, const GeometryStep& geometry) {
  const TrackGeo& geo = geometry.geo;
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
  const bool resume = (f & G4HB200_F_MSC_SUBSTEP) != 0u;
  int nXS = 0, iDProc = -1;
  double lam, stepLimitLeft, totalEloss, stepPreEkin, stepPreLogEkin;
  if (!resume) {
    // .cc:430-445: interaction lengths, the discrete step limit, the winner MSC may replace, SavePreStepEKin
    nXS = ResampleNumIALeftWindow(s.nIA, dw);
    lam = HowFarToDiscreteInteractionILP(tv, s);
    iDProc         = s.winner;
    stepLimitLeft  = s.pStep;
    totalEloss     = 0.0;
    stepPreEkin    = s.ekin;
    stepPreLogEkin = s.logEkin;
  } else {
    // .cc:574-596: what the previous round left
    const Pair m01 = LoadPair(b.mfp01, i);
    const Pair m23 = LoadPair(b.mfp23, i);
    const Pair sl  = LoadPair(geo.sub_left_eloss, i);
    const Pair sp  = LoadPair(geo.sub_pre, i);
    const Pair sr  = LoadPair(geo.sub_range_proc, i);
    s.mfp[0] = m01.a; s.mfp[1] = m01.b; s.mfp[2] = m23.a; s.mfp[3] = m23.b;
    stepLimitLeft  = sl.a;
    totalEloss     = sl.b;
    stepPreEkin    = sp.a;
    stepPreLogEkin = sp.b;
    s.range  = sr.a;
    iDProc   = static_cast<int>(sr.b);
    s.pStep  = stepLimitLeft;
    s.gStep  = stepLimitLeft;
    s.winner = iDProc;
    const double le = GetLogEKin(s);
    lam = TransportMFP(tv.el[s.isPositron ? 1 : 0], G4H_LD(tv.mcImat + s.imc), s.ekin, le);
  }
  s.lambtr1 = MSCStepLimitApplies(s.pStep, s.ekin) ? lam : 0.0;
  const double uA = nXS == 0 ? dw.u[0] : nXS == 1 ? dw.u[1] : nXS == 2 ? dw.u[2] : nXS == 3 ? dw.u[3] : dw.u[4];
  const double uB = nXS == 0 ? dw.u[1] : nXS == 1 ? dw.u[2] : nXS == 2 ? dw.u[3] : nXS == 3 ? dw.u[4] : dw.Sixth();
  const int nMSC = HowFarMSCCore(tv, s, hasGauss, gauss, uA, uB, dw.k0, dw.k1, static_cast<uint32_t>(m.draw + nXS));
  const int callerFlags = static_cast<int>(G4H_LD(tv.regionPars + 8 * G4H_LD(tv.mcIreg + s.imc) + kRCallerFlags));
  bool continueStepping = (callerFlags & 1) != 0 && s.winner == -2;  // .cc:427,450-453
  geometry(i, s, dzs.a);
  if (s.onBoundary) continueStepping = false;  // geometryLimitedStep (.cc:468-469)
  const int how = AlongStepPhysics(tv, s);
  if (how == kASStopped || how == kASMsc || how == kASNoMsc) totalEloss += s.edep;  // .cc:514
  if (how == kASStopped || how == kASNoStep || how == kASZeroStep) continueStepping = false;
  int route;
  if (continueStepping) {
    // .cc:574-596: the deposit is accumulated, the winner MSC replaced comes back, the rest of the step limit and of the range
    stepLimitLeft -= s.pStep;
    StorePair(geo.sub_left_eloss, i, stepLimitLeft, totalEloss);
    StorePair(geo.sub_pre, i, stepPreEkin, stepPreLogEkin);
    StorePair(geo.sub_range_proc, i, s.range - s.pStep, static_cast<double>(iDProc));
    f |= G4HB200_F_MSC_SUBSTEP;
    s.edep   = 0.0;
    s.winner = iDProc;
    route    = how == kASMsc ? (s.isPositron ? kQMscPos : kQMscEl) : -1;
    StorePair(prestep, i, s.preStepEkin, s.preStepLogEkin);
  } else {
    f &= ~G4HB200_F_MSC_SUBSTEP;
    if (how != kASNoStep) s.edep = totalEloss;  // .cc:600
    route = AlongStepRoute(tv, s, how, stepPreEkin);
    // the MSC stage reads the energy at the beginning of this round and leaves that of the step behind it
    const bool toMsc = route == kQMscEl || route == kQMscPos;
    StorePair(prestep, i, toMsc ? s.preStepEkin : stepPreEkin, toMsc ? s.preStepLogEkin : stepPreLogEkin);
  }
  StorePair(steppre, i, stepPreEkin, stepPreLogEkin);
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
  return route == -2 ? -1 : route;
}

// ---- the loop's steps as single persistent launches (g4h_fused.cuh) with the geometry step inside the head -------------
struct MakeSlabGeometry {
  const SlabGeom& g;
  const TrackGeo& geo;
  const double* dirx_diry;
  __device__ __forceinline__ SlabGeometryStep operator()() const { return SlabGeometryStep{g, geo, dirx_diry}; }
};
struct MakeSlabGammaGeometry {
  const SlabGeom& g;
  const TrackGeo& geo;
  __device__ __forceinline__ SlabGammaGeometryStep operator()() const { return SlabGammaGeometryStep{g, geo}; }
};

__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_FUSED)
ShowerElectronFusedKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b, double* prestep,
                          double* steppre, const __grid_constant__ G4HB200SecondaryQueue sq, uint64_t seed,
                          const __grid_constant__ SlabGeom g, const __grid_constant__ TrackGeo geo) {
  ElFusedBody<kHeadLoop>(tv, b, prestep, steppre, sq, seed, MakeSlabGeometry{g, geo, b.dirx_diry});
}

__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_FUSED)
ShowerGammaFusedKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200GammaBatch b,
                       const __grid_constant__ G4HB200SecondaryQueue sq, uint64_t seed, const __grid_constant__ SlabGeom g,
                       const __grid_constant__ TrackGeo geo) {
  GammaFusedBody<2>(tv, b, sq, seed, MakeSlabGammaGeometry{g, geo});
}

// ---- geometry step: between HowFar and Perform ------------------------------------------------------------------
// kGamma: the batch is a gamma batch (groups gstep_mfp0 / dirz_nia0), else an e-/e+ batch (gstep_pstep / dirz_safety)
template <bool kGamma>
__global__ void __launch_bounds__(kThreadsPerBlock)
ShowerGeomKernel(const __grid_constant__ SlabGeom g, int64_t n, const double* __restrict__ dirx_diry,
                 const double* __restrict__ dirz_x, double* __restrict__ gstep_x, int32_t* __restrict__ meta,
                 const __grid_constant__ TrackGeo geo) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Pair dxy = LoadPair(dirx_diry, i);
    const Pair dz  = LoadPair(dirz_x, i);
    const Pair gs  = LoadPair(gstep_x, i);
    const Pair pxy = LoadPair(geo.posx_posy, i);
    const Pair pz  = LoadPair(geo.posz_pad, i);
    const int vol  = geo.vol[i];
    const double dir[3] = {dxy.a, dxy.b, dz.a};
    double pos[3] = {pxy.a, pxy.b, pz.a};
    int nextVol;
    const double dist = DistanceToBoundary(g, vol, pos, dir, nextVol);
    double step = gs.a;
    const bool onBoundary = dist < step;
    if (onBoundary) step = dist;
    pos[0] += step * dir[0];
    pos[1] += step * dir[1];
    pos[2] += step * dir[2];
    StorePair(gstep_x, i, step, gs.b);
    StorePair(geo.posx_posy, i, pos[0], pos[1]);
    StorePair(geo.posz_pad, i, pos[2], 0.0);
    geo.nextVol[i] = nextVol;
    int f = meta[4 * i + 1];
    f = onBoundary ? (f | static_cast<int>(G4HB200_F_ON_BOUNDARY)) : (f & ~static_cast<int>(G4HB200_F_ON_BOUNDARY));
    meta[4 * i + 1] = f;
  }
}

// per CTA histogram in shared memory, flushed with one atomicAdd per touched bin
struct CtaHist {
  static constexpr int kMaxBins = 512;
  double bin[kMaxBins];
  __device__ void Init(int nbins) {
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) bin[k] = 0.0;
    __syncthreads();
  }
  __device__ void Flush(double* global, int nbins) {
    __syncthreads();
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) {
      if (bin[k] != 0.0) atomicAdd(global + k, bin[k]);
    }
  }
};

// ---- after Perform: e-/e+ ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsPerBlock)
ShowerElectronPostKernel(const __grid_constant__ SlabGeom g, const __grid_constant__ G4HB200ElectronBatch b,
                         const __grid_constant__ TrackGeo geo, const __grid_constant__ G4HB200ElectronBatch nb,
                         const __grid_constant__ TrackGeo ngeo, const __grid_constant__ ShowerScore sc) {
  __shared__ CtaHist hist;
  __shared__ CtaCounters<1> cc;
  __shared__ double sLeak;
  const int nbins = g.numLayers * g.numAbsorbers;
  hist.Init(nbins);
  cc.Init();
  if (threadIdx.x == 0) sLeak = 0.0;
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nLive  = geo.n_dev != nullptr ? *geo.n_dev : b.n;
  const int64_t nRound = RoundUpToCta(nLive);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    bool alive = false;
    Meta m{0, 0, 0, 0};
    Pair e{0, 0}, dxy{0, 0}, dzs{0, 0};
    double pos[3] = {0, 0, 0};
    int vol = -1;
    if (i < nLive) {
      m   = LoadMeta(b.meta, i);
      e   = LoadPair(b.ekin_logekin, i);
      dxy = LoadPair(b.dirx_diry, i);
      dzs = LoadPair(b.dirz_safety, i);
      const Pair ed  = LoadPair(b.edep_dispx, i);
      const Pair dyz = LoadPair(b.dispy_dispz, i);
      const Pair pxy = LoadPair(geo.posx_posy, i);
      const Pair pz  = LoadPair(geo.posz_pad, i);
      vol = geo.vol[i];
      pos[0] = pxy.a; pos[1] = pxy.b; pos[2] = pz.a;
      const bool onBoundary = (static_cast<uint32_t>(m.flags) & G4HB200_F_ON_BOUNDARY) != 0u;
      // MSC displacement (G4HepEmTrackingManager.cc:527-568)
      if (!onBoundary) {
        const double disp[3] = {ed.b, dyz.a, dyz.b};
        const double dLength2 = disp[0] * disp[0] + disp[1] * disp[1] + disp[2] * disp[2];
        const double kGeomMinLength = 5.0e-8;
        if (dLength2 > kGeomMinLength * kGeomMinLength) {
          const double dispR = sqrt(dLength2);
          const double postSafety = 0.99 * SlabSafety(g, vol, pos);
          double scale = 0.0;
          if (postSafety > 0.0 && dispR <= postSafety) {
            scale = 1.0;
          } else if (dispR < postSafety) {
            scale = 1.0;
          } else if (postSafety > kGeomMinLength) {
            scale = postSafety / dispR;
          }
          if (scale > 0.0) {
            pos[0] += disp[0] * scale;
            pos[1] += disp[1] * scale;
            pos[2] += disp[2] * scale;
          }
        }
      }
      // scoring: the deposit of this step belongs to the volume the step was made in
      if (ed.a != 0.0) atomicAdd(&hist.bin[vol], ed.a);
      // relocation
      int newVol = vol;
      int imc = m.imc;
      if (onBoundary) {
        newVol = geo.nextVol[i];
        if (newVol >= 0) imc = g.couple[newVol % g.numAbsorbers];
      }
      alive = e.a > 0.0 && newVol >= 0;
      if (e.a > 0.0 && newVol < 0) atomicAdd(&sLeak, e.a);
      // keep the post-step point for the secondaries of this track
      StorePair(geo.posx_posy, i, pos[0], pos[1]);
      StorePair(geo.posz_pad, i, pos[2], 0.0);
      if (alive) {
        // between the MSC sub-steps of one step the caller does not update the safety (it is set once per step,
        // G4HepEmTrackingManager.cc:419-423)
        const bool subStep = (static_cast<uint32_t>(m.flags) & G4HB200_F_MSC_SUBSTEP) != 0u;
        if (!subStep) dzs.b = onBoundary ? 0.0 : SlabSafety(g, newVol, pos);
        m.imc = imc;
        vol   = newVol;
      }
    }
    // survivors -> next-step store
    const unsigned mk = __ballot_sync(0xffffffffu, alive);
    int wb = 0;
    const int lane = threadIdx.x & 31;
    if (lane == 0 && mk != 0u) wb = atomicAdd(&cc.count[0], __popc(mk));
    wb = __shfl_sync(0xffffffffu, wb, 0);
    __syncthreads();
    if (threadIdx.x == 0) {
      const int t = cc.count[0];
      cc.base[0]  = t > 0 ? atomicAdd(sc.nextCount + 0, t) : 0;
      cc.count[0] = 0;
    }
    __syncthreads();
    if (alive) {
      const int64_t o = static_cast<int64_t>(cc.base[0]) + wb + __popc(mk & ((1u << lane) - 1u));
      if (o >= sc.capacity) {
        *sc.overflow = 1;
        continue;
      }
      StorePair(nb.ekin_logekin, o, e.a, e.b);
      StorePair(nb.dirx_diry, o, dxy.a, dxy.b);
      StorePair(nb.dirz_safety, o, dzs.a, dzs.b);
      const Pair n01 = LoadPair(b.nia01, i), n23 = LoadPair(b.nia23, i);
      const Pair ir = LoadPair(b.msc_irange_dynrf, i), tg = LoadPair(b.msc_tlimmin_gauss, i);
      StorePair(nb.nia01, o, n01.a, n01.b);
      StorePair(nb.nia23, o, n23.a, n23.b);
      StorePair(nb.msc_irange_dynrf, o, ir.a, ir.b);
      StorePair(nb.msc_tlimmin_gauss, o, tg.a, tg.b);
      StorePair(nb.edep_dispx, o, 0.0, 0.0);
      StoreMeta(nb.meta, o, m);
      nb.winner[o] = -1;
      StorePair(ngeo.posx_posy, o, pos[0], pos[1]);
      StorePair(ngeo.posz_pad, o, pos[2], 0.0);
      ngeo.vol[o] = vol;
      if ((static_cast<uint32_t>(m.flags) & G4HB200_F_MSC_SUBSTEP) != 0u) {
        // the step goes on in the next iteration: the mean free paths of its first round and the caller's local variables
        const Pair m01 = LoadPair(b.mfp01, i), m23 = LoadPair(b.mfp23, i);
        const Pair sl = LoadPair(geo.sub_left_eloss, i), sp = LoadPair(geo.sub_pre, i), sr = LoadPair(geo.sub_range_proc, i);
        StorePair(nb.mfp01, o, m01.a, m01.b);
        StorePair(nb.mfp23, o, m23.a, m23.b);
        StorePair(ngeo.sub_left_eloss, o, sl.a, sl.b);
        StorePair(ngeo.sub_pre, o, sp.a, sp.b);
        StorePair(ngeo.sub_range_proc, o, sr.a, sr.b);
      }
    }
  }
  hist.Flush(sc.hist, nbins);
  if (threadIdx.x == 0 && sLeak != 0.0) atomicAdd(sc.leak + 0, sLeak);
}

// ---- after Perform: gamma --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsPerBlock)
ShowerGammaPostKernel(const __grid_constant__ SlabGeom g, const __grid_constant__ G4HB200GammaBatch b,
                      const __grid_constant__ TrackGeo geo, const __grid_constant__ G4HB200GammaBatch nb,
                      const __grid_constant__ TrackGeo ngeo, const __grid_constant__ ShowerScore sc) {
  __shared__ CtaHist hist;
  __shared__ CtaCounters<1> cc;
  __shared__ double sLeak;
  const int nbins = g.numLayers * g.numAbsorbers;
  hist.Init(nbins);
  cc.Init();
  if (threadIdx.x == 0) sLeak = 0.0;
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nLive  = geo.n_dev != nullptr ? *geo.n_dev : b.n;
  const int64_t nRound = RoundUpToCta(nLive);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    bool alive = false;
    Meta m{0, 0, 0, 0};
    Pair e{0, 0}, dxy{0, 0}, dzn{0, 0}, pxy{0, 0}, pz{0, 0};
    int vol = -1;
    if (i < nLive) {
      m   = LoadMeta(b.meta, i);
      e   = LoadPair(b.ekin_logekin, i);
      dxy = LoadPair(b.dirx_diry, i);
      dzn = LoadPair(b.dirz_nia0, i);
      const Pair ep = LoadPair(b.edep_pemxsec, i);
      pxy = LoadPair(geo.posx_posy, i);
      pz  = LoadPair(geo.posz_pad, i);
      vol = geo.vol[i];
      const bool onBoundary = (static_cast<uint32_t>(m.flags) & G4HB200_F_ON_BOUNDARY) != 0u;
      if (ep.a != 0.0) atomicAdd(&hist.bin[vol], ep.a);
      int newVol = vol;
      if (onBoundary) {
        newVol = geo.nextVol[i];
        if (newVol >= 0) m.imc = g.couple[newVol % g.numAbsorbers];
      }
      alive = e.a > 0.0 && newVol >= 0;
      if (e.a > 0.0 && newVol < 0) atomicAdd(&sLeak, e.a);
      vol = newVol;
    }
    const unsigned mk = __ballot_sync(0xffffffffu, alive);
    int wb = 0;
    const int lane = threadIdx.x & 31;
    if (lane == 0 && mk != 0u) wb = atomicAdd(&cc.count[0], __popc(mk));
    wb = __shfl_sync(0xffffffffu, wb, 0);
    __syncthreads();
    if (threadIdx.x == 0) {
      const int t = cc.count[0];
      cc.base[0]  = t > 0 ? atomicAdd(sc.nextCount + 1, t) : 0;
      cc.count[0] = 0;
    }
    __syncthreads();
    if (alive) {
      const int64_t o = static_cast<int64_t>(cc.base[0]) + wb + __popc(mk & ((1u << lane) - 1u));
      if (o >= sc.capacity) {
        *sc.overflow = 1;
        continue;
      }
      StorePair(nb.ekin_logekin, o, e.a, e.b);
      StorePair(nb.dirx_diry, o, dxy.a, dxy.b);
      StorePair(nb.dirz_nia0, o, dzn.a, dzn.b);
      const Pair ep = LoadPair(b.edep_pemxsec, i);
      StorePair(nb.edep_pemxsec, o, 0.0, ep.b);
      StoreMeta(nb.meta, o, m);
      nb.winner[o] = b.winner[i];
      StorePair(ngeo.posx_posy, o, pxy.a, pxy.b);
      StorePair(ngeo.posz_pad, o, pz.a, 0.0);
      ngeo.vol[o] = vol;
    }
  }
  hist.Flush(sc.hist, nbins);
  if (threadIdx.x == 0 && sLeak != 0.0) atomicAdd(sc.leak + 1, sLeak);
}

// ---- secondaries -> new tracks ----------------------------------------------------------------------------------------
// (id, first draw) of the k-th secondary of a parent: words of one Philox block keyed by the global seed
__device__ __forceinline__ void ChildStream(uint64_t seed, int parentId, int parentDraw, int slot, int& id, int& draw) {
  const Philox4 r = PhiloxBlock(static_cast<uint32_t>(seed) ^ 0x5EC0DA2Au, static_cast<uint32_t>(seed >> 32),
                                static_cast<uint32_t>(parentId), static_cast<uint32_t>(parentDraw));
  const uint32_t a = slot == 0 ? r.x : r.z;
  const uint32_t b = slot == 0 ? r.y : r.w;
  id   = static_cast<int>(a);
  draw = static_cast<int>(b & 0x3FFFFFFEu);
}

// the parents sit in a batch of one kind: only their meta (id, draw counter), position and volume are needed
// Secondary production cuts of the caller (StackSecondaries(..., isApplyCuts), G4HepEmTrackingManager.cc:1254-1326, with
// fIsApplyCuts of the region of the parent's pre-step couple, .cc:426): a secondary below the cut of its kind is not
// tracked, its kinetic energy (+ 2 m_e c^2 for a positron) is deposited where the parent's step deposits.
// returns the energy to deposit (> 0: the secondary is dropped)
G4H_FN double SecondaryBelowCut(const TablesView& tv, int parentImc, int kind, double secEKin) {
  const int ireg = G4H_LD(tv.mcIreg + parentImc);
  const bool isApplyCuts = (static_cast<int>(G4H_LD(tv.regionPars + 8 * ireg + kRCallerFlags)) & 2) != 0;
  if (!isApplyCuts) return 0.0;
  const double* cuts = tv.mcCuts + 4 * parentImc;
  if (kind == G4HB200_SEC_ELECTRON) return secEKin < G4H_LD(cuts + kCElCut) ? secEKin : 0.0;
  if (kind == G4HB200_SEC_POSITRON) {
    return (kElectronMassC2 < G4H_LD(cuts + kCGamCut) && secEKin < G4H_LD(cuts + kCPosCut)) ? secEKin + 2 * kElectronMassC2 : 0.0;
  }
  return secEKin < G4H_LD(cuts + kCGamCut) ? secEKin : 0.0;
}

__global__ void __launch_bounds__(kThreadsPerBlock)
ShowerSecondaryKernel(const __grid_constant__ TablesView tv, const __grid_constant__ SlabGeom g, uint64_t seed,
                      const __grid_constant__ G4HB200SecondaryQueue q, const int32_t* __restrict__ parentMeta,
                      const __grid_constant__ TrackGeo pgeo, const __grid_constant__ G4HB200ElectronBatch ne,
                      const __grid_constant__ TrackGeo negeo, const __grid_constant__ G4HB200GammaBatch ng,
                      const __grid_constant__ TrackGeo nggeo, const __grid_constant__ ShowerScore sc) {
  __shared__ CtaCounters<2> cc;
  __shared__ CtaHist hist;
  const int nbins = g.numLayers * g.numAbsorbers;
  cc.Init();
  if (g.inheritCouple == 0) hist.Init(nbins);
  const int cnt = q.count[0];
  const int nRound = static_cast<int>(RoundUpToCta(cnt));
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nRound; j += stride) {
    int route = -1;  // 0: e-/e+ store, 1: gamma store
    Pair dxy{0, 0}, dze{0, 0}, pxy{0, 0}, pz{0, 0};
    int kind = 0, id = 0, draw = 0, vol = -1, parentImc = 0;
    if (j < cnt) {
      dxy = LoadPair(q.dirx_diry, j);
      dze = LoadPair(q.dirz_ekin, j);
      const int2 pk = reinterpret_cast<const int2*>(q.parent_kind)[j];
      const int2 ps = reinterpret_cast<const int2*>(q.parent_slot)[j];
      kind = pk.y;
      const int p = ps.x;
      const int parentDraw = parentMeta[4 * p + 3];
      parentImc = parentMeta[4 * p + 0];
      ChildStream(seed, pk.x, parentDraw, ps.y, id, draw);
      pxy = LoadPair(pgeo.posx_posy, p);
      pz  = LoadPair(pgeo.posz_pad, p);
      vol = pgeo.vol[p];
      route = kind == G4HB200_SEC_GAMMA ? 1 : 0;
      if (g.inheritCouple == 0) {
        const double cutEdep = SecondaryBelowCut(tv, parentImc, kind, dze.b);
        if (cutEdep > 0.0) {
          atomicAdd(&hist.bin[vol], cutEdep);
          route = -1;
        }
      }
    }
    // reserve slots in the two next-step stores
    const unsigned active = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int offset = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const unsigned mk = __ballot_sync(active, route == k);
      if (mk != 0u) {
        int wb = 0;
        if (lane == 0) wb = atomicAdd(&cc.count[k], __popc(mk));
        wb = __shfl_sync(active, wb, 0);
        if (route == k) offset = wb + __popc(mk & ((1u << lane) - 1u));
      }
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      const int t = cc.count[threadIdx.x];
      cc.base[threadIdx.x]  = t > 0 ? atomicAdd(sc.nextCount + threadIdx.x, t) : 0;
      cc.count[threadIdx.x] = 0;
    }
    __syncthreads();
    if (route < 0) continue;
    const int64_t o = static_cast<int64_t>(cc.base[route]) + offset;
    if (o >= sc.capacity) {
      *sc.overflow = 1;
      continue;
    }
    const double pos[3] = {pxy.a, pxy.b, pz.a};
    const int imc = g.inheritCouple ? parentImc : g.couple[vol % g.numAbsorbers];
    if (route == 0) {
      // G4HepEmElectronTrack::ReSet() state (G4HepEmTrack.hh:175-206, G4HepEmMSCTrackData.hh:54-78)
      StorePair(ne.ekin_logekin, o, dze.b, 100.0);
      StorePair(ne.dirx_diry, o, dxy.a, dxy.b);
      StorePair(ne.dirz_safety, o, dze.a, SlabSafety(g, vol, pos));
      StorePair(ne.nia01, o, -1.0, -1.0);
      StorePair(ne.nia23, o, -1.0, -1.0);
      StorePair(ne.msc_irange_dynrf, o, 1.0e+21, 0.04);
      StorePair(ne.msc_tlimmin_gauss, o, 1.0e-7, 0.0);
      StorePair(ne.edep_dispx, o, 0.0, 0.0);
      const int flags = static_cast<int>(G4HB200_F_MSC_FIRST_STEP | (kind == G4HB200_SEC_POSITRON ? G4HB200_F_POSITRON : 0u));
      StoreMeta(ne.meta, o, Meta{imc, flags, id, draw});
      ne.winner[o] = -1;
      StorePair(negeo.posx_posy, o, pos[0], pos[1]);
      StorePair(negeo.posz_pad, o, pos[2], 0.0);
      negeo.vol[o] = vol;
    } else {
      StorePair(ng.ekin_logekin, o, dze.b, 100.0);
      StorePair(ng.dirx_diry, o, dxy.a, dxy.b);
      StorePair(ng.dirz_nia0, o, dze.a, -1.0);
      StorePair(ng.edep_pemxsec, o, 0.0, 0.0);
      StoreMeta(ng.meta, o, Meta{imc, 0, id, draw});
      ng.winner[o] = -1;
      StorePair(nggeo.posx_posy, o, pos[0], pos[1]);
      StorePair(nggeo.posz_pad, o, pos[2], 0.0);
      nggeo.vol[o] = vol;
    }
  }
  if (g.inheritCouple == 0) hist.Flush(sc.hist, nbins);
}

// ---- the top of an iteration of the graph-driven tail: what the host does between two iterations otherwise --------------------
// (populations of the iteration <- what the previous one left in the next-step stores; counters back to zero; statistics)
// elQueueCount / gmQueueCount: the interaction-queue counters of the two pipelines (they zero them again with a memset
// node of their own; written here too so that every byte a kernel of the graph reads has been written by a kernel --
// compute-sanitizer's initcheck does not follow memset nodes)
__global__ void ShowerIterKernel(ShowerCtrl* ctrl, int32_t* nextCount, int32_t* secElCount, int32_t* secGmCount,
                                 const int32_t* overflow, int32_t* elQueueCount, int32_t* gmQueueCount) {
  if (blockIdx.x == 0 && threadIdx.x < kNumElQueues) {
    elQueueCount[threadIdx.x] = 0;
    gmQueueCount[threadIdx.x] = 0;
  }
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ctrl->sumSec += static_cast<long long>(*secElCount) + *secGmCount;  // secondaries the previous iteration created
  // a store that ran out of capacity ends the run (the host reports it at its next look): nothing more is stepped
  const bool dead = *overflow != 0;
  const int nEl = dead ? 0 : nextCount[0], nGm = dead ? 0 : nextCount[1];
  ctrl->cur[0]  = nEl;
  ctrl->cur[1]  = nGm;
  nextCount[0]  = 0;
  nextCount[1]  = 0;
  *secElCount   = 0;
  *secGmCount   = 0;
  if (nEl > 0 || nGm > 0) {
    ctrl->steps += 1;
    ctrl->sumEl += nEl;
    ctrl->sumGm += nGm;
    if (nEl > ctrl->peakEl) ctrl->peakEl = nEl;
    if (nGm > ctrl->peakGm) ctrl->peakGm = nGm;
  }
}

// primaries: at the front face of the calorimeter, along +x, entering (on the boundary)
__global__ void __launch_bounds__(kThreadsPerBlock)
ShowerPrimaryKernel(const __grid_constant__ SlabGeom g, int64_t n, int kind, double ekin, int firstId,
                    const __grid_constant__ G4HB200ElectronBatch ne, const __grid_constant__ TrackGeo negeo,
                    const __grid_constant__ G4HB200GammaBatch ng, const __grid_constant__ TrackGeo nggeo) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < n; o += stride) {
    const int id = firstId + static_cast<int>(o);
    const int imc = g.couple[0];
    if (kind != G4HB200_SEC_GAMMA) {
      StorePair(ne.ekin_logekin, o, ekin, 100.0);
      StorePair(ne.dirx_diry, o, 1.0, 0.0);
      StorePair(ne.dirz_safety, o, 0.0, 0.0);
      StorePair(ne.nia01, o, -1.0, -1.0);
      StorePair(ne.nia23, o, -1.0, -1.0);
      StorePair(ne.msc_irange_dynrf, o, 1.0e+21, 0.04);
      StorePair(ne.msc_tlimmin_gauss, o, 1.0e-7, 0.0);
      StorePair(ne.edep_dispx, o, 0.0, 0.0);
      const int flags = static_cast<int>(G4HB200_F_MSC_FIRST_STEP | G4HB200_F_ON_BOUNDARY |
                                         (kind == G4HB200_SEC_POSITRON ? G4HB200_F_POSITRON : 0u));
      StoreMeta(ne.meta, o, Meta{imc, flags, id, 0});
      ne.winner[o] = -1;
      StorePair(negeo.posx_posy, o, g.xFront, 0.0);
      StorePair(negeo.posz_pad, o, 0.0, 0.0);
      negeo.vol[o] = 0;
    } else {
      StorePair(ng.ekin_logekin, o, ekin, 100.0);
      StorePair(ng.dirx_diry, o, 1.0, 0.0);
      StorePair(ng.dirz_nia0, o, 0.0, -1.0);
      StorePair(ng.edep_pemxsec, o, 0.0, 0.0);
      StoreMeta(ng.meta, o, Meta{imc, static_cast<int>(G4HB200_F_ON_BOUNDARY), id, 0});
      ng.winner[o] = -1;
      StorePair(nggeo.posx_posy, o, g.xFront, 0.0);
      StorePair(nggeo.posz_pad, o, 0.0, 0.0);
      nggeo.vol[o] = 0;
    }
  }
}
// BASELINE configs[3]: n tracks, one third each e-, e+, gamma, contiguous per particle and, inside a particle,
// per couple (the queue order a stepping loop keeps); E log-uniform in [emin, emax], isotropic directions, first-step
// state, not on a boundary.  The population is a pure function of (seed, track index) built from +, *, /, sqrt and the VDT
// exponential only, so that the CPU driver of the parity test (tests/shower_oracle.py: mixed_population) generates the
// very same tracks: uniforms of the stream keyed (seed ^ 0x1A2B3C4D, track index): draw 0 energy, 1 cos(theta), 2 safety,
// 3.. the azimuth by von Neumann's rejection (a point in the unit disc -> cos / sin of twice its polar angle).
// lmin = log(emin), lrange = log(emax / emin): evaluated by the host (one value for all tracks).
__global__ void __launch_bounds__(kThreadsPerBlock)
MixedPopulationKernel(int64_t nEl, int64_t nGm, int numCouples, double lmin, double lrange, uint64_t seed,
                      const __grid_constant__ G4HB200ElectronBatch ne, const __grid_constant__ TrackGeo negeo,
                      const __grid_constant__ G4HB200GammaBatch ng, const __grid_constant__ TrackGeo nggeo) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n = nEl + nGm;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += stride) {
    const bool isGamma = t >= nEl;
    const int64_t o = isGamma ? t - nEl : t;
    const int64_t half = nEl / 2;
    const bool isPositron = !isGamma && o >= half;
    const int64_t inKind = isGamma ? o : (isPositron ? o - half : o);
    const int64_t kindSize = isGamma ? nGm : (isPositron ? nEl - half : half);
    const int imc = static_cast<int>((inKind * numCouples) / (kindSize > 0 ? kindSize : 1));
    const int id = static_cast<int>(t);
    Rng gen;
    gen.Init(seed ^ 0x1A2B3C4DULL, static_cast<uint32_t>(id), 0u, false, 0.0);
    const double ekin   = Exp(lmin + gen.Flat() * lrange);
    const double cost   = 2.0 * gen.Flat() - 1.0;
    const double safety = gen.Flat();
    const double sint   = sqrt((1.0 - cost) * (1.0 + cost));
    double vx, vy, r2;
    do {
      vx = 2.0 * gen.Flat() - 1.0;
      vy = 2.0 * gen.Flat() - 1.0;
      r2 = vx * vx + vy * vy;
    } while (r2 > 1.0 || r2 == 0.0);
    const double cphi = (vx * vx - vy * vy) / r2;
    const double sphi = 2.0 * vx * vy / r2;
    if (!isGamma) {
      StorePair(ne.ekin_logekin, o, ekin, 100.0);
      StorePair(ne.dirx_diry, o, sint * cphi, sint * sphi);
      StorePair(ne.dirz_safety, o, cost, safety);
      StorePair(ne.nia01, o, -1.0, -1.0);
      StorePair(ne.nia23, o, -1.0, -1.0);
      StorePair(ne.msc_irange_dynrf, o, 1.0e+21, 0.04);
      StorePair(ne.msc_tlimmin_gauss, o, 1.0e-7, 0.0);
      StorePair(ne.edep_dispx, o, 0.0, 0.0);
      StoreMeta(ne.meta, o, Meta{imc, static_cast<int>(G4HB200_F_MSC_FIRST_STEP | (isPositron ? G4HB200_F_POSITRON : 0u)), id, 0});
      ne.winner[o] = -1;
      StorePair(negeo.posx_posy, o, 0.0, 0.0);
      StorePair(negeo.posz_pad, o, 0.0, 0.0);
      negeo.vol[o] = 0;
    } else {
      StorePair(ng.ekin_logekin, o, ekin, 100.0);
      StorePair(ng.dirx_diry, o, sint * cphi, sint * sphi);
      StorePair(ng.dirz_nia0, o, cost, -1.0);
      StorePair(ng.edep_pemxsec, o, 0.0, 0.0);
      StoreMeta(ng.meta, o, Meta{imc, 0, id, 0});
      ng.winner[o] = -1;
      StorePair(nggeo.posx_posy, o, 0.0, 0.0);
      StorePair(nggeo.posz_pad, o, 0.0, 0.0);
      nggeo.vol[o] = 0;
    }
  }
}
#endif  // __CUDACC__

}  // namespace g4h
#endif
