// g4h_trackops.cuh -- the track-level statics of the reference's managers as batch kernels.
//
// The production caller of G4HepEm does not use the two-call HowFar / Perform protocol: G4HepEmTrackingManager::TrackElectron
// (G4HepEm/G4HepEm/src/G4HepEmTrackingManager.cc:428-665) drives the pieces one by one with geometry in between --
//   HowFarToDiscreteInteraction, loop { HowFarToMSC, [geometry], UpdatePStepLength, UpdateNumIALeft, ApplyMeanEnergyLoss,
//   SampleMSC, [displacement] }, SampleLossFluctuations, PerformDiscrete | annihilation at rest
// (G4HepEmRun/include/G4HepEmElectronManager.hh:90-206), and TrackGamma (.cc:985-1140) drives HowFar, UpdateNumIALeft,
// SelectInteraction and Perform of G4HepEmGammaManager (G4HepEmGammaManager.hh:32-55) separately.  Every one of those is an
// entry point here: one launch applies ONE of them to every track of a device batch, in place, exactly like the static
// function works in place on its track object.  The whole track (persistent, result and hand-over groups, and the
// pre-step energy of G4HepEmElectronTrack::fPreStepEKin / fPreStepLogEKin in `prestep`) is the state between calls.
//
// These are the drop-in pieces, not the fast path: a caller that can batch whole steps uses g4hb200_electron_step.
#ifndef G4H_TRACKOPS_CUH
#define G4H_TRACKOPS_CUH

#include "g4h_kernels.cuh"

namespace g4h {

// values of the `op` argument of g4hb200_electron_track_op / g4hb200_gamma_track_op (include/g4hepem_b200.h)
enum ElTrackOp {
  kOpHowFarDiscrete = 0, kOpHowFarMSC, kOpUpdatePStep, kOpUpdateNIA, kOpMeanELoss, kOpSampleMSC, kOpLossFluct, kOpDiscrete,
  kOpAnnihilateAtRest, kOpPerformContinuous, kOpResampleNIA, kNumElTrackOps
};
enum GmTrackOp { kGOpHowFarTrack = 0, kGOpUpdateNIA, kGOpSelectInteraction, kGOpPerformSelected, kNumGmTrackOps };

// UpdateNumIALeft (G4HepEmElectronManager.icc:205-214)
G4H_FN void UpdateNumIALeft(ElectronState& s) {
  const double pStepLength = s.pStep;
  s.nIA[0] -= pStepLength / s.mfp[0];
  s.nIA[1] -= pStepLength / s.mfp[1];
  s.nIA[2] -= pStepLength / s.mfp[2];
  s.nIA[3] -= pStepLength / s.mfp[3];
}

#if defined(__CUDACC__)
// flag[i]: the bool the reference's function returns (stopped / delta interaction); untouched for the void ones
template <int kOp>
__global__ void __launch_bounds__(kThreadsPerBlock)
ElTrackOpKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                const __grid_constant__ G4HB200SecondaryQueue q, uint64_t seed, int32_t* __restrict__ flag) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nRound = RoundUpToCta(b.n);
  constexpr bool kHasSecondaries = kOp == kOpDiscrete || kOp == kOpAnnihilateAtRest;
  __shared__ CtaCounters<1> cc;
  if (kHasSecondaries) cc.Init();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    const bool valid = i < b.n;
    ElectronState s;
    Rng rng;
    Secondaries sec;
    sec.n = 0;
    if (valid) {
      LoadElectron(b, i, seed, s, rng);
      LoadElectronHandOver(b, i, s);
      if (b.prestep != nullptr) {
        const Pair pre = LoadPair(b.prestep, i);
        s.preStepEkin    = pre.a;
        s.preStepLogEkin = pre.b;
      }
      bool result = false;
      if (kOp == kOpResampleNIA) ResampleNumIALeft(s, rng);  // the loop at the top of HowFar (.icc:39-43) / TrackElectron (.cc:430-434)
      if (kOp == kOpHowFarDiscrete) HowFarToDiscreteInteraction(tv, s);
      if (kOp == kOpHowFarMSC) HowFarToMSC(tv, s, rng);
      if (kOp == kOpUpdatePStep) UpdatePStepLength(s);
      if (kOp == kOpUpdateNIA) UpdateNumIALeft(s);
      if (kOp == kOpMeanELoss) result = ApplyMeanEnergyLoss(tv, s);
      if (kOp == kOpSampleMSC) SampleMSC(tv, s, rng);
      if (kOp == kOpLossFluct) result = SampleLossFluctuations(tv, s, rng);
      if (kOp == kOpPerformContinuous) result = PerformContinuous(tv, s, rng);
      if (kOp == kOpDiscrete) PerformDiscrete(tv, s, rng, sec);
      if (kOp == kOpAnnihilateAtRest) AnnihilateAtRest(rng, sec);
      StoreElectron(b, i, s, rng);
      StoreElectronHandOver(b, i, s);
      if (b.prestep != nullptr) StorePair(b.prestep, i, s.preStepEkin, s.preStepLogEkin);
      if (flag != nullptr && (kOp == kOpMeanELoss || kOp == kOpLossFluct || kOp == kOpPerformContinuous)) flag[i] = result ? 1 : 0;
    }
    if (kHasSecondaries) AppendSecondaries(cc, q, sec, valid ? s.id : 0, i);
  }
}

// CheckDelta(data, track, rand) with a caller supplied uniform (G4HepEmElectronManager.icc:408-423): flag[i] = delta interaction
__global__ void __launch_bounds__(kThreadsPerBlock)
ElCheckDeltaKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b, const double* __restrict__ urnd,
                   int32_t* __restrict__ flag) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < b.n; i += stride) {
    ElectronState s;
    Rng rng;
    LoadElectron(b, i, 0, s, rng);
    LoadElectronHandOver(b, i, s);
    flag[i] = CheckDelta(tv, s, urnd[i]) ? 1 : 0;
    // CheckDelta caches the logarithm of the energy in the track (GetLogEKin)
    StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
  }
}

template <int kOp>
__global__ void __launch_bounds__(kThreadsPerBlock)
GammaTrackOpKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200GammaBatch b,
                   const __grid_constant__ G4HB200SecondaryQueue q, uint64_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nRound = RoundUpToCta(b.n);
  __shared__ CtaCounters<1> cc;
  if (kOp == kGOpPerformSelected) cc.Init();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    const bool valid = i < b.n;
    GammaState s;
    Rng rng;
    Secondaries sec;
    sec.n = 0;
    if (valid) {
      LoadGamma(b, i, seed, s, rng);
      const int flags = b.meta[4 * i + 1];
      LoadGammaHandOver(b, i, s);
      if (kOp == kGOpHowFarTrack) {
        // HowFar(data, pars, gammaTrack) (G4HepEmGammaManager.icc:38-48): no resampling of the interaction length
        const double lekin    = GetLogEKin(s);
        const double totMXSec = GammaTotalMacXSec(tv, G4H_LD(tv.mcImat + s.imc), s.ekin, lekin, s.peMXsec);
        const double totalMFP = (totMXSec > 0.) ? 1. / totMXSec : kALargeValue;
        s.mfp0  = totalMFP;
        s.gStep = totalMFP * s.nIA0;
      }
      if (kOp == kGOpUpdateNIA) s.nIA0 -= s.gStep / s.mfp0;  // UpdateNumIALeft (.icc:97-105)
      if (kOp == kGOpSelectInteraction) {
        // SelectInteraction (.icc:173-177) -> SampleInteraction (.icc:186-219)
        const double urnd = rng.Flat();
        s.nIA0 = -1.0;
        const double lekin = (s.ekin > tv.gmEMax1) ? GetLogEKin(s) : 0.0;
        s.winner = GammaSampleInteraction(tv, G4H_LD(tv.mcImat + s.imc), s.ekin, lekin, s.mfp0, urnd, s.peMXsec);
      }
      if (kOp == kGOpPerformSelected) {
        // Perform (.icc:54-94) for a track whose interaction has been selected by the caller
        s.nIA0 -= s.gStep / s.mfp0;
        s.edep = 0.0;
        if (!s.onBoundary) {
          const int iDProc = s.winner;
          if (iDProc == 0) s.nIA0 = -1.0;  // SetNumIALeft(-1, iDProc): only slot 0 is live state of a gamma
          if (iDProc == 0) PerformConversion(tv, s, rng, sec);
          if (iDProc == 1) PerformCompton(tv, s, rng, sec);
          if (iDProc == 2) PerformPhotoelectric(tv, s, rng, sec);
          const double finalEkin = s.ekin;
          if (finalEkin > 0.0 && finalEkin <= tv.gammaTrackingCut) {
            SetEKin(s, 0.0);
            s.edep += finalEkin;
          }
        }
      }
      StoreGamma(b, i, s, rng, flags);
    }
    if (kOp == kGOpPerformSelected) AppendSecondaries(cc, q, sec, valid ? s.id : 0, i);
  }
}
#endif  // __CUDACC__

}  // namespace g4h
#endif
