// g4h_pipeline.cuh -- G4HepEmElectronManager::Perform as a pipeline of small kernels over interaction queues.
//
// Why: one thread running the whole Perform of one track is 150 KB of SASS; warps of a CTA sit in different
// branches of it, the instruction cache thrashes (87 % of issue slots stalled on "no instruction",
// profiles/r01_electron_step_monolith.md) and 11 of 32 lanes are active on average.  Here every kernel is a
// few thousand instructions, and a track only enters the kernels it needs:
//
//   ElContinuousKernel   all tracks: UpdatePStepLength, UpdateNumIALeft, ApplyMeanEnergyLoss, SampleMSC
//                        (.icc:375-405 up to the fluctuation) -> queues: fluctuation | discrete | at-rest
//   ElFluctuationKernel  queue: SampleLossFluctuations (.icc:324-368)           -> queues: discrete | at-rest
//   ElDiscreteKernel     queue: PerformDiscrete head (.icc:425-441): reset nIA, CheckDelta, model choice
//                                                                               -> queues: one per model
//   ElSamplerKernel<K>   queue K: Moller | Bhabha | Seltzer-Berger | rel. brem | annihilation in flight | at rest
//
// Queues are arrays of track indices in the handle's workspace, filled with one global atomicAdd per CTA
// and queue (warp ballots + a shared-memory counter, g4h_kernels.cuh); results go back to the track's own slot, so the outcome does not depend on queue order.
// Between kernels the track lives in its batch groups plus one workspace group (the pre-step energy).
// The uniform stream is keyed by (seed, track id, draw counter), the counter travels in meta[3].
#ifndef G4H_PIPELINE_CUH
#define G4H_PIPELINE_CUH

#include "g4h_kernels.cuh"
#include "g4h_stages.cuh"

namespace g4h {

enum ElQueue { kQFluct = 0, kQDiscrete, kQAtRest, kQMoller, kQBhabha, kQSB, kQRB, kQAnnih, kQConvRange, kNumElQueues };

struct ElectronWork {
  double* prestep;                // [n] pairs {preStepEkin, preStepLogEkin}
  int32_t* queue[kNumElQueues];   // [n] track indices each
  int32_t* count;                 // [kNumElQueues]
};

// ---- HowFar in two stages (g4h_stages.cuh) ---------------------------------------------------------------------
#ifndef G4H_MINB_XS
#define G4H_MINB_XS 3
#endif
#ifndef G4H_MINB_MSCLIM
#define G4H_MINB_MSCLIM 4
#endif
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_XS)
ElHowFarXSKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b, uint64_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < b.n; i += stride) {
    StageHowFarXS(tv, b, i, seed);
  }
}

template <bool kStoreResults>
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_MSCLIM)
ElHowFarMSCKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                  const __grid_constant__ ElectronWork w, uint64_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nRound = RoundUpToCta(b.n);
  __shared__ CtaCounters<1> cc;
  cc.Init();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    const bool queued = i < b.n && StageHowFarMSC<kStoreResults>(tv, b, i, seed);
    RouteToQueues<1>(cc, queued ? 0 : -1, static_cast<int32_t>(i), w.queue + kQConvRange, w.count + kQConvRange);
  }
}

__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_QUEUE)
ElHowFarMSCRangeKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                       const __grid_constant__ ElectronWork w) {
  const int cnt = w.count[kQConvRange];
  const int stride = gridDim.x * blockDim.x;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += stride) {
    StageHowFarMSCRange(tv, b, w.queue[kQConvRange][q]);
  }
}

// ---- continuous part for every track --------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_CONT)
ElContinuousKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                   const __grid_constant__ ElectronWork w, uint64_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nRound = RoundUpToCta(b.n);
  __shared__ CtaCounters<3> cc;
  cc.Init();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    int route = -1;
    if (i < b.n) {
      ElectronState s;
      Rng rng;
      LoadElectron(b, i, seed, s, rng);
      LoadElectronHandOver(b, i, s);
      // G4HepEmElectronManager::Perform (.icc:461-470)
      s.edep  = 0;
      s.pStep = s.gStep;
      if (s.gStep > 0.) {
        // PerformContinuous (.icc:375-405)
        s.preStepEkin    = s.ekin;
        s.preStepLogEkin = GetLogEKin(s);
        UpdatePStepLength(s);
        bool stopped = false;
        bool toDiscrete = false;
        if (s.pStep <= 0.0) {
          toDiscrete = true;
        } else {
          s.nIA[0] -= s.pStep / s.mfp[0];
          s.nIA[1] -= s.pStep / s.mfp[1];
          s.nIA[2] -= s.pStep / s.mfp[2];
          s.nIA[3] -= s.pStep / s.mfp[3];
          stopped = ApplyMeanEnergyLoss(tv, s);
          if (!stopped) {
            SampleMSC(tv, s, rng);
            if (LossFluctuationIsSampled(tv, s)) {
              route = kQFluct;
            } else {
              double ekin, edep;
              stopped = LossFluctuationFinish(tv, s.preStepEkin, s.ekin, s.edep, ekin, edep);
              SetEKin(s, ekin);
              s.edep = edep;
              toDiscrete = !stopped;
            }
          }
        }
        if (stopped && s.isPositron) route = kQAtRest;
        if (toDiscrete && s.winner >= 0 && !s.onBoundary) route = kQDiscrete;
        StorePair(w.prestep, i, s.preStepEkin, s.preStepLogEkin);
      }
      StoreElectron(b, i, s, rng);
      StorePair(b.tstep_zpath, i, s.trueStep, s.zPath);
    }
    // kQFluct, kQDiscrete, kQAtRest are queues 0..2
    RouteToQueues<3>(cc, route, static_cast<int32_t>(i), w.queue, w.count);
  }
}

// ---- energy loss fluctuation over its queue ---------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_QUEUE)
ElFluctuationKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                    const __grid_constant__ ElectronWork w, uint64_t seed) {
  const int cnt = w.count[kQFluct];
  const int nRound = static_cast<int>(RoundUpToCta(cnt));
  const int stride = gridDim.x * blockDim.x;
  __shared__ CtaCounters<2> cc;
  cc.Init();
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nRound; q += stride) {
    int route = -1;
    int32_t i = 0;
    if (q < cnt) {
      i = w.queue[kQFluct][q];
      const Meta m  = LoadMeta(b.meta, i);
      const Pair pre = LoadPair(w.prestep, i);
      const Pair ed  = LoadPair(b.edep_dispx, i);
      const Pair tg  = LoadPair(b.msc_tlimmin_gauss, i);
      const uint32_t f = static_cast<uint32_t>(m.flags);
      const bool isPositron = (f & G4HB200_F_POSITRON) != 0u;
      Rng rng;
      rng.Init(seed, static_cast<uint32_t>(m.id), static_cast<uint32_t>(m.draw), (f & G4HB200_F_GAUSS_CACHED) != 0u, tg.b);
      double finalEkin, eloss;
      LossFluctuationSample(tv, m.imc, !isPositron, pre.a, ed.a, rng, finalEkin, eloss);
      double ekin, edep;
      const bool stopped = LossFluctuationFinish(tv, pre.a, finalEkin, eloss, ekin, edep);
      StorePair(b.ekin_logekin, i, ekin, 100.0);
      StorePair(b.edep_dispx, i, edep, ed.b);
      StorePair(b.msc_tlimmin_gauss, i, tg.a, rng.gauss);
      const uint32_t fl = (f & ~G4HB200_F_GAUSS_CACHED) | (rng.hasGauss ? G4HB200_F_GAUSS_CACHED : 0u);
      StoreMeta(b.meta, i, Meta{m.imc, static_cast<int>(fl), m.id, static_cast<int>(rng.draw)});
      if (stopped) {
        if (isPositron) route = kQAtRest;
      } else if (b.winner[i] >= 0 && (f & G4HB200_F_ON_BOUNDARY) == 0u) {
        route = kQDiscrete;
      }
    }
    RouteToQueues<2>(cc, route - kQDiscrete, i, w.queue + kQDiscrete, w.count + kQDiscrete);
  }
}

// ---- head of PerformDiscrete: real or delta interaction, which model -----------------------------------------------
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_QUEUE)
ElDiscreteKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                 const __grid_constant__ ElectronWork w, uint64_t seed) {
  const int cnt = w.count[kQDiscrete];
  const int nRound = static_cast<int>(RoundUpToCta(cnt));
  const int stride = gridDim.x * blockDim.x;
  __shared__ CtaCounters<5> cc;
  cc.Init();
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nRound; q += stride) {
    int route = -1;
    int32_t i = 0;
    if (q < cnt) {
      i = w.queue[kQDiscrete][q];
      const Meta m = LoadMeta(b.meta, i);
      const Pair e = LoadPair(b.ekin_logekin, i);
      const int iDProc = b.winner[i];
      double* niaGroup = iDProc < 2 ? b.nia01 : b.nia23;
      const double* mfpGroup = iDProc < 2 ? b.mfp01 : b.mfp23;
      Pair nia = LoadPair(niaGroup, i);
      const Pair mfp = LoadPair(mfpGroup, i);
      ElectronState s;
      s.ekin = e.a; s.logEkin = e.b;
      s.imc = m.imc; s.id = m.id; s.winner = iDProc;
      s.isPositron = (static_cast<uint32_t>(m.flags) & G4HB200_F_POSITRON) != 0u;
      Rng rng;
      rng.Init(seed, static_cast<uint32_t>(m.id), static_cast<uint32_t>(m.draw), false, 0.0);
      // s.nIA[iDProc] = -1 (.icc:437)
      if (iDProc & 1) nia.b = -1.0; else nia.a = -1.0;
      const bool isDelta = CheckDeltaWith(tv, s, (iDProc & 1) ? mfp.b : mfp.a, rng.Flat());
      StorePair(niaGroup, i, nia.a, nia.b);
      StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
      StoreMeta(b.meta, i, Meta{m.imc, m.flags, m.id, static_cast<int>(rng.draw)});
      if (!isDelta) {
        if (iDProc == 0) {
          // Ioni::Perform (Ioni.icc:24-27): nothing happens when the maximum transfer is below the cut
          const double elCut = G4H_LD(tv.mcCuts + 4 * m.imc + kCElCut);
          const double maxETransfer = s.isPositron ? s.ekin : 0.5 * s.ekin;
          if (maxETransfer > elCut) route = s.isPositron ? kQBhabha : kQMoller;
        } else if (iDProc == 1) {
          const double gamCut = G4H_LD(tv.mcCuts + 4 * m.imc + kCGamCut);
          if (s.ekin > gamCut) route = s.ekin < tv.bremModelLim ? kQSB : kQRB;
        } else if (iDProc == 2) {
          route = kQAnnih;
        }
      }
    }
    // kQMoller .. kQAnnih are five consecutive queues
    RouteToQueues<5>(cc, route < 0 ? -1 : route - kQMoller, i, w.queue + kQMoller, w.count + kQMoller);
  }
}

// ---- final state samplers, one kernel per model -----------------------------------------------------------------------
template <int kQueue>
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_QUEUE)
ElSamplerKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                const __grid_constant__ ElectronWork w, const __grid_constant__ G4HB200SecondaryQueue sq, uint64_t seed) {
  const int cnt = w.count[kQueue];
  const int nRound = static_cast<int>(RoundUpToCta(cnt));
  const int stride = gridDim.x * blockDim.x;
  __shared__ CtaCounters<1> cc;
  cc.Init();
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nRound; q += stride) {
    Secondaries sec;
    sec.n = 0;
    int32_t i = 0;
    int id = 0;
    if (q < cnt) {
      i = w.queue[kQueue][q];
      const Meta m = LoadMeta(b.meta, i);
      id = m.id;
      Rng rng;
      rng.Init(seed, static_cast<uint32_t>(m.id), static_cast<uint32_t>(m.draw), false, 0.0);
      if (kQueue == kQAtRest) {
        AnnihilateAtRest(rng, sec);
      } else {
        const Pair e   = LoadPair(b.ekin_logekin, i);
        const Pair dxy = LoadPair(b.dirx_diry, i);
        const Pair dzs = LoadPair(b.dirz_safety, i);
        ElectronState s;
        s.ekin = e.a; s.logEkin = e.b;
        s.dir[0] = dxy.a; s.dir[1] = dxy.b; s.dir[2] = dzs.a;
        s.imc = m.imc; s.id = m.id;
        s.isPositron = (static_cast<uint32_t>(m.flags) & G4HB200_F_POSITRON) != 0u;
        if (kQueue == kQMoller || kQueue == kQBhabha) PerformIoni(tv, s, rng, sec);
        if (kQueue == kQSB) PerformBrem(tv, s, rng, sec, true);
        if (kQueue == kQRB) PerformBrem(tv, s, rng, sec, false);
        if (kQueue == kQAnnih) AnnihilateInFlight(s, rng, sec);
        StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
        StorePair(b.dirx_diry, i, s.dir[0], s.dir[1]);
        StorePair(b.dirz_safety, i, s.dir[2], dzs.b);
      }
      StoreMeta(b.meta, i, Meta{m.imc, m.flags, m.id, static_cast<int>(rng.draw)});
    }
    AppendSecondaries(cc, sq, sec, id, i);
  }
}

// ---- gamma step: head over every track, one sampler kernel per process over its queue --------------------------------
enum GmQueue { kGQConversion = 0, kGQCompton, kGQPhotoelectric, kNumGmQueues };

template <int kMode>
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_QUEUE)
GammaHeadKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200GammaBatch b,
                const __grid_constant__ ElectronWork w, uint64_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nRound = RoundUpToCta(b.n);
  __shared__ CtaCounters<3> cc;
  cc.Init();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    const int route = i < b.n ? StageGammaHead<kMode>(tv, b, i, seed) : -1;
    RouteToQueues<3>(cc, route, static_cast<int32_t>(i), w.queue, w.count);
  }
}

template <int kProc>
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_QUEUE)
GammaInteractKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200GammaBatch b,
                    const __grid_constant__ ElectronWork w, const __grid_constant__ G4HB200SecondaryQueue sq, uint64_t seed) {
  const int cnt = w.count[kProc];
  const int nRound = static_cast<int>(RoundUpToCta(cnt));
  const int stride = gridDim.x * blockDim.x;
  __shared__ CtaCounters<1> cc;
  cc.Init();
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nRound; q += stride) {
    Secondaries sec;
    sec.n = 0;
    int32_t i = 0;
    int id = 0;
    if (q < cnt) {
      i = w.queue[kProc][q];
      StageGammaInteract<kProc>(tv, b, i, seed, sec, id);
    }
    AppendSecondaries(cc, sq, sec, id, i);
  }
}

}  // namespace g4h
#endif
