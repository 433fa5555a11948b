// g4h_pipeline.cuh -- G4HepEmElectronManager::Perform as a pipeline of small kernels over interaction queues.
//
// Why: one thread running the whole Perform of one track is 150 KB of SASS; warps of a CTA sit in different
// branches of it, the instruction cache thrashes (87 % of issue slots stalled on "no instruction",
// profiles/r01_electron_step_monolith.md) and 11 of 32 lanes are active on average.  Here every kernel is a
// few thousand instructions, and a track only enters the kernels it needs:
//
//   ElAlongStepKernel    all tracks: UpdatePStepLength, UpdateNumIALeft, ApplyMeanEnergyLoss (.icc:375-392)
//                                                          -> queues: MSC e- | MSC e+ | fluctuation | discrete | at-rest
//   ElMSCSampleKernel<P> queue of one particle type: SampleMSC (.icc:261-322) in lock step (g4h_perform_stages.cuh)
//                                                          -> queues: fluctuation | discrete | at-rest
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
#include "g4h_msc_f32.cuh"
#include "g4h_perform_stages.cuh"
#include "g4h_refill.cuh"
#include "g4h_stages.cuh"

namespace g4h {

// ---- HowFar in two stages (g4h_stages.cuh) ---------------------------------------------------------------------
#ifndef G4H_MINB_XS
#define G4H_MINB_XS 3
#endif
#ifndef G4H_MINB_MSCLIM
#define G4H_MINB_MSCLIM 3
#endif
#ifndef G4H_MINB_HEAD
#define G4H_MINB_HEAD 3
#endif
#ifndef G4H_MINB_ALONG
#define G4H_MINB_ALONG 3
#endif
// draw windows (uniforms per track pre-generated at full warp width, g4h_rng.cuh): the fluctuation sampler takes
// ~9 draws per track on average, the rejection samplers 3-11
#ifndef G4H_WINDOW_FLUCT
#define G4H_WINDOW_FLUCT 16
#endif
#ifndef G4H_WINDOW_SAMPLER
#define G4H_WINDOW_SAMPLER 8
#endif
#ifndef G4H_MINB_MSC
#define G4H_MINB_MSC 3
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
ElHowFarMSCKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b, uint64_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < b.n; i += stride) {
    StageHowFarMSC<kStoreResults>(tv, b, i, seed);
  }
}

// ---- along-step part for every track (g4h_perform_stages.cuh) -------------------------------------------------------
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_ALONG)
ElAlongStepKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                  const __grid_constant__ ElectronWork w) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nRound = RoundUpToCta(b.n);
  __shared__ CtaCounters<5> cc;
  cc.Init();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    const int route = i < b.n ? StageAlongStep(tv, b, w.prestep, i) : -1;
    // kQMscEl, kQMscPos, kQFluct, kQDiscrete, kQAtRest are queues 0..4
    RouteToQueues<5>(cc, route, static_cast<int32_t>(i), w.queue, w.count);
  }
}

// ---- the fused step: HowFar + along-step part of Perform for every track (g4h_stages.cuh) ---------------------------------
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_HEAD)
ElStepHeadKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                 const __grid_constant__ ElectronWork w, uint64_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nRound = RoundUpToCta(b.n);
  __shared__ CtaCounters<5> cc;
  cc.Init();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    const int route = i < b.n ? StageStepHead(tv, b, w.prestep, i, seed, NoGeometryStep{}) : -1;
    RouteToQueues<5>(cc, route, static_cast<int32_t>(i), w.queue, w.count);
  }
}

// ---- multiple scattering over the queue of one particle type, in lock step ---------------------------------------------
template <bool kPositron>
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_MSC)
ElMSCSampleKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                  const __grid_constant__ ElectronWork w, uint64_t seed) {
  const int cnt = w.count[kPositron ? kQMscPos : kQMscEl];
  const int32_t* queue = w.queue[kPositron ? kQMscPos : kQMscEl];
  const int nRound = static_cast<int>(RoundUpToCta(cnt));
  const int stride = gridDim.x * blockDim.x;
  __shared__ CtaCounters<3> cc;
  cc.Init();
  const double cbeta1 = MscCBeta1();
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nRound; q += stride) {
    int route = -1;
    int32_t i = 0;
    if (q < cnt) {
      i = queue[q];
      route = StageMSCSample<kPositron>(tv, b, w.prestep, i, seed, cbeta1, w.steppre);
    }
    // kQFluct, kQDiscrete, kQAtRest are three consecutive queues
    RouteToQueues<3>(cc, route < 0 ? -1 : route - kQFluct, i, w.queue + kQFluct, w.count + kQFluct);
  }
}

// the offered single-precision variant (g4h_msc_f32.cuh; g4hb200_set_msc_precision(h, 32))
template <bool kPositron>
__global__ void __launch_bounds__(kThreadsPerBlock, 4)
ElMSCSampleF32Kernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                     const __grid_constant__ ElectronWork w, uint64_t seed) {
  const int cnt = w.count[kPositron ? kQMscPos : kQMscEl];
  const int32_t* queue = w.queue[kPositron ? kQMscPos : kQMscEl];
  const int nRound = static_cast<int>(RoundUpToCta(cnt));
  const int stride = gridDim.x * blockDim.x;
  __shared__ CtaCounters<3> cc;
  cc.Init();
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nRound; q += stride) {
    int route = -1;
    int32_t i = 0;
    if (q < cnt) {
      i = queue[q];
      route = StageMSCSampleF32<kPositron>(tv, b, w.prestep, i, seed, w.steppre);
    }
    RouteToQueues<3>(cc, route < 0 ? -1 : route - kQFluct, i, w.queue + kQFluct, w.count + kQFluct);
  }
}

// ---- energy loss fluctuation over its queue ---------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_QUEUE)
ElFluctuationKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                    const __grid_constant__ ElectronWork w, uint64_t seed) {
  const int cnt = w.count[kQFluct];
  const int nRound = static_cast<int>(RoundUpToCta(cnt));
  const int stride = gridDim.x * blockDim.x;
  __shared__ CtaCounters<6> cc;
  __shared__ double window[G4H_WINDOW_FLUCT * kThreadsPerBlock];
  cc.Init();
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nRound; q += stride) {
    int route = -1;
    int32_t i = 0;
    if (q < cnt) {
      i = w.queue[kQFluct][q];
      route = StageFluctuation(tv, b, w.prestep, i, seed, window + threadIdx.x, kThreadsPerBlock, G4H_WINDOW_FLUCT);
    }
    // kQAtRest, kQMoller .. kQAnnih are six consecutive queues
    RouteToQueues<6>(cc, route < 0 ? -1 : route - kQAtRest, i, w.queue + kQAtRest, w.count + kQAtRest);
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
      route = StageDiscrete(tv, b, i, seed);
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
  constexpr uint32_t kSlots = kQueue == kQAtRest ? 2 : (kQueue == kQBhabha ? 2 * G4H_WINDOW_SAMPLER : G4H_WINDOW_SAMPLER);
  __shared__ double window[kSlots * kThreadsPerBlock];
  cc.Init();
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nRound; q += stride) {
    Secondaries sec;
    sec.n = 0;
    int32_t i = 0;
    int id = 0;
    if (q < cnt) {
      i = w.queue[kQueue][q];
      StageSampler<kQueue>(tv, b, i, seed, sec, id, window + threadIdx.x, kThreadsPerBlock, kSlots);
    }
    AppendSecondaries(cc, sq, sec, id, i);
  }
}

// ---- gamma step: head over every track, one sampler kernel per process over its queue --------------------------------
enum GmQueue { kGQConversion = 0, kGQCompton, kGQPhotoelectric, kNumGmQueues };

template <int kMode>
__global__ void __launch_bounds__(kGammaThreads, 2 * G4H_MINB_QUEUE)
GammaHeadKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200GammaBatch b,
                const __grid_constant__ ElectronWork w, uint64_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nRound = RoundUpToCta(b.n);
  __shared__ CtaCounters<3> cc;
  cc.Init();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    const int route = i < b.n ? StageGammaHead<kMode>(tv, b, i, seed, NoGeometryStep{}) : -1;
    RouteToQueues<3>(cc, route, static_cast<int32_t>(i), w.queue, w.count);
  }
}

template <int kProc>
__global__ void __launch_bounds__(kGammaThreads, 2 * G4H_MINB_QUEUE)
GammaInteractKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200GammaBatch b,
                    const __grid_constant__ ElectronWork w, const __grid_constant__ G4HB200SecondaryQueue sq, uint64_t seed) {
  const int cnt = w.count[kProc];
  const int nRound = static_cast<int>(RoundUpToCta(cnt));
  const int stride = gridDim.x * blockDim.x;
  __shared__ CtaCounters<1> cc;
  // draws per track (tools, reference run): conversion 11-12 (p90), Compton 4-7, photoelectric 1 (median; most
  // photons are absorbed below the K edge without an electron) .. 9 (p90)
  constexpr uint32_t kSlots = kProc == kGQCompton ? G4H_WINDOW_SAMPLER : (kProc == kGQConversion ? 2 * G4H_WINDOW_SAMPLER : 4);
  __shared__ double window[kSlots * kGammaThreads];
  cc.Init();
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nRound; q += stride) {
    Secondaries sec;
    sec.n = 0;
    int32_t i = 0;
    int id = 0;
    if (q < cnt) {
      i = w.queue[kProc][q];
      StageGammaInteract<kProc>(tv, b, i, seed, sec, id, window + threadIdx.x, kGammaThreads, kSlots);
    }
    AppendSecondaries(cc, sq, sec, id, i);
  }
}

// ---- the rejection samplers with lane refill (g4h_refill.cuh): one warp = one stream of queue entries -----------------------
template <int kQueue> struct ElRefillSampler;
template <> struct ElRefillSampler<kQMoller> { using type = MollerSampler; };
template <> struct ElRefillSampler<kQBhabha> { using type = BhabhaSampler; };
template <> struct ElRefillSampler<kQSB>     { using type = SBSampler; };
template <> struct ElRefillSampler<kQRB>     { using type = RBSampler; };
template <int kProc> struct GammaRefillSampler;
template <> struct GammaRefillSampler<kGQConversion>    { using type = ConversionSampler; };
template <> struct GammaRefillSampler<kGQCompton>       { using type = ComptonSampler; };
template <> struct GammaRefillSampler<kGQPhotoelectric> { using type = PhotoelectricSampler; };

// dynamic shared memory of a refill kernel: one RefillWarpStore per warp
template <class S>
constexpr size_t RefillSmemBytes() { return sizeof(RefillWarpStore<S>) * kWarpsPerBlock; }

template <int kQueue>
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_QUEUE)
ElRefillSamplerKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
                      const __grid_constant__ ElectronWork w, const __grid_constant__ G4HB200SecondaryQueue sq, uint64_t seed,
                      int chunksPerWarp) {
  using S = typename ElRefillSampler<kQueue>::type;
  extern __shared__ __align__(16) unsigned char refillSmem[];
  RefillWarpStore<S>* store = reinterpret_cast<RefillWarpStore<S>*>(refillSmem);
  RefillSamplerWarp<S, ElectronSamplerIO>(tv, b, w.queue[kQueue], w.count[kQueue], sq, seed, chunksPerWarp, store[threadIdx.x >> 5]);
}

template <int kProc>
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_QUEUE)
GammaRefillKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200GammaBatch b,
                  const __grid_constant__ ElectronWork w, const __grid_constant__ G4HB200SecondaryQueue sq, uint64_t seed,
                  int chunksPerWarp) {
  using S = typename GammaRefillSampler<kProc>::type;
  extern __shared__ __align__(16) unsigned char refillSmem[];
  RefillWarpStore<S>* store = reinterpret_cast<RefillWarpStore<S>*>(refillSmem);
  RefillSamplerWarp<S, GammaSamplerIO>(tv, b, w.queue[kProc], w.count[kProc], sq, seed, chunksPerWarp, store[threadIdx.x >> 5]);
}

}  // namespace g4h
#endif
