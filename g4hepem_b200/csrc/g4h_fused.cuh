// g4h_fused.cuh -- the whole e-/e+ (or gamma) step as ONE persistent launch with CTA-local interaction queues.
//
// Round 1 ran the step as a pipeline of 11 launches over global queues (g4h_pipeline.cuh): every stage streamed the
// track state in from HBM and back out (1 145 MB per 1M-track step against 340 MB algorithmic, profiles/r01c_*), paid
// a grid-wide ramp and tail per launch, and the small final-state queues left most of a one-wave grid idle.  Here a
// CTA owns tiles c, c+G, c+2G ... of the batch (G CTAs, one wave) and keeps its own queues of track indices in shared
// memory.  It is a small scheduler: as soon as any queue holds a full CTA's worth of entries (256) the CTA runs
// that stage on exactly 256 tracks -- every lane busy, all warps of the CTA in the same code -- otherwise it
// takes its next tile through the head stage; at the end it drains what is left, upstream queues first.
//
//   head (all tracks of a tile)  -> MSC e- | MSC e+ | fluctuation | discrete | at-rest
//   MSC e- / e+                  -> fluctuation | discrete | at-rest
//   fluctuation, discrete        -> at-rest | Moller | Bhabha | Seltzer-Berger | rel. brem | annihilation
//   final-state samplers         -> secondaries (CTA-aggregated append to the global secondary queue)
//
// The stage functions are those of g4h_stages.cuh / g4h_perform_stages.cuh, unchanged: the state of a track is handed
// from stage to stage through its slots in the batch, which a CTA re-reads within a few microseconds of writing
// them -- out of L2, not HBM (a CTA's in-flight tracks are ~0.3 MB, all CTAs' ~130 MB of traffic per "round" against
// 126 MB of L2; measured DRAM traffic: profiles/r02_*).  Shared memory stays small (queues + draw windows: 42 KB per
// CTA) so that L1 keeps the tables: the stage kernels gather ~40 table words per track and their L1 hit rate (82 %)
// is what an smem-resident track tile (260 B per track + 128 B of draw window) would have given up.
//
// Deadlock / overflow freedom: a stage run appends at most 256 entries to any downstream queue and only runs when
// every queue downstream of it holds fewer than 256, so no queue ever exceeds 511 entries (capacity 512).
#ifndef G4H_FUSED_CUH
#define G4H_FUSED_CUH

#include "g4h_kernels.cuh"
#include "g4h_pipeline.cuh"

namespace g4h {

constexpr int kFusedListCap = 2 * kThreadsPerBlock;

#ifndef G4H_MINB_FUSED
#define G4H_MINB_FUSED 3
#endif
// draw-window slots per thread of the fused kernels (doubles): the fluctuation sampler and the Bhabha / conversion
// samplers use all of them, the other samplers the first G4H_WINDOW_SAMPLER
constexpr int kFusedWindowSlots = G4H_WINDOW_FLUCT > 2 * G4H_WINDOW_SAMPLER ? G4H_WINDOW_FLUCT : 2 * G4H_WINDOW_SAMPLER;

struct FusedShared {
  uint16_t list[kNumElQueues][kFusedListCap];  // entry = (ordinal of the CTA's tile) * 256 + index inside the tile
  int cnt[kNumElQueues];
  int nextTile;                                // ordinal of the next tile the head stage takes
  CtaCounters<1> cc;                           // CTA-aggregated append of secondaries
  double window[kFusedWindowSlots * kThreadsPerBlock];
};

// most tiles one CTA can take: the entry code is 16 bits
constexpr int kFusedMaxTilesPerCta = 65536 / kThreadsPerBlock;

__device__ __forceinline__ int64_t FusedTrackIndex(uint32_t entry) {
  return (static_cast<int64_t>(entry >> 8) * gridDim.x + blockIdx.x) * kThreadsPerBlock + (entry & 255u);
}

// append `entry` to queue `route` (in [kLo, kLo + K)) or to none (route outside): one shared-memory atomic per warp and queue
template <int kLo, int K>
__device__ __forceinline__ void FusedRoute(FusedShared& sh, int route, uint32_t entry) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const unsigned m = __ballot_sync(0xffffffffu, route == kLo + k);
    if (m != 0u) {
      const int leader = __ffs(m) - 1;
      int wb = 0;
      if (static_cast<int>(lane) == leader) wb = atomicAdd(&sh.cnt[kLo + k], __popc(m));
      wb = __shfl_sync(0xffffffffu, wb, leader);
      if (route == kLo + k) sh.list[kLo + k][wb + __popc(m & ((1u << lane) - 1u))] = static_cast<uint16_t>(entry);
    }
  }
}

// take up to 256 entries off the top of queue q (one per thread); true if this thread got one
__device__ __forceinline__ bool FusedPop(FusedShared& sh, int q, uint32_t& entry) {
  const int cnt   = sh.cnt[q];
  const int take  = cnt < kThreadsPerBlock ? cnt : kThreadsPerBlock;
  const int first = cnt - take;
  const bool has  = static_cast<int>(threadIdx.x) < take;
  entry = has ? sh.list[q][first + threadIdx.x] : 0u;
  __syncthreads();  // everybody has read the count and its entry
  if (threadIdx.x == 0) sh.cnt[q] = first;
  return has;
}

// What the CTA does next (the same value in every thread: computed from shared memory after a barrier).
//   >= 0: run the stage of that queue; kFusedHead: next tile through the head; kFusedDone: nothing left
constexpr int kFusedHead = -1, kFusedDone = -2;
template <int kNumQueues>
__device__ __forceinline__ int FusedChoose(const FusedShared& sh, int tilesOfCta) {
  // a full CTA's worth waiting somewhere: the most downstream such queue first (keeps every queue below 512)
#pragma unroll
  for (int q = kNumQueues - 1; q >= 0; --q) {
    if (sh.cnt[q] >= kThreadsPerBlock) return q;
  }
  if (sh.nextTile < tilesOfCta) return kFusedHead;
  // the tail: drain upstream queues first (their output tops up the queues downstream)
#pragma unroll
  for (int q = 0; q < kNumQueues; ++q) {
    if (sh.cnt[q] > 0) return q;
  }
  return kFusedDone;
}

// number of tiles CTA c of a G-CTA grid takes out of numTiles (tiles c, c+G, ...)
__device__ __forceinline__ int FusedTilesOfCta(int64_t n) {
  const int64_t numTiles = (n + kThreadsPerBlock - 1) / kThreadsPerBlock;
  const int64_t c = blockIdx.x, g = gridDim.x;
  return c < numTiles ? static_cast<int>((numTiles - c + g - 1) / g) : 0;
}

// ---- e-/e+ ----------------------------------------------------------------------------------------------------------
// Geometry: NoGeometryStep (g4hb200_electron_step) or the slab step of the stepping loop (g4h_shower.cuh)
// kHead: kHeadStep = HowFar + along-step (g4hb200_electron_step); kHeadPerform = HowFar ran as its own call
// (g4hb200_electron_perform): the along-step stage alone; kHeadLoop = the head of the stepping loop with the geometry step and the
// MSC sub-step accounting inside (g4h_shower.cuh: StageLoopHead; steppre: its second workspace)
enum FusedHead { kHeadStep = 0, kHeadPerform, kHeadLoop };
template <class GeometryStep>
G4H_FN int StageLoopHead(const TablesView& tv, const G4HB200ElectronBatch& b, double* prestep, double* steppre, int64_t i,
                         uint64_t seed, const GeometryStep& geometry);

template <int kHead, class MakeGeometry>
__device__ __forceinline__ void ElFusedBody(const TablesView& tv, const G4HB200ElectronBatch& b, double* prestep, double* steppre,
                                            const G4HB200SecondaryQueue& sq, uint64_t seed, const MakeGeometry& makeGeometry) {
  __shared__ FusedShared sh;
  if (threadIdx.x < kNumElQueues) sh.cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) sh.nextTile = 0;
  sh.cc.Init();  // contains the barrier
  const int tilesOfCta = FusedTilesOfCta(b.n);
  const double cbeta1  = MscCBeta1();
  double* window = sh.window + threadIdx.x;
  for (;;) {
    const int action = FusedChoose<kNumElQueues>(sh, tilesOfCta);
    if (action == kFusedDone) break;
    uint32_t entry = 0u;
    bool has = false;
    if (action == kFusedHead) {
      const int k = sh.nextTile;
      __syncthreads();
      if (threadIdx.x == 0) sh.nextTile = k + 1;
      entry = static_cast<uint32_t>(k) * kThreadsPerBlock + threadIdx.x;
      has   = FusedTrackIndex(entry) < b.n;
    } else {
      has = FusedPop(sh, action, entry);
    }
    const int64_t i = FusedTrackIndex(entry);
    switch (action) {
      case kFusedHead: {
        int route = -1;
        if (has) {
          if constexpr (kHead == kHeadPerform) route = StageAlongStep(tv, b, prestep, i);
          if constexpr (kHead == kHeadStep) route = StageStepHead(tv, b, prestep, i, seed, makeGeometry());
          if constexpr (kHead == kHeadLoop) route = StageLoopHead(tv, b, prestep, steppre, i, seed, makeGeometry());
        }
        FusedRoute<kQMscEl, 5>(sh, route, entry);
        break;
      }
      case kQMscEl: {
        const int route = has ? StageMSCSample<false>(tv, b, prestep, i, seed, cbeta1, steppre) : -1;
        FusedRoute<kQFluct, 3>(sh, route, entry);
        break;
      }
      case kQMscPos: {
        const int route = has ? StageMSCSample<true>(tv, b, prestep, i, seed, cbeta1, steppre) : -1;
        FusedRoute<kQFluct, 3>(sh, route, entry);
        break;
      }
      case kQFluct: {
        const int route = has ? StageFluctuation(tv, b, prestep, i, seed, window, kThreadsPerBlock, G4H_WINDOW_FLUCT) : -1;
        FusedRoute<kQAtRest, 6>(sh, route, entry);
        break;
      }
      case kQDiscrete: {
        const int route = has ? StageDiscrete(tv, b, i, seed) : -1;
        FusedRoute<kQMoller, 5>(sh, route, entry);
        break;
      }
      default: {
        Secondaries sec;
        sec.n  = 0;
        int id = 0;
        if (has) {
          switch (action) {
            case kQAtRest: StageSampler<kQAtRest>(tv, b, i, seed, sec, id, window, kThreadsPerBlock, 2); break;
            case kQMoller: StageSampler<kQMoller>(tv, b, i, seed, sec, id, window, kThreadsPerBlock, G4H_WINDOW_SAMPLER); break;
            case kQBhabha: StageSampler<kQBhabha>(tv, b, i, seed, sec, id, window, kThreadsPerBlock, 2 * G4H_WINDOW_SAMPLER); break;
            case kQSB: StageSampler<kQSB>(tv, b, i, seed, sec, id, window, kThreadsPerBlock, G4H_WINDOW_SAMPLER); break;
            case kQRB: StageSampler<kQRB>(tv, b, i, seed, sec, id, window, kThreadsPerBlock, G4H_WINDOW_SAMPLER); break;
            default: StageSampler<kQAnnih>(tv, b, i, seed, sec, id, window, kThreadsPerBlock, G4H_WINDOW_SAMPLER); break;
          }
        }
        AppendSecondaries(sh.cc, sq, sec, id, i);
        break;
      }
    }
    __syncthreads();  // queue counts and the track state written by this stage are visible to the whole CTA
  }
}

struct MakeNoGeometry {
  __device__ __forceinline__ NoGeometryStep operator()() const { return NoGeometryStep{}; }
};

template <bool kPerformOnly>
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_FUSED)
ElFusedStepKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b, double* prestep,
                  const __grid_constant__ G4HB200SecondaryQueue sq, uint64_t seed) {
  ElFusedBody<kPerformOnly ? kHeadPerform : kHeadStep>(tv, b, prestep, nullptr, sq, seed, MakeNoGeometry{});
}

// ---- gamma ----------------------------------------------------------------------------------------------------------
// kMode 1: SelectInteraction + Perform, 2: HowFar first (the fused step)
template <int kMode, class MakeGeometry>
__device__ __forceinline__ void GammaFusedBody(const TablesView& tv, const G4HB200GammaBatch& b, const G4HB200SecondaryQueue& sq,
                                               uint64_t seed, const MakeGeometry& makeGeometry) {
  __shared__ FusedShared sh;
  if (threadIdx.x < kNumElQueues) sh.cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) sh.nextTile = 0;
  sh.cc.Init();
  const int tilesOfCta = FusedTilesOfCta(b.n);
  double* window = sh.window + threadIdx.x;
  for (;;) {
    const int action = FusedChoose<kNumGmQueues>(sh, tilesOfCta);
    if (action == kFusedDone) break;
    uint32_t entry = 0u;
    bool has = false;
    if (action == kFusedHead) {
      const int k = sh.nextTile;
      __syncthreads();
      if (threadIdx.x == 0) sh.nextTile = k + 1;
      entry = static_cast<uint32_t>(k) * kThreadsPerBlock + threadIdx.x;
      has   = FusedTrackIndex(entry) < b.n;
    } else {
      has = FusedPop(sh, action, entry);
    }
    const int64_t i = FusedTrackIndex(entry);
    if (action == kFusedHead) {
      const int route = has ? StageGammaHead<kMode>(tv, b, i, seed, makeGeometry()) : -1;
      FusedRoute<kGQConversion, 3>(sh, route, entry);
    } else {
      Secondaries sec;
      sec.n  = 0;
      int id = 0;
      if (has) {
        switch (action) {
          case kGQConversion: StageGammaInteract<kGQConversion>(tv, b, i, seed, sec, id, window, kThreadsPerBlock, 2 * G4H_WINDOW_SAMPLER); break;
          case kGQCompton: StageGammaInteract<kGQCompton>(tv, b, i, seed, sec, id, window, kThreadsPerBlock, G4H_WINDOW_SAMPLER); break;
          default: StageGammaInteract<kGQPhotoelectric>(tv, b, i, seed, sec, id, window, kThreadsPerBlock, 4); break;
        }
      }
      AppendSecondaries(sh.cc, sq, sec, id, i);
    }
    __syncthreads();
  }
}

template <int kMode>
__global__ void __launch_bounds__(kThreadsPerBlock, G4H_MINB_FUSED)
GammaFusedStepKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200GammaBatch b,
                     const __grid_constant__ G4HB200SecondaryQueue sq, uint64_t seed) {
  GammaFusedBody<kMode>(tv, b, sq, seed, MakeNoGeometry{});
}

}  // namespace g4h
#endif
