// g4h_refill.cuh -- the rejection samplers a warp at a time, with lane refill.
//
// One thread per track through a rejection loop runs at the width of the slowest lane: Bhabha (4.8 passes on
// average) ran at 12.6 of 32 lanes, the photoelectric sampler at 8.4 (most photons end without an electron), the
// others at 15-24 (profiles/r02_pipeline_full.md).  Here a warp owns a stream of queue entries and keeps three
// populations of them in shared memory, each served 32 at a time:
//
//   fresh queue entries --Setup (32 lanes)--> pending ring --Trial (32 lanes)--> finished ring --Finish (32 lanes)--> batch
//                                                 ^  rejected  |
//                                                 +------------+
//
// A trial is ONE pass of the rejection loop for 32 pending entries, whichever pass it is for each of them; the
// rejected go back to the ring, the accepted move on.  Entries that need no sampling at all never enter a ring.
// The uniforms of a pass are generated inside the pass by every lane (one or two Philox blocks), so no uniform is
// generated that is not consumed.  What travels through shared memory per entry: the sampler's Pars (7-11 doubles),
// track index, track id, draw counter and the unused half of the last Philox block.
//
// Results do not depend on any of this: a track's uniforms are a function of (seed, track id, draw index), and its
// passes see them in the order of the reference's loop.
//
// Measured (profiles/r02b_refill_ab.log, r02b_variants_full.md): lanes go up (Bhabha 12.7 -> 24.6, Moller 17.9 -> 26.3) and
// every kernel but the photoelectric one gets SLOWER -- these kernels are bound by the latency of one warp's instruction
// stream, a pass through the rings costs more than the lanes it fills, and Setup + Finish outweigh the loop.  The executor
// is therefore opt-in (G4HB200_REFILL=k); the default runs RunSampler one thread per track (g4h_pipeline.cuh).
#ifndef G4H_REFILL_CUH
#define G4H_REFILL_CUH

#include "g4h_kernels.cuh"
#include "g4h_perform_stages.cuh"

namespace g4h {

constexpr int kRefillSlots   = 64;  // per warp; pending + finished never exceed it
constexpr int kWarpsPerBlock = kThreadsPerBlock / 32;

template <class S>
struct RefillWarpStore {
  static constexpr int kNumPars = static_cast<int>(sizeof(typename S::Pars) / sizeof(double));
  double par[kNumPars][kRefillSlots];
  double spare[kRefillSlots];   // uniform number `draw` when draw is odd
  int32_t track[kRefillSlots];
  uint32_t id[kRefillSlots];
  uint32_t draw[kRefillSlots];
  uint8_t pend[kRefillSlots];   // ring of slot numbers
  uint8_t fin[kRefillSlots];    // ring of slot numbers
  uint8_t freeStack[kRefillSlots];
};

template <class Pars>
__device__ __forceinline__ void StorePars(double (*par)[kRefillSlots], int slot, const Pars& p) {
  constexpr int n = static_cast<int>(sizeof(Pars) / sizeof(double));
  const double* src = reinterpret_cast<const double*>(&p);
#pragma unroll
  for (int k = 0; k < n; ++k) par[k][slot] = src[k];
}
// the trailing kResults fields only (what a trial writes)
template <int kResults, class Pars>
__device__ __forceinline__ void StoreResults(double (*par)[kRefillSlots], int slot, const Pars& p) {
  constexpr int n = static_cast<int>(sizeof(Pars) / sizeof(double));
  const double* src = reinterpret_cast<const double*>(&p);
#pragma unroll
  for (int k = n - kResults; k < n; ++k) par[k][slot] = src[k];
}
template <class Pars>
__device__ __forceinline__ void LoadPars(const double (*par)[kRefillSlots], int slot, Pars& p) {
  constexpr int n = static_cast<int>(sizeof(Pars) / sizeof(double));
  double* dst = reinterpret_cast<double*>(&p);
#pragma unroll
  for (int k = 0; k < n; ++k) dst[k] = par[k][slot];
}

// the uniforms of one pass: draws d .. d + kDraws - 1 of the track; spare (in: uniform d when d is odd; out: uniform
// d + kDraws when that is odd) saves the half block a pass leaves over
template <int kDraws>
__device__ __forceinline__ void TrialUniforms(uint32_t k0, uint32_t k1, uint32_t id, uint32_t d, double& spare, double* u) {
  const bool odd = (d & 1u) != 0u;
  const uint32_t blk = (d + 1u) >> 1;
  const Philox4 a = PhiloxBlockInl(k0, k1, id, blk);
  const double e0 = ToUniform(a.x, a.y), e1 = ToUniform(a.z, a.w);
  if (kDraws == 2) {
    u[0] = odd ? spare : e0;
    u[1] = odd ? e0 : e1;
    spare = e1;
  } else {
    double f0 = 0.0, f1 = 0.0;
    if (!odd) {
      const Philox4 b = PhiloxBlock(k0, k1, id, blk + 1u);
      f0 = ToUniform(b.x, b.y);
      f1 = ToUniform(b.z, b.w);
    }
    u[0] = odd ? spare : e0;
    u[1] = odd ? e0 : e1;
    u[2] = odd ? e1 : f0;
    spare = f1;
  }
}

// ---- track <-> sampler state, e-/e+ ---------------------------------------------------------------------------------------
struct ElectronSamplerIO {
  using Batch = G4HB200ElectronBatch;
  using Track = ElectronState;
  struct Keep {  // what Finish stores back unchanged
    Meta m;
    double safety;
  };
  __device__ __forceinline__ static void LoadForSetup(const Batch& b, int64_t i, Track& s, Keep& k) {
    k.m = LoadMeta(b.meta, i);
    const Pair e = LoadPair(b.ekin_logekin, i);
    s.ekin = e.a; s.logEkin = e.b;
    s.imc = k.m.imc; s.id = k.m.id;
    s.isPositron = (static_cast<uint32_t>(k.m.flags) & G4HB200_F_POSITRON) != 0u;
  }
  // nothing was sampled (the cut is above the maximum transfer): what the per-track stage stored in that case
  __device__ __forceinline__ static void StoreDone(const TablesView&, const Batch& b, int64_t i, Track& s, const Keep& k, uint32_t draw) {
    StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
    StoreMeta(b.meta, i, Meta{k.m.imc, k.m.flags, k.m.id, static_cast<int>(draw)});
  }
  __device__ __forceinline__ static void LoadForFinish(const Batch& b, int64_t i, Track& s, Keep& k) {
    LoadForSetup(b, i, s, k);
    const Pair dxy = LoadPair(b.dirx_diry, i);
    const Pair dzs = LoadPair(b.dirz_safety, i);
    s.dir[0] = dxy.a; s.dir[1] = dxy.b; s.dir[2] = dzs.a;
    k.safety = dzs.b;
  }
  __device__ __forceinline__ static void StoreFinish(const TablesView&, const Batch& b, int64_t i, Track& s, const Keep& k, uint32_t draw) {
    StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
    StorePair(b.dirx_diry, i, s.dir[0], s.dir[1]);
    StorePair(b.dirz_safety, i, s.dir[2], k.safety);
    StoreMeta(b.meta, i, Meta{k.m.imc, k.m.flags, k.m.id, static_cast<int>(draw)});
  }
};

// ---- track <-> sampler state, gamma (+ the tracking cut behind the interaction, G4HepEmGammaManager.icc:88-93) ----------------
struct GammaSamplerIO {
  using Batch = G4HB200GammaBatch;
  using Track = GammaState;
  struct Keep {
    Meta m;
    double nia0;
  };
  __device__ __forceinline__ static void LoadForSetup(const Batch& b, int64_t i, Track& s, Keep& k) {
    k.m = LoadMeta(b.meta, i);
    const Pair e  = LoadPair(b.ekin_logekin, i);
    const Pair ep = LoadPair(b.edep_pemxsec, i);
    s.ekin = e.a; s.logEkin = e.b;
    s.imc = k.m.imc; s.id = k.m.id;
    s.edep = ep.a; s.peMXsec = ep.b;
  }
  __device__ __forceinline__ static void TrackingCut(const TablesView& tv, Track& s) {
    const double finalEkin = s.ekin;
    if (finalEkin > 0.0 && finalEkin <= tv.gammaTrackingCut) {
      SetEKin(s, 0.0);
      s.edep += finalEkin;
    }
  }
  __device__ __forceinline__ static void StoreDone(const TablesView& tv, const Batch& b, int64_t i, Track& s, const Keep& k, uint32_t draw) {
    TrackingCut(tv, s);
    StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
    StorePair(b.edep_pemxsec, i, s.edep, s.peMXsec);
    StoreMeta(b.meta, i, Meta{k.m.imc, k.m.flags, k.m.id, static_cast<int>(draw)});
  }
  __device__ __forceinline__ static void LoadForFinish(const Batch& b, int64_t i, Track& s, Keep& k) {
    LoadForSetup(b, i, s, k);
    const Pair dxy = LoadPair(b.dirx_diry, i);
    const Pair dzn = LoadPair(b.dirz_nia0, i);
    s.dir[0] = dxy.a; s.dir[1] = dxy.b; s.dir[2] = dzn.a;
    k.nia0 = dzn.b;
  }
  __device__ __forceinline__ static void StoreFinish(const TablesView& tv, const Batch& b, int64_t i, Track& s, const Keep& k, uint32_t draw) {
    TrackingCut(tv, s);
    StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
    StorePair(b.dirx_diry, i, s.dir[0], s.dir[1]);
    StorePair(b.dirz_nia0, i, s.dir[2], k.nia0);
    StorePair(b.edep_pemxsec, i, s.edep, s.peMXsec);
    StoreMeta(b.meta, i, Meta{k.m.imc, k.m.flags, k.m.id, static_cast<int>(draw)});
  }
};

// secondaries of up to 32 tracks: one atomicAdd per warp; every lane of the warp calls (sec.n = 0 for idle lanes)
__device__ __forceinline__ void AppendSecondariesWarp(const G4HB200SecondaryQueue& q, const Secondaries& sec, int parentId,
                                                      int64_t parentIndex) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned has1 = __ballot_sync(full, sec.n >= 1);
  const unsigned has2 = __ballot_sync(full, sec.n >= 2);
  const int total = __popc(has1) + __popc(has2);
  if (total == 0) return;
  const unsigned below = (1u << lane) - 1u;
  const int excl = __popc(has1 & below) + __popc(has2 & below);
  int warpBase = 0;
  if (lane == 0) warpBase = atomicAdd(q.count, total);
  warpBase = __shfl_sync(full, warpBase, 0);
  const int64_t base = static_cast<int64_t>(warpBase) + excl;
  for (int k = 0; k < sec.n; ++k) {
    const int64_t slot = base + k;
    if (slot < q.capacity) {
      reinterpret_cast<double2*>(q.dirx_diry)[slot] = make_double2(sec.s[k].dir[0], sec.s[k].dir[1]);
      reinterpret_cast<double2*>(q.dirz_ekin)[slot] = make_double2(sec.s[k].dir[2], sec.s[k].ekin);
      reinterpret_cast<int2*>(q.parent_kind)[slot]  = make_int2(parentId, sec.s[k].kind);
      reinterpret_cast<int2*>(q.parent_slot)[slot]  = make_int2(static_cast<int>(parentIndex) + q.parent_base, k);
    }
  }
}

// ---- the executor: one warp, its share of a queue -----------------------------------------------------------------------
// chunksPerWarp: how many 32-entry chunks a warp should get at least (refill needs a supply); short queues are
// therefore served by fewer warps, spread over the CTAs of the grid
template <class S, class IO>
__device__ __forceinline__ void RefillSamplerWarp(const TablesView& tv, const typename IO::Batch& b, const int32_t* __restrict__ queue,
                                                  int cnt, const G4HB200SecondaryQueue& sq, uint64_t seed, int chunksPerWarp,
                                                  RefillWarpStore<S>& st) {
  using Pars  = typename S::Pars;
  using Track = typename IO::Track;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  // worker number: warp w of CTA c is worker w * gridDim.x + c, so that few workers still spread over all SMs
  const int numChunks  = (cnt + 31) >> 5;
  const int allWarps   = static_cast<int>(gridDim.x) * kWarpsPerBlock;
  int workers = numChunks / (chunksPerWarp > 0 ? chunksPerWarp : 1);
  workers = workers < 1 ? 1 : (workers > allWarps ? allWarps : workers);
  const int worker = static_cast<int>(threadIdx.x >> 5) * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x);
  if (worker >= workers) return;
  int chunk = worker;
  bool exhausted = chunk >= numChunks;
  int nPend = 0, pHead = 0, nFin = 0, fHead = 0, nFree = kRefillSlots;
  st.freeStack[lane]      = static_cast<uint8_t>(lane);
  st.freeStack[lane + 32] = static_cast<uint8_t>(lane + 32);
  __syncwarp();
  for (;;) {
    int action;  // 0 top up, 1 trial, 2 finish, 3 leave
    int m = 0;
    if (nFin >= 32) {
      action = 2; m = 32;
    } else if (nPend < 32 && !exhausted && nFree >= 32) {
      action = 0;
    } else if (nPend < 32 && !exhausted && nFin >= nPend) {
      action = 2; m = nFin;
    } else if (nPend > 0) {
      action = 1; m = nPend < 32 ? nPend : 32;
    } else if (nFin > 0) {
      action = 2; m = nFin;
    } else {
      action = 3;
    }
    if (action == 3) break;
    if (action == 0) {
      // ---- Setup for the next 32 queue entries
      const int q = chunk * 32 + lane;
      chunk += workers;
      exhausted = chunk >= numChunks;
      int next = kSamplerDone;
      Pars p;
      int32_t i = 0;
      uint32_t draw = 0u, id = 0u;
      double spare = 0.0;
      if (q < cnt) {
        i = queue[q];
        Track s;
        typename IO::Keep keep;
        IO::LoadForSetup(b, i, s, keep);
        Rng rng;
        rng.Init(seed, static_cast<uint32_t>(keep.m.id), static_cast<uint32_t>(keep.m.draw), false, 0.0);
        next = S::Setup(tv, s, rng, p);
        draw = rng.draw;
        id   = static_cast<uint32_t>(keep.m.id);
        if (next == kSamplerDone) {
          IO::StoreDone(tv, b, i, s, keep, draw);
        } else if ((draw & 1u) != 0u) {
          spare = rng.hasNext ? rng.next : UniformPair(k0, k1, id, draw >> 1).b;
        }
      }
      const unsigned toPend = __ballot_sync(full, next == kSamplerLoop);
      const unsigned toFin  = __ballot_sync(full, next == kSamplerFinish);
      const unsigned taking = toPend | toFin;
      if (next != kSamplerDone) {
        const int slot = st.freeStack[nFree - 1 - __popc(taking & below)];
        StorePars(st.par, slot, p);
        st.spare[slot] = spare;
        st.track[slot] = i;
        st.id[slot]    = id;
        st.draw[slot]  = draw;
        if (next == kSamplerLoop) {
          st.pend[(pHead + nPend + __popc(toPend & below)) & (kRefillSlots - 1)] = static_cast<uint8_t>(slot);
        } else {
          st.fin[(fHead + nFin + __popc(toFin & below)) & (kRefillSlots - 1)] = static_cast<uint8_t>(slot);
        }
      }
      nFree -= __popc(taking);
      nPend += __popc(toPend);
      nFin  += __popc(toFin);
      __syncwarp();
    } else if (action == 1) {
      // ---- one pass of the rejection loop for m pending entries
      bool accepted = false;
      int slot = 0;
      if (lane < m) {
        slot = st.pend[(pHead + lane) & (kRefillSlots - 1)];
        Pars p;
        LoadPars(st.par, slot, p);
        const uint32_t d = st.draw[slot];
        double spare = st.spare[slot];
        double u[3];
        TrialUniforms<S::kDraws>(k0, k1, st.id[slot], d, spare, u);
        accepted = S::Trial(tv, p, u);
        StoreResults<S::kNumResults>(st.par, slot, p);
        st.draw[slot]  = d + static_cast<uint32_t>(S::kDraws);
        st.spare[slot] = spare;
      }
      __syncwarp();
      pHead = (pHead + m) & (kRefillSlots - 1);
      nPend -= m;
      const unsigned acc = __ballot_sync(full, lane < m && accepted);
      const unsigned rej = __ballot_sync(full, lane < m && !accepted);
      if (lane < m) {
        if (accepted) {
          st.fin[(fHead + nFin + __popc(acc & below)) & (kRefillSlots - 1)] = static_cast<uint8_t>(slot);
        } else {
          st.pend[(pHead + nPend + __popc(rej & below)) & (kRefillSlots - 1)] = static_cast<uint8_t>(slot);
        }
      }
      nFin  += __popc(acc);
      nPend += __popc(rej);
      __syncwarp();
    } else {
      // ---- Finish for m accepted entries
      Secondaries sec;
      sec.n = 0;
      int32_t i = 0;
      int id = 0;
      if (lane < m) {
        const int slot = st.fin[(fHead + lane) & (kRefillSlots - 1)];
        Pars p;
        LoadPars(st.par, slot, p);
        i  = st.track[slot];
        id = static_cast<int>(st.id[slot]);
        const uint32_t d = st.draw[slot];
        Rng rng;
        rng.Init(seed, static_cast<uint32_t>(id), d, false, 0.0);
        if ((d & 1u) != 0u) {
          rng.hasNext = true;
          rng.next    = st.spare[slot];
        }
        Track s;
        typename IO::Keep keep;
        IO::LoadForFinish(b, i, s, keep);
        S::Finish(tv, s, p, rng, sec);
        IO::StoreFinish(tv, b, i, s, keep, rng.draw);
        st.freeStack[nFree + lane] = static_cast<uint8_t>(slot);
      }
      fHead = (fHead + m) & (kRefillSlots - 1);
      nFin -= m;
      nFree += m;
      AppendSecondariesWarp(sq, sec, id, i);
      __syncwarp();
    }
  }
}

}  // namespace g4h
#endif
