// g4h_kernels.cuh -- the sm_100a kernels: one track per thread over paired-SoA batches.
//
// Launch shape: 256-thread CTAs, grid = a multiple of the SM count (148 on B200) with a grid-stride
// loop, so every SM gets the same number of resident CTAs regardless of the batch size.
// Tables are read through the read-only path (__ldg): the whole set is < 1 MB and stays L2 resident.
// Secondaries are appended to the queue with one atomicAdd per warp (ballot + prefix popcount).
#ifndef G4H_KERNELS_CUH
#define G4H_KERNELS_CUH

#include <cuda_runtime.h>

#include "g4h_batch_io.cuh"
#include "g4h_tma.cuh"

namespace g4h {

constexpr int kThreadsPerBlock = 256;
// the gamma step's kernels run as 128-thread CTAs, six per SM: measured 3 % faster than 256 x 3 for the 1M-photon step (the
// e-/e+ kernels are indifferent: 2.00e9 either way; profiles/r02b_schedule_ab.log)
constexpr int kGammaThreads = 128;
// resident CTAs per SM the queue kernels are compiled for (register cap = 65536 / (256 * k)); tuned on the B200,
// see profiles/
#ifndef G4H_MINB_QUEUE
#define G4H_MINB_QUEUE 3
#endif

// ---- CTA aggregated appends ----------------------------------------------------------------------------------
// Same-address global atomics serialise in L2: with one atomicAdd per warp the queue counters were the
// hottest lines of the queue kernels (40 % of the stall samples of ElDiscreteKernel, profiles/r01_*).  Appends
// are therefore aggregated per CTA: warps reserve their share in a shared-memory counter, one thread
// reserves the CTA's range with a single global atomicAdd.  Every thread of the CTA must call these (loops
// run a CTA-uniform number of iterations), they contain two __syncthreads().

template <int K>
struct CtaCounters {
  int count[K];  // zero between calls
  int base[K];
  __device__ __forceinline__ void Init() {
    if (threadIdx.x < K) count[threadIdx.x] = 0;
    __syncthreads();
  }
};

// loop bound of a grid-stride loop in which every thread of a CTA runs the same number of iterations
__device__ __forceinline__ int64_t RoundUpToCta(int64_t n) {
  return ((n + blockDim.x - 1) / blockDim.x) * blockDim.x;
}

// secondaries: sec.n in {0,1,2} per thread
__device__ __forceinline__ void AppendSecondaries(CtaCounters<1>& cc, const G4HB200SecondaryQueue& q, const Secondaries& sec,
                                                  int parentId, int64_t parentIndex) {
  const unsigned active = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned has1 = __ballot_sync(active, sec.n >= 1);
  const unsigned has2 = __ballot_sync(active, sec.n >= 2);
  const unsigned below = (1u << lane) - 1u;
  const int excl  = __popc(has1 & below) + __popc(has2 & below);
  const int total = __popc(has1) + __popc(has2);
  int warpBase = 0;
  if (lane == 0 && total > 0) warpBase = atomicAdd(&cc.count[0], total);
  warpBase = __shfl_sync(active, warpBase, 0);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = cc.count[0];
    cc.base[0]  = t > 0 ? atomicAdd(q.count, t) : 0;
    cc.count[0] = 0;
  }
  __syncthreads();
  const int64_t base = static_cast<int64_t>(cc.base[0]) + warpBase + excl;
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

// track index -> one of K queues: route in [0,K) or -1 (none); the k-th queue is queues[first + k] with its
// global counter counts + first + k
template <int K>
__device__ __forceinline__ void RouteToQueues(CtaCounters<K>& cc, int route, int32_t value, int32_t* const* queues,
                                              int32_t* counts) {
  const unsigned active = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int offset = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const unsigned m = __ballot_sync(active, route == k);
    if (m != 0u) {
      int wb = 0;
      if (lane == 0) wb = atomicAdd(&cc.count[k], __popc(m));
      wb = __shfl_sync(active, wb, 0);
      if (route == k) offset = wb + __popc(m & ((1u << lane) - 1u));
    }
  }
  __syncthreads();
  if (threadIdx.x < K) {
    const int t = cc.count[threadIdx.x];
    cc.base[threadIdx.x]  = t > 0 ? atomicAdd(counts + threadIdx.x, t) : 0;
    cc.count[threadIdx.x] = 0;
  }
  __syncthreads();
  if (route >= 0) queues[route][cc.base[route] + offset] = value;
}

// ---- gamma HowFar (G4HepEmGammaManager::HowFar(data, pars, tlData), .icc:27-48): one pass over every track --------------------
__global__ void __launch_bounds__(kThreadsPerBlock)
GammaHowFarKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200GammaBatch b, uint64_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < b.n; i += stride) {
    GammaState s;
    Rng rng;
    LoadGamma(b, i, seed, s, rng);
    const int flags = b.meta[4 * i + 1];
    GammaHowFar(tv, s, rng);
    StoreGamma(b, i, s, rng, flags);
  }
}

// ---- look-up kernels (BASELINE config 1 and the reference-style table tests) -----------------------------------------
__global__ void __launch_bounds__(kThreadsPerBlock)
ElectronLookupsKernel(const __grid_constant__ TablesView tv, int64_t n, const int32_t* __restrict__ imc,
                      const double* __restrict__ ekin, const double* __restrict__ lekin, int particle, double* __restrict__ out) {
  const ElectronTablesView& ed = tv.el[particle];
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = imc[i];
    const double e = ekin[i], le = lekin[i];
    const int imat = __ldg(tv.mcImat + c);
    const double range = RestRange(ed, c, e, le);
    out[0 * n + i] = range;
    out[1 * n + i] = RestDEDX(ed, c, e, le);
    out[2 * n + i] = InvRange(ed, c, range);
    out[3 * n + i] = RestMacXSec(ed, c, e, le, true);
    out[4 * n + i] = RestMacXSec(ed, c, e, le, false);
    out[5 * n + i] = MacXSecNuclear(ed, imat, e, le);
    out[6 * n + i] = TransportMFP(ed, imat, e, le);
  }
}

// the same with the particle's hot tables (loss, restricted cross sections, nuclear, transport: one contiguous block
// of the arena) staged in shared memory by a TMA bulk copy (g4h_tma.cuh): a look-up gathers 8 B from up to 32 different 32 B sectors of L1 per
// instruction; shared memory serves scattered 8 B words at bank rate
// the shared-memory copy of *p: the address is derived from the shared array, not from the global pointer (the
// compiler picks the load instruction from the provenance of the address)
template <class T>
__device__ __forceinline__ const T* StagedPtr(const void* smemBase, const char* globalBase, const T* p) {
  return reinterpret_cast<const T*>(static_cast<const char*>(smemBase) + (reinterpret_cast<const char*>(p) - globalBase));
}

__global__ void __launch_bounds__(1024, 1)
ElectronLookupsSmemKernel(const __grid_constant__ TablesView tv, int64_t n, const int32_t* __restrict__ imc,
                          const double* __restrict__ ekin, const double* __restrict__ lekin, int particle, double* __restrict__ out) {
  extern __shared__ double2 smemTables[];
  const ElectronTablesView& ed = tv.el[particle];
  const char* lo = reinterpret_cast<const char*>(ed.lossEGrid);
  const char* hi = reinterpret_cast<const char*>(ed.tr1Data + 2 * ed.numLoss * tv.numMat);
  const int n16  = static_cast<int>((hi - lo + 15) / 16);
  // the block travels through the TMA unit (one cp.async.bulk issued by one thread, an mbarrier counts the bytes in)
  __shared__ uint64_t tablesArrived;
  StageThroughTma(smemTables, lo, static_cast<uint32_t>(n16) * 16u, &tablesArrived);
  ElectronTablesView es = ed;
  es.lossEGrid = StagedPtr(smemTables, lo, ed.lossEGrid);
  es.lossData  = StagedPtr(smemTables, lo, ed.lossData);
  es.resStart  = StagedPtr(smemTables, lo, ed.resStart);
  es.resData   = StagedPtr(smemTables, lo, ed.resData);
  es.enucEGrid = StagedPtr(smemTables, lo, ed.enucEGrid);
  es.enucData  = StagedPtr(smemTables, lo, ed.enucData);
  es.tr1Data   = StagedPtr(smemTables, lo, ed.tr1Data);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = imc[i];
    const double e = ekin[i], le = lekin[i];
    const int imat = __ldg(tv.mcImat + c);
    const double range = RestRange(es, c, e, le);
    out[0 * n + i] = range;
    out[1 * n + i] = RestDEDX(es, c, e, le);
    out[2 * n + i] = InvRange(es, c, range);
    out[3 * n + i] = RestMacXSec(es, c, e, le, true);
    out[4 * n + i] = RestMacXSec(es, c, e, le, false);
    out[5 * n + i] = MacXSecNuclear(es, imat, e, le);
    out[6 * n + i] = TransportMFP(es, imat, e, le);
  }
}

__global__ void __launch_bounds__(kThreadsPerBlock)
ElectronSteppingXSecsKernel(const __grid_constant__ TablesView tv, int64_t n, const int32_t* __restrict__ imc,
                            const double* __restrict__ ekin, const double* __restrict__ lekin, int particle,
                            double* __restrict__ out) {
  const ElectronTablesView& ed = tv.el[particle];
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = imc[i];
    const double e = ekin[i], le = lekin[i];
    const int imat = __ldg(tv.mcImat + c);
    out[0 * n + i] = RestMacXSecForStepping(ed, c, e, le, true);
    out[1 * n + i] = RestMacXSecForStepping(ed, c, e, le, false);
    out[2 * n + i] = MacXSecNuclear(ed, imat, e, le);
    out[3 * n + i] = MacXSecAnnihilation(0.8 * e, __ldg(tv.matPars + 16 * imat + kMElectronDensity));
  }
}

__global__ void __launch_bounds__(kThreadsPerBlock)
GammaLookupsKernel(const __grid_constant__ TablesView tv, int64_t n, const int32_t* __restrict__ imc,
                   const double* __restrict__ ekin, const double* __restrict__ lekin, const double* __restrict__ urnd,
                   double* __restrict__ outMxsec, int32_t* __restrict__ outPid) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int imat = __ldg(tv.mcImat + imc[i]);
    double pe = 0.0;
    const double mx  = GammaTotalMacXSec(tv, imat, ekin[i], lekin[i], pe);
    const double mfp = mx > 0.0 ? 1.0 / mx : kALargeValue;
    outMxsec[i] = mx;
    outPid[i]   = GammaSampleInteraction(tv, imat, ekin[i], lekin[i], mfp, urnd[i], pe);
  }
}

__global__ void __launch_bounds__(kThreadsPerBlock)
SelectTargetElementKernel(const __grid_constant__ TablesView tv, int kind, int particle, int64_t n,
                          const int32_t* __restrict__ imc, const double* __restrict__ ekin, const double* __restrict__ lekin,
                          const double* __restrict__ urnd, int32_t* __restrict__ outElem) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    outElem[i] = kind == 2 ? SelectTargetAtomConversion(tv, imc[i], ekin[i], lekin[i], urnd[i])
                           : SelectTargetAtomBrem(tv.el[particle], imc[i], ekin[i], lekin[i], urnd[i], kind == 0);
  }
}

__global__ void __launch_bounds__(kThreadsPerBlock)
VdtLogExpKernel(int64_t n, const double* __restrict__ x, double* __restrict__ outLog, double* __restrict__ outExp) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    outLog[i] = Log(x[i]);
    outExp[i] = Exp(x[i]);
  }
}

__global__ void __launch_bounds__(kThreadsPerBlock)
RngUniformsKernel(uint64_t seed, int64_t n, const int32_t* __restrict__ trackId, int ndraw, double* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    Rng rng;
    rng.Init(seed, static_cast<uint32_t>(trackId[i]), 0u, false, 0.0);
    for (int j = 0; j < ndraw; ++j) out[i * ndraw + j] = rng.Flat();
  }
}

}  // namespace g4h
#endif
