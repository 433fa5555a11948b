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

namespace g4h {

constexpr int kThreadsPerBlock = 256;
// resident CTAs per SM the big kernels are compiled for (register cap = 65536 / (256 * k)); tuned on the B200,
// see profiles/
#ifndef G4H_MINB_HOWFAR
#define G4H_MINB_HOWFAR 2
#endif
#ifndef G4H_MINB_CONT
#define G4H_MINB_CONT 2
#endif
#ifndef G4H_MINB_QUEUE
#define G4H_MINB_QUEUE 2
#endif

// ---- warp aggregated append to the secondary queue -------------------------------------------------------
// called by all 32 lanes of a warp (sec.n may be 0): ballots give each lane its offset, one atomicAdd per
// warp reserves the slots
__device__ __forceinline__ void AppendSecondaries(const G4HB200SecondaryQueue& q, const Secondaries& sec, int parentId,
                                                  int64_t parentIndex) {
  const unsigned active = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned has1 = __ballot_sync(active, sec.n >= 1);
  const unsigned has2 = __ballot_sync(active, sec.n >= 2);
  const unsigned below = (1u << lane) - 1u;
  const int excl  = __popc(has1 & below) + __popc(has2 & below);
  const int total = __popc(has1) + __popc(has2);
  if (total == 0) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(q.count, total);
  base = __shfl_sync(active, base, 0);
  for (int k = 0; k < sec.n; ++k) {
    const int64_t slot = static_cast<int64_t>(base) + excl + k;
    if (slot < q.capacity) {
      reinterpret_cast<double2*>(q.dirx_diry)[slot] = make_double2(sec.s[k].dir[0], sec.s[k].dir[1]);
      reinterpret_cast<double2*>(q.dirz_ekin)[slot] = make_double2(sec.s[k].dir[2], sec.s[k].ekin);
      reinterpret_cast<int2*>(q.parent_kind)[slot]  = make_int2(parentId, sec.s[k].kind);
      reinterpret_cast<int2*>(q.parent_slot)[slot]  = make_int2(static_cast<int>(parentIndex) + q.parent_base, k);
    }
  }
}

// ---- e-/e+ ---------------------------------------------------------------------------------------------------
// mode 0: HowFar, 1: Perform, 2: fused HowFar + Perform
template <int kMode>
__global__ void __launch_bounds__(kThreadsPerBlock, kMode == 0 ? G4H_MINB_HOWFAR : 1)
ElectronKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200ElectronBatch b,
               const __grid_constant__ G4HB200SecondaryQueue q, uint64_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // the loop bound is rounded up to a full warp so that whole warps reach the aggregated append
  const int64_t nRound = (b.n + 31) & ~static_cast<int64_t>(31);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    const bool valid = i < b.n;
    ElectronState s;
    Rng rng;
    Secondaries sec;
    sec.n = 0;
    if (valid) {
      LoadElectron(b, i, seed, s, rng);
      if (kMode == 1) LoadElectronHandOver(b, i, s);
      if (kMode != 1) {
        ResampleNumIALeft(s, rng);
        HowFarToDiscreteInteraction(tv, s);
        HowFarToMSC(tv, s, rng);
      }
      if (kMode != 0) {
        ElectronPerform(tv, s, rng, sec);
      }
      StoreElectron(b, i, s, rng);
      if (kMode == 0) StoreElectronHandOver(b, i, s);
      // Perform re-converts the geometrical step when geometry cut it (UpdatePStepLength): fTrueStepLength / fZPathLength
      if (kMode == 1) StorePair(b.tstep_zpath, i, s.trueStep, s.zPath);
    }
    if (kMode != 0) {
      __syncwarp();
      AppendSecondaries(q, sec, valid ? s.id : 0, i);
    }
  }
}

// ---- gamma -------------------------------------------------------------------------------------------------------
template <int kMode>
__global__ void __launch_bounds__(kThreadsPerBlock)
GammaKernel(const __grid_constant__ TablesView tv, const __grid_constant__ G4HB200GammaBatch b,
            const __grid_constant__ G4HB200SecondaryQueue q, uint64_t seed) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nRound = (b.n + 31) & ~static_cast<int64_t>(31);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nRound; i += stride) {
    const bool valid = i < b.n;
    GammaState s;
    Rng rng;
    Secondaries sec;
    sec.n = 0;
    if (valid) {
      LoadGamma(b, i, seed, s, rng);
      const int flags = b.meta[4 * i + 1];
      if (kMode == 1) LoadGammaHandOver(b, i, s);
      if (kMode != 1) GammaHowFar(tv, s, rng);
      if (kMode != 0) GammaPerform(tv, s, rng, sec);
      StoreGamma(b, i, s, rng, flags);
    }
    if (kMode != 0) {
      __syncwarp();
      AppendSecondaries(q, sec, valid ? s.id : 0, i);
    }
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
