// g4h_lookups_f32.cuh -- single precision variant of the e-/e+ look-up set (BASELINE configs[0]).
//
// SURVEY par. 8f rank 4 / north_star: "within a stated bound if an FP32 variant is offered".  The FP64 path is the
// product (bit-identical to the reference); this kernel answers what single precision buys for the table look-ups:
// the same GetSplineLog / GetInvRange / GetRestMacXSec arithmetic (G4HepEmRunUtils.icc:49-120,
// G4HepEmElectronManager.icc:486-582), operation by operation, in float, on a float copy of the particle's hot tables
// staged in shared memory.  Stated bound (asserted by tests/test_gpu_parity.py::test_electron_lookups_f32_bound):
// |f32 - f64| <= 2e-5 |f64| + 1e-6 max|f64| per output (measured: 4e-6 relative) for the material couples; the vacuum
// couple's ranges square to more than single precision holds.  Inputs and outputs are float arrays.
// Measured on the B200: 0.0270 ms per 1M look-up sets against 0.0520 ms in FP64 (tools/lookups_f32_probe.py).
#ifndef G4H_LOOKUPS_F32_CUH
#define G4H_LOOKUPS_F32_CUH

#include "g4h_kernels.cuh"

namespace g4h {

struct LookupsF32Layout {
  // offsets (in floats) of the arrays inside the shared-memory block; the block mirrors the contiguous piece of the
  // arena that starts at lossEGrid, 1 float per double
  int lossEGrid, lossData, resData, enucEGrid, enucData, tr1Data, total;
};

__device__ __forceinline__ float SplineF(float x1, float x2, float y1, float y2, float sd1, float sd2, float x) {
  const float dl = x2 - x1;
  const float b  = fmaxf(0.f, fminf(1.f, (x - x1) / dl));
  const float os = 0.166666666667f;
  const float c0 = (2.0f - b) * sd1;
  const float c1 = (1.0f + b) * sd2;
  return y1 + b * (y2 - y1) + (b * (b - 1.0f)) * (c0 + c1) * (dl * dl * os);
}

__device__ __forceinline__ int LogBinF(float logx, float logxmin, float invLDBin, int ndata) {
  return static_cast<int>(fmaxf(0.f, fminf((logx - logxmin) * invLDBin, ndata - 2.f)));
}

__device__ __forceinline__ float SplineLogYSDF(int ndata, const float* xdata, const float* ydata, float x, float logx, float logxmin,
                                               float invLDBin) {
  const float xv = fmaxf(xdata[0], fminf(xdata[ndata - 1], x));
  const int idx  = LogBinF(logx, logxmin, invLDBin, ndata);
  const int idx2 = 2 * idx;
  return SplineF(xdata[idx], xdata[idx + 1], ydata[idx2], ydata[idx2 + 2], ydata[idx2 + 1], ydata[idx2 + 3], xv);
}

__device__ __forceinline__ float SplineLogXYSDF(int ndata, const float* data, float x, float logx, float logxmin, float invLDBin) {
  const float xv = fmaxf(data[0], fminf(data[3 * (ndata - 1)], x));
  const int idx  = LogBinF(logx, logxmin, invLDBin, ndata);
  const int idx3 = 3 * idx;
  return SplineF(data[idx3], data[idx3 + 3], data[idx3 + 1], data[idx3 + 4], data[idx3 + 2], data[idx3 + 5], xv);
}

// GetRestMacXSec (.icc:522-532) on the float copy; the per-couple start offsets stay integers (bit copies)
__device__ __forceinline__ float RestMacXSecF(const float* resData, const int* resStart, int imc, float ekin, float lekin, bool isIoni) {
  const int iIoni   = resStart[imc];
  const int numIoni = static_cast<int>(resData[iIoni]);
  const int iStart  = isIoni ? iIoni : iIoni + 3 * numIoni + 5;
  const float* d    = resData + iStart;
  const int numData = static_cast<int>(d[0]);
  if (ekin < d[5]) return 0.0f;
  return fmaxf(0.0f, SplineLogXYSDF(numData, d + 5, ekin, lekin, d[3], d[4]));
}

__global__ void __launch_bounds__(1024, 1)
ElectronLookupsF32Kernel(const __grid_constant__ TablesView tv, const __grid_constant__ LookupsF32Layout lay, int64_t n,
                         const int32_t* __restrict__ imc, const float* __restrict__ ekin, const float* __restrict__ lekin,
                         int particle, float* __restrict__ out) {
  extern __shared__ float smemF[];
  const ElectronTablesView& ed = tv.el[particle];
  // stage: double -> float for the whole block (the slots under the integer start offsets are not used); the start
  // offsets themselves are copied as integers behind the block
  const double* lo = ed.lossEGrid;
  for (int k = threadIdx.x; k < lay.total; k += blockDim.x) smemF[k] = static_cast<float>(__ldg(lo + k));
  int* resStartS = reinterpret_cast<int*>(smemF + lay.total);
  for (int k = threadIdx.x; k < tv.numMatCut; k += blockDim.x) resStartS[k] = __ldg(ed.resStart + k);
  __syncthreads();
  const float* lossEGrid = smemF + lay.lossEGrid;
  const float* lossData  = smemF + lay.lossData;
  const float* resData   = smemF + lay.resData;
  const float* enucEGrid = smemF + lay.enucEGrid;
  const float* enucData  = smemF + lay.enucData;
  const float* tr1Data   = smemF + lay.tr1Data;
  const int nl = ed.numLoss;
  const float lossLogMin = static_cast<float>(ed.lossLogMinEkin), lossILD = static_cast<float>(ed.lossEILDelta);
  const float enucLogMin = static_cast<float>(ed.enucLogMinEkin), enucILD = static_cast<float>(ed.enucEILDelta);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = imc[i];
    const float e = ekin[i], le = lekin[i];
    const int imat = __ldg(tv.mcImat + c);
    const float* rdata = lossData + 5 * nl * c;
    const float range = fmaxf(0.0f, SplineLogYSDF(nl, lossEGrid, rdata, e, le, lossLogMin, lossILD));
    out[0 * n + i] = range;
    out[1 * n + i] = fmaxf(0.0f, SplineLogYSDF(nl, lossEGrid, rdata + 2 * nl, e, le, lossLogMin, lossILD));
    // GetInvRange (.icc:504-519)
    float inv;
    const float minRange = rdata[0];
    if (range < minRange) {
      const float dum = range / minRange;
      inv = fmaxf(0.0f, lossEGrid[0] * dum * dum);
    } else {
      int ml = -1, mu = nl - 1;
      while (mu - ml > 1) {
        const int mav = (ml + mu) >> 1;
        if (range < rdata[2 * mav]) mu = mav; else ml = mav;
      }
      const int j = mu > 0 ? mu - 1 : 0;
      const float* sd = rdata + 4 * nl;
      inv = fmaxf(0.0f, SplineF(rdata[2 * j], rdata[2 * (j + 1)], lossEGrid[j], lossEGrid[j + 1], sd[j], sd[j + 1], range));
    }
    out[2 * n + i] = inv;
    out[3 * n + i] = RestMacXSecF(resData, resStartS, c, e, le, true);
    out[4 * n + i] = RestMacXSecF(resData, resStartS, c, e, le, false);
    out[5 * n + i] = e < enucEGrid[0] ? 0.0f
                                      : fmaxf(0.0f, SplineLogYSDF(128, enucEGrid, enucData + imat * 2 * 128, e, le, enucLogMin, enucILD));
    const float tr1 = fmaxf(0.0f, SplineLogYSDF(nl, lossEGrid, tr1Data + 2 * nl * imat, e, le, lossLogMin, lossILD));
    out[6 * n + i] = tr1 > 0.f ? 1.f / tr1 : 1.0e20f;
  }
}

}  // namespace g4h
#endif
