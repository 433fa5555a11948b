// g4h_tma.cuh -- bulk asynchronous copies global -> shared memory through the TMA unit (cp.async.bulk, sm_90+; UBLKCP in SASS)
// with an mbarrier that counts the bytes as they land.  Used to stage the hot tables of a particle in shared memory
// (g4h_kernels.cuh, g4h_lookups_f32.cuh): one thread issues the copy of the whole block, every thread of the CTA waits on
// the barrier's phase -- no registers are tied up by the data in flight, and the first table look-ups wait for the arrival
// of bytes instead of a __syncthreads behind a copy loop.
#ifndef G4H_TMA_CUH
#define G4H_TMA_CUH

#include <cstdint>

namespace g4h {

__device__ __forceinline__ uint32_t SharedAddress(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void MbarrierInit(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SharedAddress(bar)), "r"(arrivals) : "memory");
  // make the initialised barrier visible to the async proxy before a bulk copy signals it
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// one arrival + the number of bytes the barrier's current phase waits for
__device__ __forceinline__ void MbarrierArriveExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SharedAddress(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void MbarrierWait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (done == 0) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(SharedAddress(bar)), "r"(parity)
        : "memory");
  }
}

// bytes: a multiple of 16; both addresses 16-byte aligned.  Completion is signalled on `bar` (complete_tx of `bytes`)
__device__ __forceinline__ void BulkCopyGlobalToShared(void* smemDst, const void* globalSrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(SharedAddress(smemDst)),
               "l"(globalSrc), "r"(bytes), "r"(SharedAddress(bar))
               : "memory");
}

// Stage `bytes` (rounded up to 16 by the caller; the source must be readable that far) at smemDst; every thread of the
// CTA calls this and returns when the data is there.  bar: a __shared__ uint64_t of the caller.
__device__ __forceinline__ void StageThroughTma(void* smemDst, const void* globalSrc, uint32_t bytes, uint64_t* bar) {
  if (threadIdx.x == 0) MbarrierInit(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    // a single bulk copy moves at most 2^20 - 16 bytes here (tx-count range of the barrier): larger blocks go in pieces
    constexpr uint32_t kPiece = 1u << 16;
    MbarrierArriveExpectTx(bar, bytes);
    for (uint32_t off = 0; off < bytes; off += kPiece) {
      const uint32_t len = bytes - off < kPiece ? bytes - off : kPiece;
      BulkCopyGlobalToShared(static_cast<char*>(smemDst) + off, static_cast<const char*>(globalSrc) + off, len, bar);
    }
  }
  MbarrierWait(bar, 0);
}

}  // namespace g4h
#endif
