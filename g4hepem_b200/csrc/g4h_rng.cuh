// g4h_rng.cuh -- per-track counter based uniform stream + the reference's Gauss / Poisson on top.
//
// Stream definition (must match oracle/g4h_rng_host.h, the stream injected into the CPU
// reference through G4HepEmRandomEngine::flat/flatArray):
//   u(seed, id, j) = (2k+1) * 2^-53,  k = top 52 bits of word pair (j&1) of
//   Philox4x32-10(counter = {j>>1, 0, id, 0}, key = {seed lo, seed hi})
// A Philox block yields two uniforms; the second one is kept in registers so consecutive draws
// cost one block per pair.  Gauss()/Poisson() restate G4HepEmRandomEngine.hh:50-97, including
// the cached second Box-Muller variate, which is per-track state here (the reference keeps it in
// the per-worker engine).
#ifndef G4H_RNG_CUH
#define G4H_RNG_CUH

#include "g4h_math.cuh"

namespace g4h {

G4H_FN uint32_t MulHi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

struct Philox4 {
  uint32_t x, y, z, w;
};

// Philox4x32-10 block: counter = {blk, 0, id, 0}, key = {k0, k1}
// PhiloxBlockInl: the body, for straight-line stage code (the integer rounds interleave with FP64 chains there);
// PhiloxBlock: the out-of-line copy everything else calls
G4H_FN Philox4 PhiloxBlockInl(uint32_t k0, uint32_t k1, uint32_t id, uint32_t blk) {
  uint32_t x0 = blk, x1 = 0u, x2 = id, x3 = 0u;
  uint32_t ka = k0, kb = k1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = MulHi32(0xD2511F53u, x0);
    const uint32_t lo0 = 0xD2511F53u * x0;
    const uint32_t hi1 = MulHi32(0xCD9E8D57u, x2);
    const uint32_t lo1 = 0xCD9E8D57u * x2;
    const uint32_t n0 = hi1 ^ x1 ^ ka;
    const uint32_t n2 = hi0 ^ x3 ^ kb;
    x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
    ka += 0x9E3779B9u;
    kb += 0xBB67AE85u;
  }
  return Philox4{x0, x1, x2, x3};
}

G4H_LEAF Philox4 PhiloxBlock(uint32_t k0, uint32_t k1, uint32_t id, uint32_t blk) { return PhiloxBlockInl(k0, k1, id, blk); }

struct Uniform2 {
  double a, b;
};

// (2k+1) * 2^-53 with k the top 52 bits of hi:lo -- exact: build 1.m in [1,2) and subtract (1 - 2^-53)
G4H_FN double ToUniform(uint32_t lo, uint32_t hi) {
  const uint64_t bits = (static_cast<uint64_t>(hi) << 32) | lo;
  const double d = FromBits(0x3FF0000000000000ULL | (bits >> 12));
  return d - 0.99999999999999988897769753748;  // 1 - 2^-53
}

// the two uniforms of block blk (draws 2*blk and 2*blk+1) of track id
G4H_LEAF Uniform2 UniformPair(uint32_t k0, uint32_t k1, uint32_t id, uint32_t blk) {
  const Philox4 r = PhiloxBlock(k0, k1, id, blk);
  return Uniform2{ToUniform(r.x, r.y), ToUniform(r.z, r.w)};
}

// The next uniforms of a track as a window u[0..4] = draws first .. first+4, generated in one go by every lane of
// the warp (three Philox blocks cover the five draws whatever the parity of `first`): the lock-step stages take
// their draws from it instead of calling Rng::Flat() from divergent sites.  Draw(5) (the sixth uniform) is in the
// window when `first` is even and costs one more block otherwise.
struct DrawWindow {
  double u[5];
  double v5;        // uniform 2*(blk0+2)+1: the sixth draw when `first` is even
  uint32_t k0, k1, id, first;

  G4H_MFN void Init(uint64_t seed, uint32_t trackId, uint32_t firstDraw) {
    k0 = static_cast<uint32_t>(seed);
    k1 = static_cast<uint32_t>(seed >> 32);
    id = trackId;
    first = firstDraw;
    const uint32_t blk = firstDraw >> 1;
    const Philox4 a = PhiloxBlockInl(k0, k1, id, blk);
    const Philox4 b = PhiloxBlockInl(k0, k1, id, blk + 1u);
    const Philox4 c = PhiloxBlockInl(k0, k1, id, blk + 2u);
    const double v0 = ToUniform(a.x, a.y), v1 = ToUniform(a.z, a.w);
    const double v2 = ToUniform(b.x, b.y), v3 = ToUniform(b.z, b.w);
    const double v4 = ToUniform(c.x, c.y);
    v5 = ToUniform(c.z, c.w);
    const bool odd = (firstDraw & 1u) != 0u;
    u[0] = odd ? v1 : v0;
    u[1] = odd ? v2 : v1;
    u[2] = odd ? v3 : v2;
    u[3] = odd ? v4 : v3;
    u[4] = odd ? v5 : v4;
  }
  // the sixth uniform of the window (draw first+5)
  G4H_MFN double Sixth() const {
    if ((first & 1u) == 0u) return v5;
    const Philox4 d = PhiloxBlock(k0, k1, id, (first >> 1) + 3u);
    return ToUniform(d.x, d.y);
  }
};

struct Rng {
  uint32_t k0, k1;  // key: global seed
  uint32_t id;      // track id
  uint32_t draw;    // index of the next uniform
  bool hasNext;     // `next` holds the uniform with index `draw` (only ever true while draw is odd)
  double next;
  bool hasGauss;    // G4HepEmRandomEngine::fIsGauss
  double gauss;     // G4HepEmRandomEngine::fGauss
  // optional window of pre-generated uniforms (queue kernels, shared memory): win[k * winStride] is draw winFirst + k.
  // Rejection and Poisson loops consume a data dependent number of draws per track; generated one block at a time
  // from inside those loops the Philox rounds ran at 8 of 32 lanes (42 % of the instructions of the fluctuation
  // kernel, profiles/r01c_*), generated up front every lane of the warp works.
  const double* win;
  uint32_t winFirst, winCount, winStride;

  G4H_MFN void Init(uint64_t seed, uint32_t trackId, uint32_t firstDraw, bool isGauss, double gaussVal) {
    k0 = static_cast<uint32_t>(seed);
    k1 = static_cast<uint32_t>(seed >> 32);
    id = trackId;
    draw = firstDraw;
    hasNext = false;
    next = 0.0;
    hasGauss = isGauss;
    gauss = gaussVal;
    win = nullptr;
    winFirst = 0u;
    winCount = 0u;
    winStride = 0u;
  }

  // fill `slots` (even) window entries starting at the block that holds the next draw; call right after Init
  G4H_MFN void FillWindow(double* window, uint32_t stride, uint32_t slots) {
    const uint32_t blk0 = draw >> 1;
    uint32_t k = 0;
    // four blocks at a time, inlined: the ten rounds of one block are a dependent chain of integer multiplies
#pragma unroll 1
    for (; k + 8u <= slots; k += 8u) {
      const uint32_t b = blk0 + (k >> 1);
      const Philox4 p0 = PhiloxBlockInl(k0, k1, id, b);
      const Philox4 p1 = PhiloxBlockInl(k0, k1, id, b + 1u);
      const Philox4 p2 = PhiloxBlockInl(k0, k1, id, b + 2u);
      const Philox4 p3 = PhiloxBlockInl(k0, k1, id, b + 3u);
      window[k * stride]        = ToUniform(p0.x, p0.y);
      window[(k + 1u) * stride] = ToUniform(p0.z, p0.w);
      window[(k + 2u) * stride] = ToUniform(p1.x, p1.y);
      window[(k + 3u) * stride] = ToUniform(p1.z, p1.w);
      window[(k + 4u) * stride] = ToUniform(p2.x, p2.y);
      window[(k + 5u) * stride] = ToUniform(p2.z, p2.w);
      window[(k + 6u) * stride] = ToUniform(p3.x, p3.y);
      window[(k + 7u) * stride] = ToUniform(p3.z, p3.w);
    }
#pragma unroll 1
    for (; k < slots; k += 2u) {
      const Uniform2 u = UniformPair(k0, k1, id, blk0 + (k >> 1));
      window[k * stride]        = u.a;
      window[(k + 1u) * stride] = u.b;
    }
    win = window;
    winFirst = blk0 << 1;
    winCount = slots;
    winStride = stride;
  }

  // G4HepEmRandomEngine::flat(): draws are consumed strictly in order, so the second uniform of a block is
  // always the next one asked for
  G4H_MFN double Flat() {
    const uint32_t j = draw++;
    const uint32_t k = j - winFirst;
    if (k < winCount) return win[k * winStride];
    if (hasNext) {
      hasNext = false;
      return next;
    }
    const Uniform2 u = UniformPair(k0, k1, id, j >> 1);
    if (j & 1u) return u.b;
    next = u.b;
    hasNext = true;
    return u.a;
  }

  // G4HepEmRandomEngine::Gauss (G4HepEmRandomEngine.hh:50-67): polar Box-Muller, second variate cached
  G4H_MFN double Gauss(double mean, double stDev) {
    if (hasGauss) {
      hasGauss = false;
      return gauss * stDev + mean;
    }
    double r, v1, v2;
    do {
      const double u0 = Flat();
      const double u1 = Flat();
      v1 = 2. * u0 - 1.;
      v2 = 2. * u1 - 1.;
      r = v1 * v1 + v2 * v2;
    } while (r > 1.);
    const double fac = sqrt(-2. * Log(r) / r);
    gauss = v1 * fac;
    hasGauss = true;
    return v2 * fac * stDev + mean;
  }

  // G4HepEmRandomEngine::Poisson (G4HepEmRandomEngine.hh:74-97)
  G4H_MFN int Poisson(double mean) {
    const int border = 16;
    const double limit = 2.E+9;
    int number = 0;
    if (mean <= border) {
      const double position = Flat();
      double poissonValue = Exp(-mean);
      double poissonSum = poissonValue;
      while (poissonSum <= position) {
        ++number;
        poissonValue *= FastDiv(mean, static_cast<double>(number));  // mean in (0, 16], number a small integer
        poissonSum += poissonValue;
      }
      return number;
    }
    const double u0 = Flat();
    const double u1 = Flat();
    const double t = sqrt(-2. * Log(u0)) * Cos(k2Pi * u1);
    const double value = mean + t * sqrt(mean) + 0.5;
    return value < 0. ? 0 : value >= limit ? static_cast<int>(limit) : static_cast<int>(value);
  }
};

}  // namespace g4h
#endif
