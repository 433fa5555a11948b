/* G4HepEmB200Stream.h -- host definition of the per-track counter based uniform stream (plain C).
 *
 * The reference leaves G4HepEmRandomEngine::flat()/flatArray() to the consumer
 * (G4HepEm/G4HepEmRun/include/G4HepEmRandomEngine.hh:21-28,33-47).  The kernels draw from a stream that is a pure
 * function of (seed, track id, draw index), so that results do not depend on the order tracks are processed in:
 *
 *     u(seed, track_id, j) = ((bits64 >> 12) * 2 + 1) * 2^-53  in (0,1)
 *     bits64 = word pair (j & 1) of Philox4x32-10(counter = {j>>1, 0, track_id, 0}, key = {seed lo, seed hi})
 *
 * Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11 (public algorithm;
 * known-answer vectors checked in tests/test_rng.py).  The device implementation of the same definition is
 * g4hepem_b200/csrc/g4h_rng.cuh.  A host application that wants its G4HepEmRandomEngine to produce the same numbers
 * as the kernels (the drop-in managers of G4HepEmB200DropIn.hh need that; so does the parity harness) points
 * G4HepEmRandomEngine::fObject at a G4HepEmB200Stream and implements flat() with G4HepEmB200StreamNext.
 */
#ifndef G4HEPEMB200_STREAM_H
#define G4HEPEMB200_STREAM_H
#include <stdint.h>

static inline void G4HepEmB200Philox4x32_10(uint32_t c[4], const uint32_t key[2]) {
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

static inline double G4HepEmB200BitsToUniform(uint64_t bits) {
  /* (2k+1) * 2^-53 with k = top 52 bits: exactly representable, never 0 or 1 */
  return (double)(((bits >> 12) << 1) | 1ull) * 1.1102230246251565e-16;
}

static inline double G4HepEmB200Uniform(uint64_t seed, uint32_t track_id, uint32_t draw) {
  uint32_t c[4] = {draw >> 1, 0u, track_id, 0u};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  G4HepEmB200Philox4x32_10(c, key);
  const uint64_t bits = (draw & 1u) ? (((uint64_t)c[3] << 32) | c[2]) : (((uint64_t)c[1] << 32) | c[0]);
  return G4HepEmB200BitsToUniform(bits);
}

/* the state of one track's stream: what a G4HepEmRandomEngine's fObject points to */
typedef struct G4HepEmB200Stream {
  uint64_t seed;
  uint32_t track_id;
  uint32_t draw;     /* index of the next uniform */
  /* the Gauss cache of G4HepEmRandomEngine (fIsGauss / fGauss, private there) as the device kernels keep it per track;
   * used by the drop-in managers only */
  int32_t is_gauss;
  double gauss;
} G4HepEmB200Stream;

static inline double G4HepEmB200StreamNext(G4HepEmB200Stream* s) { return G4HepEmB200Uniform(s->seed, s->track_id, s->draw++); }

#endif
