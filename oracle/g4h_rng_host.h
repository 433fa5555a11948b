/* g4h_rng_host.h -- host definition of the injected uniform stream (TEST INFRASTRUCTURE).
 *
 * The reference leaves G4HepEmRandomEngine::flat()/flatArray() to the consumer
 * (G4HepEm/G4HepEmRun/include/G4HepEmRandomEngine.hh:21-28,33-47).  Parity needs the CPU
 * reference and the GPU kernels to consume the *same* uniforms per track, independent of the
 * order tracks are processed in, so the stream is counter based:
 *
 *     u(seed, track_id, j) = ((bits64 >> 12) * 2 + 1) * 2^-53  in (0,1)
 *     bits64 = word pair (j & 1) of Philox4x32-10(counter = {j>>1 lo, j>>1 hi, track_id, 0},
 *                                                 key = {seed lo, seed hi})
 *
 * Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11 (public algorithm; known-answer vectors checked in tests/test_rng.py).
 * The GPU implementation of the same definition lives in g4hepem_b200/csrc/g4h_rng.cuh.
 */
#ifndef G4H_RNG_HOST_H
#define G4H_RNG_HOST_H
#include <stdint.h>

static inline void g4h_philox4x32_10(uint32_t c[4], const uint32_t key[2]) {
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

static inline double g4h_bits_to_uniform(uint64_t bits) {
  /* (2k+1) * 2^-53 with k = top 52 bits: exactly representable, never 0 or 1 */
  return (double)(((bits >> 12) << 1) | 1ull) * 1.1102230246251565e-16;
}

static inline double g4h_uniform(uint64_t seed, uint32_t track_id, uint32_t draw) {
  uint32_t c[4] = {draw >> 1, 0u, track_id, 0u};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  g4h_philox4x32_10(c, key);
  const uint64_t bits = (draw & 1u) ? (((uint64_t)c[3] << 32) | c[2]) : (((uint64_t)c[1] << 32) | c[0]);
  return g4h_bits_to_uniform(bits);
}

typedef struct G4HStream {
  uint64_t seed;
  uint32_t track_id;
  uint32_t draw; /* next draw index */
} G4HStream;

static inline double g4h_stream_next(G4HStream* s) { return g4h_uniform(s->seed, s->track_id, s->draw++); }

#endif
