/* g4h_rng_host.h -- the injected uniform stream of the CPU reference (TEST INFRASTRUCTURE).
 *
 * The definition of the stream is the product's (g4hepem_b200/host/G4HepEmB200Stream.h: Philox4x32-10 keyed by
 * (seed, track id, draw index)); the oracle injects exactly that into the reference's G4HepEmRandomEngine::flat()
 * (oracle/ref_shim.cc), so that the CPU reference and the kernels consume the same uniforms per track.
 */
#ifndef G4H_RNG_HOST_H
#define G4H_RNG_HOST_H
#include "../g4hepem_b200/host/G4HepEmB200Stream.h"

typedef G4HepEmB200Stream G4HStream;
static inline void g4h_philox4x32_10(uint32_t c[4], const uint32_t key[2]) { G4HepEmB200Philox4x32_10(c, key); }
static inline double g4h_uniform(uint64_t seed, uint32_t track_id, uint32_t draw) { return G4HepEmB200Uniform(seed, track_id, draw); }
static inline double g4h_stream_next(G4HStream* s) { return G4HepEmB200StreamNext(s); }

#endif
