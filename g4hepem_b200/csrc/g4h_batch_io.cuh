// g4h_batch_io.cuh -- SoA batch <-> register state.  Every double group is an array of {a,b} pairs:
// one 128-bit load/store per thread and group, a warp touches 512 contiguous bytes.
#ifndef G4H_BATCH_IO_CUH
#define G4H_BATCH_IO_CUH

#include "../../include/g4hepem_b200.h"
#include "g4h_interactions.cuh"

namespace g4h {

struct Pair {
  double a, b;
};

G4H_FN Pair LoadPair(const double* group, int64_t i) {
#if defined(__CUDA_ARCH__)
  // track state streams through once per kernel: evict-first keeps the tables in L1/L2
  const double2 v = __ldcs(reinterpret_cast<const double2*>(group) + i);
  return Pair{v.x, v.y};
#else
  return Pair{group[2 * i], group[2 * i + 1]};
#endif
}

G4H_FN void StorePair(double* group, int64_t i, double a, double b) {
#if defined(__CUDA_ARCH__)
  __stcs(reinterpret_cast<double2*>(group) + i, make_double2(a, b));
#else
  group[2 * i]     = a;
  group[2 * i + 1] = b;
#endif
}

struct Meta {
  int imc, flags, id, draw;
};

G4H_FN Meta LoadMeta(const int32_t* meta, int64_t i) {
#if defined(__CUDA_ARCH__)
  const int4 v = __ldcs(reinterpret_cast<const int4*>(meta) + i);
  return Meta{v.x, v.y, v.z, v.w};
#else
  return Meta{meta[4 * i], meta[4 * i + 1], meta[4 * i + 2], meta[4 * i + 3]};
#endif
}

G4H_FN void StoreMeta(int32_t* meta, int64_t i, const Meta& m) {
#if defined(__CUDA_ARCH__)
  __stcs(reinterpret_cast<int4*>(meta) + i, make_int4(m.imc, m.flags, m.id, m.draw));
#else
  meta[4 * i] = m.imc; meta[4 * i + 1] = m.flags; meta[4 * i + 2] = m.id; meta[4 * i + 3] = m.draw;
#endif
}

// ---- e-/e+ ------------------------------------------------------------------------------------------
// persistent groups -> state (what G4HepEmElectronManager::HowFar reads)
G4H_FN void LoadElectron(const G4HB200ElectronBatch& b, int64_t i, uint64_t seed, ElectronState& s, Rng& rng) {
  const Meta m = LoadMeta(b.meta, i);
  const Pair e = LoadPair(b.ekin_logekin, i);
  const Pair dxy = LoadPair(b.dirx_diry, i);
  const Pair dzs = LoadPair(b.dirz_safety, i);
  const Pair n01 = LoadPair(b.nia01, i);
  const Pair n23 = LoadPair(b.nia23, i);
  const Pair ir = LoadPair(b.msc_irange_dynrf, i);
  const Pair tg = LoadPair(b.msc_tlimmin_gauss, i);
  s.ekin = e.a; s.logEkin = e.b;
  s.dir[0] = dxy.a; s.dir[1] = dxy.b; s.dir[2] = dzs.a;
  s.safety = dzs.b;
  s.nIA[0] = n01.a; s.nIA[1] = n01.b; s.nIA[2] = n23.a; s.nIA[3] = n23.b;
  s.initialRange = ir.a; s.dynRangeFactor = ir.b; s.tlimitMin = tg.a;
  s.imc = m.imc; s.id = m.id;
  const uint32_t f = static_cast<uint32_t>(m.flags);
  s.isPositron   = (f & G4HB200_F_POSITRON) != 0u;
  s.onBoundary   = (f & G4HB200_F_ON_BOUNDARY) != 0u;
  s.mscFirstStep = (f & G4HB200_F_MSC_FIRST_STEP) != 0u;
  s.mscActive    = (f & G4HB200_F_MSC_ACTIVE) != 0u;
  s.mscDisplace  = (f & G4HB200_F_MSC_DISPLACE) != 0u;
  s.mscNoScatter = (f & G4HB200_F_MSC_NO_SCATTER) != 0u;
  rng.Init(seed, static_cast<uint32_t>(m.id), static_cast<uint32_t>(m.draw), (f & G4HB200_F_GAUSS_CACHED) != 0u, tg.b);
  // defaults of the fields HowFar defines (G4HepEmTrack::ReSet / G4HepEmMSCTrackData::ReSet values); the energy
  // deposit is not one of them: HowFar leaves what the previous Perform wrote
  s.winner = -1; s.gStep = 0.0; s.pStep = 0.0; s.edep = LoadPair(b.edep_dispx, i).a; s.range = 0.0;
  s.mfp[0] = s.mfp[1] = s.mfp[2] = s.mfp[3] = -1.0;
  s.lambtr1 = 0.0; s.trueStep = 0.0; s.zPath = 0.0;
  s.disp[0] = s.disp[1] = s.disp[2] = 0.0;
  s.par1 = -1.0; s.par2 = 0.0; s.par3 = 0.0;
  s.preStepEkin = 0.0; s.preStepLogEkin = 0.0;
}

// hand-over groups written by HowFar -> state (what G4HepEmElectronManager::Perform additionally reads)
G4H_FN void LoadElectronHandOver(const G4HB200ElectronBatch& b, int64_t i, ElectronState& s) {
  const Pair gp = LoadPair(b.gstep_pstep, i);
  const Pair ed = LoadPair(b.edep_dispx, i);
  const Pair dyz = LoadPair(b.dispy_dispz, i);
  const Pair m01 = LoadPair(b.mfp01, i);
  const Pair m23 = LoadPair(b.mfp23, i);
  const Pair rl = LoadPair(b.range_lambtr1, i);
  const Pair tz = LoadPair(b.tstep_zpath, i);
  const Pair p12 = LoadPair(b.par12, i);
  const Pair p3 = LoadPair(b.par3_pad, i);
  s.gStep = gp.a; s.pStep = gp.b;
  s.edep = ed.a;
  s.disp[0] = ed.b; s.disp[1] = dyz.a; s.disp[2] = dyz.b;
  s.mfp[0] = m01.a; s.mfp[1] = m01.b; s.mfp[2] = m23.a; s.mfp[3] = m23.b;
  s.range = rl.a; s.lambtr1 = rl.b;
  s.trueStep = tz.a; s.zPath = tz.b;
  s.par1 = p12.a; s.par2 = p12.b; s.par3 = p3.a;
  s.winner = b.winner[i];
}

G4H_FN int ElectronFlags(const ElectronState& s, const Rng& rng) {
  uint32_t f = 0u;
  if (s.isPositron) f |= G4HB200_F_POSITRON;
  if (s.onBoundary) f |= G4HB200_F_ON_BOUNDARY;
  if (s.mscFirstStep) f |= G4HB200_F_MSC_FIRST_STEP;
  if (s.mscActive) f |= G4HB200_F_MSC_ACTIVE;
  if (s.mscDisplace) f |= G4HB200_F_MSC_DISPLACE;
  if (s.mscNoScatter) f |= G4HB200_F_MSC_NO_SCATTER;
  if (rng.hasGauss) f |= G4HB200_F_GAUSS_CACHED;
  return static_cast<int>(f);
}

// state -> persistent + result groups
G4H_FN void StoreElectron(const G4HB200ElectronBatch& b, int64_t i, const ElectronState& s, const Rng& rng) {
  StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
  StorePair(b.dirx_diry, i, s.dir[0], s.dir[1]);
  StorePair(b.dirz_safety, i, s.dir[2], s.safety);
  StorePair(b.nia01, i, s.nIA[0], s.nIA[1]);
  StorePair(b.nia23, i, s.nIA[2], s.nIA[3]);
  StorePair(b.msc_irange_dynrf, i, s.initialRange, s.dynRangeFactor);
  StorePair(b.msc_tlimmin_gauss, i, s.tlimitMin, rng.gauss);
  StoreMeta(b.meta, i, Meta{s.imc, ElectronFlags(s, rng), s.id, static_cast<int>(rng.draw)});
  StorePair(b.gstep_pstep, i, s.gStep, s.pStep);
  StorePair(b.edep_dispx, i, s.edep, s.disp[0]);
  StorePair(b.dispy_dispz, i, s.disp[1], s.disp[2]);
  b.winner[i] = s.winner;
}

G4H_FN void StoreElectronHandOver(const G4HB200ElectronBatch& b, int64_t i, const ElectronState& s) {
  StorePair(b.mfp01, i, s.mfp[0], s.mfp[1]);
  StorePair(b.mfp23, i, s.mfp[2], s.mfp[3]);
  StorePair(b.range_lambtr1, i, s.range, s.lambtr1);
  StorePair(b.tstep_zpath, i, s.trueStep, s.zPath);
  StorePair(b.par12, i, s.par1, s.par2);
  StorePair(b.par3_pad, i, s.par3, 0.0);
}

// ---- gamma --------------------------------------------------------------------------------------------
G4H_FN void LoadGamma(const G4HB200GammaBatch& b, int64_t i, uint64_t seed, GammaState& s, Rng& rng) {
  const Meta m = LoadMeta(b.meta, i);
  const Pair e = LoadPair(b.ekin_logekin, i);
  const Pair dxy = LoadPair(b.dirx_diry, i);
  const Pair dzn = LoadPair(b.dirz_nia0, i);
  s.ekin = e.a; s.logEkin = e.b;
  s.dir[0] = dxy.a; s.dir[1] = dxy.b; s.dir[2] = dzn.a;
  s.nIA0 = dzn.b;
  s.imc = m.imc; s.id = m.id;
  s.onBoundary = (static_cast<uint32_t>(m.flags) & G4HB200_F_ON_BOUNDARY) != 0u;
  rng.Init(seed, static_cast<uint32_t>(m.id), static_cast<uint32_t>(m.draw), false, 0.0);
  // fields HowFar does not define keep what the previous step left in the track (the reference's managers work
  // in place on a persistent G4HepEmGammaTrack): winner index, deposit, fPEmxSec
  const Pair ep = LoadPair(b.edep_pemxsec, i);
  s.mfp0 = -1.0; s.gStep = 0.0; s.edep = ep.a; s.peMXsec = ep.b; s.winner = b.winner[i];
}

G4H_FN void LoadGammaHandOver(const G4HB200GammaBatch& b, int64_t i, GammaState& s) {
  const Pair gm = LoadPair(b.gstep_mfp0, i);
  const Pair ep = LoadPair(b.edep_pemxsec, i);
  s.gStep = gm.a; s.mfp0 = gm.b;
  s.edep = ep.a; s.peMXsec = ep.b;
  s.winner = b.winner[i];
}

G4H_FN void StoreGamma(const G4HB200GammaBatch& b, int64_t i, const GammaState& s, const Rng& rng, int flags) {
  StorePair(b.ekin_logekin, i, s.ekin, s.logEkin);
  StorePair(b.dirx_diry, i, s.dir[0], s.dir[1]);
  StorePair(b.dirz_nia0, i, s.dir[2], s.nIA0);
  StoreMeta(b.meta, i, Meta{s.imc, flags, s.id, static_cast<int>(rng.draw)});
  StorePair(b.gstep_mfp0, i, s.gStep, s.mfp0);
  StorePair(b.edep_pemxsec, i, s.edep, s.peMXsec);
  b.winner[i] = s.winner;
}

}  // namespace g4h
#endif
