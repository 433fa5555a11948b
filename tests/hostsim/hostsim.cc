// hostsim.cc -- PRE-FLIGHT TEST HARNESS, never linked into the product.
//
// Builds the per-track device functions of g4hepem_b200/csrc/*.cuh for the host (g++ -ffp-contract=off)
// so that their logic can be compared with the reference in a container without a GPU.  It exists to
// catch transcription mistakes before GPU time is spent; the parity claims rest on the `-m gpu` tests,
// which run the CUDA kernels through the C-ABI.  Nothing under g4hepem_b200/ loads this library.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../g4hepem_b200/csrc/g4h_batch_io.cuh"
#include "../../g4hepem_b200/csrc/g4h_perform_stages.cuh"
#include "../../g4hepem_b200/csrc/g4h_stages.cuh"
#include "../../g4hepem_b200/csrc/g4h_view.cuh"

using namespace g4h;

namespace {
void AppendSec(G4HB200SecondaryQueue* q, const Secondaries& sec, int parentId, int64_t parentIndex) {
  for (int k = 0; k < sec.n; ++k) {
    const int64_t n = q->count[0]++;
    q->dirx_diry[2 * n]       = sec.s[k].dir[0];
    q->dirx_diry[2 * n + 1]   = sec.s[k].dir[1];
    q->dirz_ekin[2 * n]       = sec.s[k].dir[2];
    q->dirz_ekin[2 * n + 1]   = sec.s[k].ekin;
    q->parent_kind[2 * n]     = parentId;
    q->parent_kind[2 * n + 1] = sec.s[k].kind;
    q->parent_slot[2 * n]     = static_cast<int32_t>(parentIndex);
    q->parent_slot[2 * n + 1] = k;
  }
}
}  // namespace

extern "C" {

void g4hsim_electron_lookups(const G4HB200Tables* t, int64_t n, const int32_t* imc, const double* ekin, const double* lekin,
                             int isElectron, double* out) {
  const TablesView tv = MakeView(*t);
  const ElectronTablesView& ed = tv.el[isElectron ? 0 : 1];
  for (int64_t i = 0; i < n; ++i) {
    const int imat = tv.mcImat[imc[i]];
    const double range = RestRange(ed, imc[i], ekin[i], lekin[i]);
    out[0 * n + i] = range;
    out[1 * n + i] = RestDEDX(ed, imc[i], ekin[i], lekin[i]);
    out[2 * n + i] = InvRange(ed, imc[i], range);
    out[3 * n + i] = RestMacXSec(ed, imc[i], ekin[i], lekin[i], true);
    out[4 * n + i] = RestMacXSec(ed, imc[i], ekin[i], lekin[i], false);
    out[5 * n + i] = MacXSecNuclear(ed, imat, ekin[i], lekin[i]);
    out[6 * n + i] = TransportMFP(ed, imat, ekin[i], lekin[i]);
  }
}

void g4hsim_electron_stepping_xsecs(const G4HB200Tables* t, int64_t n, const int32_t* imc, const double* ekin,
                                    const double* lekin, int isElectron, double* out) {
  const TablesView tv = MakeView(*t);
  const ElectronTablesView& ed = tv.el[isElectron ? 0 : 1];
  for (int64_t i = 0; i < n; ++i) {
    const int imat = tv.mcImat[imc[i]];
    out[0 * n + i] = RestMacXSecForStepping(ed, imc[i], ekin[i], lekin[i], true);
    out[1 * n + i] = RestMacXSecForStepping(ed, imc[i], ekin[i], lekin[i], false);
    out[2 * n + i] = MacXSecNuclear(ed, imat, ekin[i], lekin[i]);
    out[3 * n + i] = MacXSecAnnihilation(0.8 * ekin[i], tv.matPars[16 * imat + kMElectronDensity]);
  }
}

void g4hsim_gamma_lookups(const G4HB200Tables* t, int64_t n, const int32_t* imc, const double* ekin, const double* lekin,
                          const double* urnd, double* outMxsec, int32_t* outPid) {
  const TablesView tv = MakeView(*t);
  for (int64_t i = 0; i < n; ++i) {
    const int imat = tv.mcImat[imc[i]];
    double pe = 0.0;
    const double mx = GammaTotalMacXSec(tv, imat, ekin[i], lekin[i], pe);
    const double mfp = mx > 0.0 ? 1.0 / mx : 1.0e20;
    outMxsec[i] = mx;
    outPid[i]   = GammaSampleInteraction(tv, imat, ekin[i], lekin[i], mfp, urnd[i], pe);
  }
}

void g4hsim_select_target_element(const G4HB200Tables* t, int kind, int isElectron, int64_t n, const int32_t* imc,
                                  const double* ekin, const double* lekin, const double* urnd, int32_t* outElem) {
  const TablesView tv = MakeView(*t);
  for (int64_t i = 0; i < n; ++i) {
    outElem[i] = kind == 2 ? SelectTargetAtomConversion(tv, imc[i], ekin[i], lekin[i], urnd[i])
                           : SelectTargetAtomBrem(tv.el[isElectron ? 0 : 1], imc[i], ekin[i], lekin[i], urnd[i], kind == 0);
  }
}

void g4hsim_vdt_log_exp(int64_t n, const double* x, double* outLog, double* outExp) {
  for (int64_t i = 0; i < n; ++i) {
    outLog[i] = Log(x[i]);
    outExp[i] = Exp(x[i]);
  }
}

void g4hsim_rng_uniforms(uint64_t seed, int64_t n, const int32_t* trackId, int32_t ndraw, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    Rng rng;
    rng.Init(seed, static_cast<uint32_t>(trackId[i]), 0u, false, 0.0);
    for (int j = 0; j < ndraw; ++j) out[i * ndraw + j] = rng.Flat();
  }
}

// mode: 0 = HowFar, 1 = Perform, 2 = fused step
int g4hsim_electron(const G4HB200Tables* t, G4HB200ElectronBatch* b, G4HB200SecondaryQueue* q, uint64_t seed, int mode) {
  const TablesView tv = MakeView(*t);
  for (int64_t i = 0; i < b->n; ++i) {
    ElectronState s;
    Rng rng;
    Secondaries sec;
    sec.n = 0;
    LoadElectron(*b, i, seed, s, rng);
    if (mode == 1) LoadElectronHandOver(*b, i, s);
    if (mode != 1) {
      ResampleNumIALeft(s, rng);
      HowFarToDiscreteInteraction(tv, s);
      HowFarToMSC(tv, s, rng);
    }
    if (mode != 0) {
      ElectronPerform(tv, s, rng, sec);
      AppendSec(q, sec, s.id, i);
    }
    StoreElectron(*b, i, s, rng);
    if (b->mfp01 != nullptr) StoreElectronHandOver(*b, i, s);
  }
  return 0;
}

// the staged HowFar (g4h_stages.cuh), stage after stage over the whole batch like the kernels do
int g4hsim_electron_howfar_staged(const G4HB200Tables* t, G4HB200ElectronBatch* b, uint64_t seed) {
  const TablesView tv = MakeView(*t);
  for (int64_t i = 0; i < b->n; ++i) StageHowFarXS(tv, *b, i, seed);
  for (int64_t i = 0; i < b->n; ++i) StageHowFarMSC<true>(tv, *b, i, seed);
  return 0;
}

// the staged Perform (g4h_perform_stages.cuh), stage after stage over the queues like the kernels do
// fused != 0: the fused step (StageStepHead instead of HowFar + StageAlongStep)
int g4hsim_electron_perform_staged(const G4HB200Tables* t, G4HB200ElectronBatch* b, G4HB200SecondaryQueue* q, uint64_t seed,
                                   int fused) {
  const TablesView tv = MakeView(*t);
  std::vector<double> prestep(2 * static_cast<size_t>(b->n) + 2);
  std::vector<int64_t> queue[kNumElQueues];
  for (int64_t i = 0; i < b->n; ++i) {
    const int r = fused ? StageStepHead(tv, *b, prestep.data(), i, seed, NoGeometryStep{}) : StageAlongStep(tv, *b, prestep.data(), i);
    if (r >= 0) queue[r].push_back(i);
  }
  if (std::getenv("G4HSIM_ROUTES") != nullptr) {
    std::fprintf(stderr, "after head: mscEl=%zu mscPos=%zu fluct=%zu discrete=%zu atRest=%zu\n", queue[kQMscEl].size(),
                 queue[kQMscPos].size(), queue[kQFluct].size(), queue[kQDiscrete].size(), queue[kQAtRest].size());
  }
  const double cbeta1 = MscCBeta1();
  for (int64_t i : queue[kQMscEl]) {
    const int r = StageMSCSample<false>(tv, *b, prestep.data(), i, seed, cbeta1);
    if (r >= 0) queue[r].push_back(i);
  }
  for (int64_t i : queue[kQMscPos]) {
    const int r = StageMSCSample<true>(tv, *b, prestep.data(), i, seed, cbeta1);
    if (r >= 0) queue[r].push_back(i);
  }
  if (std::getenv("G4HSIM_ROUTES") != nullptr) {
    std::fprintf(stderr, "after msc: fluct=%zu discrete=%zu atRest=%zu\n", queue[kQFluct].size(), queue[kQDiscrete].size(),
                 queue[kQAtRest].size());
  }
  for (int64_t i : queue[kQFluct]) {
    const int r = StageFluctuation(tv, *b, prestep.data(), i, seed);
    if (r >= 0) queue[r].push_back(i);
  }
  if (std::getenv("G4HSIM_ROUTES") != nullptr) {
    std::fprintf(stderr, "routes: n=%ld mscEl=%zu mscPos=%zu fluct=%zu discrete=%zu atRest=%zu\n", static_cast<long>(b->n),
                 queue[kQMscEl].size(), queue[kQMscPos].size(), queue[kQFluct].size(), queue[kQDiscrete].size(), queue[kQAtRest].size());
  }
  for (int64_t i : queue[kQDiscrete]) {
    const int r = StageDiscrete(tv, *b, i, seed);
    if (r >= 0) queue[r].push_back(i);
  }
  for (int k : {static_cast<int>(kQMoller), static_cast<int>(kQBhabha), static_cast<int>(kQSB), static_cast<int>(kQRB),
                static_cast<int>(kQAnnih), static_cast<int>(kQAtRest)}) {
    for (int64_t i : queue[k]) {
      Secondaries sec;
      sec.n = 0;
      int id = 0;
      if (k == kQMoller) StageSampler<kQMoller>(tv, *b, i, seed, sec, id);
      if (k == kQBhabha) StageSampler<kQBhabha>(tv, *b, i, seed, sec, id);
      if (k == kQSB) StageSampler<kQSB>(tv, *b, i, seed, sec, id);
      if (k == kQRB) StageSampler<kQRB>(tv, *b, i, seed, sec, id);
      if (k == kQAnnih) StageSampler<kQAnnih>(tv, *b, i, seed, sec, id);
      if (k == kQAtRest) StageSampler<kQAtRest>(tv, *b, i, seed, sec, id);
      AppendSec(q, sec, id, i);
    }
  }
  return 0;
}

// the staged gamma step / Perform (mode 2 / 1): head over the batch, then the three process queues
int g4hsim_gamma_staged(const G4HB200Tables* t, G4HB200GammaBatch* b, G4HB200SecondaryQueue* q, uint64_t seed, int mode) {
  const TablesView tv = MakeView(*t);
  std::vector<int64_t> queue[3];
  for (int64_t i = 0; i < b->n; ++i) {
    const int route = mode == 1 ? StageGammaHead<1>(tv, *b, i, seed, NoGeometryStep{}) : StageGammaHead<2>(tv, *b, i, seed, NoGeometryStep{});
    if (route >= 0) queue[route].push_back(i);
  }
  for (int k = 0; k < 3; ++k) {
    for (int64_t i : queue[k]) {
      Secondaries sec;
      sec.n = 0;
      int id = 0;
      if (k == 0) StageGammaInteract<0>(tv, *b, i, seed, sec, id);
      if (k == 1) StageGammaInteract<1>(tv, *b, i, seed, sec, id);
      if (k == 2) StageGammaInteract<2>(tv, *b, i, seed, sec, id);
      AppendSec(q, sec, id, i);
    }
  }
  return 0;
}

int g4hsim_gamma(const G4HB200Tables* t, G4HB200GammaBatch* b, G4HB200SecondaryQueue* q, uint64_t seed, int mode) {
  const TablesView tv = MakeView(*t);
  for (int64_t i = 0; i < b->n; ++i) {
    GammaState s;
    Rng rng;
    Secondaries sec;
    sec.n = 0;
    LoadGamma(*b, i, seed, s, rng);
    const int flags = b->meta[4 * i + 1];
    if (mode == 1) LoadGammaHandOver(*b, i, s);
    if (mode != 1) GammaHowFar(tv, s, rng);
    if (mode != 0) {
      GammaPerform(tv, s, rng, sec);
      AppendSec(q, sec, s.id, i);
    }
    StoreGamma(*b, i, s, rng, flags);
  }
  return 0;
}

}  // extern "C"
