// capi.cu -- implementation of the C-ABI (include/g4hepem_b200.h): table arena, batch memory and
// kernel launches.  Built for sm_100a only (see __graft_entry__.build()).  No CPU fallback: every
// compute entry point launches a kernel or returns an error.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/g4hepem_b200.h"
#include "g4h_kernels.cuh"
#include "g4h_pipeline.cuh"
#include "g4h_fused.cuh"
#include "g4h_lookups_f32.cuh"
#include "g4h_shower.cuh"
#include "g4h_trackops.cuh"
#include "g4h_view.cuh"

using namespace g4h;

namespace {

thread_local std::string g_lastError;

int Fail(int code, const char* what, cudaError_t err = cudaSuccess) {
  g_lastError = what;
  if (err != cudaSuccess) {
    g_lastError += ": ";
    g_lastError += cudaGetErrorString(err);
  }
  return code;
}

#define G4H_CUDA(call)                                              \
  do {                                                              \
    const cudaError_t err__ = (call);                               \
    if (err__ != cudaSuccess) return Fail(G4HB200_ECUDA, #call, err__); \
  } while (0)

// collects host arrays, packs them into one arena (doubles, then int32s, 16-byte aligned pieces)
struct ArenaBuilder {
  struct Piece {
    const void* src;
    size_t bytes;
    size_t offset;
    const void** dst;  // where the device pointer goes
  };
  std::vector<Piece> pieces;
  size_t total = 0;
  template <class T>
  void Add(const T*& field, size_t count) {
    const T* src = field;
    if (src == nullptr || count == 0) {
      field = nullptr;
      return;
    }
    const size_t bytes = count * sizeof(T);
    pieces.push_back(Piece{src, bytes, total, reinterpret_cast<const void**>(&field)});
    total += (bytes + 255) & ~static_cast<size_t>(255);
  }
};

void AddElectron(ArenaBuilder& ab, G4HB200ElectronTables& e, int numMatCut, int numMat) {
  ab.Add(e.loss_egrid, e.num_loss);
  ab.Add(e.loss_data, static_cast<size_t>(5) * e.num_loss * numMatCut);
  ab.Add(e.resmx_start, numMatCut);
  ab.Add(e.resmx_data, e.num_resmx);
  ab.Add(e.enuc_egrid, 128);
  ab.Add(e.enuc_data, static_cast<size_t>(2) * 128 * numMat);
  ab.Add(e.tr1_data, static_cast<size_t>(2) * e.num_loss * numMat);
  ab.Add(e.sel_ioni_start, numMatCut);
  ab.Add(e.sel_ioni_data, e.num_sel_ioni);
  ab.Add(e.sel_sb_start, numMatCut);
  ab.Add(e.sel_sb_data, e.num_sel_sb);
  ab.Add(e.sel_rb_start, numMatCut);
  ab.Add(e.sel_rb_data, e.num_sel_rb);
}

int GridFor(int64_t n, int smCount, int ctasPerSM, int threads = kThreadsPerBlock) {
  const int64_t want = (n + threads - 1) / threads;
  const int64_t full = static_cast<int64_t>(smCount) * ctasPerSM;
  if (want <= 0) return 1;
  if (want >= full) return static_cast<int>(full);
  // round up to a multiple of the SM count so that every SM gets the same number of CTAs
  const int64_t rounded = ((want + smCount - 1) / smCount) * smCount;
  return static_cast<int>(rounded < full ? rounded : full);
}

}  // namespace

struct G4HB200 {
  int device = 0;
  int smCount = 148;
  void* arena = nullptr;
  size_t arenaBytes = 0;
  G4HB200Tables desc;  // descriptor with device pointers
  TablesView view;
  cudaStream_t stream = nullptr;  // internal stream of the *_host entry points
  int64_t launches = 0;
  // device scratch of the *_host entry points
  G4HB200ElectronBatch elDev;
  G4HB200GammaBatch gmDev;
  G4HB200SecondaryQueue secDev;
  int64_t elCap = 0, gmCap = 0, secCap = 0;
  // workspace of the pipelined Perform (interaction queues, pre-step energies)
  struct WorkSlot {
    ElectronWork work;
    double* steppreMem = nullptr;
    void* mem = nullptr;
    int64_t cap = 0;
    cudaStream_t stream = nullptr;   // chunk stream of the host entry points
    cudaEvent_t counted = nullptr;   // the chunk's secondary count has landed in pinnedCount
    // the final-state samplers work on disjoint queues: they run side by side on these streams (fork / join)
    static constexpr int kNumAux = 5;
    cudaStream_t aux[kNumAux] = {};
    cudaEvent_t fork = nullptr;
    cudaEvent_t join[kNumAux] = {};
  };
  static constexpr int kNumSlots = 4;
  WorkSlot slots[kNumSlots];         // slot 0 also serves the device-batch entry points
  WorkSlot gmSlot;                   // queues of the gamma pipeline
  WorkSlot gmSlot2;                  // ... of the second part-batch (LaunchGammaPipelineHalves)
  int32_t* pinnedCounts = nullptr;   // [kMaxChunks] secondary counts of the chunks of a host call
  int32_t* chunkCounters = nullptr;  // device, [kMaxChunks]
  static constexpr int kMaxChunks = 256;
  // Two ways to run a step: as the pipeline of stage kernels over global queues (g4h_pipeline.cuh; the default) and as ONE
  // persistent launch with CTA-local queues (g4h_fused.cuh) for batches below fusedBelow tracks.  Measured on the B200
  // (tools/size_probe.py, profiles/r02_fused_*): the single launch moves 28 % less through HBM (827 against 1 145 MB per
  // 1M-track step) but is slower at EVERY batch size (256 tracks: 144 against 114 us, 1M: 0.78 against 0.61 ms): a CTA runs
  // its stages one after the other with barriers in between, and each stage costs its dependency latency (~3 000 dependent
  // instructions for the head), while the pipeline runs six samplers side by side and overlaps stages of two half batches.
  // It stays available for A/B runs: G4HB200_FUSED=1 (always), G4HB200_FUSED_BELOW=n (below n tracks).
  int64_t fusedBelow = 0;
  // the stepping loop runs as CUDA graphs with the populations on the device once both populations are below tailBelow
  // tracks (capi_shower.inl: RunGraphTail); G4HB200_GRAPH_TAIL=0 turns it off, G4HB200_TAIL_BELOW=n moves the threshold
  bool graphTail = true;
  int64_t tailBelow = 1 << 20;
  bool Fused(int64_t n) const { return n < fusedBelow; }
  // the single-launch e-/e+ step has no single-precision SampleMSC: with that variant on, the pipeline runs
  bool FusedElectron(int64_t n) const { return n < fusedBelow && !mscF32; }
  // device batches of at least this many tracks run as two half-batch pipelines side by side (G4HB200_SPLIT_MIN)
  int64_t splitThreshold = 1 << 18;
  int splitParts = 2;  // G4HB200_SPLIT_PARTS, at most kNumSlots
  cudaStream_t loopStream = nullptr;  // gamma chain of the stepping loops (capi_shower.inl)
  cudaEvent_t loopFork = nullptr, loopJoin = nullptr;
  cudaEvent_t splitFork = nullptr, splitJoin[4] = {};
  // per-kernel timing (g4hb200_set_kernel_timing): one event row per timed pipeline call
  bool timing = false;
  struct TimedCall {
    cudaEvent_t ev[2 * G4HB200_NUM_STAGES];  // {before, after} per stage
    cudaEvent_t done;
    bool ran[G4HB200_NUM_STAGES];
    int32_t* counts;  // pinned copy of the queue counters of that call
    int64_t n;
  };
  std::vector<TimedCall> timed;
  std::unordered_map<const void*, int> residentCtas;  // per kernel: CTAs of kThreadsPerBlock threads that fit on one SM
  // The table arena (< 1 MB) as a persisting L2 access-policy window on every stream the library launches on: track state
  // streams through L2 at hundreds of MB per step, the tables must not be evicted by it.  G4HB200_L2_PERSIST=0 turns it off.
  // rejection samplers with lane refill (g4h_refill.cuh): 32-entry chunks a warp gets at least; 0: the one-thread-per-track
  // samplers (G4HB200_REFILL)
  int refillChunks = 0;
  bool discreteAside = true;  // ElDiscreteKernel beside ElFluctuationKernel (G4HB200_DISCRETE_ASIDE=0: behind it)
  // SampleMSC in single precision (g4h_msc_f32.cuh), an offered variant: g4hb200_set_msc_precision(h, 32)
  bool mscF32 = false;
  bool l2Persist = false;
  std::unordered_set<cudaStream_t> pinnedStreams;
  void PinTables(cudaStream_t st) {
    if (!l2Persist || st == nullptr || pinnedStreams.count(st) != 0) return;
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr  = arena;
    attr.accessPolicyWindow.num_bytes = arenaBytes;
    attr.accessPolicyWindow.hitRatio  = 1.0f;
    attr.accessPolicyWindow.hitProp   = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp  = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    pinnedStreams.insert(st);
  }
};

namespace {

// One wave: as many CTAs as are resident at once (SM count x occupancy of that kernel), never more than the
// work needs.  Every kernel here is a grid-stride loop, and a CTA costs ~2.5 us of launch + first-load latency
// whatever it does: with the former 8 CTAs per SM a queue kernel over a thousand tracks took 11 us.
template <class K>
int OneWave(G4HB200* h, K kernel, int64_t n, int threads = kThreadsPerBlock) {
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = h->residentCtas.find(key);
  if (it == h->residentCtas.end()) {
    int perSM = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, threads, 0) != cudaSuccess || perSM < 1) perSM = 1;
    it = h->residentCtas.emplace(key, perSM).first;
  }
  return GridFor(n, h->smCount, it->second, threads);
}

// the same for a kernel with dynamic shared memory (the refill samplers, g4h_refill.cuh)
template <class K>
int OneWaveSmem(G4HB200* h, K kernel, int64_t n, size_t smemBytes) {
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = h->residentCtas.find(key);
  if (it == h->residentCtas.end()) {
    int perSM = 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smemBytes)) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, kThreadsPerBlock, smemBytes) != cudaSuccess || perSM < 1) {
      cudaGetLastError();
      perSM = 1;
    }
    it = h->residentCtas.emplace(key, perSM).first;
  }
  return GridFor(n, h->smCount, it->second);
}

template <class T>
int DevAlloc(T*& p, size_t count) {
  void* q = nullptr;
  const cudaError_t err = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
  if (err != cudaSuccess) return Fail(err == cudaErrorMemoryAllocation ? G4HB200_ENOMEM : G4HB200_ECUDA, "cudaMalloc", err);
  p = static_cast<T*>(q);
  return 0;
}

constexpr int kNumElGroups = 17;  // double-pair groups of G4HB200ElectronBatch
void ElectronDoubleGroups(G4HB200ElectronBatch* b, double** out[kNumElGroups]) {
  double** g[kNumElGroups] = {&b->ekin_logekin, &b->dirx_diry, &b->dirz_safety, &b->nia01, &b->nia23, &b->msc_irange_dynrf,
                              &b->msc_tlimmin_gauss, &b->gstep_pstep, &b->edep_dispx, &b->dispy_dispz, &b->mfp01, &b->mfp23,
                              &b->range_lambtr1, &b->tstep_zpath, &b->par12, &b->par3_pad, &b->prestep};
  for (int i = 0; i < kNumElGroups; ++i) out[i] = g[i];
}

void GammaDoubleGroups(G4HB200GammaBatch* b, double** out[5]) {
  double** g[5] = {&b->ekin_logekin, &b->dirx_diry, &b->dirz_nia0, &b->gstep_mfp0, &b->edep_pemxsec};
  for (int i = 0; i < 5; ++i) out[i] = g[i];
}

int CopyGroup(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
  if (dst == nullptr || src == nullptr || bytes == 0) return 0;
  G4H_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, st));
  return 0;
}

int CopyElectron(const G4HB200ElectronBatch* from, G4HB200ElectronBatch* to, cudaMemcpyKind kind, cudaStream_t st,
                 int firstGroup, int lastGroup, bool withMeta, bool withWinner) {
  const int64_t n = from->n;
  double** gf[kNumElGroups];
  double** gt[kNumElGroups];
  ElectronDoubleGroups(const_cast<G4HB200ElectronBatch*>(from), gf);
  ElectronDoubleGroups(to, gt);
  for (int i = firstGroup; i < lastGroup; ++i) {
    const int rc = CopyGroup(*gt[i], *gf[i], static_cast<size_t>(n) * 16, kind, st);
    if (rc != 0) return rc;
  }
  if (withMeta) {
    const int rc = CopyGroup(to->meta, from->meta, static_cast<size_t>(n) * 16, kind, st);
    if (rc != 0) return rc;
  }
  if (withWinner) {
    const int rc = CopyGroup(to->winner, from->winner, static_cast<size_t>(n) * 4, kind, st);
    if (rc != 0) return rc;
  }
  to->n = n;
  return 0;
}

int CopyGamma(const G4HB200GammaBatch* from, G4HB200GammaBatch* to, cudaMemcpyKind kind, cudaStream_t st, int firstGroup,
              int lastGroup, bool withMeta, bool withWinner) {
  const int64_t n = from->n;
  double** gf[5];
  double** gt[5];
  GammaDoubleGroups(const_cast<G4HB200GammaBatch*>(from), gf);
  GammaDoubleGroups(to, gt);
  for (int i = firstGroup; i < lastGroup; ++i) {
    const int rc = CopyGroup(*gt[i], *gf[i], static_cast<size_t>(n) * 16, kind, st);
    if (rc != 0) return rc;
  }
  if (withMeta) {
    const int rc = CopyGroup(to->meta, from->meta, static_cast<size_t>(n) * 16, kind, st);
    if (rc != 0) return rc;
  }
  if (withWinner) {
    const int rc = CopyGroup(to->winner, from->winner, static_cast<size_t>(n) * 4, kind, st);
    if (rc != 0) return rc;
  }
  to->n = n;
  return 0;
}

int CheckHandle(G4HB200* h) {
  if (h == nullptr) return Fail(G4HB200_EINVAL, "null handle");
  G4H_CUDA(cudaSetDevice(h->device));
  return 0;
}

G4HB200SecondaryQueue NullQueue() {
  G4HB200SecondaryQueue q;
  std::memset(&q, 0, sizeof(q));
  return q;
}

int LaunchGammaHowFar(G4HB200* h, G4HB200GammaBatch* dev, uint64_t seed, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || dev->n < 0) return Fail(G4HB200_EINVAL, "bad gamma batch");
  if (dev->n == 0) return 0;
  h->PinTables(static_cast<cudaStream_t>(stream));
  GammaHowFarKernel<<<OneWave(h, GammaHowFarKernel, dev->n), kThreadsPerBlock, 0, static_cast<cudaStream_t>(stream)>>>(h->view, *dev, seed);
  ++h->launches;
  G4H_CUDA(cudaGetLastError());
  return 0;
}

// interaction queues + pre-step energies for n tracks: one allocation, carved up
int EnsureElectronWork(G4HB200::WorkSlot& slot, int64_t n) {
  if (n <= slot.cap) return 0;
  if (slot.mem != nullptr) {
    G4H_CUDA(cudaDeviceSynchronize());
    cudaFree(slot.mem);
    slot.mem = nullptr;
    slot.cap = 0;
  }
  const size_t cap = static_cast<size_t>((n + 255) & ~static_cast<int64_t>(255));
  const size_t bytes = 2 * cap * 16 + static_cast<size_t>(kNumElQueues) * cap * 4 + 256;
  const cudaError_t err = cudaMalloc(&slot.mem, bytes);
  if (err != cudaSuccess) return Fail(G4HB200_ENOMEM, "cudaMalloc(workspace)", err);
  unsigned char* p = static_cast<unsigned char*>(slot.mem);
  slot.work.prestep = reinterpret_cast<double*>(p);
  p += cap * 16;
  slot.work.steppre = nullptr;  // set by the stepping loop only (MSC sub-steps: the energy at the beginning of the whole step)
  slot.steppreMem   = reinterpret_cast<double*>(p);
  p += cap * 16;
  for (int k = 0; k < kNumElQueues; ++k) {
    slot.work.queue[k] = reinterpret_cast<int32_t*>(p);
    p += cap * 4;
  }
  slot.work.count = reinterpret_cast<int32_t*>(p);
  slot.cap = static_cast<int64_t>(cap);
  return 0;
}

// pipeline stages of the e-/e+ step, in launch order (g4h_pipeline.cuh)
enum ElStage {
  kSHowFarXS = 0, kSHowFarMSC, kSAlongStep, kSStepHead, kSMscEl, kSMscPos, kSFluct, kSDiscrete, kSMoller, kSBhabha, kSSB, kSRB,
  kSAnnih, kSAtRest, kSGammaHead, kSGammaConversion, kSGammaCompton, kSGammaPhotoelectric, kSElFused, kSGammaFused,
  kNumElStages
};
static_assert(kNumElStages <= G4HB200_NUM_STAGES, "G4HB200_NUM_STAGES too small");
// pipeline stage -> queue that feeds it (-1: every track of the batch)
const int kStageQueue[G4HB200_NUM_STAGES] = {-1, -1, -1, -1, kQMscEl, kQMscPos, kQFluct, kQDiscrete, kQMoller, kQBhabha, kQSB,
                                             kQRB, kQAnnih, kQAtRest, -1, kGQConversion, kGQCompton, kGQPhotoelectric, -1, -1};
const char* const kStageName[G4HB200_NUM_STAGES] = {
    "ElHowFarXSKernel", "ElHowFarMSCKernel", "ElAlongStepKernel", "ElStepHeadKernel", "ElMSCSampleKernel<e->",
    "ElMSCSampleKernel<e+>", "ElFluctuationKernel", "ElDiscreteKernel",
    "ElSamplerKernel<Moller>", "ElSamplerKernel<Bhabha>", "ElSamplerKernel<SeltzerBerger>", "ElSamplerKernel<RelBrem>",
    "ElSamplerKernel<Annihilation>", "ElSamplerKernel<AtRest>", "GammaHeadKernel", "GammaInteractKernel<Conversion>",
    "GammaInteractKernel<Compton>", "GammaInteractKernel<Photoelectric>", "ElFusedStepKernel", "GammaFusedStepKernel"};

struct StageTimer {
  G4HB200* h;
  cudaStream_t st;
  G4HB200::TimedCall* tc = nullptr;
  cudaError_t Begin(int64_t n) {
    if (!h->timing) return cudaSuccess;
    h->timed.emplace_back();
    tc = &h->timed.back();
    tc->n = n;
    tc->counts = nullptr;
    for (auto& r : tc->ran) r = false;
    for (auto& e : tc->ev) {
      const cudaError_t err = cudaEventCreate(&e);
      if (err != cudaSuccess) return err;
    }
    cudaError_t err = cudaEventCreate(&tc->done);
    if (err != cudaSuccess) return err;
    err = cudaMallocHost(reinterpret_cast<void**>(&tc->counts), kNumElQueues * sizeof(int32_t));
    if (err != cudaSuccess) return err;
    for (int k = 0; k < kNumElQueues; ++k) tc->counts[k] = 0;
    return cudaSuccess;
  }
  // bracket one launch: Before(stage) ... kernel ... After(stage)
  cudaError_t Before(int stage) { return Before(stage, st); }
  cudaError_t After(int stage) { return After(stage, st); }
  cudaError_t Before(int stage, cudaStream_t on) { return tc != nullptr ? cudaEventRecord(tc->ev[2 * stage], on) : cudaSuccess; }
  cudaError_t After(int stage, cudaStream_t on) {
    ++h->launches;
    if (tc == nullptr) return cudaGetLastError();
    tc->ran[stage] = true;
    return cudaEventRecord(tc->ev[2 * stage + 1], on);
  }
};

// G4HepEmElectronManager::HowFar as two kernels (g4h_stages.cuh)
template <bool kStoreResults>
int LaunchHowFarStages(G4HB200* h, G4HB200ElectronBatch* dev, const ElectronWork& w, uint64_t seed, cudaStream_t st,
                       StageTimer& t) {
  const int64_t n = dev->n;
  G4H_CUDA(t.Before(kSHowFarXS));
  ElHowFarXSKernel<<<OneWave(h, ElHowFarXSKernel, n), kThreadsPerBlock, 0, st>>>(h->view, *dev, seed);
  G4H_CUDA(t.After(kSHowFarXS));
  G4H_CUDA(t.Before(kSHowFarMSC));
  ElHowFarMSCKernel<kStoreResults><<<OneWave(h, ElHowFarMSCKernel<kStoreResults>, n), kThreadsPerBlock, 0, st>>>(h->view, *dev, seed);
  G4H_CUDA(t.After(kSHowFarMSC));
  return 0;
}

int LaunchElectronHowFar(G4HB200* h, G4HB200ElectronBatch* dev, uint64_t seed, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || dev->n < 0) return Fail(G4HB200_EINVAL, "bad electron batch");
  if (dev->n == 0) return 0;
  if (dev->n > 0x7fffffff) return Fail(G4HB200_EINVAL, "batch too large (track indices are 32 bit)");
  if ((rc = EnsureElectronWork(h->slots[0], dev->n)) != 0) return rc;
  const ElectronWork& w = h->slots[0].work;
  StageTimer t{h, static_cast<cudaStream_t>(stream)};
  h->PinTables(t.st);
  G4H_CUDA(cudaMemsetAsync(w.count, 0, kNumElQueues * sizeof(int32_t), t.st));
  rc = LaunchHowFarStages<true>(h, dev, w, seed, t.st, t);
  if (rc != 0) return rc;
  G4H_CUDA(cudaGetLastError());
  return 0;
}

int EnsureAuxStreams(G4HB200::WorkSlot& slot);

// G4HepEmElectronManager::Perform as a pipeline (g4h_pipeline.cuh); kFused: HowFar first
// geometry step inside the fused head (the stepping loop over the slab calorimeter, capi_shower.inl)
struct SlabHead {
  SlabGeom g;
  TrackGeo geo;
};

template <bool kFused>
int LaunchElectronPipeline(G4HB200* h, G4HB200ElectronBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream,
                           int slotIndex = 0, const SlabHead* slab = nullptr) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || dev->n < 0) return Fail(G4HB200_EINVAL, "bad electron batch");
  if (sec == nullptr) return Fail(G4HB200_EINVAL, "secondary queue required");
  if (dev->n == 0) return 0;
  if (dev->n > 0x7fffffff) return Fail(G4HB200_EINVAL, "batch too large (track indices are 32 bit)");
  if ((rc = EnsureElectronWork(h->slots[slotIndex], dev->n)) != 0) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  h->PinTables(st);
  const int64_t n = dev->n;
  ElectronWork w = h->slots[slotIndex].work;
  if (slab != nullptr) w.steppre = h->slots[slotIndex].steppreMem;
  G4H_CUDA(cudaMemsetAsync(w.count, 0, kNumElQueues * sizeof(int32_t), st));
  StageTimer t{h, st};
  G4H_CUDA(t.Begin(dev->n));
#define G4H_STAGE(stage, ...)       \
  G4H_CUDA(t.Before(stage));        \
  __VA_ARGS__;                      \
  G4H_CUDA(t.After(stage))
  G4HB200::WorkSlot& slot = h->slots[slotIndex];
  if ((rc = EnsureAuxStreams(slot)) != 0) return rc;
  for (auto& a : slot.aux) h->PinTables(a);
  if (kFused && slab != nullptr) {
    G4H_STAGE(kSStepHead, ShowerElectronHeadKernel<<<OneWave(h, ShowerElectronHeadKernel, n), kThreadsPerBlock, 0, st>>>(
                              h->view, *dev, w, seed, slab->g, slab->geo));
  } else if (kFused) {
    G4H_STAGE(kSStepHead, ElStepHeadKernel<<<OneWave(h, ElStepHeadKernel, n), kThreadsPerBlock, 0, st>>>(h->view, *dev, w, seed));
  } else {
    G4H_STAGE(kSAlongStep, ElAlongStepKernel<<<OneWave(h, ElAlongStepKernel, n), kThreadsPerBlock, 0, st>>>(h->view, *dev, w));
  }
  // the two particle types are scattered side by side (disjoint queues and tracks).  With per-kernel timing on,
  // every kernel runs alone on the caller's stream instead: the CUDA-event durations are then those of the kernels
  // themselves (comparable with an ncu launch list), not of kernels waiting for each other's SMs.
  const bool alone = t.tc != nullptr;
  cudaStream_t side[G4HB200::WorkSlot::kNumAux];
  for (int k = 0; k < G4HB200::WorkSlot::kNumAux; ++k) side[k] = alone ? st : slot.aux[k];
  if (!alone) {
    G4H_CUDA(cudaEventRecord(slot.fork, st));
    G4H_CUDA(cudaStreamWaitEvent(slot.aux[0], slot.fork, 0));
  }
  if (h->mscF32) {
    G4H_STAGE(kSMscEl, ElMSCSampleF32Kernel<false><<<OneWave(h, ElMSCSampleF32Kernel<false>, n), kThreadsPerBlock, 0, st>>>(h->view, *dev, w, seed));
    G4H_CUDA(t.Before(kSMscPos, side[0]));
    ElMSCSampleF32Kernel<true><<<OneWave(h, ElMSCSampleF32Kernel<true>, n), kThreadsPerBlock, 0, side[0]>>>(h->view, *dev, w, seed);
    G4H_CUDA(t.After(kSMscPos, side[0]));
  } else {
    G4H_STAGE(kSMscEl, ElMSCSampleKernel<false><<<OneWave(h, ElMSCSampleKernel<false>, n), kThreadsPerBlock, 0, st>>>(h->view, *dev, w, seed));
    G4H_CUDA(t.Before(kSMscPos, side[0]));
    ElMSCSampleKernel<true><<<OneWave(h, ElMSCSampleKernel<true>, n), kThreadsPerBlock, 0, side[0]>>>(h->view, *dev, w, seed);
    G4H_CUDA(t.After(kSMscPos, side[0]));
  }
  if (!alone) {
    G4H_CUDA(cudaEventRecord(slot.join[0], slot.aux[0]));
    G4H_CUDA(cudaStreamWaitEvent(st, slot.join[0], 0));
  }
  // the fluctuation sampler and the head of PerformDiscrete for the tracks that skip it read different queues and only meet
  // in the (atomic) counters of the sampler queues: side by side
  if (!alone && h->discreteAside) {
    G4H_CUDA(cudaEventRecord(slot.fork, st));
    G4H_CUDA(cudaStreamWaitEvent(slot.aux[1], slot.fork, 0));
  }
  cudaStream_t sd = (alone || !h->discreteAside) ? st : slot.aux[1];
  G4H_STAGE(kSFluct, ElFluctuationKernel<<<OneWave(h, ElFluctuationKernel, n), kThreadsPerBlock, 0, st>>>(h->view, *dev, w, seed));
  G4H_CUDA(t.Before(kSDiscrete, sd));
  ElDiscreteKernel<<<OneWave(h, ElDiscreteKernel, n), kThreadsPerBlock, 0, sd>>>(h->view, *dev, w, seed);
  G4H_CUDA(t.After(kSDiscrete, sd));
  if (sd != st) {
    G4H_CUDA(cudaEventRecord(slot.join[1], sd));
    G4H_CUDA(cudaStreamWaitEvent(st, slot.join[1], 0));
  }
#undef G4H_STAGE
  // fork: the six samplers read disjoint queues and write disjoint tracks (+ atomic appends of secondaries)
  if (!alone) {
    G4H_CUDA(cudaEventRecord(slot.fork, st));
    for (int k = 0; k < G4HB200::WorkSlot::kNumAux; ++k) G4H_CUDA(cudaStreamWaitEvent(slot.aux[k], slot.fork, 0));
  }
#define G4H_STAGE(stage, on, ...)   \
  G4H_CUDA(t.Before(stage, on));    \
  __VA_ARGS__;                      \
  G4H_CUDA(t.After(stage, on))
  const int rc4 = h->refillChunks;
#define G4H_EL_SAMPLER(stage, on, Q)                                                                                                     \
  if (rc4 > 0) {                                                                                                                         \
    constexpr size_t smem = RefillSmemBytes<ElRefillSampler<Q>::type>();                                                                 \
    G4H_STAGE(stage, on, ElRefillSamplerKernel<Q><<<OneWaveSmem(h, ElRefillSamplerKernel<Q>, n, smem), kThreadsPerBlock, smem, on>>>(   \
                             h->view, *dev, w, *sec, seed, rc4));                                                                        \
  } else {                                                                                                                               \
    G4H_STAGE(stage, on, ElSamplerKernel<Q><<<OneWave(h, ElSamplerKernel<Q>, n), kThreadsPerBlock, 0, on>>>(h->view, *dev, w, *sec, seed)); \
  }
  G4H_EL_SAMPLER(kSRB, st, kQRB)
  G4H_EL_SAMPLER(kSSB, side[0], kQSB)
  G4H_EL_SAMPLER(kSBhabha, side[1], kQBhabha)
  G4H_EL_SAMPLER(kSMoller, side[2], kQMoller)
#undef G4H_EL_SAMPLER
  G4H_STAGE(kSAnnih, side[3], ElSamplerKernel<kQAnnih><<<OneWave(h, ElSamplerKernel<kQAnnih>, n), kThreadsPerBlock, 0, side[3]>>>(h->view, *dev, w, *sec, seed));
  G4H_STAGE(kSAtRest, side[4], ElSamplerKernel<kQAtRest><<<OneWave(h, ElSamplerKernel<kQAtRest>, n), kThreadsPerBlock, 0, side[4]>>>(h->view, *dev, w, *sec, seed));
  if (!alone) {
    for (int k = 0; k < G4HB200::WorkSlot::kNumAux; ++k) {
      G4H_CUDA(cudaEventRecord(slot.join[k], slot.aux[k]));
      G4H_CUDA(cudaStreamWaitEvent(st, slot.join[k], 0));
    }
  }
#undef G4H_STAGE
  if (t.tc != nullptr) {
    G4H_CUDA(cudaMemcpyAsync(t.tc->counts, w.count, kNumElQueues * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    G4H_CUDA(cudaEventRecord(t.tc->done, st));
  }
  G4H_CUDA(cudaGetLastError());
  return 0;
}

// fork / join streams of a work slot (created on first use)
int EnsureAuxStreams(G4HB200::WorkSlot& slot) {
  if (slot.fork != nullptr) return 0;
  G4H_CUDA(cudaEventCreateWithFlags(&slot.fork, cudaEventDisableTiming));
  for (int k = 0; k < G4HB200::WorkSlot::kNumAux; ++k) {
    G4H_CUDA(cudaStreamCreateWithFlags(&slot.aux[k], cudaStreamNonBlocking));
    G4H_CUDA(cudaEventCreateWithFlags(&slot.join[k], cudaEventDisableTiming));
  }
  return 0;
}

// G4HepEmGammaManager::[HowFar +] SelectInteraction + Perform as a pipeline: head over every track, then the three
// final state samplers side by side over their queues (g4h_pipeline.cuh); kMode 1: Perform, 2: fused step
template <int kMode>
int LaunchGammaPipeline(G4HB200* h, G4HB200GammaBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream,
                        bool secondSlot = false, const SlabHead* slab = nullptr) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || dev->n < 0) return Fail(G4HB200_EINVAL, "bad gamma batch");
  if (sec == nullptr) return Fail(G4HB200_EINVAL, "secondary queue required");
  if (dev->n == 0) return 0;
  if (dev->n > 0x7fffffff) return Fail(G4HB200_EINVAL, "batch too large (track indices are 32 bit)");
  G4HB200::WorkSlot& slot = secondSlot ? h->gmSlot2 : h->gmSlot;
  if ((rc = EnsureElectronWork(slot, dev->n)) != 0) return rc;
  if ((rc = EnsureAuxStreams(slot)) != 0) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  h->PinTables(st);
  for (auto& a : slot.aux) h->PinTables(a);
  const int64_t n = dev->n;
  const ElectronWork& w = slot.work;
  G4H_CUDA(cudaMemsetAsync(w.count, 0, kNumElQueues * sizeof(int32_t), st));
  StageTimer t{h, st};
  G4H_CUDA(t.Begin(n));
  G4H_CUDA(t.Before(kSGammaHead));
  if (kMode == 2 && slab != nullptr) {
    ShowerGammaHeadKernel<<<OneWave(h, ShowerGammaHeadKernel, n, kGammaThreads), kGammaThreads, 0, st>>>(h->view, *dev, w, seed, slab->g, slab->geo);
  } else {
    GammaHeadKernel<kMode><<<OneWave(h, GammaHeadKernel<kMode>, n, kGammaThreads), kGammaThreads, 0, st>>>(h->view, *dev, w, seed);
  }
  G4H_CUDA(t.After(kSGammaHead));
  // the three samplers side by side; alone on the caller's stream when per-kernel timing is on
  const bool alone = t.tc != nullptr;
  cudaStream_t side[2] = {alone ? st : slot.aux[0], alone ? st : slot.aux[1]};
  if (!alone) {
    G4H_CUDA(cudaEventRecord(slot.fork, st));
    for (int k = 0; k < 2; ++k) G4H_CUDA(cudaStreamWaitEvent(slot.aux[k], slot.fork, 0));
  }
  const int rc4 = h->refillChunks;
#define G4H_GM_SAMPLER(stage, on, P)                                                                                                  \
  G4H_CUDA(t.Before(stage, on));                                                                                                      \
  if (rc4 > 0) {                                                                                                                      \
    constexpr size_t smem = RefillSmemBytes<GammaRefillSampler<P>::type>();                                                           \
    GammaRefillKernel<P><<<OneWaveSmem(h, GammaRefillKernel<P>, n, smem), kThreadsPerBlock, smem, on>>>(h->view, *dev, w, *sec, seed, rc4); \
  } else {                                                                                                                            \
    GammaInteractKernel<P><<<OneWave(h, GammaInteractKernel<P>, n, kGammaThreads), kGammaThreads, 0, on>>>(h->view, *dev, w, *sec, seed); \
  }                                                                                                                                   \
  G4H_CUDA(t.After(stage, on));
  G4H_GM_SAMPLER(kSGammaCompton, st, kGQCompton)
  G4H_GM_SAMPLER(kSGammaConversion, side[0], kGQConversion)
  G4H_GM_SAMPLER(kSGammaPhotoelectric, side[1], kGQPhotoelectric)
#undef G4H_GM_SAMPLER
  if (!alone) {
    for (int k = 0; k < 2; ++k) {
      G4H_CUDA(cudaEventRecord(slot.join[k], slot.aux[k]));
      G4H_CUDA(cudaStreamWaitEvent(st, slot.join[k], 0));
    }
  }
  if (t.tc != nullptr) {
    G4H_CUDA(cudaMemcpyAsync(t.tc->counts, w.count, kNumElQueues * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    G4H_CUDA(cudaEventRecord(t.tc->done, st));
  }
  G4H_CUDA(cudaGetLastError());
  return 0;
}


// A sub-range of a batch (same struct, pointers advanced).
G4HB200ElectronBatch ElectronBatchView(const G4HB200ElectronBatch& full, int64_t lo, int64_t len);
G4HB200GammaBatch GammaBatchView(const G4HB200GammaBatch& full, int64_t lo, int64_t len);

// One wave of the fused kernels: SM count x occupancy CTAs, never more than there are tiles
template <class K>
int FusedGrid(G4HB200* h, K kernel, int64_t n) {
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = h->residentCtas.find(key);
  if (it == h->residentCtas.end()) {
    int perSM = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, kThreadsPerBlock, 0) != cudaSuccess || perSM < 1) perSM = 1;
    it = h->residentCtas.emplace(key, perSM).first;
  }
  const int64_t tiles = (n + kThreadsPerBlock - 1) / kThreadsPerBlock;
  const int64_t full  = static_cast<int64_t>(h->smCount) * it->second;
  return static_cast<int>(tiles < full ? (tiles > 0 ? tiles : 1) : full);
}

// The e-/e+ step (kPerformOnly: Perform alone) as one persistent launch (g4h_fused.cuh); a batch of more tiles than
// one wave can index with 16-bit queue entries (19M tracks on a B200) goes in several launches.
template <bool kPerformOnly>
int LaunchElectronFused(G4HB200* h, G4HB200ElectronBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream,
                        const SlabHead* slab = nullptr, int slotIndex = 0) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || dev->n < 0) return Fail(G4HB200_EINVAL, "bad electron batch");
  if (sec == nullptr) return Fail(G4HB200_EINVAL, "secondary queue required");
  if (dev->n == 0) return 0;
  if (dev->n > 0x7fffffff) return Fail(G4HB200_EINVAL, "batch too large (track indices are 32 bit)");
  if ((rc = EnsureElectronWork(h->slots[slotIndex], dev->n)) != 0) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StageTimer t{h, st};
  G4H_CUDA(t.Begin(dev->n));
  const int fullGrid = FusedGrid(h, ElFusedStepKernel<kPerformOnly>, dev->n);
  const int64_t perLaunch = static_cast<int64_t>(fullGrid) * kFusedMaxTilesPerCta * kThreadsPerBlock;
  G4H_CUDA(t.Before(kSElFused));
  for (int64_t lo = 0; lo < dev->n; lo += perLaunch) {
    const int64_t len = dev->n - lo < perLaunch ? dev->n - lo : perLaunch;
    G4HB200ElectronBatch part = ElectronBatchView(*dev, lo, len);
    G4HB200SecondaryQueue q = *sec;
    q.parent_base = sec->parent_base + static_cast<int32_t>(lo);
    double* prestep = h->slots[slotIndex].work.prestep + 2 * lo;
    double* steppre = h->slots[slotIndex].steppreMem + 2 * lo;
    if (slab != nullptr) {
      TrackGeo geo = slab->geo;
      geo.posx_posy += 2 * lo;
      geo.posz_pad += 2 * lo;
      geo.vol += lo;
      geo.nextVol += lo;
      geo.sub_left_eloss += 2 * lo;
      geo.sub_pre += 2 * lo;
      geo.sub_range_proc += 2 * lo;
      ShowerElectronFusedKernel<<<FusedGrid(h, ShowerElectronFusedKernel, len), kThreadsPerBlock, 0, st>>>(h->view, part, prestep, steppre, q,
                                                                                                       seed, slab->g, geo);
    } else {
      ElFusedStepKernel<kPerformOnly><<<FusedGrid(h, ElFusedStepKernel<kPerformOnly>, len), kThreadsPerBlock, 0, st>>>(h->view, part, prestep,
                                                                                                                 q, seed);
    }
    ++h->launches;
  }
  --h->launches;  // After() counts one
  G4H_CUDA(t.After(kSElFused));
  if (t.tc != nullptr) G4H_CUDA(cudaEventRecord(t.tc->done, st));
  G4H_CUDA(cudaGetLastError());
  return 0;
}

template <int kMode>
int LaunchGammaFused(G4HB200* h, G4HB200GammaBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream,
                     const SlabHead* slab = nullptr) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || dev->n < 0) return Fail(G4HB200_EINVAL, "bad gamma batch");
  if (sec == nullptr) return Fail(G4HB200_EINVAL, "secondary queue required");
  if (dev->n == 0) return 0;
  if (dev->n > 0x7fffffff) return Fail(G4HB200_EINVAL, "batch too large (track indices are 32 bit)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StageTimer t{h, st};
  G4H_CUDA(t.Begin(dev->n));
  const int fullGrid = FusedGrid(h, GammaFusedStepKernel<kMode>, dev->n);
  const int64_t perLaunch = static_cast<int64_t>(fullGrid) * kFusedMaxTilesPerCta * kThreadsPerBlock;
  G4H_CUDA(t.Before(kSGammaFused));
  for (int64_t lo = 0; lo < dev->n; lo += perLaunch) {
    const int64_t len = dev->n - lo < perLaunch ? dev->n - lo : perLaunch;
    G4HB200GammaBatch part = GammaBatchView(*dev, lo, len);
    G4HB200SecondaryQueue q = *sec;
    q.parent_base = sec->parent_base + static_cast<int32_t>(lo);
    if (kMode == 2 && slab != nullptr) {
      TrackGeo geo = slab->geo;
      geo.posx_posy += 2 * lo;
      geo.posz_pad += 2 * lo;
      geo.vol += lo;
      geo.nextVol += lo;
      ShowerGammaFusedKernel<<<FusedGrid(h, ShowerGammaFusedKernel, len), kThreadsPerBlock, 0, st>>>(h->view, part, q, seed, slab->g, geo);
    } else {
      GammaFusedStepKernel<kMode><<<FusedGrid(h, GammaFusedStepKernel<kMode>, len), kThreadsPerBlock, 0, st>>>(h->view, part, q, seed);
    }
    ++h->launches;
  }
  --h->launches;
  G4H_CUDA(t.After(kSGammaFused));
  if (t.tc != nullptr) G4H_CUDA(cudaEventRecord(t.tc->done, st));
  G4H_CUDA(cudaGetLastError());
  return 0;
}

// A sub-range of a batch (same struct, pointers advanced).
G4HB200ElectronBatch ElectronBatchView(const G4HB200ElectronBatch& full, int64_t lo, int64_t len) {
  G4HB200ElectronBatch v = full;
  double** g[kNumElGroups];
  ElectronDoubleGroups(&v, g);
  for (int k = 0; k < kNumElGroups; ++k) if (*g[k] != nullptr) *g[k] += 2 * lo;
  if (v.meta != nullptr) v.meta += 4 * lo;
  if (v.winner != nullptr) v.winner += lo;
  v.n = len;
  return v;
}

// The pipeline of a large device batch as a few part-batch pipelines side by side (the caller's stream and internal
// streams, fork / join by events).  The queue kernels at the end of a pipeline (discrete head, final state
// samplers: 20-35 % of the issue slots, bound by gather latency and rejection loops) leave most of the machine idle;
// with two halves in flight they run next to the other half's arithmetic-bound head / MSC / fluctuation kernels.
// Tracks are independent and the uniform stream is keyed per track, so the result does not depend on the split.
template <bool kFused>
int LaunchElectronPipelineHalves(G4HB200* h, G4HB200ElectronBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream,
                                 const SlabHead* slab = nullptr) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || sec == nullptr || dev->n < h->splitThreshold || h->timing || h->splitParts < 2) {
    return LaunchElectronPipeline<kFused>(h, dev, sec, seed, stream, 0, slab);
  }
  if (dev->n > 0x7fffffff) return Fail(G4HB200_EINVAL, "batch too large (track indices are 32 bit)");
  const int parts = h->splitParts;
  for (int p = 1; p < parts; ++p) {
    G4HB200::WorkSlot& other = h->slots[p];
    if (other.stream == nullptr) {
      G4H_CUDA(cudaStreamCreateWithFlags(&other.stream, cudaStreamNonBlocking));
      G4H_CUDA(cudaEventCreateWithFlags(&other.counted, cudaEventDisableTiming));
    }
  }
  if (h->splitFork == nullptr) {
    G4H_CUDA(cudaEventCreateWithFlags(&h->splitFork, cudaEventDisableTiming));
    for (auto& e : h->splitJoin) G4H_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t per = ((dev->n / parts + kThreadsPerBlock - 1) / kThreadsPerBlock) * kThreadsPerBlock;
  G4H_CUDA(cudaEventRecord(h->splitFork, st));
  for (int p = 1; p < parts; ++p) G4H_CUDA(cudaStreamWaitEvent(h->slots[p].stream, h->splitFork, 0));
  for (int p = 0; p < parts; ++p) {
    const int64_t lo = p * per;
    if (lo >= dev->n) break;
    const int64_t len = (p == parts - 1 || lo + per > dev->n) ? dev->n - lo : per;
    G4HB200ElectronBatch part = ElectronBatchView(*dev, lo, len);
    G4HB200SecondaryQueue q = *sec;
    q.parent_base = sec->parent_base + static_cast<int32_t>(lo);
    cudaStream_t ps = p == 0 ? st : h->slots[p].stream;
    SlabHead partSlab;
    if (slab != nullptr) {
      partSlab = *slab;
      partSlab.geo.posx_posy += 2 * lo;
      partSlab.geo.posz_pad += 2 * lo;
      partSlab.geo.vol += lo;
      partSlab.geo.nextVol += lo;
      partSlab.geo.sub_left_eloss += 2 * lo;
      partSlab.geo.sub_pre += 2 * lo;
      partSlab.geo.sub_range_proc += 2 * lo;
    }
    if ((rc = LaunchElectronPipeline<kFused>(h, &part, &q, seed, ps, p, slab != nullptr ? &partSlab : nullptr)) != 0) return rc;
  }
  for (int p = 1; p < parts; ++p) {
    G4H_CUDA(cudaEventRecord(h->splitJoin[p], h->slots[p].stream));
    G4H_CUDA(cudaStreamWaitEvent(st, h->splitJoin[p], 0));
  }
  return 0;
}

G4HB200GammaBatch GammaBatchView(const G4HB200GammaBatch& full, int64_t lo, int64_t len) {
  G4HB200GammaBatch v = full;
  double** g[5];
  GammaDoubleGroups(&v, g);
  for (int k = 0; k < 5; ++k) if (*g[k] != nullptr) *g[k] += 2 * lo;
  if (v.meta != nullptr) v.meta += 4 * lo;
  if (v.winner != nullptr) v.winner += lo;
  v.n = len;
  return v;
}

// the gamma pipeline of a large device batch as two half-batch pipelines side by side (see the e-/e+ one above)
template <int kMode>
int LaunchGammaPipelineHalves(G4HB200* h, G4HB200GammaBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream,
                              const SlabHead* slab = nullptr) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || sec == nullptr || dev->n < h->splitThreshold || h->timing || h->splitParts < 2) {
    return LaunchGammaPipeline<kMode>(h, dev, sec, seed, stream, false, slab);
  }
  if (dev->n > 0x7fffffff) return Fail(G4HB200_EINVAL, "batch too large (track indices are 32 bit)");
  G4HB200::WorkSlot& other = h->gmSlot2;
  if (other.stream == nullptr) {
    G4H_CUDA(cudaStreamCreateWithFlags(&other.stream, cudaStreamNonBlocking));
    G4H_CUDA(cudaEventCreateWithFlags(&other.counted, cudaEventDisableTiming));
  }
  if (h->splitFork == nullptr) {
    G4H_CUDA(cudaEventCreateWithFlags(&h->splitFork, cudaEventDisableTiming));
    for (auto& e : h->splitJoin) G4H_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n0 = ((dev->n / 2 + kThreadsPerBlock - 1) / kThreadsPerBlock) * kThreadsPerBlock;
  G4HB200GammaBatch first  = GammaBatchView(*dev, 0, n0);
  G4HB200GammaBatch second = GammaBatchView(*dev, n0, dev->n - n0);
  G4HB200SecondaryQueue q1 = *sec;
  q1.parent_base = sec->parent_base + static_cast<int32_t>(n0);
  G4H_CUDA(cudaEventRecord(h->splitFork, st));
  G4H_CUDA(cudaStreamWaitEvent(other.stream, h->splitFork, 0));
  SlabHead slab1;
  if (slab != nullptr) {
    slab1 = *slab;
    slab1.geo.posx_posy += 2 * n0;
    slab1.geo.posz_pad += 2 * n0;
    slab1.geo.vol += n0;
    slab1.geo.nextVol += n0;
  }
  if ((rc = LaunchGammaPipeline<kMode>(h, &first, sec, seed, st, false, slab)) != 0) return rc;
  if ((rc = LaunchGammaPipeline<kMode>(h, &second, &q1, seed, other.stream, true, slab != nullptr ? &slab1 : nullptr)) != 0) return rc;
  G4H_CUDA(cudaEventRecord(h->splitJoin[0], other.stream));
  G4H_CUDA(cudaStreamWaitEvent(st, h->splitJoin[0], 0));
  return 0;
}

}  // namespace

namespace {
template <int kOp>
int LaunchElectronTrackOp(G4HB200* h, G4HB200ElectronBatch* dev, const G4HB200SecondaryQueue& q, uint64_t seed, int32_t* flag,
                          cudaStream_t st) {
  ElTrackOpKernel<kOp><<<OneWave(h, ElTrackOpKernel<kOp>, dev->n), kThreadsPerBlock, 0, st>>>(h->view, *dev, q, seed, flag);
  ++h->launches;
  G4H_CUDA(cudaGetLastError());
  return 0;
}
template <int kOp>
int LaunchGammaTrackOp(G4HB200* h, G4HB200GammaBatch* dev, const G4HB200SecondaryQueue& q, uint64_t seed, cudaStream_t st) {
  GammaTrackOpKernel<kOp><<<OneWave(h, GammaTrackOpKernel<kOp>, dev->n), kThreadsPerBlock, 0, st>>>(h->view, *dev, q, seed);
  ++h->launches;
  G4H_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

extern "C" {

const char* g4hb200_last_error(void) { return g_lastError.c_str(); }

int g4hb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int g4hb200_create(const G4HB200Tables* tables, int device, G4HB200** out) {
  if (tables == nullptr || out == nullptr) return Fail(G4HB200_EINVAL, "null argument");
  if (tables->num_matcut <= 0 || tables->num_mat <= 0 || tables->num_regions <= 0)
    return Fail(G4HB200_EINVAL, "empty table set");
  if (tables->electron.num_loss < 2 || tables->positron.num_loss < 2) return Fail(G4HB200_EINVAL, "bad e-loss grid");
  for (int i = 0; i < tables->num_matcut; ++i) {
    if (tables->mc_imat[i] < 0 || tables->mc_imat[i] >= tables->num_mat || tables->mc_ireg[i] < 0 ||
        tables->mc_ireg[i] >= tables->num_regions)
      return Fail(G4HB200_EINVAL, "couple refers to an unknown material / region");
  }
  if (g4hb200_device_count() <= 0) return Fail(G4HB200_ENODEVICE, "no CUDA device");
  G4H_CUDA(cudaSetDevice(device));
  G4HB200* h = new (std::nothrow) G4HB200;
  if (h == nullptr) return Fail(G4HB200_ENOMEM, "host allocation");
  h->device = device;
  {
    int smCount = 0;
    const cudaError_t err = cudaDeviceGetAttribute(&smCount, cudaDevAttrMultiProcessorCount, device);
    if (err != cudaSuccess) {
      delete h;
      return Fail(G4HB200_ECUDA, "cudaDeviceGetAttribute", err);
    }
    h->smCount = smCount;
  }
  h->desc = *tables;
  G4HB200Tables& d = h->desc;
  ArenaBuilder ab;
  const int nmc = d.num_matcut, nmat = d.num_mat;
  int nElemTot = 0;
  for (int i = 0; i < nmat; ++i) nElemTot += tables->mat_num_elem[i];
  ab.Add(d.region_pars, static_cast<size_t>(8) * d.num_regions);
  ab.Add(d.mc_cuts, static_cast<size_t>(4) * nmc);
  ab.Add(d.mc_imat, nmc);
  ab.Add(d.mc_ireg, nmc);
  ab.Add(d.mat_num_elem, nmat);
  ab.Add(d.mat_elem_start, nmat);
  ab.Add(d.mat_elem_z, nElemTot);
  ab.Add(d.mat_elem_natoms, nElemTot);
  ab.Add(d.mat_pars, static_cast<size_t>(16) * nmat);
  ab.Add(d.mat_sandia_num, nmat);
  ab.Add(d.mat_sandia_start, nmat);
  ab.Add(d.elem_pars, static_cast<size_t>(12) * 121);
  ab.Add(d.elem_sandia_num, 121);
  ab.Add(d.elem_sandia_start, 121);
  ab.Add(d.sandia_energies, d.num_sandia);
  ab.Add(d.sandia_cof, static_cast<size_t>(4) * d.num_sandia);
  AddElectron(ab, d.electron, nmc, nmat);
  AddElectron(ab, d.positron, nmc, nmat);
  ab.Add(d.sb_el_energy, 65);
  ab.Add(d.sb_lel_energy, 65);
  ab.Add(d.sb_lkappa, 54);
  ab.Add(d.sb_gcut_start, nmc);
  ab.Add(d.sb_gcut_indices, d.num_sb_gcut);
  ab.Add(d.sb_start_per_z, 121);
  ab.Add(d.sb_data, d.num_sb_data);
  ab.Add(d.gm_mxsec, static_cast<size_t>(nmat) * d.gm_data_per_mat);
  ab.Add(d.gm_conv_start, nmat);
  ab.Add(d.gm_conv_egrid, d.gm_conv_egrid_size);
  ab.Add(d.gm_conv_data, d.num_gm_conv);
  // one allocation, one staged copy
  h->arenaBytes = ab.total + 256;
  {
    const cudaError_t err = cudaMalloc(&h->arena, h->arenaBytes);
    if (err != cudaSuccess) {
      delete h;
      return Fail(G4HB200_ENOMEM, "cudaMalloc(arena)", err);
    }
  }
  std::vector<unsigned char> staging(h->arenaBytes, 0);
  for (const auto& p : ab.pieces) {
    std::memcpy(staging.data() + p.offset, p.src, p.bytes);
    *p.dst = static_cast<unsigned char*>(h->arena) + p.offset;
  }
  {
    const cudaError_t err = cudaMemcpy(h->arena, staging.data(), h->arenaBytes, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
      cudaFree(h->arena);
      delete h;
      return Fail(G4HB200_ECUDA, "cudaMemcpy(arena)", err);
    }
  }
  h->view = MakeView(d);
  {
    const char* env = std::getenv("G4HB200_L2_PERSIST");
    int maxPersist = 0;
    if (!(env != nullptr && env[0] == '0') &&
        cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, device) == cudaSuccess && maxPersist > 0) {
      const size_t want = h->arenaBytes < static_cast<size_t>(maxPersist) ? h->arenaBytes : static_cast<size_t>(maxPersist);
      size_t have = 0;
      cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize);
      if (have >= want || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) h->l2Persist = true;
      cudaGetLastError();
    }
  }
  if (const char* da = std::getenv("G4HB200_DISCRETE_ASIDE")) h->discreteAside = da[0] != '0';
  if (const char* mf = std::getenv("G4HB200_MSC_F32")) h->mscF32 = mf[0] == '1';
  if (const char* rf = std::getenv("G4HB200_REFILL")) h->refillChunks = std::atoi(rf) < 0 ? 0 : std::atoi(rf);
  {
    if (const char* fu = std::getenv("G4HB200_FUSED")) h->fusedBelow = fu[0] != '0' ? (int64_t{1} << 62) : 0;
    if (const char* fb = std::getenv("G4HB200_FUSED_BELOW")) h->fusedBelow = std::atoll(fb);
    if (const char* gt = std::getenv("G4HB200_GRAPH_TAIL")) h->graphTail = gt[0] != '0';
    if (const char* tb = std::getenv("G4HB200_TAIL_BELOW")) {
      const long long v = std::atoll(tb);
      if (v >= 256) h->tailBelow = v;
    }
    if (const char* sp = std::getenv("G4HB200_SPLIT_MIN")) h->splitThreshold = std::atoll(sp);
    if (const char* sp = std::getenv("G4HB200_SPLIT_PARTS")) {
      const int v = std::atoi(sp);
      if (v >= 1 && v <= G4HB200::kNumSlots) h->splitParts = v;
    }
  }
  {
    const cudaError_t err = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) {
      cudaFree(h->arena);
      delete h;
      return Fail(G4HB200_ECUDA, "cudaStreamCreateWithFlags", err);
    }
  }
  std::memset(&h->elDev, 0, sizeof(h->elDev));
  std::memset(&h->gmDev, 0, sizeof(h->gmDev));
  std::memset(&h->secDev, 0, sizeof(h->secDev));
  *out = h;
  return 0;
}

int g4hb200_destroy(G4HB200* h) {
  if (h == nullptr) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  // per-kernel timing rows nobody collected (g4hb200_kernel_times)
  for (auto& tc : h->timed) {
    for (auto& e : tc.ev) cudaEventDestroy(e);
    cudaEventDestroy(tc.done);
    if (tc.counts != nullptr) cudaFreeHost(tc.counts);
  }
  h->timed.clear();
  if (h->elCap > 0) g4hb200_electron_batch_free(h, &h->elDev);
  if (h->gmCap > 0) g4hb200_gamma_batch_free(h, &h->gmDev);
  if (h->secCap > 0) g4hb200_secondary_queue_free(h, &h->secDev);
  for (G4HB200::WorkSlot* sp : {&h->slots[0], &h->slots[1], &h->slots[2], &h->slots[3], &h->gmSlot, &h->gmSlot2}) {
    G4HB200::WorkSlot& slot = *sp;
    if (slot.mem != nullptr) cudaFree(slot.mem);
    if (slot.stream != nullptr) cudaStreamDestroy(slot.stream);
    if (slot.counted != nullptr) cudaEventDestroy(slot.counted);
    for (auto& a : slot.aux) if (a != nullptr) cudaStreamDestroy(a);
    for (auto& e : slot.join) if (e != nullptr) cudaEventDestroy(e);
    if (slot.fork != nullptr) cudaEventDestroy(slot.fork);
  }
  if (h->loopStream != nullptr) {
    cudaStreamDestroy(h->loopStream);
    cudaEventDestroy(h->loopFork);
    cudaEventDestroy(h->loopJoin);
  }
  if (h->splitFork != nullptr) {
    cudaEventDestroy(h->splitFork);
    for (auto& e : h->splitJoin) cudaEventDestroy(e);
  }
  if (h->pinnedCounts != nullptr) cudaFreeHost(h->pinnedCounts);
  if (h->chunkCounters != nullptr) cudaFree(h->chunkCounters);
  if (h->stream != nullptr) cudaStreamDestroy(h->stream);
  if (h->arena != nullptr) cudaFree(h->arena);
  delete h;
  return 0;
}

int g4hb200_electron_batch_alloc(G4HB200* h, int64_t capacity, G4HB200ElectronBatch* out) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (out == nullptr || capacity < 0) return Fail(G4HB200_EINVAL, "bad argument");
  std::memset(out, 0, sizeof(*out));
  double** g[kNumElGroups];
  ElectronDoubleGroups(out, g);
  for (int i = 0; i < kNumElGroups; ++i) {
    rc = DevAlloc(*g[i], static_cast<size_t>(capacity) * 2);
    if (rc != 0) return rc;
  }
  rc = DevAlloc(out->meta, static_cast<size_t>(capacity) * 4);
  if (rc != 0) return rc;
  rc = DevAlloc(out->winner, static_cast<size_t>(capacity));
  if (rc != 0) return rc;
  out->n = 0;
  return 0;
}

int g4hb200_electron_batch_free(G4HB200* h, G4HB200ElectronBatch* dev) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr) return 0;
  double** g[kNumElGroups];
  ElectronDoubleGroups(dev, g);
  for (int i = 0; i < kNumElGroups; ++i) cudaFree(*g[i]);
  cudaFree(dev->meta);
  cudaFree(dev->winner);
  std::memset(dev, 0, sizeof(*dev));
  return 0;
}

int g4hb200_gamma_batch_alloc(G4HB200* h, int64_t capacity, G4HB200GammaBatch* out) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (out == nullptr || capacity < 0) return Fail(G4HB200_EINVAL, "bad argument");
  std::memset(out, 0, sizeof(*out));
  double** g[5];
  GammaDoubleGroups(out, g);
  for (int i = 0; i < 5; ++i) {
    rc = DevAlloc(*g[i], static_cast<size_t>(capacity) * 2);
    if (rc != 0) return rc;
  }
  rc = DevAlloc(out->meta, static_cast<size_t>(capacity) * 4);
  if (rc != 0) return rc;
  rc = DevAlloc(out->winner, static_cast<size_t>(capacity));
  if (rc != 0) return rc;
  out->n = 0;
  return 0;
}

int g4hb200_gamma_batch_free(G4HB200* h, G4HB200GammaBatch* dev) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr) return 0;
  double** g[5];
  GammaDoubleGroups(dev, g);
  for (int i = 0; i < 5; ++i) cudaFree(*g[i]);
  cudaFree(dev->meta);
  cudaFree(dev->winner);
  std::memset(dev, 0, sizeof(*dev));
  return 0;
}

int g4hb200_secondary_queue_alloc(G4HB200* h, int64_t capacity, G4HB200SecondaryQueue* out) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (out == nullptr || capacity < 0) return Fail(G4HB200_EINVAL, "bad argument");
  std::memset(out, 0, sizeof(*out));
  out->capacity = capacity;
  if ((rc = DevAlloc(out->dirx_diry, static_cast<size_t>(capacity) * 2)) != 0) return rc;
  if ((rc = DevAlloc(out->dirz_ekin, static_cast<size_t>(capacity) * 2)) != 0) return rc;
  if ((rc = DevAlloc(out->parent_kind, static_cast<size_t>(capacity) * 2)) != 0) return rc;
  if ((rc = DevAlloc(out->parent_slot, static_cast<size_t>(capacity) * 2)) != 0) return rc;
  if ((rc = DevAlloc(out->count, 4)) != 0) return rc;
  G4H_CUDA(cudaMemset(out->count, 0, 16));
  return 0;
}

int g4hb200_secondary_queue_free(G4HB200* h, G4HB200SecondaryQueue* dev) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr) return 0;
  cudaFree(dev->dirx_diry);
  cudaFree(dev->dirz_ekin);
  cudaFree(dev->parent_kind);
  cudaFree(dev->parent_slot);
  cudaFree(dev->count);
  std::memset(dev, 0, sizeof(*dev));
  return 0;
}

int g4hb200_secondary_queue_reset(G4HB200* h, G4HB200SecondaryQueue* dev, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || dev->count == nullptr) return Fail(G4HB200_EINVAL, "bad queue");
  G4H_CUDA(cudaMemsetAsync(dev->count, 0, sizeof(int32_t), static_cast<cudaStream_t>(stream)));
  return 0;
}

int g4hb200_electron_batch_upload(G4HB200* h, const G4HB200ElectronBatch* host, G4HB200ElectronBatch* dev, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (host == nullptr || dev == nullptr) return Fail(G4HB200_EINVAL, "null batch");
  return CopyElectron(host, dev, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream), 0, kNumElGroups, true, true);
}

int g4hb200_electron_batch_download(G4HB200* h, const G4HB200ElectronBatch* dev, G4HB200ElectronBatch* host, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (host == nullptr || dev == nullptr) return Fail(G4HB200_EINVAL, "null batch");
  return CopyElectron(dev, host, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream), 0, kNumElGroups, true, true);
}

int g4hb200_gamma_batch_upload(G4HB200* h, const G4HB200GammaBatch* host, G4HB200GammaBatch* dev, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (host == nullptr || dev == nullptr) return Fail(G4HB200_EINVAL, "null batch");
  return CopyGamma(host, dev, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream), 0, 5, true, true);
}

int g4hb200_gamma_batch_download(G4HB200* h, const G4HB200GammaBatch* dev, G4HB200GammaBatch* host, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (host == nullptr || dev == nullptr) return Fail(G4HB200_EINVAL, "null batch");
  return CopyGamma(dev, host, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream), 0, 5, true, true);
}

int g4hb200_secondary_queue_download(G4HB200* h, const G4HB200SecondaryQueue* dev, G4HB200SecondaryQueue* host, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (host == nullptr || dev == nullptr) return Fail(G4HB200_EINVAL, "null queue");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  G4H_CUDA(cudaMemcpyAsync(host->count, dev->count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  G4H_CUDA(cudaStreamSynchronize(st));
  int64_t n = host->count[0];
  if (n > dev->capacity) {
    host->count[0] = static_cast<int32_t>(dev->capacity);
    return Fail(G4HB200_ECAPACITY, "secondary queue overflow on the device");
  }
  if (n > host->capacity) return Fail(G4HB200_ECAPACITY, "host secondary queue too small");
  G4H_CUDA(cudaMemcpyAsync(host->dirx_diry, dev->dirx_diry, static_cast<size_t>(n) * 16, cudaMemcpyDeviceToHost, st));
  G4H_CUDA(cudaMemcpyAsync(host->dirz_ekin, dev->dirz_ekin, static_cast<size_t>(n) * 16, cudaMemcpyDeviceToHost, st));
  G4H_CUDA(cudaMemcpyAsync(host->parent_kind, dev->parent_kind, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, st));
  G4H_CUDA(cudaMemcpyAsync(host->parent_slot, dev->parent_slot, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, st));
  return 0;
}

int g4hb200_sync(G4HB200* h, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  G4H_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return 0;
}

int g4hb200_device_alloc(G4HB200* h, size_t bytes, void** out) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (out == nullptr) return Fail(G4HB200_EINVAL, "null argument");
  const cudaError_t err = cudaMalloc(out, bytes > 0 ? bytes : 16);
  if (err != cudaSuccess) return Fail(err == cudaErrorMemoryAllocation ? G4HB200_ENOMEM : G4HB200_ECUDA, "cudaMalloc", err);
  return 0;
}

int g4hb200_device_free(G4HB200* h, void* p) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (p != nullptr) G4H_CUDA(cudaFree(p));
  return 0;
}

int g4hb200_memcpy(G4HB200* h, void* dst, const void* src, size_t bytes, int to_device, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (bytes == 0) return 0;
  if (dst == nullptr || src == nullptr) return Fail(G4HB200_EINVAL, "null argument");
  G4H_CUDA(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  return 0;
}

int g4hb200_electron_lookups(G4HB200* h, int64_t n, const int32_t* imc, const double* ekin, const double* logekin,
                             int is_electron, double* out, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (n < 0 || (n > 0 && (!imc || !ekin || !logekin || !out))) return Fail(G4HB200_EINVAL, "bad argument");
  if (n == 0) return 0;
  const ElectronTablesView& ed = h->view.el[is_electron ? 0 : 1];
  const size_t hot = reinterpret_cast<const char*>(ed.tr1Data + 2 * ed.numLoss * h->view.numMat) - reinterpret_cast<const char*>(ed.lossEGrid) + 16;
  // the hot tables of one particle (66 KB for six couples) staged in shared memory, one 1024-thread CTA per SM: 7 %
  // faster than gathering them through L1 (0.0583 -> 0.0546 ms per 1M look-up sets; G4HB200_LOOKUPS_SMEM=0: the L1 kernel)
  const char* smemEnv = std::getenv("G4HB200_LOOKUPS_SMEM");
  if (!(smemEnv != nullptr && smemEnv[0] == '0') && hot <= 200 * 1024) {
    cudaFuncSetAttribute(ElectronLookupsSmemKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hot));
    ElectronLookupsSmemKernel<<<h->smCount, 1024, hot, static_cast<cudaStream_t>(stream)>>>(h->view, n, imc, ekin, logekin,
                                                                                            is_electron ? 0 : 1, out);
  } else {
    ElectronLookupsKernel<<<OneWave(h, ElectronLookupsKernel, n), kThreadsPerBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        h->view, n, imc, ekin, logekin, is_electron ? 0 : 1, out);
  }
  ++h->launches;
  G4H_CUDA(cudaGetLastError());
  return 0;
}

int g4hb200_electron_lookups_f32(G4HB200* h, int64_t n, const int32_t* imc, const float* ekin, const float* logekin,
                                 int is_electron, float* out, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (n < 0 || (n > 0 && (!imc || !ekin || !logekin || !out))) return Fail(G4HB200_EINVAL, "bad argument");
  if (n == 0) return 0;
  const ElectronTablesView& ed = h->view.el[is_electron ? 0 : 1];
  LookupsF32Layout lay;
  lay.lossEGrid = 0;
  lay.lossData  = static_cast<int>(ed.lossData - ed.lossEGrid);
  lay.resData   = static_cast<int>(ed.resData - ed.lossEGrid);
  lay.enucEGrid = static_cast<int>(ed.enucEGrid - ed.lossEGrid);
  lay.enucData  = static_cast<int>(ed.enucData - ed.lossEGrid);
  lay.tr1Data   = static_cast<int>(ed.tr1Data - ed.lossEGrid);
  lay.total     = lay.tr1Data + 2 * ed.numLoss * h->view.numMat;
  const size_t bytes = (static_cast<size_t>(lay.total) + h->view.numMatCut) * sizeof(float);
  if (bytes > 200 * 1024) return Fail(G4HB200_EINVAL, "table set too large for the shared-memory look-up kernel");
  G4H_CUDA(cudaFuncSetAttribute(ElectronLookupsF32Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  ElectronLookupsF32Kernel<<<h->smCount, 1024, bytes, static_cast<cudaStream_t>(stream)>>>(h->view, lay, n, imc, ekin, logekin,
                                                                                         is_electron ? 0 : 1, out);
  ++h->launches;
  G4H_CUDA(cudaGetLastError());
  return 0;
}

int g4hb200_electron_stepping_xsecs(G4HB200* h, int64_t n, const int32_t* imc, const double* ekin, const double* logekin,
                                    int is_electron, double* out, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (n < 0 || (n > 0 && (!imc || !ekin || !logekin || !out))) return Fail(G4HB200_EINVAL, "bad argument");
  if (n == 0) return 0;
  ElectronSteppingXSecsKernel<<<OneWave(h, ElectronSteppingXSecsKernel, n), kThreadsPerBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      h->view, n, imc, ekin, logekin, is_electron ? 0 : 1, out);
  ++h->launches;
  G4H_CUDA(cudaGetLastError());
  return 0;
}

int g4hb200_gamma_lookups(G4HB200* h, int64_t n, const int32_t* imc, const double* ekin, const double* logekin,
                          const double* urnd, double* out_mxsec, int32_t* out_pid, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (n < 0 || (n > 0 && (!imc || !ekin || !logekin || !urnd || !out_mxsec || !out_pid))) return Fail(G4HB200_EINVAL, "bad argument");
  if (n == 0) return 0;
  GammaLookupsKernel<<<OneWave(h, GammaLookupsKernel, n), kThreadsPerBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      h->view, n, imc, ekin, logekin, urnd, out_mxsec, out_pid);
  ++h->launches;
  G4H_CUDA(cudaGetLastError());
  return 0;
}

int g4hb200_select_target_element(G4HB200* h, int kind, int is_electron, int64_t n, const int32_t* imc,
                                  const double* ekin, const double* logekin, const double* urnd, int32_t* out_elem,
                                  void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (kind < 0 || kind > 2 || n < 0 || (n > 0 && (!imc || !ekin || !logekin || !urnd || !out_elem)))
    return Fail(G4HB200_EINVAL, "bad argument");
  if (n == 0) return 0;
  SelectTargetElementKernel<<<OneWave(h, SelectTargetElementKernel, n), kThreadsPerBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      h->view, kind, is_electron ? 0 : 1, n, imc, ekin, logekin, urnd, out_elem);
  ++h->launches;
  G4H_CUDA(cudaGetLastError());
  return 0;
}

int g4hb200_vdt_log_exp(G4HB200* h, int64_t n, const double* x, double* out_log, double* out_exp, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (n < 0 || (n > 0 && (!x || !out_log || !out_exp))) return Fail(G4HB200_EINVAL, "bad argument");
  if (n == 0) return 0;
  VdtLogExpKernel<<<OneWave(h, VdtLogExpKernel, n), kThreadsPerBlock, 0, static_cast<cudaStream_t>(stream)>>>(n, x, out_log, out_exp);
  ++h->launches;
  G4H_CUDA(cudaGetLastError());
  return 0;
}

int g4hb200_rng_uniforms(G4HB200* h, uint64_t seed, int64_t n, const int32_t* track_id, int32_t ndraw, double* out,
                         void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (n < 0 || ndraw < 0 || (n > 0 && (!track_id || !out))) return Fail(G4HB200_EINVAL, "bad argument");
  if (n == 0 || ndraw == 0) return 0;
  RngUniformsKernel<<<OneWave(h, RngUniformsKernel, n), kThreadsPerBlock, 0, static_cast<cudaStream_t>(stream)>>>(seed, n, track_id, ndraw, out);
  ++h->launches;
  G4H_CUDA(cudaGetLastError());
  return 0;
}

int g4hb200_electron_howfar(G4HB200* h, G4HB200ElectronBatch* dev, uint64_t seed, void* stream) {
  return LaunchElectronHowFar(h, dev, seed, stream);
}
int g4hb200_electron_perform(G4HB200* h, G4HB200ElectronBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream) {
  if (h != nullptr && dev != nullptr && h->FusedElectron(dev->n)) return LaunchElectronFused<true>(h, dev, sec, seed, stream);
  return LaunchElectronPipelineHalves<false>(h, dev, sec, seed, stream);
}
int g4hb200_electron_step(G4HB200* h, G4HB200ElectronBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream) {
  if (h != nullptr && dev != nullptr && h->FusedElectron(dev->n)) return LaunchElectronFused<false>(h, dev, sec, seed, stream);
  return LaunchElectronPipelineHalves<true>(h, dev, sec, seed, stream);
}
int g4hb200_gamma_howfar(G4HB200* h, G4HB200GammaBatch* dev, uint64_t seed, void* stream) {
  return LaunchGammaHowFar(h, dev, seed, stream);
}
int g4hb200_gamma_perform(G4HB200* h, G4HB200GammaBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream) {
  if (h != nullptr && dev != nullptr && h->Fused(dev->n)) return LaunchGammaFused<1>(h, dev, sec, seed, stream);
  return LaunchGammaPipelineHalves<1>(h, dev, sec, seed, stream);
}
int g4hb200_gamma_step(G4HB200* h, G4HB200GammaBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream) {
  if (h != nullptr && dev != nullptr && h->Fused(dev->n)) return LaunchGammaFused<2>(h, dev, sec, seed, stream);
  return LaunchGammaPipelineHalves<2>(h, dev, sec, seed, stream);
}


int g4hb200_electron_track_op(G4HB200* h, int op, G4HB200ElectronBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed,
                              int32_t* out_flag, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || dev->n < 0) return Fail(G4HB200_EINVAL, "bad electron batch");
  if (op < 0 || op >= kNumElTrackOps) return Fail(G4HB200_EINVAL, "unknown track-level op");
  if ((op == kOpDiscrete || op == kOpAnnihilateAtRest) && sec == nullptr) return Fail(G4HB200_EINVAL, "secondary queue required");
  if (dev->mfp01 == nullptr || dev->mfp23 == nullptr || dev->range_lambtr1 == nullptr || dev->tstep_zpath == nullptr ||
      dev->par12 == nullptr || dev->par3_pad == nullptr)
    return Fail(G4HB200_EINVAL, "the track-level calls need the hand-over groups of the batch");
  if (dev->n == 0) return 0;
  const G4HB200SecondaryQueue q = sec != nullptr ? *sec : NullQueue();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (op) {
    case kOpHowFarDiscrete: return LaunchElectronTrackOp<kOpHowFarDiscrete>(h, dev, q, seed, out_flag, st);
    case kOpHowFarMSC: return LaunchElectronTrackOp<kOpHowFarMSC>(h, dev, q, seed, out_flag, st);
    case kOpUpdatePStep: return LaunchElectronTrackOp<kOpUpdatePStep>(h, dev, q, seed, out_flag, st);
    case kOpUpdateNIA: return LaunchElectronTrackOp<kOpUpdateNIA>(h, dev, q, seed, out_flag, st);
    case kOpMeanELoss: return LaunchElectronTrackOp<kOpMeanELoss>(h, dev, q, seed, out_flag, st);
    case kOpSampleMSC: return LaunchElectronTrackOp<kOpSampleMSC>(h, dev, q, seed, out_flag, st);
    case kOpLossFluct: return LaunchElectronTrackOp<kOpLossFluct>(h, dev, q, seed, out_flag, st);
    case kOpDiscrete: return LaunchElectronTrackOp<kOpDiscrete>(h, dev, q, seed, out_flag, st);
    case kOpAnnihilateAtRest: return LaunchElectronTrackOp<kOpAnnihilateAtRest>(h, dev, q, seed, out_flag, st);
    case kOpPerformContinuous: return LaunchElectronTrackOp<kOpPerformContinuous>(h, dev, q, seed, out_flag, st);
    default: return LaunchElectronTrackOp<kOpResampleNIA>(h, dev, q, seed, out_flag, st);
  }
}

int g4hb200_electron_check_delta(G4HB200* h, G4HB200ElectronBatch* dev, const double* urnd, int32_t* out_flag, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || dev->n < 0 || dev->mfp01 == nullptr || dev->mfp23 == nullptr) return Fail(G4HB200_EINVAL, "bad electron batch");
  if (dev->n > 0 && (urnd == nullptr || out_flag == nullptr)) return Fail(G4HB200_EINVAL, "null argument");
  if (dev->n == 0) return 0;
  ElCheckDeltaKernel<<<OneWave(h, ElCheckDeltaKernel, dev->n), kThreadsPerBlock, 0, static_cast<cudaStream_t>(stream)>>>(h->view, *dev, urnd,
                                                                                                                         out_flag);
  ++h->launches;
  G4H_CUDA(cudaGetLastError());
  return 0;
}

int g4hb200_gamma_track_op(G4HB200* h, int op, G4HB200GammaBatch* dev, G4HB200SecondaryQueue* sec, uint64_t seed, void* stream) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (dev == nullptr || dev->n < 0) return Fail(G4HB200_EINVAL, "bad gamma batch");
  if (op < 0 || op >= kNumGmTrackOps) return Fail(G4HB200_EINVAL, "unknown track-level op");
  if (op == kGOpPerformSelected && sec == nullptr) return Fail(G4HB200_EINVAL, "secondary queue required");
  if (dev->n == 0) return 0;
  const G4HB200SecondaryQueue q = sec != nullptr ? *sec : NullQueue();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (op) {
    case kGOpHowFarTrack: return LaunchGammaTrackOp<kGOpHowFarTrack>(h, dev, q, seed, st);
    case kGOpUpdateNIA: return LaunchGammaTrackOp<kGOpUpdateNIA>(h, dev, q, seed, st);
    case kGOpSelectInteraction: return LaunchGammaTrackOp<kGOpSelectInteraction>(h, dev, q, seed, st);
    default: return LaunchGammaTrackOp<kGOpPerformSelected>(h, dev, q, seed, st);
  }
}

// Host buffers in, host buffers out.  The batch is cut into chunks that travel on kNumSlots streams: while chunk c
// computes, chunk c+1 uploads and chunk c-1 downloads, so the two PCIe directions and the kernels overlap.  Every
// chunk has its own region of the device secondary queue (2 records per track, its own counter); the regions are
// copied back to back into the caller's queue, parent indices already rebased by the kernels (parent_base).
int g4hb200_electron_step_host(G4HB200* h, G4HB200ElectronBatch* host, G4HB200SecondaryQueue* hostSec, uint64_t seed) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (host == nullptr || hostSec == nullptr) return Fail(G4HB200_EINVAL, "null argument");
  const int64_t n = host->n;
  if (n < 0 || n > 0x3fffffff) return Fail(G4HB200_EINVAL, "bad batch size");
  if (hostSec->count == nullptr) return Fail(G4HB200_EINVAL, "secondary queue without a counter");
  hostSec->count[0] = 0;
  if (n == 0) return 0;
  // a step creates at most two secondaries per track: checked BEFORE anything is uploaded or launched, so that a queue
  // that is too small never costs the caller the step (the primaries would already have lost their energy)
  if (hostSec->capacity < 2 * n) return Fail(G4HB200_ECAPACITY, "host secondary queue must hold 2 records per track");
  if (n > h->elCap) {
    if (h->elCap > 0) g4hb200_electron_batch_free(h, &h->elDev);
    h->elCap = 0;
    if ((rc = g4hb200_electron_batch_alloc(h, n, &h->elDev)) != 0) return rc;
    h->elCap = n;
  }
  if (2 * n > h->secCap) {
    if (h->secCap > 0) g4hb200_secondary_queue_free(h, &h->secDev);
    h->secCap = 0;
    if ((rc = g4hb200_secondary_queue_alloc(h, 2 * n, &h->secDev)) != 0) return rc;
    h->secCap = 2 * n;
  }
  if (h->pinnedCounts == nullptr) {
    G4H_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->pinnedCounts), G4HB200::kMaxChunks * sizeof(int32_t)));
    G4H_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->chunkCounters), G4HB200::kMaxChunks * sizeof(int32_t)));
  }
  for (auto& slot : h->slots) {
    if (slot.stream == nullptr) {
      G4H_CUDA(cudaStreamCreateWithFlags(&slot.stream, cudaStreamNonBlocking));
      G4H_CUDA(cudaEventCreateWithFlags(&slot.counted, cudaEventDisableTiming));
    }
  }
  // chunk size: about a quarter of the batch, between 64k and 256k tracks (measured on the B200 box for 1M tracks:
  // 32k 7.7 ms, 64k 7.0, 128k 5.9, 256k 5.7, 512k 6.3, unchunked 7.5 -- small chunks are bound by the ~3.5 us the
  // host needs to enqueue each of the ~35 operations of a chunk, large ones overlap too little)
  int64_t chunk = ((n / 4 + 32767) / 32768) * 32768;
  if (chunk < 65536) chunk = 65536;
  if (chunk > 262144) chunk = 262144;
  if (const char* env = std::getenv("G4HB200_HOST_CHUNK")) {
    const long v = std::atol(env);
    if (v >= 1024) chunk = v;
  }
  while ((n + chunk - 1) / chunk > G4HB200::kMaxChunks) chunk *= 2;
  // (equal chunks: a ramp of small first chunks, meant to start the device -> host direction earlier, measured 3 % slower --
  // more chunks are more operations to enqueue, and the host's enqueue rate is what limits this call, profiles/r02b_ab2.log)
  const int numChunks = static_cast<int>((n + chunk - 1) / chunk);
  G4H_CUDA(cudaMemsetAsync(h->chunkCounters, 0, numChunks * sizeof(int32_t), h->slots[0].stream));
  G4H_CUDA(cudaStreamSynchronize(h->slots[0].stream));
  std::vector<cudaEvent_t> counted(numChunks);
  auto view = ElectronBatchView;
  for (int c = 0; c < numChunks; ++c) {
    G4HB200::WorkSlot& slot = h->slots[c % G4HB200::kNumSlots];
    const int64_t lo = c * chunk, len = (lo + chunk <= n) ? chunk : n - lo;
    G4HB200ElectronBatch hv = view(*host, lo, len);
    G4HB200ElectronBatch dv = view(h->elDev, lo, len);
    G4HB200SecondaryQueue q = h->secDev;
    q.capacity = 2 * len;
    q.dirx_diry += 4 * lo;
    q.dirz_ekin += 4 * lo;
    q.parent_kind += 4 * lo;
    q.parent_slot += 4 * lo;
    q.count = h->chunkCounters + c;
    q.parent_base = static_cast<int32_t>(lo);
    // H2D: the 7 persistent groups + meta (128 B / track)
    if ((rc = CopyElectron(&hv, &dv, cudaMemcpyHostToDevice, slot.stream, 0, 7, true, false)) != 0) return rc;
    if ((rc = h->FusedElectron(len) ? LaunchElectronFused<false>(h, &dv, &q, seed, slot.stream, nullptr, c % G4HB200::kNumSlots)
                       : LaunchElectronPipeline<true>(h, &dv, &q, seed, slot.stream, c % G4HB200::kNumSlots)) != 0)
      return rc;
    // D2H: persistent + result groups + meta + winner (180 B / track)
    if ((rc = CopyElectron(&dv, &hv, cudaMemcpyDeviceToHost, slot.stream, 0, 10, true, true)) != 0) return rc;
    G4H_CUDA(cudaMemcpyAsync(h->pinnedCounts + c, h->chunkCounters + c, sizeof(int32_t), cudaMemcpyDeviceToHost, slot.stream));
    G4H_CUDA(cudaEventCreateWithFlags(&counted[c], cudaEventDisableTiming));
    G4H_CUDA(cudaEventRecord(counted[c], slot.stream));
  }
  // secondaries: as soon as a chunk's count is known, its records follow on the chunk's stream
  int64_t total = 0;
  int rcSec = 0;
  for (int c = 0; c < numChunks; ++c) {
    G4HB200::WorkSlot& slot = h->slots[c % G4HB200::kNumSlots];
    G4H_CUDA(cudaEventSynchronize(counted[c]));
    cudaEventDestroy(counted[c]);
    const int64_t lo = c * chunk;
    const int64_t cnt = h->pinnedCounts[c];
    if (rcSec != 0) continue;
    if (total + cnt > hostSec->capacity) {
      rcSec = Fail(G4HB200_ECAPACITY, "host secondary queue too small");
      continue;
    }
    const size_t off2 = static_cast<size_t>(2 * total);
    G4H_CUDA(cudaMemcpyAsync(hostSec->dirx_diry + off2, h->secDev.dirx_diry + 4 * lo, static_cast<size_t>(cnt) * 16, cudaMemcpyDeviceToHost, slot.stream));
    G4H_CUDA(cudaMemcpyAsync(hostSec->dirz_ekin + off2, h->secDev.dirz_ekin + 4 * lo, static_cast<size_t>(cnt) * 16, cudaMemcpyDeviceToHost, slot.stream));
    G4H_CUDA(cudaMemcpyAsync(hostSec->parent_kind + off2, h->secDev.parent_kind + 4 * lo, static_cast<size_t>(cnt) * 8, cudaMemcpyDeviceToHost, slot.stream));
    G4H_CUDA(cudaMemcpyAsync(hostSec->parent_slot + off2, h->secDev.parent_slot + 4 * lo, static_cast<size_t>(cnt) * 8, cudaMemcpyDeviceToHost, slot.stream));
    total += cnt;
  }
  for (auto& slot : h->slots) G4H_CUDA(cudaStreamSynchronize(slot.stream));
  if (rcSec != 0) return rcSec;
  hostSec->count[0] = static_cast<int32_t>(total);
  return 0;
}

int g4hb200_gamma_step_host(G4HB200* h, G4HB200GammaBatch* host, G4HB200SecondaryQueue* hostSec, uint64_t seed) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (host == nullptr || hostSec == nullptr) return Fail(G4HB200_EINVAL, "null argument");
  const int64_t n = host->n;
  if (n < 0 || n > 0x3fffffff) return Fail(G4HB200_EINVAL, "bad batch size");
  if (hostSec->count == nullptr) return Fail(G4HB200_EINVAL, "secondary queue without a counter");
  hostSec->count[0] = 0;
  if (n == 0) return 0;
  if (hostSec->capacity < 2 * n) return Fail(G4HB200_ECAPACITY, "host secondary queue must hold 2 records per track");
  if (n > h->gmCap) {
    if (h->gmCap > 0) g4hb200_gamma_batch_free(h, &h->gmDev);
    h->gmCap = 0;
    if ((rc = g4hb200_gamma_batch_alloc(h, n, &h->gmDev)) != 0) return rc;
    h->gmCap = n;
  }
  if (2 * n > h->secCap) {
    if (h->secCap > 0) g4hb200_secondary_queue_free(h, &h->secDev);
    h->secCap = 0;
    if ((rc = g4hb200_secondary_queue_alloc(h, 2 * n, &h->secDev)) != 0) return rc;
    h->secCap = 2 * n;
  }
  // small batches (and the single-launch option): one upload, one pipeline, one download
  if (n < 131072 || h->Fused(n)) {
    cudaStream_t st = h->stream;
    // the winner index travels too: a photon that ends its step on a boundary keeps the one it came with
    if ((rc = CopyGamma(host, &h->gmDev, cudaMemcpyHostToDevice, st, 0, 3, true, true)) != 0) return rc;
    if ((rc = g4hb200_secondary_queue_reset(h, &h->secDev, st)) != 0) return rc;
    if ((rc = h->Fused(n) ? LaunchGammaFused<2>(h, &h->gmDev, &h->secDev, seed, st) : LaunchGammaPipeline<2>(h, &h->gmDev, &h->secDev, seed, st)) != 0)
      return rc;
    if ((rc = CopyGamma(&h->gmDev, host, cudaMemcpyDeviceToHost, st, 0, 5, true, true)) != 0) return rc;
    if ((rc = g4hb200_secondary_queue_download(h, &h->secDev, hostSec, st)) != 0) return rc;
    G4H_CUDA(cudaStreamSynchronize(st));
    return 0;
  }
  // large ones in chunks on the two gamma work slots, like g4hb200_electron_step_host: while chunk c computes, chunk c+1
  // uploads and chunk c-1 downloads (68 B in, 100 B + secondaries out per track)
  if (h->pinnedCounts == nullptr) {
    G4H_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->pinnedCounts), G4HB200::kMaxChunks * sizeof(int32_t)));
    G4H_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->chunkCounters), G4HB200::kMaxChunks * sizeof(int32_t)));
  }
  G4HB200::WorkSlot* gslot[2] = {&h->gmSlot, &h->gmSlot2};
  for (auto* slot : gslot) {
    if (slot->stream == nullptr) {
      G4H_CUDA(cudaStreamCreateWithFlags(&slot->stream, cudaStreamNonBlocking));
      G4H_CUDA(cudaEventCreateWithFlags(&slot->counted, cudaEventDisableTiming));
    }
  }
  int64_t chunk = ((n / 4 + 32767) / 32768) * 32768;
  if (chunk < 65536) chunk = 65536;
  if (chunk > 262144) chunk = 262144;
  while ((n + chunk - 1) / chunk > G4HB200::kMaxChunks) chunk *= 2;
  const int numChunks = static_cast<int>((n + chunk - 1) / chunk);
  G4H_CUDA(cudaMemsetAsync(h->chunkCounters, 0, numChunks * sizeof(int32_t), gslot[0]->stream));
  G4H_CUDA(cudaStreamSynchronize(gslot[0]->stream));
  std::vector<cudaEvent_t> counted(numChunks);
  for (int c = 0; c < numChunks; ++c) {
    cudaStream_t cs = gslot[c & 1]->stream;
    const int64_t lo = c * chunk, len = (lo + chunk <= n) ? chunk : n - lo;
    G4HB200GammaBatch hv = GammaBatchView(*host, lo, len);
    G4HB200GammaBatch dv = GammaBatchView(h->gmDev, lo, len);
    G4HB200SecondaryQueue q = h->secDev;
    q.capacity = 2 * len;
    q.dirx_diry += 4 * lo;
    q.dirz_ekin += 4 * lo;
    q.parent_kind += 4 * lo;
    q.parent_slot += 4 * lo;
    q.count = h->chunkCounters + c;
    q.parent_base = static_cast<int32_t>(lo);
    if ((rc = CopyGamma(&hv, &dv, cudaMemcpyHostToDevice, cs, 0, 3, true, true)) != 0) return rc;
    if ((rc = LaunchGammaPipeline<2>(h, &dv, &q, seed, cs, (c & 1) != 0)) != 0) return rc;
    if ((rc = CopyGamma(&dv, &hv, cudaMemcpyDeviceToHost, cs, 0, 5, true, true)) != 0) return rc;
    G4H_CUDA(cudaMemcpyAsync(h->pinnedCounts + c, h->chunkCounters + c, sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
    G4H_CUDA(cudaEventCreateWithFlags(&counted[c], cudaEventDisableTiming));
    G4H_CUDA(cudaEventRecord(counted[c], cs));
  }
  int64_t total = 0;
  for (int c = 0; c < numChunks; ++c) {
    cudaStream_t cs = gslot[c & 1]->stream;
    G4H_CUDA(cudaEventSynchronize(counted[c]));
    cudaEventDestroy(counted[c]);
    const int64_t lo = c * chunk;
    const int64_t cnt = h->pinnedCounts[c];
    const size_t off2 = static_cast<size_t>(2 * total);
    G4H_CUDA(cudaMemcpyAsync(hostSec->dirx_diry + off2, h->secDev.dirx_diry + 4 * lo, static_cast<size_t>(cnt) * 16, cudaMemcpyDeviceToHost, cs));
    G4H_CUDA(cudaMemcpyAsync(hostSec->dirz_ekin + off2, h->secDev.dirz_ekin + 4 * lo, static_cast<size_t>(cnt) * 16, cudaMemcpyDeviceToHost, cs));
    G4H_CUDA(cudaMemcpyAsync(hostSec->parent_kind + off2, h->secDev.parent_kind + 4 * lo, static_cast<size_t>(cnt) * 8, cudaMemcpyDeviceToHost, cs));
    G4H_CUDA(cudaMemcpyAsync(hostSec->parent_slot + off2, h->secDev.parent_slot + 4 * lo, static_cast<size_t>(cnt) * 8, cudaMemcpyDeviceToHost, cs));
    total += cnt;
  }
  for (auto* slot : gslot) G4H_CUDA(cudaStreamSynchronize(slot->stream));
  hostSec->count[0] = static_cast<int32_t>(total);
  return 0;
}

int64_t g4hb200_launch_count(const G4HB200* h) { return h != nullptr ? h->launches : 0; }

const char* g4hb200_stage_name(int k) { return (k >= 0 && k < kNumElStages) ? kStageName[k] : ""; }

int g4hb200_set_kernel_timing(G4HB200* h, int enable) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  h->timing = enable != 0;
  return 0;
}

int g4hb200_set_msc_precision(G4HB200* h, int bits) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (bits != 32 && bits != 64) return Fail(G4HB200_EINVAL, "MSC precision must be 32 or 64");
  h->mscF32 = bits == 32;
  return 0;
}

int g4hb200_kernel_times(G4HB200* h, double* ms_sum, int64_t* launches, int64_t* items) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (ms_sum == nullptr || launches == nullptr || items == nullptr) return Fail(G4HB200_EINVAL, "null argument");
  for (int k = 0; k < G4HB200_NUM_STAGES; ++k) {
    ms_sum[k] = 0.0;
    launches[k] = 0;
    items[k] = 0;
  }
  for (auto& tc : h->timed) {
    G4H_CUDA(cudaEventSynchronize(tc.done));
    for (int k = 0; k < G4HB200_NUM_STAGES; ++k) {
      if (!tc.ran[k]) continue;
      float ms = 0.f;
      G4H_CUDA(cudaEventElapsedTime(&ms, tc.ev[2 * k], tc.ev[2 * k + 1]));
      ms_sum[k] += ms;
      launches[k] += 1;
      items[k] += kStageQueue[k] < 0 ? tc.n : tc.counts[kStageQueue[k]];
    }
    for (auto& e : tc.ev) cudaEventDestroy(e);
    cudaEventDestroy(tc.done);
    cudaFreeHost(tc.counts);
  }
  h->timed.clear();
  return 0;
}

}  // extern "C"

#include "capi_shower.inl"
