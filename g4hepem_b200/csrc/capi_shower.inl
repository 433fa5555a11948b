// capi_shower.inl -- host side of g4hb200_shower_run: the stepping loop over device-resident track stores
// (kernels: g4h_shower.cuh).  Included by capi.cu.
namespace {

struct ShowerStore {
  G4HB200ElectronBatch el[2];
  G4HB200GammaBatch gm[2];
  TrackGeo elGeo[2], gmGeo[2];
  G4HB200SecondaryQueue secEl, secGm;
  void* geoMem = nullptr;
  void* scoreMem = nullptr;
  ShowerScore score;
  ShowerCtrl* ctrl = nullptr;       // device bookkeeping of the graph-driven tail (inside scoreMem)
  ShowerCtrl* pinnedCtrl = nullptr; // host copy
  int32_t* pinned = nullptr;  // {next e-, next gamma, overflow, secondaries e-, secondaries gamma}
};

// withSubSteps: the e-/e+ stores also carry the MSC sub-step state (TrackGeo::sub_*)
int CarveGeo(unsigned char*& p, int64_t cap, TrackGeo& g, bool withSubSteps) {
  g.posx_posy = reinterpret_cast<double*>(p);
  p += cap * 16;
  g.posz_pad = reinterpret_cast<double*>(p);
  p += cap * 16;
  g.sub_left_eloss = g.sub_pre = g.sub_range_proc = nullptr;
  if (withSubSteps) {
    g.sub_left_eloss = reinterpret_cast<double*>(p);
    p += cap * 16;
    g.sub_pre = reinterpret_cast<double*>(p);
    p += cap * 16;
    g.sub_range_proc = reinterpret_cast<double*>(p);
    p += cap * 16;
  }
  g.vol = reinterpret_cast<int32_t*>(p);
  p += cap * 4;
  g.nextVol = reinterpret_cast<int32_t*>(p);
  p += cap * 4;
  g.n_dev = nullptr;
  return 0;
}

void FreeShowerStore(G4HB200* h, ShowerStore& s) {
  for (int k = 0; k < 2; ++k) {
    g4hb200_electron_batch_free(h, &s.el[k]);
    g4hb200_gamma_batch_free(h, &s.gm[k]);
  }
  g4hb200_secondary_queue_free(h, &s.secEl);
  g4hb200_secondary_queue_free(h, &s.secGm);
  if (s.geoMem != nullptr) cudaFree(s.geoMem);
  if (s.scoreMem != nullptr) cudaFree(s.scoreMem);
  if (s.pinned != nullptr) cudaFreeHost(s.pinned);
  if (s.pinnedCtrl != nullptr) cudaFreeHost(s.pinnedCtrl);
}

}  // namespace

namespace {
struct MixedSpec {
  bool enabled = false;
  int64_t nEl = 0, nGm = 0;
  double emin = 0, emax = 0;
};
int RunStepLoop(G4HB200* h, const G4HB200SlabGeometry* geom, int64_t numPrimaries, int32_t primaryKind, double primaryEkin,
                uint64_t seed, int32_t firstTrackId, int64_t capacity, int32_t maxSteps, double* edepOut,
                G4HB200ShowerStats* stats, const MixedSpec& mixed);
}  // namespace

extern "C" int g4hb200_shower_run(G4HB200* h, const G4HB200SlabGeometry* geom, int64_t numPrimaries, int32_t primaryKind,
                                  double primaryEkin, uint64_t seed, int32_t firstTrackId, int64_t capacity, int32_t maxSteps,
                                  double* edepOut, G4HB200ShowerStats* stats) {
  return RunStepLoop(h, geom, numPrimaries, primaryKind, primaryEkin, seed, firstTrackId, capacity, maxSteps, edepOut, stats,
                     MixedSpec());
}

extern "C" int g4hb200_mixed_run(G4HB200* h, int64_t numElectrons, int64_t numGammas, double emin, double emax, uint64_t seed,
                                 int64_t capacity, int32_t numSteps, double* edepTotal, G4HB200ShowerStats* stats) {
  if (h == nullptr) return Fail(G4HB200_EINVAL, "null handle");
  if (numElectrons < 0 || numGammas < 0 || !(emin > 0.0) || !(emax > emin) || numSteps < 1)
    return Fail(G4HB200_EINVAL, "bad mixed workload");
  // no geometry: a single scoring cell; a track's volume index is never used to look a couple up (secondaries
  // inherit their parent's) and no step ends on a boundary
  G4HB200SlabGeometry g;
  std::memset(&g, 0, sizeof(g));
  g.num_layers = 1;
  g.num_absorbers = 1;
  g.absorber_thickness[0] = 1.0;
  g.absorber_couple[0] = 0;
  g.half_yz = 1.0;
  MixedSpec m;
  m.enabled = true;
  m.nEl = numElectrons;
  m.nGm = numGammas;
  m.emin = emin;
  m.emax = emax;
  return RunStepLoop(h, &g, numElectrons + numGammas, G4HB200_SEC_ELECTRON, 0.0, seed, 0, capacity, numSteps, edepTotal,
                     stats, m);
}

namespace {
// One iteration of the loop with the populations on the device (cur -> the other store), enqueued on st / sg: what the
// body of RunStepLoop's loop does, for stream capture.
int EnqueueTailIteration(G4HB200* h, const SlabGeom& g, ShowerStore& s, uint64_t seed, int cur, cudaStream_t st, cudaStream_t sg) {
  const int nxt = cur ^ 1;
  // sizes the grids (a full wave: the populations may still grow); the kernels read the live counts from ShowerCtrl
  const int64_t nMax = s.score.capacity;
  int rc = 0;
  ShowerIterKernel<<<1, 32, 0, st>>>(s.ctrl, s.score.nextCount, s.secEl.count, s.secGm.count, s.score.overflow, h->slots[0].work.count,
                                     h->gmSlot.work.count);
  ++h->launches;
  G4H_CUDA(cudaEventRecord(h->loopFork, st));
  G4H_CUDA(cudaStreamWaitEvent(sg, h->loopFork, 0));
  {
    G4HB200ElectronBatch b = s.el[cur];
    b.n = nMax;
    TrackGeo geo = s.elGeo[cur];
    geo.n_dev = &s.ctrl->cur[0];
    const SlabHead slab{g, geo};
    if ((rc = LaunchElectronPipeline<true>(h, &b, &s.secEl, seed, st, 0, &slab)) != 0) return rc;
    ShowerElectronPostKernel<<<OneWave(h, ShowerElectronPostKernel, nMax), kThreadsPerBlock, 0, st>>>(g, b, geo, s.el[nxt], s.elGeo[nxt],
                                                                                                  s.score);
    ShowerSecondaryKernel<<<OneWave(h, ShowerSecondaryKernel, 2 * nMax), kThreadsPerBlock, 0, st>>>(
        h->view, g, seed, s.secEl, s.el[cur].meta, s.elGeo[cur], s.el[nxt], s.elGeo[nxt], s.gm[nxt], s.gmGeo[nxt], s.score);
    h->launches += 2;
  }
  {
    G4HB200GammaBatch b = s.gm[cur];
    b.n = nMax;
    TrackGeo geo = s.gmGeo[cur];
    geo.n_dev = &s.ctrl->cur[1];
    const SlabHead slab{g, geo};
    if ((rc = LaunchGammaPipeline<2>(h, &b, &s.secGm, seed, sg, false, &slab)) != 0) return rc;
    ShowerGammaPostKernel<<<OneWave(h, ShowerGammaPostKernel, nMax), kThreadsPerBlock, 0, sg>>>(g, b, geo, s.gm[nxt], s.gmGeo[nxt], s.score);
    ShowerSecondaryKernel<<<OneWave(h, ShowerSecondaryKernel, 2 * nMax), kThreadsPerBlock, 0, sg>>>(
        h->view, g, seed, s.secGm, s.gm[cur].meta, s.gmGeo[cur], s.el[nxt], s.elGeo[nxt], s.gm[nxt], s.gmGeo[nxt], s.score);
    h->launches += 2;
  }
  G4H_CUDA(cudaEventRecord(h->loopJoin, sg));
  G4H_CUDA(cudaStreamWaitEvent(st, h->loopJoin, 0));
  return 0;
}

// The rest of the loop from store `cur` on: graphs of two iterations (cur -> other -> cur), polled every kTailPoll launches.
int RunGraphTail(G4HB200* h, const SlabGeom& g, ShowerStore& s, uint64_t seed, int cur, cudaStream_t st, cudaStream_t sg,
                 G4HB200ShowerStats* stats) {
  constexpr int kTailPoll = 4;
  // the counters of the last host-driven iteration have been booked by the host: the tail's bookkeeping starts from zero
  G4H_CUDA(cudaMemsetAsync(s.secEl.count, 0, sizeof(int32_t), st));
  G4H_CUDA(cudaMemsetAsync(s.secGm.count, 0, sizeof(int32_t), st));
  G4H_CUDA(cudaMemsetAsync(s.ctrl, 0, sizeof(ShowerCtrl), st));
  G4H_CUDA(cudaStreamSynchronize(st));
  const int64_t launchesBefore = h->launches;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  G4H_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  int rc = EnqueueTailIteration(h, g, s, seed, cur, st, sg);
  if (rc == 0) rc = EnqueueTailIteration(h, g, s, seed, cur ^ 1, st, sg);
  const cudaError_t endErr = cudaStreamEndCapture(st, &graph);
  if (rc != 0 || endErr != cudaSuccess || graph == nullptr) {
    if (graph != nullptr) cudaGraphDestroy(graph);
    return rc != 0 ? rc : Fail(G4HB200_ECUDA, "cudaStreamEndCapture", endErr);
  }
  const int64_t launchesPerGraph = h->launches - launchesBefore;
  h->launches = launchesBefore;
  cudaError_t err = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (err != cudaSuccess) return Fail(G4HB200_ECUDA, "cudaGraphInstantiate", err);
  int status = 0;
  for (int poll = 0; poll < 250000; ++poll) {
    for (int k = 0; k < kTailPoll; ++k) {
      if ((err = cudaGraphLaunch(exec, st)) != cudaSuccess) break;
      h->launches += launchesPerGraph;
    }
    if (err != cudaSuccess) {
      status = Fail(G4HB200_ECUDA, "cudaGraphLaunch", err);
      break;
    }
    cudaMemcpyAsync(s.pinned, s.score.nextCount, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if ((err = cudaStreamSynchronize(st)) != cudaSuccess) {
      status = Fail(G4HB200_ECUDA, "shower tail", err);
      break;
    }
    if (s.pinned[2] != 0) {
      status = Fail(G4HB200_ECAPACITY, "shower track store capacity exceeded");
      break;
    }
    if (s.pinned[0] == 0 && s.pinned[1] == 0) break;
  }
  cudaGraphExecDestroy(exec);
  if (status != 0) return status;
  int32_t lastSec[2] = {0, 0};
  G4H_CUDA(cudaMemcpyAsync(s.pinnedCtrl, s.ctrl, sizeof(ShowerCtrl), cudaMemcpyDeviceToHost, st));
  G4H_CUDA(cudaMemcpyAsync(&lastSec[0], s.secEl.count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  G4H_CUDA(cudaMemcpyAsync(&lastSec[1], s.secGm.count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  G4H_CUDA(cudaStreamSynchronize(st));
  const ShowerCtrl& c = *s.pinnedCtrl;
  stats->num_steps += c.steps;
  stats->electron_track_steps += c.sumEl;
  stats->gamma_track_steps += c.sumGm;
  stats->secondaries += c.sumSec + lastSec[0] + lastSec[1];
  if (c.peakEl > stats->peak_electrons) stats->peak_electrons = c.peakEl;
  if (c.peakGm > stats->peak_gammas) stats->peak_gammas = c.peakGm;
  return 0;
}

int RunStepLoop(G4HB200* h, const G4HB200SlabGeometry* geom, int64_t numPrimaries, int32_t primaryKind, double primaryEkin,
                uint64_t seed, int32_t firstTrackId, int64_t capacity, int32_t maxSteps, double* edepOut,
                G4HB200ShowerStats* stats, const MixedSpec& mixed) {
  int rc = CheckHandle(h);
  if (rc != 0) return rc;
  if (geom == nullptr || edepOut == nullptr || stats == nullptr) return Fail(G4HB200_EINVAL, "null argument");
  if (geom->num_layers < 1 || geom->num_absorbers < 1 || geom->num_absorbers > kMaxAbsorbers ||
      geom->num_layers * geom->num_absorbers > CtaHist::kMaxBins)
    return Fail(G4HB200_EINVAL, "bad slab geometry");
  if (numPrimaries < 0 || (mixed.enabled ? (mixed.nEl > capacity || mixed.nGm > capacity) : numPrimaries > capacity) ||
      capacity > 0x3fffffff)
    return Fail(G4HB200_EINVAL, "bad primary count / capacity");
  for (int k = 0; k < geom->num_absorbers; ++k) {
    if (geom->absorber_couple[k] < 0 || geom->absorber_couple[k] >= h->view.numMatCut || !(geom->absorber_thickness[k] > 0.0))
      return Fail(G4HB200_EINVAL, "bad absorber");
  }
  SlabGeom g;
  std::memset(&g, 0, sizeof(g));
  g.numLayers = geom->num_layers;
  g.numAbsorbers = geom->num_absorbers;
  g.absFront[0] = 0.0;
  for (int k = 0; k < g.numAbsorbers; ++k) {
    g.thickness[k] = geom->absorber_thickness[k];
    g.couple[k] = geom->absorber_couple[k];
    g.absFront[k + 1] = g.absFront[k] + g.thickness[k];
  }
  g.halfYZ = geom->half_yz;
  g.xFront = -0.5 * (g.numLayers * g.absFront[g.numAbsorbers]);
  g.inheritCouple = mixed.enabled ? 1 : 0;
  g.wdtOn = (!mixed.enabled && geom->woodcock_on != 0) ? 1 : 0;
  g.wdtCouple = geom->woodcock_couple;
  g.wdtEkinMin = geom->woodcock_ekin_min;
  if (g.wdtOn != 0 && (g.wdtCouple < 0 || g.wdtCouple >= h->view.numMatCut)) return Fail(G4HB200_EINVAL, "bad Woodcock couple");
  const int nbins = g.numLayers * g.numAbsorbers;
  std::memset(stats, 0, sizeof(*stats));
  for (int k = 0; k < nbins; ++k) edepOut[k] = 0.0;
  if (numPrimaries == 0) return 0;

  ShowerStore s;
  std::memset(&s, 0, sizeof(s));
  auto fail = [&](int code) {
    FreeShowerStore(h, s);
    return code;
  };
  for (int k = 0; k < 2; ++k) {
    if ((rc = g4hb200_electron_batch_alloc(h, capacity, &s.el[k])) != 0) return fail(rc);
    if ((rc = g4hb200_gamma_batch_alloc(h, capacity, &s.gm[k])) != 0) return fail(rc);
  }
  if ((rc = g4hb200_secondary_queue_alloc(h, 2 * capacity, &s.secEl)) != 0) return fail(rc);
  if ((rc = g4hb200_secondary_queue_alloc(h, 2 * capacity, &s.secGm)) != 0) return fail(rc);
  {
    const size_t per = static_cast<size_t>(capacity) * 40;
    const size_t sub = static_cast<size_t>(capacity) * 48;
    if (cudaMalloc(&s.geoMem, 4 * per + 2 * sub) != cudaSuccess) return fail(Fail(G4HB200_ENOMEM, "cudaMalloc(shower geo)"));
    unsigned char* p = static_cast<unsigned char*>(s.geoMem);
    CarveGeo(p, capacity, s.elGeo[0], true);
    CarveGeo(p, capacity, s.elGeo[1], true);
    CarveGeo(p, capacity, s.gmGeo[0], false);
    CarveGeo(p, capacity, s.gmGeo[1], false);
    const size_t sbytes = static_cast<size_t>(nbins) * 8 + 2 * 8 + 4 * 4 + sizeof(ShowerCtrl);
    if (cudaMalloc(&s.scoreMem, sbytes) != cudaSuccess) return fail(Fail(G4HB200_ENOMEM, "cudaMalloc(shower score)"));
    unsigned char* q = static_cast<unsigned char*>(s.scoreMem);
    s.score.hist = reinterpret_cast<double*>(q);
    s.score.leak = s.score.hist + nbins;
    s.score.nextCount = reinterpret_cast<int32_t*>(s.score.leak + 2);
    s.score.overflow = s.score.nextCount + 2;
    s.score.capacity = capacity;
    s.ctrl = reinterpret_cast<ShowerCtrl*>(s.score.nextCount + 4);  // 8-byte aligned: nbins * 8 + 16 + 16
    if (cudaMemset(s.scoreMem, 0, sbytes) != cudaSuccess) return fail(Fail(G4HB200_ECUDA, "cudaMemset(score)"));
    if (cudaMallocHost(reinterpret_cast<void**>(&s.pinned), 8 * sizeof(int32_t)) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&s.pinnedCtrl), sizeof(ShowerCtrl)) != cudaSuccess)
      return fail(Fail(G4HB200_ENOMEM, "cudaMallocHost"));
  }
  // size the pipelines' workspaces once (they grow on demand otherwise: a reallocation per iteration while the
  // shower develops)
  if ((rc = EnsureElectronWork(h->slots[0], capacity)) != 0) return fail(rc);
  // large populations run as part-batch pipelines side by side (LaunchElectronPipelineHalves): their slots too
  for (int p = 1; p < h->splitParts; ++p) {
    if ((rc = EnsureElectronWork(h->slots[p], capacity / h->splitParts + 2 * kThreadsPerBlock)) != 0) return fail(rc);
  }
  if ((rc = EnsureElectronWork(h->gmSlot, capacity)) != 0) return fail(rc);
  if (h->splitParts > 1 && (rc = EnsureElectronWork(h->gmSlot2, capacity / 2 + 2 * kThreadsPerBlock)) != 0) return fail(rc);
  cudaStream_t st = h->stream;
  // the e-/e+ chain and the gamma chain of an iteration are independent until the host reads the counts: they run
  // side by side (st / sg).  Most iterations of a shower are in its tail, where a chain is a dozen launches over a
  // few thousand tracks and bound by launch latency.
  if (h->loopStream == nullptr) {
    if (cudaStreamCreateWithFlags(&h->loopStream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->loopFork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->loopJoin, cudaEventDisableTiming) != cudaSuccess) {
      return fail(Fail(G4HB200_ECUDA, "loop stream"));
    }
  }
  cudaStream_t sg = h->loopStream;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEventCreate(&ev0);
  cudaEventCreate(&ev1);
  const int64_t launches0 = h->launches;
  int cur = 0;
  int64_t nEl = primaryKind == G4HB200_SEC_GAMMA ? 0 : numPrimaries;
  int64_t nGm = primaryKind == G4HB200_SEC_GAMMA ? numPrimaries : 0;
  if (mixed.enabled) {
    nEl = mixed.nEl;
    nGm = mixed.nGm;
    MixedPopulationKernel<<<OneWave(h, MixedPopulationKernel, numPrimaries), kThreadsPerBlock, 0, st>>>(
        nEl, nGm, h->view.numMatCut, std::log(mixed.emin), std::log(mixed.emax / mixed.emin), seed, s.el[0], s.elGeo[0], s.gm[0],
        s.gmGeo[0]);
    cudaStreamSynchronize(st);  // the population is an input: keep it out of the timed loop
    cudaEventRecord(ev0, st);
  } else {
    cudaEventRecord(ev0, st);
    ShowerPrimaryKernel<<<OneWave(h, ShowerPrimaryKernel, numPrimaries), kThreadsPerBlock, 0, st>>>(
        g, numPrimaries, primaryKind, primaryEkin, firstTrackId, s.el[0], s.elGeo[0], s.gm[0], s.gmGeo[0]);
  }
  ++h->launches;
  int status = 0;
  const int stepLimit = maxSteps > 0 ? maxSteps : 1000000;
  for (int step = 0; (nEl > 0 || nGm > 0); ++step) {
    if (step >= stepLimit) break;
    // Most iterations of a shower step a few thousand tracks: a chain of ~25 small kernels whose ~100 API calls and one
    // host synchronisation cost more than the kernels run.  Below tailBelow tracks per population (from the second
    // iteration on: the first one creates the streams and sizes the launches) the loop therefore runs as CUDA graphs of
    // two iterations each with the populations on the device (ShowerCtrl, TrackGeo::n_dev), full-wave grids -- the
    // populations may still grow -- and the host looks at the counters once per kTailPoll graph launches.  Above it the
    // host-driven iteration with its two half-batch pipelines side by side is the faster one.
    if (step > 0 && maxSteps == 0 && !mixed.enabled && h->graphTail && !h->timing && nEl < h->tailBelow && nGm < h->tailBelow) {
      status = RunGraphTail(h, g, s, seed, cur, st, sg, stats);
      nEl = nGm = 0;
      break;
    }
    const int nxt = cur ^ 1;
    stats->num_steps += 1;
    stats->electron_track_steps += nEl;
    stats->gamma_track_steps += nGm;
    if (nEl > stats->peak_electrons) stats->peak_electrons = nEl;
    if (nGm > stats->peak_gammas) stats->peak_gammas = nGm;
    cudaMemsetAsync(s.score.nextCount, 0, 2 * sizeof(int32_t), st);
    cudaEventRecord(h->loopFork, st);
    cudaStreamWaitEvent(sg, h->loopFork, 0);
    s.el[cur].n = nEl;
    s.gm[cur].n = nGm;
    if (nEl > 0) {
      G4HB200ElectronBatch& b = s.el[cur];
      g4hb200_secondary_queue_reset(h, &s.secEl, st);
      if (mixed.enabled) {
        // no geometry: the fused step (the proposed step is accepted)
        if ((status = g4hb200_electron_step(h, &b, &s.secEl, seed, st)) != 0) break;
      } else {
        // HowFar + geometry step + Perform: the head of the pipeline does the first two and the along-step part of
        // the third in one pass (ShowerElectronHeadKernel), the queue kernels of Perform follow
        const SlabHead slab{g, s.elGeo[cur]};
        if ((status = h->FusedElectron(nEl) ? LaunchElectronFused<false>(h, &b, &s.secEl, seed, st, &slab)
                               : LaunchElectronPipelineHalves<true>(h, &b, &s.secEl, seed, st, &slab)) != 0)
          break;
      }
      ShowerElectronPostKernel<<<OneWave(h, ShowerElectronPostKernel, nEl), kThreadsPerBlock, 0, st>>>(
          g, b, s.elGeo[cur], s.el[nxt], s.elGeo[nxt], s.score);
      ++h->launches;
    }
    if (nEl > 0) {
      ShowerSecondaryKernel<<<OneWave(h, ShowerSecondaryKernel, 2 * nEl), kThreadsPerBlock, 0, st>>>(
          h->view, g, seed, s.secEl, s.el[cur].meta, s.elGeo[cur], s.el[nxt], s.elGeo[nxt], s.gm[nxt], s.gmGeo[nxt], s.score);
      ++h->launches;
      cudaMemcpyAsync(s.pinned + 3, s.secEl.count, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    } else {
      s.pinned[3] = 0;
    }
    if (nGm > 0) {
      G4HB200GammaBatch& b = s.gm[cur];
      g4hb200_secondary_queue_reset(h, &s.secGm, sg);
      if (mixed.enabled) {
        if ((status = g4hb200_gamma_step(h, &b, &s.secGm, seed, sg)) != 0) break;
      } else {
        // HowFar + geometry step + SelectInteraction / Perform: one head kernel (ShowerGammaHeadKernel), then the samplers
        const SlabHead slab{g, s.gmGeo[cur]};
        if ((status = h->Fused(nGm) ? LaunchGammaFused<2>(h, &b, &s.secGm, seed, sg, &slab)
                               : LaunchGammaPipelineHalves<2>(h, &b, &s.secGm, seed, sg, &slab)) != 0)
          break;
      }
      ShowerGammaPostKernel<<<OneWave(h, ShowerGammaPostKernel, nGm), kThreadsPerBlock, 0, sg>>>(
          g, b, s.gmGeo[cur], s.gm[nxt], s.gmGeo[nxt], s.score);
      ++h->launches;
      ShowerSecondaryKernel<<<OneWave(h, ShowerSecondaryKernel, 2 * nGm), kThreadsPerBlock, 0, sg>>>(
          h->view, g, seed, s.secGm, s.gm[cur].meta, s.gmGeo[cur], s.el[nxt], s.elGeo[nxt], s.gm[nxt], s.gmGeo[nxt], s.score);
      ++h->launches;
      cudaMemcpyAsync(s.pinned + 4, s.secGm.count, sizeof(int32_t), cudaMemcpyDeviceToHost, sg);
    } else {
      s.pinned[4] = 0;
    }
    cudaEventRecord(h->loopJoin, sg);
    cudaStreamWaitEvent(st, h->loopJoin, 0);
    cudaMemcpyAsync(s.pinned, s.score.nextCount, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    const cudaError_t err = cudaStreamSynchronize(st);
    if (err != cudaSuccess) {
      status = Fail(G4HB200_ECUDA, "shower step", err);
      break;
    }
    if (s.pinned[2] != 0) {
      status = Fail(G4HB200_ECAPACITY, "shower track store capacity exceeded");
      break;
    }
    stats->secondaries += s.pinned[3] + s.pinned[4];
    nEl = s.pinned[0];
    nGm = s.pinned[1];
    cur = nxt;
  }
  cudaEventRecord(ev1, st);
  cudaEventSynchronize(ev1);
  // whatever was forked onto the gamma stream and the part-batch streams is done before the stores go away (an error
  // path leaves the loop between a fork and its join)
  cudaStreamSynchronize(sg);
  for (auto& slot : h->slots) {
    if (slot.stream != nullptr) cudaStreamSynchronize(slot.stream);
    for (auto& a : slot.aux) if (a != nullptr) cudaStreamSynchronize(a);
  }
  if (h->gmSlot2.stream != nullptr) cudaStreamSynchronize(h->gmSlot2.stream);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev0, ev1);
  stats->device_ms = ms;
  stats->kernel_launches = h->launches - launches0;
  if (status == 0 && (nEl > 0 || nGm > 0)) {
    // stopped on max_steps: what is left alive
    stats->remaining_electrons = nEl;
    stats->remaining_gammas    = nGm;
    std::vector<double> e(static_cast<size_t>(2 * (nEl > nGm ? nEl : nGm)));
    double left = 0.0;
    if (nEl > 0 && cudaMemcpy(e.data(), s.el[cur].ekin_logekin, static_cast<size_t>(nEl) * 16, cudaMemcpyDeviceToHost) == cudaSuccess) {
      for (int64_t k = 0; k < nEl; ++k) left += e[2 * k];
    }
    if (nGm > 0 && cudaMemcpy(e.data(), s.gm[cur].ekin_logekin, static_cast<size_t>(nGm) * 16, cudaMemcpyDeviceToHost) == cudaSuccess) {
      for (int64_t k = 0; k < nGm; ++k) left += e[2 * k];
    }
    stats->remaining_ekin = left;
  }
  if (status == 0) {
    double leak[2];
    cudaMemcpy(edepOut, s.score.hist, static_cast<size_t>(nbins) * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(leak, s.score.leak, 16, cudaMemcpyDeviceToHost);
    stats->leak_electron = leak[0];
    stats->leak_gamma = leak[1];
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  FreeShowerStore(h, s);
  if (status == 0) {
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return Fail(G4HB200_ECUDA, "shower", err);
  }
  return status;
}
}  // namespace
