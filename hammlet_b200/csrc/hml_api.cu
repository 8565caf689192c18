// hammlet_b200 — C ABI (include/hammlet_b200.h): context, load orchestration, sweeps, getters.
#include <dlfcn.h>
#include <float.h>
#include <math.h>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges appear when a profiler injects its library, else no-ops
#include <nccl.h>  // types only: the library is loaded with dlopen in hml_comm_init (single-GPU use needs no NCCL)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "../../include/hammlet_b200.h"
#include "hml_common.cuh"
#include "hml_kernels.h"

using namespace hml;

namespace {
std::string g_create_error;
struct Stage {
  const char* name;
  cudaEvent_t ev;
};
// device/pinned result block, 64-bit words (the last term: per-state sums of the further dimensions of multivariate data)
constexpr size_t kOutWords = 2 + 32 + 32 * 32 + 2 + 2 * 32 + 2 + 2 * 32 * (kMaxDims - 1);
constexpr int kMaxWorld = 64;
constexpr size_t kOpDoubles = 32 * 32 + 32;  // largest segment operator (mantissas + exponents)

// NCCL entry points, resolved at run time from libnccl.so.2 (the copy PyTorch already loaded, if any)
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool load() {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
      error = std::string("cannot load libnccl.so.2: ") + dlerror();
      return false;
    }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllGather || !CommDestroy || !GetErrorString) {
      error = "libnccl.so.2 lacks a required symbol";
      lib = nullptr;
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;
}  // namespace

struct hml_ctx {
  int device = 0;
  int sms = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;

  // segment mode (one sequence split over the ranks of a communicator)
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  uint64_t T_global = 0, seg_start = 0;
  uint64_t first_block = 0, global_blocks = 0;  // of the last sweep / block structure
  bool p2p = false;                             // carries travel through peer mailboxes (else NCCL all-gathers)
  unsigned char* mbox = nullptr;                // this rank's mailbox (peers write into it over NVLink)
  P2PPeers peers = {};                          // every rank's mailbox as mapped here
  uint64_t p2p_seq[kP2PSlots] = {0, 0, 0, 0};
  unsigned int* p2p_timeout_host = nullptr;     // mapped host word raised by an exchange that gave up waiting
  unsigned int* p2p_timeout_dev = nullptr;
  P2PDev* p2p_dev = nullptr;                    // device copy of what the exchanging kernels need
  double* seg_dev = nullptr;                    // send slots + gathered carries (one allocation)
  unsigned long long* stats_gather = nullptr;   // world x kOutWords
  unsigned long long* stats_gather_host = nullptr;
  // run formation across rank borders (recorded sweeps): [0] this rank's last state, [8..8+world) everybody's
  unsigned long long* run_states = nullptr;
  uint32_t* run_border = nullptr;               // [0] position 0 started a run in some recorded iteration, [1] in the last
  // merged (whole-sequence) views, filled by collective getters in segment mode
  bool g_valid = false;                         // g_seg_* hold the runs of the current states
  std::vector<uint64_t> g_seg_size;
  std::vector<int16_t> g_seg_state;
  bool g_mg_valid = false;                      // g_mg_* hold the merged marginals as of the last hml_marginals_add
  std::vector<uint64_t> g_mg_size;
  std::vector<int32_t> g_mg_counts;

  // sequence (resident since load); in segment mode T is the length of the local segment
  uint64_t T = 0;
  float* w = nullptr;       // breakpoint weights, padded to a tile multiple
  uint16_t* smax = nullptr; // bf16 max pyramid over sub-blocks of 32 weights (boundary detection reads only hot sub-blocks)
  int detect_mode = HML_DETECT_CANDIDATES;
  // candidate list (HML_DETECT_CANDIDATES): positions, ascending, and weights of everything not below cand_floor
  uint32_t* cand_pos = nullptr;
  float* cand_w = nullptr;
  double2* cand_pq = nullptr;        // integral pair of every candidate (univariate data)
  double2* spq = nullptr;            // integral pairs of the block starts, in block order (capacity + 1)
  bool spq_valid = false;            // ... of the current block structure
  uint32_t* cand_scratch = nullptr;  // per-CTA counts, offsets, ticket
  uint64_t cand_cap = 0, cand_n = 0;
  uint32_t cand_scratch_ctas = 0;
  float cand_floor = 0.f;
  float cand_too_long_floor = -1.f;  // largest floor whose candidate list was not worth keeping (> T / 2 entries)
  bool cand_valid = false;
  uint64_t cand_rebuilds = 0;
  float* coeffs = nullptr;  // maxlet coefficients (kept for hml_get_coeffs while T is small)
  double2* pq = nullptr;    // integral arrays, T+1 entries (multivariate data: D planes, pq_stride entries apart)
  double4* cell_pref = nullptr;
  int D = 1;                // data dimensions (hml_load_f32_md)
  uint64_t pq_stride = 0, cell_stride = 0;
  double sigma_hat = NAN;

  // boundary detection scratch (bit masks + per-tile / per-CTA counts)
  void* detect_scratch = nullptr;

  // block structure
  uint64_t capacity = 0;  // multiple of kTileBlocks
  uint32_t* starts = nullptr;
  uint32_t* bN = nullptr;
  double2* bS = nullptr;
  bool blocks_valid = false, stats_valid = false, states_valid = false, rows_valid = false;
  uint64_t nblocks = 0;  // host copy, valid when blocks_valid

  // per-sweep buffers (sized by capacity and KP)
  int KP = 0;
  int last_K = 0;
  double *e = nullptr, *maxE = nullptr, *alpha = nullptr;
  uint8_t *maps = nullptr, *states = nullptr, *chunk_maps = nullptr, *tile_maps = nullptr, *tile_qin = nullptr;
  uint8_t* chunk_submaps = nullptr;
  unsigned* tickets = nullptr;
  double *chunk_ops = nullptr, *tile_ops = nullptr, *tile_ain = nullptr, *group_ops = nullptr, *group_ain = nullptr;
  int *chunk_exp = nullptr, *tile_exp = nullptr, *group_exp = nullptr, *wide_exp = nullptr;
  double* wide_ops = nullptr;  // K > 8: scratch of k_fwd_chunks_wide
  double* rows = nullptr;
  uint64_t rows_cap = 0;
  double* replay_u = nullptr;
  uint64_t replay_cap = 0;
  // run-length view of the last sweep's states (hml_get_segments)
  uint32_t* seg_counts = nullptr;  // per 1024-block tile: offset of its first run; [ntiles] = number of runs
  uint32_t* seg_starts = nullptr;
  int16_t* seg_states = nullptr;
  uint64_t seg_cap = 0;            // blocks the three arrays are sized for
  bool segs_valid = false;         // seg_counts holds the offsets of the current states
  bool runs_written = false;       // seg_starts / seg_states hold the runs of the current states
  bool nsegs_known = false;        // nsegs (host) is the count of the current states
  uint64_t nsegs = 0;
  // state marginals accumulated on the device (hml_marginals_*): sorted segment starts + K counts per segment
  uint32_t* mg_pos[2] = {nullptr, nullptr};
  uint16_t* mg_cnt[2] = {nullptr, nullptr};
  uint32_t *mg_run_of_old = nullptr, *mg_olds_below = nullptr, *mg_flags = nullptr;
  uint64_t mg_cap = 0, mg_runs_cap = 0;  // segments the double buffers hold; runs the scratch holds
  int mg_cur = 0, mg_K = 0;
  uint64_t mg_n = 0, mg_iterations = 0;  // mg_n: segments as of the last completed copy of the device-side count
  uint32_t* mg_n_dev = nullptr;          // [2]: segment count of buffer 0 / 1 (the merge kernels read and write them)
  uint32_t* mg_n_host = nullptr;         // pinned mirror of the current count
  bool mg_pending = false;               // a count copy has been queued since the last stream synchronisation
  double* partials = nullptr;
  // forward filter: speculative (guessed chunk starts + repair pass) unless it keeps failing on this data
  int forward_mode = HML_FORWARD_AUTO;
  uint64_t spec_sweeps = 0, spec_failures = 0;
  uint32_t spec_skip = 0, spec_streak = 0;  // sweeps still to run through the operator scan / failures in a row
  bool spec_lowered = false;
  int spec_level = 0;                       // (piece length, warm-up) of the guesses: up after a failure, down after 64 good sweeps
  uint32_t spec_good = 0, spec_patience = 64;  // good sweeps in a row / how many it takes to try the level below again
  // HML_HOST_TIMING=1: where the host thread spends a sweep (printed when the handle is destroyed)
  uint64_t host_ns_prepare = 0, host_ns_launch = 0, host_ns_wait = 0, host_sweeps = 0;
  unsigned long long* outblk = nullptr;       // device result block: [0] nblocks, [2..] per-sweep outputs
  unsigned long long* outblk_host = nullptr;  // pinned mirror

  // the whole sweep in one persistent kernel + device-resident Gibbs chain (hml_fused.cuh)
  ChainDev* chain_dev = nullptr;
  ChainDev* chain_host = nullptr;      // pinned mirror: parameters and status as of the last synchronisation
  uint32_t* fused_cta_count = nullptr;
  unsigned* fused_barriers = nullptr;
  double* fused_qtot = nullptr;
  unsigned long long* fused_qmap = nullptr;
  unsigned long long* fused_phase_ns = nullptr;
  double* fused_subops = nullptr;
  unsigned long long* fused_submaps = nullptr;
  unsigned long long* chain_stats = nullptr;  // result block of the whole sequence (segment mode: summed over the ranks)
  bool chain_ready = false;            // hml_chain_init has been called
  uint64_t chain_fused_sweeps = 0, chain_standard_sweeps = 0;
  float last_thr = NAN;                // threshold of the current block structure

  // timing
  bool nvtx_open = false;
  bool timing = false;
  std::vector<Stage> stages;
  size_t stage_used = 0;
  std::vector<cudaEvent_t> event_pool;
  std::vector<std::string> last_names;
  std::vector<float> last_ms;
};

namespace {

int fail(hml_t* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

#define CK(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess) return fail(h, HML_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

template <typename P>
cudaError_t dev_alloc(P*& p, size_t n) {
  if (p) {
    cudaFree(p);
    p = nullptr;
  }
  if (n == 0) n = 1;
  return cudaMalloc((void**)&p, n * sizeof(P));
}
template <typename P>
void dev_free(P*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}
// device temporary of a load / getter: freed on every way out of the function (a CK() early return included)
template <typename P>
struct DevTmp {
  P* p = nullptr;
  DevTmp() = default;
  DevTmp(const DevTmp&) = delete;
  DevTmp& operator=(const DevTmp&) = delete;
  ~DevTmp() { release(); }
  cudaError_t alloc(size_t n) { return dev_alloc(p, n); }
  void release() { dev_free(p); }
  operator P*() const { return p; }
  P* operator+(size_t i) const { return p + i; }
};

// HML_NVTX=1 in the environment: every stage of a sweep is an NVTX range (nsys / ncu --nvtx show the kernels of a stage
// under its name); independent of the CUDA-event timing below
bool nvtx_enabled() {
  static const bool on = getenv("HML_NVTX") && getenv("HML_NVTX")[0] != '0';
  return on;
}

void stage_cb(void* user, const char* name) {
  hml_t* h = (hml_t*)user;
  if (nvtx_enabled()) {
    if (h->nvtx_open) nvtxRangePop();
    h->nvtx_open = strcmp(name, "end") != 0;
    if (h->nvtx_open) nvtxRangePushA(name);
  }
  if (!h->timing) return;
  if (h->stage_used == h->event_pool.size()) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    h->event_pool.push_back(ev);
  }
  cudaEvent_t ev = h->event_pool[h->stage_used++];
  cudaEventRecord(ev, h->stream);
  h->stages.push_back({name, ev});
}

void collect_timing(hml_t* h) {
  if (!h->timing) return;
  h->last_names.clear();
  h->last_ms.clear();
  for (size_t i = 0; i + 1 < h->stages.size(); ++i) {
    if (strcmp(h->stages[i].name, "end") == 0) continue;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->stages[i].ev, h->stages[i + 1].ev);
    h->last_names.push_back(h->stages[i].name);
    h->last_ms.push_back(ms);
  }
  h->stages.clear();
  h->stage_used = 0;
}

uint64_t round_up(uint64_t v, uint64_t m) { return (v + m - 1) / m * m; }

// (re)allocates everything sized by the block capacity; invalidates the block structure
int alloc_blocks(hml_t* h, uint64_t cap, int KP) {
  cap = round_up(cap < 1 ? 1 : cap, kTileBlocks);
  const uint64_t chunks = cap / kChunkLen, tiles = cap / kTileBlocks;
  const int MB = KP ? map_bytes(KP) : 8;
  if (cap != h->capacity) {
    CK(dev_alloc(h->starts, cap + 1));
    CK(dev_alloc(h->spq, h->D == 1 ? cap + 1 : 1));
    h->spq_valid = false;
    CK(dev_alloc(h->bN, cap));
    CK(dev_alloc(h->bS, cap * (uint64_t)h->D));
    CK(dev_alloc(h->states, cap));
    CK(dev_alloc(h->tile_qin, tiles));
    h->blocks_valid = h->stats_valid = h->states_valid = h->rows_valid = false;
    h->segs_valid = h->g_valid = false;
    dev_free(h->rows);
    h->rows_cap = 0;
  }
  if (KP && (cap != h->capacity || KP != h->KP)) {
    CK(dev_alloc(h->e, cap * KP));
    CK(dev_alloc(h->alpha, cap * KP));
    CK(dev_alloc(h->maxE, cap));
    CK(dev_alloc(h->maps, cap * MB));
    CK(dev_alloc(h->chunk_maps, chunks * MB));
    CK(dev_alloc(h->chunk_submaps, chunks * MB * 3));
    CK(dev_alloc(h->tile_maps, tiles * MB));
    CK(dev_alloc(h->chunk_ops, chunks * KP * KP));
    CK(dev_alloc(h->chunk_exp, chunks * KP));
    CK(dev_alloc(h->tile_ops, tiles * KP * KP));
    CK(dev_alloc(h->tile_exp, tiles * KP));
    CK(dev_alloc(h->tile_ain, tiles * KP));
    CK(dev_alloc(h->group_ops, (size_t)256 * KP * KP));  // kScanGroupsMax entries
    CK(dev_alloc(h->group_exp, (size_t)256 * KP));
    CK(dev_alloc(h->group_ain, (size_t)256 * KP));
    if (wide_scratch_doubles(KP)) {
      CK(dev_alloc(h->wide_ops, wide_scratch_doubles(KP) * h->sms * kWideCtasPerSm));
      CK(dev_alloc(h->wide_exp, wide_scratch_ints(KP) * h->sms * kWideCtasPerSm));
    } else {
      dev_free(h->wide_ops);
      dev_free(h->wide_exp);
    }
    CK(dev_alloc(h->partials, reduce_partials_doubles(KP, h->sms * 32)));
    h->KP = KP;
    h->states_valid = h->rows_valid = false;
  }
  h->capacity = cap;
  return HML_OK;
}

size_t result_words(int KP, int D);

SweepBuffers make_buffers(hml_t* h, int KP) {
  SweepBuffers b;
  memset(&b, 0, sizeof(b));
  b.pq = h->pq;
  b.cell_pref = h->cell_pref;
  b.D = h->D;
  b.pq_stride = h->pq_stride;
  b.cell_stride = h->cell_stride;
  b.starts = h->starts;
  b.spq = h->spq_valid ? h->spq : nullptr;
  b.nblocks = h->outblk;
  b.capacity = h->capacity;
  b.bN = h->bN;
  b.bS = h->bS;
  b.e = h->e;
  b.maxE = h->maxE;
  b.alpha = h->alpha;
  b.maps = h->maps;
  b.states = h->states;
  b.chunk_ops = h->chunk_ops;
  b.chunk_exp = h->chunk_exp;
  b.chunk_maps = h->chunk_maps;
  b.chunk_submaps = h->chunk_submaps;
  b.tile_maps = h->tile_maps;
  b.tile_qin = h->tile_qin;
  b.tickets = h->tickets;
  // (pinned memory is device-addressable under its own address with unified addressing; checked once, else copies)
  unsigned long long* const mirror = h->world > 1 ? h->stats_gather_host : h->outblk_host;
  void* dp = nullptr;
  b.result_host = (mirror != nullptr && cudaHostGetDevicePointer(&dp, mirror, 0) == cudaSuccess && dp == (void*)mirror)
                      ? mirror : nullptr;
  if (b.result_host == nullptr) (void)cudaGetLastError();
  b.result_words = (uint32_t)result_words(KP, h->D);
  b.tile_ops = h->tile_ops;
  b.tile_exp = h->tile_exp;
  b.tile_ain = h->tile_ain;
  b.group_ops = h->group_ops;
  b.group_exp = h->group_exp;
  b.group_ain = h->group_ain;
  b.wide_ops = h->wide_ops;
  b.wide_exp = h->wide_exp;
  b.partials = h->partials;
  b.out_u64 = h->outblk + 2;
  size_t words = (size_t)KP + (size_t)KP * KP + 1;
  words += words & 1;
  b.out_f64 = (double*)(h->outblk + 2 + words);
  if (h->world > 1) {
    double* d = h->seg_dev;
    b.seg.rank = h->rank;
    b.seg.world = h->world;
    b.seg.send_head = d;
    b.seg.send_map = (uint64_t*)(d + kHeadWords);
    b.seg.send_op = d + kHeadWords + 4;
    d += kHeadWords + 4 + kOpDoubles;
    b.seg.heads = d;
    b.seg.maps = (const uint64_t*)(d + kHeadWords * (size_t)h->world);
    b.seg.ops = d + (kHeadWords + 4) * (size_t)h->world;
    b.seg.overflow = h->outblk + 1;
    b.seg.p2p = h->p2p ? h->p2p_dev : nullptr;
    b.seg.stats_send = h->outblk;
    b.seg.stats_recv = h->stats_gather;
  }
  return b;
}

#define CKN(call)                                                                                             \
  do {                                                                                                        \
    ncclResult_t r__ = (call);                                                                                \
    if (r__ != ncclSuccess) return fail(h, HML_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r__)); \
  } while (0)

int all_gather(hml_t* h, const void* send, void* recv, size_t bytes) {
  CKN(g_nccl.AllGather(send, recv, bytes, ncclUint8, h->comm, h->stream));
  return HML_OK;
}

// per-sweep carry exchange: peer mailboxes over NVLink (one kernel) or, without peer access, an NCCL all-gather
int exchange(hml_t* h, int slot, const void* send, void* recv, size_t bytes) {
  if (!h->p2p) return all_gather(h, send, recv, bytes);
  if (bytes % 8 != 0 || bytes > kP2PPayload) return fail(h, HML_ERR_ARG, "carry payload does not fit a mailbox entry");
  launch_p2p_exchange(h->p2p_dev, slot, ++h->p2p_seq[slot], send, bytes, recv, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  return HML_OK;
}

// after a stream synchronisation: did an exchange give up waiting for a peer?
int check_exchange(hml_t* h) {
  if (h->p2p && h->p2p_timeout_host && *h->p2p_timeout_host)
    return fail(h, HML_ERR_CUDA, "carry exchange timed out waiting for a peer rank");
  return HML_OK;
}

// all-gathers one of the per-sweep carries (send slot -> gathered array), on the handle's stream
int exchange_cb(void* user, int which) {
  hml_t* h = (hml_t*)user;
  SweepBuffers b = make_buffers(h, h->KP ? h->KP : 2);
  switch (which) {
    case kExchangeHeads: return exchange(h, kSlotHeads, b.seg.send_head, (void*)b.seg.heads, kHeadWords * sizeof(double));
    case kExchangeMaps: return exchange(h, kSlotMaps, b.seg.send_map, (void*)b.seg.maps, 4 * sizeof(uint64_t));
    case kExchangeOps: {
      const size_t n = (size_t)h->KP * h->KP + h->KP;
      return exchange(h, kSlotOps, b.seg.send_op, (void*)b.seg.ops, n * sizeof(double));
    }
    default: return HML_ERR_ARG;
  }
}


unsigned long long next_seq_cb(void* user, int which) {
  hml_t* h = (hml_t*)user;
  const int slot = which == kExchangeHeads ? kSlotHeads : which == kExchangeOps ? kSlotOps
                 : which == kExchangeMaps ? kSlotMaps : kSlotStats;
  return ++h->p2p_seq[slot];
}

// 8-byte words of the result block that travel to the host (and between ranks)
size_t result_words(int KP, int D) {
  size_t words = (size_t)KP + (size_t)KP * KP + 1;
  words += words & 1;
  return 2 + words + 2 * KP + 1 + (size_t)(D - 1) * 2 * KP;
}

int alloc_blocks(hml_t* h, uint64_t cap, int KP);

// Candidate list for thresholds >= floor: one pyramid pass at the floor gives the positions (it is the block list of
// that threshold), a gather their weights.  Grows the block arrays if the list does not fit (local to this rank; no
// collective has been issued yet at this point of a sweep).
// A list that holds most of the sequence is not worth having (a tiny threshold makes every position a candidate; the
// pass over the list would read 8 bytes per candidate where the pyramid pass reads the weights once): *usable = false
// then, the floor is remembered so the pass is not repeated sweep after sweep, and the caller takes the pyramid pass at
// the threshold itself.  The block arrays are never grown beyond T / kCandMaxShare on behalf of the list (low
// compression — C5: one block per 2 to 10 observations — still gets its list).
constexpr uint64_t kCandMaxShare = 2;
int rebuild_candidates(hml_t* h, float floor, bool* usable) {
  *usable = false;
  const uint64_t limit = h->T / kCandMaxShare;
  for (int attempt = 0; attempt < 8; ++attempt) {
    h->launches += launch_detect(h->w, h->smax, h->T, floor, h->rank == 0 ? 1 : 0, h->detect_scratch, h->starts, h->capacity,
                                 h->outblk, h->stream, nullptr, nullptr);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->outblk_host, h->outblk, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const uint64_t n = h->outblk_host[0];
    if (n > limit && n > (1u << 16)) {
      h->cand_valid = false;
      h->cand_too_long_floor = floor;
      return HML_OK;
    }
    if (n > h->capacity) {
      int rc = alloc_blocks(h, n + n / 4, h->KP);
      if (rc != HML_OK) return rc;
      continue;
    }
    if (n + 4 > h->cand_cap) {
      h->cand_cap = n + n / 4 + 4096;
      CK(dev_alloc(h->cand_pos, h->cand_cap));
      CK(dev_alloc(h->cand_w, h->cand_cap));
      if (h->D == 1) CK(dev_alloc(h->cand_pq, h->cand_cap));
    }
    if (h->D != 1) dev_free(h->cand_pq);
    const uint32_t ctas = cand_ctas((uint32_t)h->cand_cap);
    if (ctas > h->cand_scratch_ctas) {
      h->cand_scratch_ctas = ctas;
      CK(dev_alloc(h->cand_scratch, 2 * (size_t)ctas + 1));
      CK(cudaMemsetAsync(h->cand_scratch, 0, (2 * (size_t)ctas + 1) * sizeof(uint32_t), h->stream));
    }
    launch_cand_gather(h->w, h->pq, h->starts, (uint32_t)n, h->cand_w, h->cand_pos, h->cand_pq, h->sms, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    h->cand_n = n;
    h->cand_floor = floor;
    h->cand_valid = true;
    h->cand_rebuilds++;
    *usable = true;
    return HML_OK;
  }
  return fail(h, HML_ERR_CAPACITY, "block capacity did not converge");
}

// Makes the candidate list serve `thr` if it can: *usable says whether the boundaries of thr are a subset of the list.
int ensure_candidates(hml_t* h, float thr, bool* usable) {
  *usable = false;
  const float floor = 0.75f * thr;
  // candidate mode needs a floor that is a normal float (a denormal or zero floor selects every position) and that is
  // well above the last floor whose list came out about as long as the sequence
  if (!(h->detect_mode == HML_DETECT_CANDIDATES && thr > 0.f && isfinite(thr) && floor >= FLT_MIN &&
        !(floor <= 1.25f * h->cand_too_long_floor)))
    return HML_OK;
  // usable: every boundary of thr is a candidate; worth keeping: the list is not much longer than the block list
  bool ok = h->cand_valid && thr >= h->cand_floor;
  if (ok && h->blocks_valid && h->cand_n > 4 * h->nblocks + 65536 && floor > 1.05f * h->cand_floor) ok = false;
  if (!ok) {
    int rc = rebuild_candidates(h, floor, &ok);
    if (rc != HML_OK) return rc;
  }
  *usable = ok && h->cand_n > 0;
  return HML_OK;
}

int run_detect(hml_t* h, float thr) {
  bool done = false, head_done = false;
  // the block structure is being overwritten: whatever the previous sweep left (states, runs, rows) no longer
  // belongs to it, also if this sweep fails half-way
  h->states_valid = h->segs_valid = h->g_valid = h->rows_valid = false;
  h->last_thr = thr;
  {
    bool usable = false;
    int rc = ensure_candidates(h, thr, &usable);
    if (rc != HML_OK) return rc;
    if (usable) {
      // split sequence over peer mailboxes: the scatter's last CTA forms the head partial and exchanges the heads
      SweepBuffers hb;
      const bool with_head = h->world > 1 && h->p2p;
      if (with_head) hb = make_buffers(h, h->KP ? h->KP : 2);
      h->launches += launch_detect_candidates(h->cand_w, h->cand_pos, h->cand_pq, (uint32_t)h->cand_n, thr, h->cand_scratch,
                                              h->cand_scratch_ctas, h->starts, h->spq, h->pq, h->capacity, h->T, h->outblk,
                                              h->stream, stage_cb, h, with_head ? &hb : nullptr,
                                              with_head ? ++h->p2p_seq[kSlotHeads] : 0ull);
      h->spq_valid = h->cand_pq != nullptr;
      done = true;
      head_done = with_head;
    }
  }
  if (!done) h->spq_valid = false;
  if (!done) {  // thresholds <= 0, NaN, inf (and an empty candidate list): every weight is looked at
    h->launches += launch_detect(h->w, h->detect_mode != HML_DETECT_STREAM ? h->smax : nullptr, h->T, thr,
                                 h->rank == 0 ? 1 : 0, h->detect_scratch, h->starts, h->capacity, h->outblk, h->stream,
                                 stage_cb, h);
  }
  CK(cudaGetLastError());
  if (h->world > 1 && !head_done) {
    // the partial block in front of each rank's first boundary joins the last block of its owner
    stage_cb(h, "seg_head");
    launch_seg_head(make_buffers(h, h->KP ? h->KP : 2), h->T, h->p2p ? ++h->p2p_seq[kSlotHeads] : 0ull, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    if (h->p2p) return HML_OK;  // the kernel exchanged the heads itself
    stage_cb(h, "exchange_heads");
    return exchange_cb(h, kExchangeHeads);
  }
  return HML_OK;
}

void load_reset(hml_t* h) {
  dev_free(h->w);
  dev_free(h->smax);
  dev_free(h->coeffs);
  dev_free(h->pq);
  dev_free(h->cell_pref);
  if (h->detect_scratch) cudaFree(h->detect_scratch);
  h->detect_scratch = nullptr;
  h->T = 0;
  h->T_global = 0;
  h->seg_start = 0;
  h->D = 1;
  h->pq_stride = h->cell_stride = 0;
  h->cand_valid = false;
  h->cand_n = 0;
  h->cand_too_long_floor = -1.f;
  h->mg_K = 0;  // marginals belong to the sequence that was loaded
  h->mg_n = h->mg_iterations = 0;
  h->blocks_valid = h->stats_valid = h->states_valid = h->rows_valid = false;
}

// level normalisers: running fp32 product of sqrt2half, includes.hpp:126-128, wavelet.hpp:144,172
void load_norms(hml_t* h) {
  float norms[64];
  const float sqrt2 = (float)sqrt(2.0);
  const float sqrt2half = (float)(sqrt2 / 2.0);
  float n = sqrt2half;
  norms[0] = 1.0f;
  for (int l = 1; l < 64; ++l) {
    norms[l] = n;
    n *= sqrt2half;
  }
  upload_level_norms(norms, h->stream);
}

// sum of the odd-index coefficients of c[0..n) (main.cpp:303-311), fp64 partials summed on the host
int load_sum_odd(hml_t* h, const float* c, uint64_t n, double* out) {
  const int nb = 512;
  DevTmp<double> part;
  CK(part.alloc(nb));
  launch_sum_odd(c, n, part, nb, h->stream);
  h->launches++;
  std::vector<double> hp(nb);
  CK(cudaMemcpyAsync(hp.data(), part, nb * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  long double s = 0;
  for (double v : hp) s += v;
  *out = (double)s;
  return HML_OK;
}

// integral arrays + double-double cell prefix of the local observations x[0..T) (data dimension `dim` of `dims`:
// every dimension has its own plane, Statistics/IntegralArray.hpp:176-182)
int load_integral(hml_t* h, const float* x_dev, uint64_t T, int dim = 0, int dims = 1) {
  const uint64_t cells = T / kCell + 1;
  if (dim == 0) {
    h->pq_stride = cells * kCell;
    h->cell_stride = cells + 1;
    CK(dev_alloc(h->pq, h->pq_stride * dims));
    CK(dev_alloc(h->cell_pref, h->cell_stride * dims));
  }
  double2* const pq = h->pq + (size_t)dim * h->pq_stride;
  double4* const cell_pref = h->cell_pref + (size_t)dim * h->cell_stride;
  DevTmp<double2> cell_tot;
  CK(cell_tot.alloc(cells));
  launch_integral_cells(x_dev, T, pq, cell_tot, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  std::vector<double2> tot(cells);
  CK(cudaMemcpyAsync(tot.data(), cell_tot, cells * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  std::vector<double4> pref(cells + 1);
  // exclusive prefix in double-double (Knuth two-sum), so differences of far-apart cells stay exact to ~1e-32
  double hx = 0, lx = 0, hq = 0, lq = 0;
  auto dd_add = [](double& hi, double& lo, double v) {
    const double s = hi + v;
    const double bb = s - hi;
    const double err = (hi - (s - bb)) + (v - bb);
    const double l2 = lo + err;
    const double s2 = s + l2;
    lo = l2 - (s2 - s);
    hi = s2;
  };
  for (uint64_t c = 0; c <= cells; ++c) {
    pref[c] = make_double4(hx, lx, hq, lq);
    if (c < cells) {
      dd_add(hx, lx, tot[c].x);
      dd_add(hq, lq, tot[c].y);
    }
  }
  CK(cudaMemcpyAsync(cell_pref, pref.data(), (cells + 1) * sizeof(double4), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return HML_OK;
}

// boundary-detection scratch and the initial block capacity (grows on demand: a sweep that overflows is re-run)
int load_finish(hml_t* h, uint64_t T) {
  CK(dev_alloc(h->smax, pyramid_entries(T)));
  launch_build_pyramid(h->w, T, h->smax, h->sms, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMalloc(&h->detect_scratch, detect_scratch_bytes(T)));
  CK(cudaMemsetAsync(h->detect_scratch, 0, detect_scratch_bytes(T), h->stream));
  h->T = T;
  uint64_t cap = T / 64;
  if (cap < (1u << 16)) cap = 1u << 16;
  if (cap > T) cap = T;
  // capacity = 0 makes alloc_blocks reallocate everything, the K-sized per-sweep buffers of an earlier sequence
  // included (h->KP is kept: a sweep with the same K must not find them at the old capacity)
  h->capacity = 0;
  int rc = alloc_blocks(h, cap, h->KP);
  if (rc != HML_OK) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return HML_OK;
}

// Haar passes above the first: `in` holds per-tile sums of the previous pass; coefficients of pass p go to
// coeffs[j * stride] (stride counted in elements of `coeffs`).
int load_upper_passes(hml_t* h, const float* in, uint64_t n_valid, uint64_t n_pos, uint64_t stride, int level0,
                      float* coeffs, float* sums[2]) {
  int which = 0;
  while (n_pos > 1) {
    launch_maxlet_level(in, n_valid, n_pos, stride, level0, coeffs, sums[which], h->stream);
    h->launches++;
    in = sums[which];
    which ^= 1;
    n_valid = n_valid / kTile;
    n_pos = (n_pos + kTile - 1) / kTile;
    stride *= kTile;
    level0 += kTileLog2;
  }
  CK(cudaGetLastError());
  return HML_OK;
}

// x_dev: T positions x D values, position-major (D = 1: the plain sequence)
int load_common(hml_t* h, const float* x_dev, uint64_t T, float mult, int D = 1) {
  if (h->world > 1) return fail(h, HML_ERR_STATE, "this handle joined a communicator: use hml_load_segment_f32");
  load_reset(h);
  h->D = D;
  const uint64_t tiles = (T + kTile - 1) / kTile;
  load_norms(h);

  // ---- maxlet coefficients, 12 levels per pass
  CK(dev_alloc(h->coeffs, tiles * kTile));
  DevTmp<float> sum0, sum1, plane, cdim;  // plane, cdim: multivariate input, one dimension at a time
  CK(sum0.alloc(tiles + 1));
  CK(sum1.alloc(tiles / kTile + 2));
  float* sums[2] = {sum0, sum1};
  int rc = HML_OK;
  if (D == 1) {
    rc = load_upper_passes(h, x_dev, T, T, 1, 0, h->coeffs, sums);
    if (rc != HML_OK) return rc;
  } else {
    CK(plane.alloc(T));
    CK(cdim.alloc(tiles * kTile));
    for (int d = 0; d < D; ++d) {
      launch_deinterleave(x_dev, T, D, d, plane, h->sms, h->stream);
      h->launches++;
      // wavelet.hpp:150-163: the coefficient is the maximum over the dimensions of the normalised |detail|
      rc = load_upper_passes(h, plane, T, T, 1, 0, d == 0 ? h->coeffs : cdim, sums);
      if (rc != HML_OK) return rc;
      if (d > 0) {
        launch_max_combine(h->coeffs, cdim, T, h->sms, h->stream);
        h->launches++;
      }
      rc = load_integral(h, plane, T, d, D);
      if (rc != HML_OK) return rc;
    }
    cdim.release();
  }
  const float inf = INFINITY;
  CK(cudaMemcpyAsync(h->coeffs, &inf, sizeof(float), cudaMemcpyHostToDevice, h->stream));  // wavelet.hpp:183

  // ---- sigma-hat (main.cpp:303-311)
  {
    double s = 0;
    rc = load_sum_odd(h, h->coeffs, T, &s);
    if (rc != HML_OK) return rc;
    const uint64_t n = T / 2;
    double est = n ? s / (double)n : NAN;
    est /= 0.797884560802865355879892119868763736951717262329869315331;
    h->sigma_hat = est;
  }

  // ---- breakpoint weights
  CK(dev_alloc(h->w, tiles * kTile));
  launch_bp_weights(h->coeffs, T, mult, h->w, h->sms, h->stream);
  h->launches++;
  CK(cudaGetLastError());

  if (D == 1) {
    rc = load_integral(h, x_dev, T);
    if (rc != HML_OK) return rc;
  }
  plane.release();
  sum0.release();
  sum1.release();
  if (T > (1ull << 26)) dev_free(h->coeffs);  // 4 B/observation is not worth keeping for big inputs
  rc = load_finish(h, T);
  h->T_global = T;
  return rc;
}

// Tiles (4096 observations) are dealt out as evenly as possible: the first tiles % world ranks hold one more.
uint64_t plan_first_tile(uint64_t tiles, int world, int rank) {
  const uint64_t q = tiles / world, rem = tiles % world;
  return (uint64_t)rank * q + ((uint64_t)rank < rem ? (uint64_t)rank : rem);
}
void segment_plan(uint64_t T, int world, int rank, uint64_t* start, uint64_t* len) {
  const uint64_t tiles = (T + kTile - 1) / kTile;
  uint64_t s = plan_first_tile(tiles, world, rank) * kTile, e = plan_first_tile(tiles, world, rank + 1) * kTile;
  if (s > T) s = T;
  if (e > T) e = T;
  *start = s;
  *len = e - s;
}

// Collective load of one segment.  Levels 1..12 of the Haar transform are local to 4096-tiles; the levels
// above are computed by every rank from the all-gathered tile sums (T/4096 floats), so no rank needs
// another rank's observations.
int load_segment_common(hml_t* h, const float* x_dev, uint64_t len, uint64_t T, float mult, int D = 1) {
  if (h->world <= 1 || !h->comm) return fail(h, HML_ERR_STATE, "hml_comm_init has not been called on this handle");
  uint64_t start = 0, plan_len = 0;
  if (T < (uint64_t)kTile * h->world) return fail(h, HML_ERR_ARG, "sequence too short to split: T < 4096 * world");
  segment_plan(T, h->world, h->rank, &start, &plan_len);
  if (len != plan_len || len == 0) return fail(h, HML_ERR_ARG, "segment length does not match hml_segment_plan");
  load_reset(h);
  h->D = D;
  load_norms(h);
  const int world = h->world;
  const uint64_t tiles_total = (T + kTile - 1) / kTile;
  const uint64_t per = (tiles_total + world - 1) / world;
  const uint64_t tiles = (len + kTile - 1) / kTile;
  const uint64_t slot = per + 16;  // floats per rank in the all-gather: tile sums + 12 edge coefficients
  const uint64_t top_tiles = (tiles_total + kTile - 1) / kTile;

  CK(dev_alloc(h->coeffs, tiles * kTile));
  DevTmp<float> send, recv, gsum, ctop, ctop_d, sum0, sum1, plane, cdim;
  CK(send.alloc(slot));
  CK(recv.alloc(slot * world));
  CK(gsum.alloc(per * world));
  CK(ctop.alloc(top_tiles * kTile));
  CK(sum0.alloc(top_tiles + 1));
  CK(sum1.alloc(top_tiles / kTile + 2));
  if (D > 1) {
    CK(plane.alloc(len));
    CK(cdim.alloc(tiles * kTile));
    CK(ctop_d.alloc(top_tiles * kTile));
  }
  std::vector<float> infs(top_tiles * kTile, INFINITY);
  int rc = HML_OK;
  // Per data dimension (wavelet.hpp:150-163: the coefficient is the maximum over the dimensions of the normalised
  // |detail|): levels 1..12 on the local observations, the tile sums of all ranks, the levels above from those.
  for (int d = 0; d < D; ++d) {
    const float* xd = x_dev;
    if (D > 1) {
      launch_deinterleave(x_dev, len, D, d, plane, h->sms, h->stream);
      h->launches++;
      xd = plane;
    }
    float* const cl = d == 0 ? h->coeffs : cdim.p;   // local coefficients of this dimension
    float* const ct = d == 0 ? ctop.p : ctop_d.p;    // coefficients at multiples of 4096 (replicated)
    CK(cudaMemsetAsync(send, 0, slot * sizeof(float), h->stream));
    launch_maxlet_level(xd, len, len, 1, 0, cl, send, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    rc = all_gather(h, send, recv, slot * sizeof(float));
    if (rc != HML_OK) return rc;
    for (int r = 0; r < world; ++r) {
      const uint64_t t0 = plan_first_tile(tiles_total, world, r), t1 = plan_first_tile(tiles_total, world, r + 1);
      if (t1 > t0)
        CK(cudaMemcpyAsync(gsum + t0, recv + (size_t)r * slot, (t1 - t0) * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    }
    CK(cudaMemcpyAsync(ct, infs.data(), infs.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    float* sums[2] = {sum0, sum1};
    rc = load_upper_passes(h, gsum, T / kTile, tiles_total, 1, kTileLog2, ct, sums);
    if (rc != HML_OK) return rc;
    if (d > 0) {
      launch_max_combine(h->coeffs, cdim, len, h->sms, h->stream);
      launch_max_combine(ctop, ctop_d, top_tiles * kTile, h->sms, h->stream);
      h->launches += 2;
    }
    if (D > 1) {
      rc = load_integral(h, plane, len, d, D);
      if (rc != HML_OK) return rc;
    }
    CK(cudaStreamSynchronize(h->stream));
  }
  // ---- the 12 edge coefficients of the (combined) local array for the next rank
  CK(cudaMemsetAsync(send, 0, slot * sizeof(float), h->stream));
  launch_pack_edge(h->coeffs, len, send + per, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  rc = all_gather(h, send, recv, slot * sizeof(float));
  if (rc != HML_OK) return rc;

  // ---- sigma-hat: odd positions are odd locally too (segments start at multiples of 4096)
  {
    double part = 0;
    rc = load_sum_odd(h, h->coeffs, len, &part);
    if (rc != HML_OK) return rc;
    DevTmp<double> ds, dr;
    CK(ds.alloc(1));
    CK(dr.alloc(world));
    CK(cudaMemcpyAsync(ds, &part, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    rc = all_gather(h, ds, dr, sizeof(double));
    if (rc != HML_OK) return rc;
    std::vector<double> parts(world);
    CK(cudaMemcpyAsync(parts.data(), dr, world * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    long double s = 0;
    for (double v : parts) s += v;
    const uint64_t n = T / 2;
    h->sigma_hat = (double)(s / (long double)n) / 0.797884560802865355879892119868763736951717262329869315331;
  }

  // ---- breakpoint weights of the local segment
  CK(dev_alloc(h->w, tiles * kTile));
  const float* halo = recv + (size_t)(h->rank > 0 ? h->rank - 1 : 0) * slot + per;
  launch_bp_weights_segment(h->coeffs, ctop, halo, start, len, T, mult, h->w, h->sms, h->stream);
  h->launches++;
  CK(cudaGetLastError());

  if (D == 1) {
    rc = load_integral(h, x_dev, len);
    if (rc != HML_OK) return rc;
  }
  CK(cudaStreamSynchronize(h->stream));
  if (len > (1ull << 26)) dev_free(h->coeffs);
  rc = load_finish(h, len);
  h->T_global = T;
  h->seg_start = start;
  return rc;
}

// P = number of emission parameters (K for univariate data), map[s * D + d] = parameter of state s in dimension d
int validate_model(hml_t* h, const hml_model* m, ModelHost& mh, int& P, std::vector<int>& map) {
  if (!m || !m->mean || !m->var || !m->A || !m->pi) return fail(h, HML_ERR_ARG, "model pointers must not be NULL");
  if (m->K < 2 || m->K > HML_MAX_STATES)
    return fail(h, HML_ERR_ARG, "number of states must be in [2, " + std::to_string(HML_MAX_STATES) + "]");
  memset(&mh, 0, sizeof(mh));
  mh.K = m->K;
  mh.D = h->D;
  mh.use_self = m->use_self_transitions ? 1 : 0;
  const int D = h->D;
  const bool mapped = m->mapping != nullptr || m->nr_dims > 1;
  if (D > 1 || mapped) {
    if (m->nr_dims != D)
      return fail(h, HML_ERR_ARG, "model.nr_dims (" + std::to_string(m->nr_dims) + ") does not match the loaded data (" +
                                      std::to_string(D) + " dimensions)");
    if (!m->mapping || m->nr_params < 1 || m->nr_params > HML_MAX_STATES)
      return fail(h, HML_ERR_ARG, "multivariate data needs model.mapping and model.nr_params in [1, 32]");
    P = m->nr_params;
    map.assign(m->mapping, m->mapping + (size_t)m->K * D);
    for (int v : map)
      if (v < 0 || v >= P) return fail(h, HML_ERR_ARG, "model.mapping refers to a parameter outside [0, nr_params)");
  } else {
    P = m->K;
    map.resize(m->K);
    for (int i = 0; i < m->K; ++i) map[i] = i;
  }
  for (int p = 0; p < P; ++p) {
    if (!isfinite(m->mean[p])) return fail(h, HML_ERR_ARG, "Mean must be set to a finite value!");
    if (!(m->var[p] > 0) || !isfinite(m->var[p])) return fail(h, HML_ERR_ARG, "Variance must be positive!");
  }
  for (int i = 0; i < m->K; ++i) {
    if (!(m->pi[i] >= 0)) return fail(h, HML_ERR_NUMERIC, "Negative backward variable!");
    for (int d = 0; d < D; ++d) {
      mh.mean_sd[d][i] = m->mean[map[(size_t)i * D + d]];
      mh.var_sd[d][i] = m->var[map[(size_t)i * D + d]];
    }
    mh.mean[i] = mh.mean_sd[0][i];
    mh.var[i] = mh.var_sd[0][i];
    mh.pi[i] = m->pi[i];
    for (int j = 0; j < m->K; ++j) {
      const double a = m->A[i * m->K + j];
      if (!(a >= 0)) return fail(h, HML_ERR_NUMERIC, "Negative backward variable!");  // FB.hpp:147-149
      mh.A[i * m->K + j] = a;
    }
  }
  return HML_OK;
}

// Copies the result block(s) of the sweep to the host.  Single handle: its own block.  Segment mode: the
// blocks of all ranks are all-gathered first; `res` then holds the rank-ordered sums (identical on every rank).
struct SweepResult {
  std::vector<unsigned long long> o64;  // [0..KP) n, [KP..KP+KP*KP) trans, [KP+KP*KP] fallbacks
  std::vector<double> of;               // [0..KP) sum, [KP..2KP) sumsq, [2KP] loglik
  uint64_t local_blocks = 0, global_blocks = 0, first_block = 0;
  bool any_overflow = false, own_overflow = false;
};

// exchanged: the sweep's last kernel already gathered the result blocks of all ranks into stats_gather
// on_host: the sweep's last kernel wrote the block(s) into the pinned mirror itself (SweepBuffers::result_host)
int fetch_result(hml_t* h, int KP, SweepResult& res, bool exchanged, bool on_host = false) {
  size_t words = (size_t)KP + (size_t)KP * KP + 1;
  words += words & 1;
  const size_t copy_words = result_words(KP, h->D);
  const int nf = 2 * KP + 1 + (h->D - 1) * 2 * KP;  // + per-state sums of the further dimensions
  res.o64.assign(words, 0);
  res.of.assign(nf, 0.0);
  const int world = h->world > 1 ? h->world : 1;
  const unsigned long long* host = h->outblk_host;
  if (world > 1) {
    if (!exchanged) {
      int rc = exchange(h, kSlotStats, h->outblk, h->stats_gather, copy_words * 8);
      if (rc != HML_OK) return rc;
    }
    if (!(on_host && exchanged))
      CK(cudaMemcpyAsync(h->stats_gather_host, h->stats_gather, world * copy_words * 8, cudaMemcpyDeviceToHost, h->stream));
    host = h->stats_gather_host;
  } else if (!on_host) {
    CK(cudaMemcpyAsync(h->outblk_host, h->outblk, copy_words * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  h->mg_pending = false;
  if (int rc = check_exchange(h)) return rc;
  res.global_blocks = res.first_block = 0;
  res.any_overflow = res.own_overflow = false;
  for (int r = 0; r < world; ++r) {
    const unsigned long long* blk = host + (size_t)r * copy_words;
    const uint64_t raw = blk[0];
    const bool over = world > 1 ? blk[1] != 0 : raw > h->capacity;
    if (r == h->rank || world == 1) {
      res.local_blocks = raw;
      res.own_overflow = over;
    }
    res.any_overflow |= over;
    if (r < h->rank) res.first_block += raw;
    res.global_blocks += raw;
    const unsigned long long* o64 = blk + 2;
    const double* of = (const double*)(blk + 2 + words);
    for (size_t i = 0; i < words; ++i) res.o64[i] += o64[i];
    for (int i = 0; i < nf; ++i) res.of[i] += of[i];
  }
  return HML_OK;
}

// ---- the fused path (hml_fused.cuh)

int chain_alloc(hml_t* h) {
  if (!h->chain_dev) {
    CK(dev_alloc(h->chain_dev, 1));
    CK(cudaMallocHost((void**)&h->chain_host, sizeof(ChainDev)));
    memset(h->chain_host, 0, sizeof(ChainDev));
    CK(dev_alloc(h->fused_cta_count, 1024));
    CK(dev_alloc(h->fused_barriers, 1 + kFusedMaxTiles));
    CK(dev_alloc(h->fused_qtot, (size_t)4 * kFusedMaxTiles * (kChainMaxStates * kChainMaxStates + kChainMaxStates)));
    CK(dev_alloc(h->fused_qmap, (size_t)4 * kFusedMaxTiles));
    CK(dev_alloc(h->fused_subops, (size_t)4 * kFusedMaxTiles * 32 * (kChainMaxStates * kChainMaxStates + kChainMaxStates)));
    CK(dev_alloc(h->fused_submaps, (size_t)4 * kFusedMaxTiles * 32));
    CK(dev_alloc(h->fused_phase_ns, 16));
    CK(cudaMemsetAsync(h->fused_phase_ns, 0, 16 * 8, h->stream));
    CK(dev_alloc(h->chain_stats, kOutWords));
  }
  return HML_OK;
}

// can this handle run sweeps of K states through the fused kernel at all?
bool fused_possible(const hml_t* h, int KP) {
  return KP >= 2 && KP <= kChainMaxStates && h->D == 1 && h->world <= 1 && h->detect_mode == HML_DETECT_CANDIDATES &&
         h->T < (1ull << 32);
}

// grid of the cooperative launch: one CTA per tile the block list can have (the candidates bound it), at most one per SM
int fused_grid(const hml_t* h, int KP) {
  const int max_grid = fused_max_grid(KP, h->sms);
  if (max_grid <= 0) return 0;
  // one CTA per quarter tile (256 blocks) the block list can have — the candidates bound it —, a multiple of four
  uint64_t g = 4 * ((h->cand_n + kTileBlocks - 1) / kTileBlocks);
  if (g < 4) g = 4;
  if (g > (uint64_t)max_grid) g = max_grid / 4 * 4;
  if (g > 1024) g = 1024;
  return (int)g;
}

// Launches `nsweeps` fused sweeps on the chain state as it stands on the device (h->chain_host is uploaded first if
// `upload`), waits, and mirrors the chain back.  The candidate list must serve the chain's threshold.
int fused_launch(hml_t* h, int KP, int nsweeps, bool sample_params, bool philox_from_chain, uint64_t seed, uint64_t sweep,
                 const double* replay_dev, bool upload) {
  ChainDev* c = h->chain_host;
  c->abort_code = 0;
  c->sweeps_done = 0;
  c->nblocks_seen = 0;
  c->cand_floor = h->cand_floor;
  for (unsigned int& f : c->phase_abort) f = 0;
  if (upload) {
    CK(cudaMemcpyAsync(h->chain_dev, c, sizeof(ChainDev), cudaMemcpyHostToDevice, h->stream));
  } else {  // only the launch status and what the host knows better (the floor of the current list)
    const size_t off = offsetof(ChainDev, abort_code);
    CK(cudaMemcpyAsync((char*)h->chain_dev + off, (const char*)c + off, sizeof(ChainDev) - off, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(&h->chain_dev->cand_floor, &c->cand_floor, sizeof(float), cudaMemcpyHostToDevice, h->stream));
  }
  h->spq_valid = true;  // the kernel fills spq from the candidates' pairs
  SweepBuffers b = make_buffers(h, KP);
  b.replay_u = replay_dev;
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.chain = h->chain_dev;
  a.cand_w = h->cand_w;
  a.cand_pos = h->cand_pos;
  a.cand_pq = h->cand_pq;
  a.nc = (uint32_t)h->cand_n;
  a.T_local = (uint32_t)h->T;
  a.cta_count = h->fused_cta_count;
  a.barriers = h->fused_barriers;
  a.qtot = h->fused_qtot;
  a.qmap = h->fused_qmap;
  a.phase_ns = h->timing ? h->fused_phase_ns : nullptr;
  a.subops = h->fused_subops;
  a.submaps = h->fused_submaps;
  CK(cudaMemsetAsync(h->fused_barriers, 0, (1 + kFusedMaxTiles) * sizeof(unsigned), h->stream));
  a.nsweeps = nsweeps;
  a.sample_params = sample_params ? 1 : 0;
  a.philox_sweep_from_chain = philox_from_chain ? 1 : 0;
  a.seed = seed;
  a.sweep = sweep;
  const int grid = fused_grid(h, KP);
  if (grid <= 0) return fail(h, HML_ERR_CUDA, "the fused sweep kernel does not fit this device");
  const int e = launch_sweep_fused(KP, b, a, grid, h->stream);
  if (e == -2) return fail(h, HML_ERR_ARG, "unsupported number of states for the fused sweep");
  if (e != 0) return fail(h, HML_ERR_CUDA, std::string("fused sweep launch: ") + cudaGetErrorString((cudaError_t)e));
  h->launches++;
  CK(cudaMemcpyAsync(c, h->chain_dev, sizeof(ChainDev), cudaMemcpyDeviceToHost, h->stream));
  return HML_OK;  // the caller synchronises (fetch_result)
}

int sweep_common(hml_t* h, const hml_model* m, uint32_t flags, float thr, uint64_t seed, uint64_t sweep,
                 const double* replay, uint64_t n_replay, hml_sweep_out* out, bool mixture) {
  if (!h) return HML_ERR_ARG;
  if (h->T == 0) return fail(h, HML_ERR_STATE, "no data loaded");
  if (!out) return fail(h, HML_ERR_ARG, "out must not be NULL");
  CK(cudaSetDevice(h->device));
  ModelHost mh;
  int P = 0;
  std::vector<int> map;
  int rc = validate_model(h, m, mh, P, map);
  if (rc != HML_OK) return rc;
  const int KP = padded_states(mh.K);
  const bool dynamic = (flags & HML_SWEEP_DYNAMIC) != 0;
  const bool seg = h->world > 1;
  if (replay && dynamic)
    return fail(h, HML_ERR_ARG, "replay uniforms need a fixed block structure: call hml_create_blocks first");
  if (!dynamic && !h->blocks_valid) return fail(h, HML_ERR_STATE, "no block structure: call hml_create_blocks first");
  const uint64_t replay_need = seg ? h->global_blocks : h->nblocks;
  if (replay && n_replay < replay_need) return fail(h, HML_ERR_ARG, "fewer replay uniforms than blocks");
  if (seg && (flags & HML_SWEEP_KEEP_ROWS))
    return fail(h, HML_ERR_ARG, "forward rows are not kept in segment mode");
  bool fused = (flags & HML_SWEEP_FUSED) != 0;
  if (fused) {
    if (mixture || (flags & (HML_SWEEP_LOGLIK | HML_SWEEP_KEEP_ROWS)) || !fused_possible(h, KP))
      return fail(h, HML_ERR_ARG, "HML_SWEEP_FUSED: forward-backward sweeps of at most 8 states on a single univariate handle "
                                  "in candidate detection mode, without log-likelihood or kept rows");
    if (!dynamic && !(h->last_thr == h->last_thr)) return fail(h, HML_ERR_STATE, "no block structure: call hml_create_blocks first");
    rc = chain_alloc(h);
    if (rc != HML_OK) return rc;
  }

  for (int attempt = 0; attempt < 8; ++attempt) {
    if (KP != h->KP) {
      // per-block statistics survive (their layout does not depend on K); only K-sized buffers change
      rc = alloc_blocks(h, h->capacity, KP);
      if (rc != HML_OK) return rc;
    }
    if ((flags & HML_SWEEP_KEEP_ROWS) && h->rows_cap < (h->capacity + 1) * (uint64_t)mh.K) {
      h->rows_cap = (h->capacity + 1) * (uint64_t)mh.K;
      CK(dev_alloc(h->rows, h->rows_cap));
    }
    if (replay) {
      if (h->replay_cap < replay_need) {
        h->replay_cap = replay_need > h->capacity ? replay_need : h->capacity;
        CK(dev_alloc(h->replay_u, h->replay_cap));
      }
      CK(cudaMemcpyAsync(h->replay_u, replay, replay_need * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    h->stages.clear();
    h->stage_used = 0;
    if (fused) {
      // one persistent kernel does the whole sweep, boundary detection included (from the candidate list)
      const float t_use = dynamic ? thr : h->last_thr;
      bool usable = false;
      rc = ensure_candidates(h, t_use, &usable);
      if (rc != HML_OK) return rc;
      if (!usable || h->cand_pq == nullptr || h->cand_n > (uint64_t)2 * kFusedMaxTiles * kTileBlocks) {
        fused = false;  // the threshold is not one the candidate list serves, or far too many blocks: the multi-kernel path
        continue;
      }
      ChainDev* c = h->chain_host;
      const bool keep_chain = h->chain_ready;  // a device-resident chain keeps its priors and RNG position
      ChainDev saved;
      if (keep_chain) saved = *c;
      for (int i = 0; i < mh.K; ++i) {
        c->mean[i] = mh.mean[i];
        c->var[i] = mh.var[i];
        c->pi[i] = mh.pi[i];
        for (int j = 0; j < mh.K; ++j) c->A[i * mh.K + j] = mh.A[i * mh.K + j];
      }
      c->K = mh.K;
      c->use_self = mh.use_self;
      c->T = (uint32_t)h->T;
      c->thr = t_use;
      chain_derive(c);
      h->last_thr = t_use;
      h->states_valid = h->segs_valid = h->g_valid = h->rows_valid = false;
      h->blocks_valid = false;
      stage_cb(h, "fused_sweep");
      rc = fused_launch(h, KP, 1, false, false, seed, sweep, replay ? h->replay_u : nullptr, true);
      stage_cb(h, "end");
      if (rc != HML_OK) return rc;
      SweepResult res;
      rc = fetch_result(h, KP, res, false);
      if (rc != HML_OK) return rc;
      const unsigned code = c->abort_code;
      const unsigned long long seen = c->nblocks_seen;
      if (keep_chain) {  // the single sweep borrowed the chain's slot for its model: put the chain back
        *c = saved;
        CK(cudaMemcpyAsync(h->chain_dev, c, sizeof(ChainDev), cudaMemcpyHostToDevice, h->stream));
      }
      if (code == kChainCapacity && seen > h->capacity && seen <= (uint64_t)kFusedMaxTiles * kTileBlocks) {
        rc = alloc_blocks(h, seen + seen / 4, KP);
        if (rc != HML_OK) return rc;
        continue;
      }
      if (code != kChainOk) {  // too many blocks for one CTA per tile, a vanished forward sum, ...: the multi-kernel path
        fused = false;
        continue;
      }
      collect_timing(h);
      h->nblocks = res.local_blocks;
      h->global_blocks = res.global_blocks;
      h->first_block = 0;
      h->blocks_valid = h->stats_valid = true;
      out->nblocks = res.global_blocks;
      out->uniform_fallbacks = 0;
      out->loglik = NAN;
      for (int s2 = 0; s2 < mh.K; ++s2) {
        if (out->counts) out->counts[s2] = res.o64[s2];
        if (out->stat_n) out->stat_n[s2] = res.o64[s2];
        if (out->stat_sum) out->stat_sum[s2] = res.of[s2];
        if (out->stat_sumsq) out->stat_sumsq[s2] = res.of[KP + s2];
        if (out->trans)
          for (int j = 0; j < mh.K; ++j) out->trans[s2 * mh.K + j] = res.o64[KP + s2 * KP + j];
      }
      h->states_valid = true;
      h->rows_valid = false;
      h->last_K = mh.K;
      return HML_OK;
    }
    const auto hc0 = std::chrono::steady_clock::now();
    bool gather = !h->stats_valid;
    if (dynamic) {
      rc = run_detect(h, thr);
      if (rc != HML_OK) return rc;
      gather = true;
      h->blocks_valid = false;
      if ((flags & HML_SWEEP_KEEP_ROWS) && h->rows_cap < (h->capacity + 1) * (uint64_t)mh.K) {
        h->rows_cap = (h->capacity + 1) * (uint64_t)mh.K;  // the candidate build may have grown the block arrays
        CK(dev_alloc(h->rows, h->rows_cap));
      }
      if (replay && h->replay_cap < h->capacity) return fail(h, HML_ERR_STATE, "replay buffer smaller than the block capacity");
    }
    SweepBuffers b = make_buffers(h, KP);
    b.rows = (flags & HML_SWEEP_KEEP_ROWS) ? h->rows : nullptr;
    b.replay_u = replay ? h->replay_u : nullptr;
    SweepLaunch l{};
    l.flags = flags;
    l.gather = gather;
    l.mixture = mixture;
    l.seed = seed;
    l.sweep = sweep;
    l.sms = h->sms;
    // upper bound of the block count for grid sizes: boundaries found among the candidates cannot outnumber them
    l.nblocks_hint = dynamic ? ((h->spq_valid && h->cand_n < h->capacity) ? h->cand_n : h->capacity) : h->nblocks;
    l.exchange = exchange_cb;
    l.next_seq = next_seq_cb;
    l.exchange_user = h;
    // (multivariate data: the sums of the further dimensions are reduced after the first, so the result blocks are
    // exchanged by a kernel of their own afterwards)
    const bool fused_stats = seg && h->p2p && h->D == 1;
    l.stats_words = fused_stats ? (uint32_t)result_words(KP, 1) : 0u;
    // the speculative forward filter: single handle, forward-backward sweeps; after a failure the operator scan takes
    // the next sweeps (1, 2, 4, ... up to 64 after failures in a row), so data on which the filter does not forget its
    // start pays the wasted pass rarely
    // (segment mode: every rank takes the same decision — the failure counter arrives summed over the ranks)
    const bool spec_ok = !mixture && !(seg && (flags & HML_SWEEP_LOGLIK));
    l.speculate = spec_ok && h->forward_mode != HML_FORWARD_OPERATORS &&
                  (h->forward_mode == HML_FORWARD_SPECULATIVE || h->spec_skip == 0);
    if (!l.speculate && h->spec_skip > 0 && spec_ok) h->spec_skip--;
    l.spec_warm = spec_warm_of(KP, h->spec_level);
    l.spec_sub = spec_sub_of(KP, h->spec_level);
    bool on_host = false;
    l.result_on_host = &on_host;
    const auto hc1 = std::chrono::steady_clock::now();
    int n = launch_sweep(mh, b, l, h->stream, stage_cb, h);
    if (n == -2) return fail(h, HML_ERR_ARG, "unsupported number of states");
    if (n < 0) return h->err.empty() ? fail(h, HML_ERR_CUDA, "carry exchange failed") : HML_ERR_CUDA;
    h->launches += n;
    CK(cudaGetLastError());
    const auto hc2 = std::chrono::steady_clock::now();
    SweepResult res;
    rc = fetch_result(h, KP, res, fused_stats, on_host);
    if (rc != HML_OK) return rc;
    {
      const auto hc3 = std::chrono::steady_clock::now();
      using ns = std::chrono::nanoseconds;
      h->host_ns_prepare += std::chrono::duration_cast<ns>(hc1 - hc0).count();
      h->host_ns_launch += std::chrono::duration_cast<ns>(hc2 - hc1).count();
      h->host_ns_wait += std::chrono::duration_cast<ns>(hc3 - hc2).count();
      h->host_sweeps++;
    }
    if (l.speculate && !(dynamic && res.any_overflow)) {
      h->spec_sweeps++;
      if (res.o64[KP + KP * KP + 1] > 0) {  // some chunk's rows did not meet the guess's: the exact operator scan
        // first remedy: longer pieces with longer warm-ups (the filter forgets, but not within a piece); past the last
        // level the operator scan takes the next 1, 3, 7, ... 63 sweeps
        h->spec_failures++;
        // failing right after stepping down a level: that level is not for this data, ask for more patience next time
        if (h->spec_lowered && h->spec_good < 8 && h->spec_patience < 16384) h->spec_patience *= 4;
        h->spec_lowered = false;
        h->spec_good = 0;
        if (h->spec_level + 1 < kSpecLevels) {
          h->spec_level++;
        } else {
          h->spec_streak = h->spec_streak < 6 ? h->spec_streak + 1 : 6;
          h->spec_skip = (1u << h->spec_streak) - 1;
        }
        l.speculate = false;
        l.gather = false;  // the block sums of this block list are in place
        l.nblocks_hint = res.local_blocks;
        h->stages.clear();
        h->stage_used = 0;
        on_host = false;
        n = launch_sweep(mh, b, l, h->stream, stage_cb, h);
        if (n < 0) return h->err.empty() ? fail(h, HML_ERR_CUDA, "carry exchange failed") : HML_ERR_CUDA;
        h->launches += n;
        CK(cudaGetLastError());
        rc = fetch_result(h, KP, res, fused_stats, on_host);
        if (rc != HML_OK) return rc;
      } else {
        h->spec_streak = 0;
        if (++h->spec_good >= h->spec_patience && h->spec_level > 0) {
          h->spec_level--;
          h->spec_good = 0;
          h->spec_lowered = true;
        }
      }
    }
    if (dynamic && res.any_overflow) {  // some rank's block arrays were too small: grow and run the sweep again
      if (res.own_overflow) {
        rc = alloc_blocks(h, res.local_blocks + res.local_blocks / 4, KP);
        if (rc != HML_OK) return rc;
      }
      continue;
    }
    const uint64_t B = res.local_blocks;
    h->nblocks = B;
    h->global_blocks = res.global_blocks;
    h->first_block = res.first_block;
    h->blocks_valid = true;
    h->stats_valid = true;
    uint64_t fallbacks = res.o64[KP + KP * KP];
    if (fallbacks > 0 && !mixture) {
      // A zero forward sum resets the filter to uniform (FB.hpp:106-111); that is not an operator
      // product, so the exact sequential recursion is run instead (still on the device).
      l.nblocks_hint = B;
      const int n2 = launch_sweep_sequential(mh, b, l, h->stream);
      if (n2 < 0) return h->err.empty() ? fail(h, HML_ERR_CUDA, "carry exchange failed") : HML_ERR_CUDA;
      h->launches += n2;
      CK(cudaGetLastError());
      rc = fetch_result(h, KP, res, fused_stats);
      if (rc != HML_OK) return rc;
      fallbacks = res.o64[KP + KP * KP];
    }
    collect_timing(h);
    out->nblocks = res.global_blocks;
    out->uniform_fallbacks = fallbacks;
    out->loglik = (flags & HML_SWEEP_LOGLIK) ? res.of[2 * KP] : NAN;
    for (int s = 0; s < mh.K; ++s) {
      if (out->counts) out->counts[s] = res.o64[s];  // occupancy: observations per state
      if (out->trans)
        for (int j = 0; j < mh.K; ++j) out->trans[s * mh.K + j] = res.o64[KP + s * KP + j];
    }
    if (h->D == 1 && P == mh.K) {
      for (int s = 0; s < mh.K; ++s) {  // univariate: state == parameter
        if (out->stat_sum) out->stat_sum[s] = res.of[s];
        if (out->stat_sumsq) out->stat_sumsq[s] = res.of[KP + s];
        if (out->stat_n) out->stat_n[s] = res.o64[s];
      }
    } else {
      // ForwardBackward.hpp:189-191: stats[mapping[state][d]].add(y.suffStat(d), N) — the per-(state, dimension)
      // sums of the device are folded into the parameters in a fixed order (state-major, dimension-minor)
      for (int p = 0; p < P; ++p) {
        double sx = 0.0, sq = 0.0;
        uint64_t n = 0;
        for (int s = 0; s < mh.K; ++s)
          for (int d = 0; d < h->D; ++d)
            if (map[(size_t)s * h->D + d] == p) {
              const double* base = d == 0 ? res.of.data() : res.of.data() + 2 * KP + 1 + (size_t)(d - 1) * 2 * KP;
              sx += base[s];
              sq += base[KP + s];
              n += res.o64[s];
            }
        if (out->stat_sum) out->stat_sum[p] = sx;
        if (out->stat_sumsq) out->stat_sumsq[p] = sq;
        if (out->stat_n) out->stat_n[p] = n;
      }
    }
    h->states_valid = true;
    h->segs_valid = h->g_valid = false;
    h->rows_valid = (flags & HML_SWEEP_KEEP_ROWS) != 0 && !mixture;
    h->last_K = mh.K;
    return HML_OK;
  }
  return fail(h, HML_ERR_CAPACITY, "block capacity did not converge");
}

// Peer mailboxes for the per-sweep carries.  Every rank allocates its mailbox, the CUDA IPC handles travel by
// one NCCL all-gather, every rank maps the others'.  Any failure (no peer access, IPC unavailable, or
// HML_EXCHANGE=nccl in the environment) leaves the NCCL all-gathers in place; the decision is collective, so
// all ranks use the same transport.
int setup_p2p(hml_t* h) {
  const int world = h->world;
  const char* env = getenv("HML_EXCHANGE");
  int ok = !(env && strcmp(env, "nccl") == 0) && world <= kP2PMaxWorld ? 1 : 0;
  const size_t bytes = p2p_mailbox_bytes(world);
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaMalloc((void**)&h->mbox, bytes) != cudaSuccess) ok = 0;
  if (ok && cudaMemsetAsync(h->mbox, 0, bytes, h->stream) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, h->mbox) != cudaSuccess) ok = 0;
  if (ok && cudaHostAlloc((void**)&h->p2p_timeout_host, sizeof(unsigned int), cudaHostAllocMapped) != cudaSuccess) ok = 0;
  if (ok) {
    *h->p2p_timeout_host = 0;
    if (cudaHostGetDevicePointer((void**)&h->p2p_timeout_dev, h->p2p_timeout_host, 0) != cudaSuccess) ok = 0;
  }
  cudaGetLastError();
  // round 1: handles + readiness of every rank
  struct Msg {
    cudaIpcMemHandle_t handle;
    int ok;
    int pad[3];
  };
  static_assert(sizeof(Msg) % 8 == 0, "Msg");
  Msg msg;
  memset(&msg, 0, sizeof(msg));
  msg.handle = mine;
  msg.ok = ok;
  DevTmp<Msg> dsend, drecv;
  CK(dsend.alloc(1));
  CK(drecv.alloc(world));
  std::vector<Msg> all(world);
  CK(cudaMemcpyAsync(dsend, &msg, sizeof(Msg), cudaMemcpyHostToDevice, h->stream));
  int rc = all_gather(h, dsend, drecv, sizeof(Msg));
  if (rc != HML_OK) return rc;
  CK(cudaMemcpyAsync(all.data(), drecv, world * sizeof(Msg), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int r = 0; r < world; ++r) ok = ok && all[r].ok;
  memset(&h->peers, 0, sizeof(h->peers));
  if (ok) {
    for (int r = 0; r < world && ok; ++r) {
      if (r == h->rank) {
        h->peers.box[r] = h->mbox;
        continue;
      }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        ok = 0;
        cudaGetLastError();
      } else {
        h->peers.box[r] = (unsigned char*)p;
      }
    }
  }
  // round 2: did every rank map every mailbox?
  msg.ok = ok;
  CK(cudaMemcpyAsync(dsend, &msg, sizeof(Msg), cudaMemcpyHostToDevice, h->stream));
  rc = all_gather(h, dsend, drecv, sizeof(Msg));
  if (rc != HML_OK) return rc;
  CK(cudaMemcpyAsync(all.data(), drecv, world * sizeof(Msg), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int r = 0; r < world; ++r) ok = ok && all[r].ok;
  if (ok) {
    P2PDev dev;
    memset(&dev, 0, sizeof(dev));
    dev.peers = h->peers;
    dev.timeout_flag = h->p2p_timeout_dev;
    dev.rank = h->rank;
    dev.world = world;
    CK(dev_alloc(h->p2p_dev, 1));
    CK(cudaMemcpyAsync(h->p2p_dev, &dev, sizeof(dev), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  h->p2p = ok != 0;
  return HML_OK;
}

void teardown_p2p(hml_t* h) {
  for (int r = 0; r < kP2PMaxWorld; ++r)
    if (h->peers.box[r] && r != h->rank) cudaIpcCloseMemHandle(h->peers.box[r]);
  memset(&h->peers, 0, sizeof(h->peers));
  if (h->mbox) cudaFree(h->mbox);
  h->mbox = nullptr;
  dev_free(h->p2p_dev);
  if (h->p2p_timeout_host) cudaFreeHost(h->p2p_timeout_host);
  h->p2p_timeout_host = nullptr;
  h->p2p = false;
}

}  // namespace

// ================================================================================================ C ABI

extern "C" {

const char* hml_version(void) { return "hammlet_b200 0.1 (sm_100a)"; }

const char* hml_last_error(const hml_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int hml_create(hml_t** out, int device) {
  if (!out) return HML_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) +
                     "); hammlet_b200 has no CPU fallback";
    return HML_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    g_create_error = "device index out of range";
    return HML_ERR_ARG;
  }
  hml_t* h = new hml_ctx();
  h->device = device;
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    delete h;
    return HML_ERR_CUDA;
  }
  if (prop.major < 10) {
    g_create_error = std::string("device ") + prop.name + " is not sm_100-class; this library is built for sm_100a only";
    cudaStreamDestroy(h->stream);
    delete h;
    return HML_ERR_CUDA;
  }
  h->sms = prop.multiProcessorCount;
  if (cudaMalloc((void**)&h->outblk, kOutWords * 8) != cudaSuccess ||
      cudaMallocHost((void**)&h->outblk_host, kOutWords * 8) != cudaSuccess) {
    g_create_error = "allocation of the result block failed";
    dev_free(h->outblk);
    if (h->outblk_host) cudaFreeHost(h->outblk_host);
    cudaStreamDestroy(h->stream);
    delete h;
    return HML_ERR_CUDA;
  }
  cudaMemset(h->outblk, 0, kOutWords * 8);
  if (cudaMalloc((void**)&h->tickets, kTickets * sizeof(unsigned)) != cudaSuccess) {
    g_create_error = "allocation of the arrival counters failed";
    hml_destroy(h);
    return HML_ERR_CUDA;
  }
  cudaMemset(h->tickets, 0, kTickets * sizeof(unsigned));
  *out = h;
  return HML_OK;
}

int hml_destroy(hml_t* h) {
  if (!h) return HML_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->host_sweeps && getenv("HML_HOST_TIMING")) {
    const double n = (double)h->host_sweeps;
    fprintf(stderr, "[hml host timing] rank %d: %llu sweeps; per sweep: detection launches %.1f us, sweep launches %.1f us, "
                    "waiting for the result %.1f us\n", h->rank, (unsigned long long)h->host_sweeps,
            h->host_ns_prepare / n * 1e-3, h->host_ns_launch / n * 1e-3, h->host_ns_wait / n * 1e-3);
  }
  dev_free(h->w);
  dev_free(h->smax);
  dev_free(h->coeffs);
  dev_free(h->pq);
  dev_free(h->cell_pref);
  if (h->detect_scratch) cudaFree(h->detect_scratch);
  dev_free(h->starts);
  dev_free(h->bN);
  dev_free(h->bS);
  dev_free(h->e);
  dev_free(h->maxE);
  dev_free(h->alpha);
  dev_free(h->maps);
  dev_free(h->states);
  dev_free(h->chunk_maps);
  dev_free(h->chunk_submaps);
  dev_free(h->tile_maps);
  dev_free(h->tile_qin);
  dev_free(h->tickets);
  dev_free(h->chunk_ops);
  dev_free(h->tile_ops);
  dev_free(h->tile_ain);
  dev_free(h->group_ops);
  dev_free(h->group_ain);
  dev_free(h->cand_pos);
  dev_free(h->cand_w);
  dev_free(h->cand_pq);
  dev_free(h->spq);
  dev_free(h->cand_scratch);
  dev_free(h->seg_counts);
  dev_free(h->seg_starts);
  dev_free(h->seg_states);
  for (int k = 0; k < 2; ++k) {
    dev_free(h->mg_pos[k]);
    dev_free(h->mg_cnt[k]);
  }
  dev_free(h->mg_n_dev);
  if (h->mg_n_host) cudaFreeHost(h->mg_n_host);
  h->mg_n_host = nullptr;
  dev_free(h->mg_run_of_old);
  dev_free(h->mg_olds_below);
  dev_free(h->mg_flags);
  dev_free(h->wide_ops);
  dev_free(h->wide_exp);
  dev_free(h->chunk_exp);
  dev_free(h->tile_exp);
  dev_free(h->group_exp);
  dev_free(h->rows);
  dev_free(h->replay_u);
  dev_free(h->partials);
  dev_free(h->outblk);
  dev_free(h->seg_dev);
  dev_free(h->stats_gather);
  dev_free(h->run_states);
  dev_free(h->run_border);
  dev_free(h->chain_dev);
  dev_free(h->fused_cta_count);
  dev_free(h->fused_barriers);
  dev_free(h->fused_qtot);
  dev_free(h->fused_qmap);
  dev_free(h->fused_phase_ns);
  dev_free(h->fused_subops);
  dev_free(h->fused_submaps);
  dev_free(h->chain_stats);
  if (h->chain_host) cudaFreeHost(h->chain_host);
  h->chain_host = nullptr;
  teardown_p2p(h);
  if (h->stats_gather_host) cudaFreeHost(h->stats_gather_host);
  if (h->comm) g_nccl.CommDestroy(h->comm);
  if (h->outblk_host) cudaFreeHost(h->outblk_host);
  for (cudaEvent_t ev : h->event_pool) cudaEventDestroy(ev);
  cudaStreamDestroy(h->stream);
  delete h;
  return HML_OK;
}

int hml_load_f32_device(hml_t* h, const float* x_dev, uint64_t T, float weight_multiplier) {
  if (!h) return HML_ERR_ARG;
  if (!x_dev) return fail(h, HML_ERR_ARG, "x must not be NULL");
  if (T == 0) return fail(h, HML_ERR_ARG, "Input vector for breakpoint weights is empty!");
  if (T >= (1ull << 32)) return fail(h, HML_ERR_ARG, "one handle holds fewer than 2^32 observations; shard the sequence");
  CK(cudaSetDevice(h->device));
  return load_common(h, x_dev, T, weight_multiplier);
}

int hml_load_f32(hml_t* h, const float* x_host, uint64_t T, float weight_multiplier) {
  if (!h) return HML_ERR_ARG;
  if (!x_host) return fail(h, HML_ERR_ARG, "x must not be NULL");
  if (T == 0) return fail(h, HML_ERR_ARG, "Input vector for breakpoint weights is empty!");
  if (T >= (1ull << 32)) return fail(h, HML_ERR_ARG, "one handle holds fewer than 2^32 observations; shard the sequence");
  CK(cudaSetDevice(h->device));
  float* xd = nullptr;
  CK(dev_alloc(xd, T));
  cudaError_t e = cudaMemcpyAsync(xd, x_host, T * sizeof(float), cudaMemcpyHostToDevice, h->stream);
  if (e != cudaSuccess) {
    dev_free(xd);
    return fail(h, HML_ERR_CUDA, cudaGetErrorString(e));
  }
  const int rc = load_common(h, xd, T, weight_multiplier);
  cudaStreamSynchronize(h->stream);
  dev_free(xd);
  return rc;
}

int hml_load_f32_device_md(hml_t* h, const float* x_dev, uint64_t T, uint32_t nr_dims, float weight_multiplier) {
  if (!h) return HML_ERR_ARG;
  if (!x_dev) return fail(h, HML_ERR_ARG, "x must not be NULL");
  if (nr_dims == 0) return fail(h, HML_ERR_ARG, "Number of dimensions must be positive!");  // wavelet.hpp:106-108
  if (nr_dims > HML_MAX_DIMS)
    return fail(h, HML_ERR_ARG, "at most " + std::to_string(HML_MAX_DIMS) + " data dimensions are supported");
  if (T == 0) return fail(h, HML_ERR_ARG, "Input vector for breakpoint weights is empty!");
  if (T >= (1ull << 32)) return fail(h, HML_ERR_ARG, "one handle holds fewer than 2^32 observations; shard the sequence");
  CK(cudaSetDevice(h->device));
  return load_common(h, x_dev, T, weight_multiplier, (int)nr_dims);
}

int hml_load_f32_md(hml_t* h, const float* x_host, uint64_t T, uint32_t nr_dims, float weight_multiplier) {
  if (!h) return HML_ERR_ARG;
  if (!x_host) return fail(h, HML_ERR_ARG, "x must not be NULL");
  if (nr_dims == 0) return fail(h, HML_ERR_ARG, "Number of dimensions must be positive!");
  if (nr_dims > HML_MAX_DIMS)
    return fail(h, HML_ERR_ARG, "at most " + std::to_string(HML_MAX_DIMS) + " data dimensions are supported");
  if (T == 0) return fail(h, HML_ERR_ARG, "Input vector for breakpoint weights is empty!");
  if (T >= (1ull << 32)) return fail(h, HML_ERR_ARG, "one handle holds fewer than 2^32 observations; shard the sequence");
  CK(cudaSetDevice(h->device));
  float* xd = nullptr;
  CK(dev_alloc(xd, T * nr_dims));
  cudaError_t e = cudaMemcpyAsync(xd, x_host, T * nr_dims * sizeof(float), cudaMemcpyHostToDevice, h->stream);
  if (e != cudaSuccess) {
    dev_free(xd);
    return fail(h, HML_ERR_CUDA, cudaGetErrorString(e));
  }
  const int rc = load_common(h, xd, T, weight_multiplier, (int)nr_dims);
  cudaStreamSynchronize(h->stream);
  dev_free(xd);
  return rc;
}

int hml_nr_dims(const hml_t* h, uint32_t* nr_dims) {
  if (!h || !nr_dims) return HML_ERR_ARG;
  *nr_dims = (uint32_t)h->D;
  return HML_OK;
}

int hml_size(const hml_t* h, uint64_t* T) {
  if (!h || !T) return HML_ERR_ARG;
  *T = h->world > 1 ? h->T_global : h->T;
  return HML_OK;
}

int hml_sigma_hat(hml_t* h, double* sigma_hat) {
  if (!h || !sigma_hat) return HML_ERR_ARG;
  if (h->T == 0) return fail(h, HML_ERR_STATE, "no data loaded");
  *sigma_hat = h->sigma_hat;
  return HML_OK;
}

int hml_get_weights(hml_t* h, float* dst, uint64_t n) {
  if (!h || !dst) return HML_ERR_ARG;
  if (h->T == 0) return fail(h, HML_ERR_STATE, "no data loaded");
  if (n < h->T) return fail(h, HML_ERR_CAPACITY, "buffer smaller than T");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(dst, h->w, h->T * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return HML_OK;
}

int hml_get_coeffs(hml_t* h, float* dst, uint64_t n) {
  if (!h || !dst) return HML_ERR_ARG;
  if (h->T == 0) return fail(h, HML_ERR_STATE, "no data loaded");
  if (!h->coeffs) return fail(h, HML_ERR_STATE, "coefficients are only kept for T <= 2^26");
  if (n < h->T) return fail(h, HML_ERR_CAPACITY, "buffer smaller than T");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(dst, h->coeffs, h->T * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return HML_OK;
}

int hml_create_blocks(hml_t* h, float threshold, uint64_t* nblocks) {
  if (!h) return HML_ERR_ARG;
  if (h->T == 0) return fail(h, HML_ERR_STATE, "no data loaded");
  CK(cudaSetDevice(h->device));
  const int world = h->world > 1 ? h->world : 1;
  for (int attempt = 0; attempt < 8; ++attempt) {
    h->stages.clear();
    h->stage_used = 0;
    int rc = run_detect(h, threshold);
    if (rc != HML_OK) return rc;
    stage_cb(h, "block_stats");
    SweepBuffers b = make_buffers(h, 2);
    launch_block_stats(b, 0, h->capacity, h->sms, h->stream);
    h->launches++;
    stage_cb(h, "end");
    CK(cudaGetLastError());
    // block counts of all ranks (a rank whose arrays were too small makes every rank repeat the call)
    const unsigned long long* host = h->outblk_host;
    if (world > 1) {
      rc = exchange(h, kSlotStats, h->outblk, h->stats_gather, 16);
      if (rc != HML_OK) return rc;
      CK(cudaMemcpyAsync(h->stats_gather_host, h->stats_gather, world * 16, cudaMemcpyDeviceToHost, h->stream));
      host = h->stats_gather_host;
    } else {
      CK(cudaMemcpyAsync(h->outblk_host, h->outblk, 8, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    if ((rc = check_exchange(h)) != HML_OK) return rc;
    const uint64_t B = host[2 * (world > 1 ? h->rank : 0)];
    bool any_over = false;
    uint64_t total = 0, first = 0;
    for (int r = 0; r < world; ++r) {
      any_over |= world > 1 ? host[2 * r + 1] != 0 : host[0] > h->capacity;
      total += host[2 * r];
      if (r < h->rank) first += host[2 * r];
    }
    if (any_over) {
      if (B > h->capacity) {
        rc = alloc_blocks(h, B + B / 4, h->KP);
        if (rc != HML_OK) return rc;
      }
      continue;
    }
    collect_timing(h);
    h->nblocks = B;
    h->global_blocks = total;
    h->first_block = first;
    h->blocks_valid = h->stats_valid = true;
    h->states_valid = h->rows_valid = false;
    if (nblocks) *nblocks = world > 1 ? total : B;
    return HML_OK;
  }
  return fail(h, HML_ERR_CAPACITY, "block capacity did not converge");
}

int hml_nr_blocks(const hml_t* h, uint64_t* nblocks) {
  if (!h || !nblocks) return HML_ERR_ARG;
  if (!h->blocks_valid) return HML_ERR_STATE;
  *nblocks = h->nblocks;
  return HML_OK;
}

int hml_get_block_sums(hml_t* h, uint32_t dim, double* sum, double* sumsq, uint64_t capacity) {
  if (!h) return HML_ERR_ARG;
  if (!h->blocks_valid) return fail(h, HML_ERR_STATE, "no block structure");
  if (dim >= (uint32_t)h->D) return fail(h, HML_ERR_ARG, "dimension out of range");
  if (!sum || !sumsq) return fail(h, HML_ERR_ARG, "sum and sumsq must be given together");
  if (capacity < h->nblocks) return fail(h, HML_ERR_CAPACITY, "buffer smaller than the number of blocks");
  CK(cudaSetDevice(h->device));
  const uint64_t B = h->nblocks;
  DevTmp<double> tmp;
  CK(tmp.alloc(2 * B));
  SweepBuffers b = make_buffers(h, 2);
  b.bS = h->bS + (size_t)dim * h->capacity;
  launch_unpermute(b, 0, B, nullptr, tmp, tmp + B, h->stream);
  h->launches++;
  CK(cudaMemcpyAsync(sum, tmp, B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(sumsq, tmp + B, B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return HML_OK;
}

int hml_get_blocks(hml_t* h, uint32_t* starts, double* sum, double* sumsq, uint64_t capacity) {
  if (!h) return HML_ERR_ARG;
  if (!h->blocks_valid) return fail(h, HML_ERR_STATE, "no block structure");
  if (capacity < h->nblocks) return fail(h, HML_ERR_CAPACITY, "buffer smaller than the number of blocks");
  CK(cudaSetDevice(h->device));
  const uint64_t B = h->nblocks;
  if (starts) {
    CK(cudaMemcpyAsync(starts, h->starts, B * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (h->seg_start) {  // segment mode: report positions in the whole sequence
      CK(cudaStreamSynchronize(h->stream));
      for (uint64_t i = 0; i < B; ++i) starts[i] += (uint32_t)h->seg_start;
    }
  }
  if (sum || sumsq) {
    if (!sum || !sumsq) return fail(h, HML_ERR_ARG, "sum and sumsq must be given together");
    DevTmp<double> tmp;
    CK(tmp.alloc(2 * B));
    SweepBuffers b = make_buffers(h, 2);
    launch_unpermute(b, 0, B, nullptr, tmp, tmp + B, h->stream);
    h->launches++;
    CK(cudaMemcpyAsync(sum, tmp, B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(sumsq, tmp + B, B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return HML_OK;
}

int hml_fb_sweep(hml_t* h, const hml_model* m, uint32_t flags, float threshold, uint64_t seed, uint64_t sweep_index,
                 const double* replay_uniforms, uint64_t n_replay, hml_sweep_out* out) {
  return sweep_common(h, m, flags, threshold, seed, sweep_index, replay_uniforms, n_replay, out, false);
}

int hml_mix_sweep(hml_t* h, const hml_model* m, uint32_t flags, float threshold, uint64_t seed, uint64_t sweep_index,
                  const double* replay_uniforms, uint64_t n_replay, hml_sweep_out* out) {
  return sweep_common(h, m, flags & ~(uint32_t)(HML_SWEEP_LOGLIK | HML_SWEEP_KEEP_ROWS), threshold, seed, sweep_index,
                      replay_uniforms, n_replay, out, true);
}

// ---- device-resident Gibbs chain

int hml_chain_init(hml_t* h, int K, const float* nig_prior, float trans, float self_trans, float alpha_pi, uint64_t seed,
                   int use_self_transitions) {
  if (!h || !nig_prior) return HML_ERR_ARG;
  if (h->T == 0) return fail(h, HML_ERR_STATE, "no data loaded");
  if (K < 2 || K > kChainMaxStates) return fail(h, HML_ERR_ARG, "a device-resident chain has 2 to 8 states");
  if (h->D != 1) return fail(h, HML_ERR_ARG, "a device-resident chain needs univariate data");
  CK(cudaSetDevice(h->device));
  int rc = chain_alloc(h);
  if (rc != HML_OK) return rc;
  ChainDev* c = h->chain_host;
  memset(c, 0, sizeof(ChainDev));
  c->K = K;
  c->use_self = use_self_transitions ? 1 : 0;
  c->T = (uint32_t)(h->world > 1 ? h->T_global : h->T);
  c->seed = seed;
  c->sweep = 0;
  for (int s = 0; s < K; ++s)
    for (int k = 0; k < 4; ++k) c->prior_theta[s][k] = nig_prior[k];
  c->prior_trans = trans;
  c->prior_self = self_trans;
  c->prior_pi = alpha_pi;
  // a leading "P" (main.cpp:393-406): theta, pi and A from their priors — the parameter phase on an empty result block
  CK(cudaMemsetAsync(h->chain_stats, 0, kOutWords * 8, h->stream));
  CK(cudaMemcpyAsync(h->chain_dev, c, sizeof(ChainDev), cudaMemcpyHostToDevice, h->stream));
  const int KP = padded_states(K);
  size_t words = (size_t)KP + (size_t)KP * KP + 1;
  words += words & 1;
  const int e = launch_chain_params(KP, h->chain_dev, h->chain_stats, (const double*)(h->chain_stats + words), h->stream);
  if (e != 0) return fail(h, HML_ERR_CUDA, "parameter kernel launch failed");
  h->launches++;
  CK(cudaMemcpyAsync(c, h->chain_dev, sizeof(ChainDev), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (c->phase_abort[6]) return fail(h, HML_ERR_NUMERIC, "Variance must be positive!");
  h->chain_ready = true;
  h->chain_fused_sweeps = h->chain_standard_sweeps = 0;
  return HML_OK;
}

int hml_chain_set(hml_t* h, const double* mean, const double* var, const double* A, const double* pi) {
  if (!h) return HML_ERR_ARG;
  if (!h->chain_ready) return fail(h, HML_ERR_STATE, "hml_chain_init has not been called");
  CK(cudaSetDevice(h->device));
  ChainDev* c = h->chain_host;
  const int K = c->K;
  for (int s = 0; s < K; ++s) {
    if (mean) c->mean[s] = mean[s];
    if (var) {
      if (!(var[s] > 0) || !isfinite(var[s])) return fail(h, HML_ERR_ARG, "Variance must be positive!");
      c->var[s] = var[s];
    }
    if (pi) c->pi[s] = pi[s];
    if (A)
      for (int j = 0; j < K; ++j) c->A[s * K + j] = A[s * K + j];
  }
  float mv = (float)c->var[0];
  for (int s = 1; s < K; ++s) mv = fminf(mv, (float)c->var[s]);
  c->thr = sqrtf(2.0f * logf((float)c->T) * mv);
  chain_derive(c);
  CK(cudaMemcpyAsync(h->chain_dev, c, sizeof(ChainDev), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return HML_OK;
}

int hml_chain_get(hml_t* h, double* mean, double* var, double* A, double* pi, float* threshold, uint64_t* sweeps) {
  if (!h) return HML_ERR_ARG;
  if (!h->chain_ready) return fail(h, HML_ERR_STATE, "hml_chain_init has not been called");
  const ChainDev* c = h->chain_host;  // current as of the last hml_chain_* call (they end with a synchronisation)
  const int K = c->K;
  for (int s = 0; s < K; ++s) {
    if (mean) mean[s] = c->mean[s];
    if (var) var[s] = c->var[s];
    if (pi) pi[s] = c->pi[s];
    if (A)
      for (int j = 0; j < K; ++j) A[s * K + j] = c->A[s * K + j];
  }
  if (threshold) *threshold = c->thr;
  if (sweeps) *sweeps = c->sweep;
  return HML_OK;
}

int hml_chain_run(hml_t* h, uint64_t nsweeps, uint64_t* fused_sweeps, hml_sweep_out* last) {
  if (!h) return HML_ERR_ARG;
  if (!h->chain_ready) return fail(h, HML_ERR_STATE, "hml_chain_init has not been called");
  if (h->T == 0) return fail(h, HML_ERR_STATE, "no data loaded");
  CK(cudaSetDevice(h->device));
  ChainDev* c = h->chain_host;
  const int K = c->K, KP = padded_states(K);
  if (KP != h->KP) {
    int rc = alloc_blocks(h, h->capacity, KP);
    if (rc != HML_OK) return rc;
  }
  size_t words = (size_t)KP + (size_t)KP * KP + 1;
  words += words & 1;
  uint64_t done = 0, fused_done = 0;
  int stalled = 0;
  bool last_fused = false;
  std::vector<double> ssum(K), ssq(K);
  std::vector<uint64_t> sn(K), trans((size_t)K * K), counts(K);
  uint64_t last_nblocks = 0;
  while (done < nsweeps) {
    const float thr = c->thr;
    bool use_fused = fused_possible(h, KP);
    if (use_fused) {
      bool usable = false;
      int rc = ensure_candidates(h, thr, &usable);
      if (rc != HML_OK) return rc;
      use_fused = usable && h->cand_pq != nullptr && h->cand_n <= (uint64_t)2 * kFusedMaxTiles * kTileBlocks;
    }
    if (use_fused && stalled < 2) {
      const uint64_t want = nsweeps - done;
      const int batch = (int)(want > (1u << 20) ? (1u << 20) : want);
      h->states_valid = h->segs_valid = h->g_valid = h->rows_valid = false;
      h->blocks_valid = false;
      h->last_thr = thr;
      int rc = fused_launch(h, KP, batch, true, true, c->seed, 0, nullptr, false);
      if (rc != HML_OK) return rc;
      CK(cudaStreamSynchronize(h->stream));
      h->mg_pending = false;
      done += c->sweeps_done;
      fused_done += c->sweeps_done;
      h->chain_fused_sweeps += c->sweeps_done;
      stalled = c->sweeps_done ? 0 : stalled + 1;
      if (c->sweeps_done) last_fused = true;
      if (c->phase_abort[6] == kChainNumeric) return fail(h, HML_ERR_NUMERIC, "Variance must be positive!");
      if (c->abort_code == kChainCapacity && c->nblocks_seen > h->capacity &&
          c->nblocks_seen <= (uint64_t)kFusedMaxTiles * kTileBlocks) {
        rc = alloc_blocks(h, c->nblocks_seen + c->nblocks_seen / 4, KP);
        if (rc != HML_OK) return rc;
        stalled = 0;
        continue;
      }
      if (c->abort_code == kChainOk || c->abort_code == kChainThreshold) continue;  // done, or the list is rebuilt above
      stalled = 2;  // this sweep needs the multi-kernel path (too many blocks, a vanished forward sum)
      continue;
    }
    // ---- one sweep through the multi-kernel path on the chain's current parameters, then the parameter phase
    hml_model m;
    memset(&m, 0, sizeof(m));
    m.K = K;
    m.use_self_transitions = c->use_self;
    m.mean = c->mean;
    m.var = c->var;
    m.A = c->A;
    m.pi = c->pi;
    hml_sweep_out o;
    memset(&o, 0, sizeof(o));
    o.stat_sum = ssum.data();
    o.stat_sumsq = ssq.data();
    o.stat_n = sn.data();
    o.trans = trans.data();
    o.counts = counts.data();
    int rc = sweep_common(h, &m, HML_SWEEP_DYNAMIC, thr, c->seed, c->sweep, nullptr, 0, &o, false);
    if (rc != HML_OK) return rc;
    // statistics of the whole sequence in the layout of the result block (segment mode: the host summed the ranks')
    std::vector<unsigned long long> blk(kOutWords, 0ull);
    for (int s = 0; s < K; ++s) {
      blk[s] = sn[s];
      for (int j = 0; j < K; ++j) blk[KP + s * KP + j] = trans[(size_t)s * K + j];
    }
    double* of = reinterpret_cast<double*>(blk.data() + words);
    for (int s = 0; s < K; ++s) {
      of[s] = ssum[s];
      of[KP + s] = ssq[s];
    }
    CK(cudaMemcpyAsync(h->chain_stats, blk.data(), (words + 2 * KP + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    for (unsigned int& f : c->phase_abort) f = 0;
    CK(cudaMemcpyAsync(h->chain_dev->phase_abort, c->phase_abort, sizeof(c->phase_abort), cudaMemcpyHostToDevice, h->stream));
    const int e = launch_chain_params(KP, h->chain_dev, h->chain_stats, (const double*)(h->chain_stats + words), h->stream);
    if (e != 0) return fail(h, HML_ERR_CUDA, "parameter kernel launch failed");
    h->launches++;
    CK(cudaMemcpyAsync(c, h->chain_dev, sizeof(ChainDev), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (c->phase_abort[6] == kChainNumeric) return fail(h, HML_ERR_NUMERIC, "Variance must be positive!");
    done++;
    h->chain_standard_sweeps++;
    stalled = 0;
    last_fused = false;
    last_nblocks = o.nblocks;
  }
  if (fused_sweeps) *fused_sweeps = fused_done;
  if (!last_fused && done && last) {
    last->nblocks = last_nblocks;
    last->uniform_fallbacks = 0;
    last->loglik = NAN;
    for (int s = 0; s < K; ++s) {
      if (last->counts) last->counts[s] = counts[s];
      if (last->stat_n) last->stat_n[s] = sn[s];
      if (last->stat_sum) last->stat_sum[s] = ssum[s];
      if (last->stat_sumsq) last->stat_sumsq[s] = ssq[s];
      if (last->trans)
        for (int j = 0; j < K; ++j) last->trans[s * K + j] = trans[(size_t)s * K + j];
    }
  }
  if (last_fused) {
    // block structure and statistics of the last sweep, as after hml_fb_sweep
    SweepResult res;
    int rc = fetch_result(h, KP, res, false);
    if (rc != HML_OK) return rc;
    h->nblocks = res.local_blocks;
    h->global_blocks = res.global_blocks;
    h->first_block = 0;
    h->blocks_valid = h->stats_valid = h->states_valid = true;
    h->last_K = K;
    if (last) {
      last->nblocks = res.global_blocks;
      last->uniform_fallbacks = 0;
      last->loglik = NAN;
      for (int s = 0; s < K; ++s) {
        if (last->counts) last->counts[s] = res.o64[s];
        if (last->stat_n) last->stat_n[s] = res.o64[s];
        if (last->stat_sum) last->stat_sum[s] = res.of[s];
        if (last->stat_sumsq) last->stat_sumsq[s] = res.of[KP + s];
        if (last->trans)
          for (int j = 0; j < K; ++j) last->trans[s * K + j] = res.o64[KP + s * KP + j];
      }
    }
  }
  return HML_OK;
}

int hml_chain_phase_ns(hml_t* h, uint64_t stamps[16]) {
  if (!h || !stamps) return HML_ERR_ARG;
  if (!h->fused_phase_ns) return fail(h, HML_ERR_STATE, "no fused sweep has been run");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(stamps, h->fused_phase_ns, 16 * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return HML_OK;
}

int hml_get_states(hml_t* h, int16_t* states, uint64_t capacity) {
  if (!h || !states) return HML_ERR_ARG;
  if (!h->states_valid) return fail(h, HML_ERR_STATE, "no sweep has been run on the current block structure");
  if (capacity < h->nblocks) return fail(h, HML_ERR_CAPACITY, "buffer smaller than the number of blocks");
  CK(cudaSetDevice(h->device));
  DevTmp<int16_t> tmp;
  CK(tmp.alloc(h->nblocks));
  SweepBuffers b = make_buffers(h, 2);
  launch_unpermute(b, 0, h->nblocks, tmp, nullptr, nullptr, h->stream);
  h->launches++;
  CK(cudaMemcpyAsync(states, tmp, h->nblocks * sizeof(int16_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return HML_OK;
}

// Equal-state runs of the last sweep, formed on the device (Records.hpp:166-188: a segment ends where the state
// changes): h->nsegs runs; with `write` their (start, state) pairs are left in h->seg_starts / h->seg_states.
// need_count: the host wants to know the number of runs (one small copy + synchronisation); the marginal merge does
// not — its kernels read the count from device memory (seg_counts[ntiles]).
static int ensure_runs(hml_t* h, bool write, bool need_count = true, bool accumulate = false) {
  const uint64_t B = h->nblocks;
  const bool seg = h->world > 1;
  uint64_t ntiles = (B + 1023) / 1024;
  if (seg && ntiles == 0) ntiles = 1;  // a rank without blocks still has its virtual run
  if (h->seg_cap < h->capacity || !h->seg_counts) {
    CK(dev_alloc(h->seg_counts, h->capacity / 1024 + 2));
    CK(dev_alloc(h->seg_starts, h->capacity + 1));
    CK(dev_alloc(h->seg_states, h->capacity + 1));
    h->seg_cap = h->capacity;
    h->segs_valid = h->g_valid = false;
  }
  SweepBuffers b = make_buffers(h, h->KP ? h->KP : 2);
  if (!h->segs_valid) {
    if (seg) {
      // the state of every rank's last block: a run that crosses a rank border is ONE segment (Records.hpp:166-188)
      launch_segments_last_state(b, h->run_states, h->stream);
      h->launches++;
      int rc = exchange(h, kSlotMaps, h->run_states, h->run_states + 8, 8);
      if (rc != HML_OK) return rc;
    }
    launch_segments_count(b, B, h->run_states ? h->run_states + 8 : nullptr, h->seg_counts, h->sms, h->stream);
    h->launches += 2;
    CK(cudaGetLastError());
    h->segs_valid = true;
    h->nsegs_known = false;
    h->runs_written = false;
  }
  if (need_count && !h->nsegs_known) {
    uint32_t n32 = 0;
    CK(cudaMemcpyAsync(&n32, h->seg_counts + ntiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->mg_pending = false;
    if (int rc = check_exchange(h)) return rc;
    h->nsegs = n32;
    h->nsegs_known = true;
  }
  if (write && (!h->runs_written || accumulate)) {
    launch_segments_write(b, B, h->run_states ? h->run_states + 8 : nullptr, h->run_border, accumulate ? 1 : 0, h->seg_counts,
                          h->seg_starts, h->seg_states, h->sms, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    h->runs_written = true;
  }
  return HML_OK;
}

// Collective over the handle's communicator: every rank contributes `bytes` bytes of host memory and receives all
// contributions in rank order (world x bytes).  NCCL on temporary device buffers; the stream is drained on return.
static int comm_allgather_host(hml_t* h, const void* send, size_t bytes, void* recv) {
  if (h->world <= 1) {
    memcpy(recv, send, bytes);
    return HML_OK;
  }
  DevTmp<unsigned char> ds, dr;
  const size_t padded = (bytes + 15) / 16 * 16;
  CK(ds.alloc(padded));
  CK(dr.alloc(padded * h->world));
  CK(cudaMemcpyAsync(ds, send, bytes, cudaMemcpyHostToDevice, h->stream));
  int rc = all_gather(h, ds, dr, padded);
  if (rc != HML_OK) return rc;
  std::vector<unsigned char> tmp(padded * h->world);
  CK(cudaMemcpyAsync(tmp.data(), dr, tmp.size(), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->mg_pending = false;
  for (int r = 0; r < h->world; ++r) memcpy((unsigned char*)recv + (size_t)r * bytes, tmp.data() + (size_t)r * padded, bytes);
  return HML_OK;
}

// the same for contributions of different lengths: counts[r] elements of `elem` bytes from rank r, concatenated
static int comm_allgatherv_host(hml_t* h, const void* send, uint64_t n, size_t elem, std::vector<unsigned char>& out,
                                std::vector<uint64_t>& counts) {
  counts.assign(h->world > 1 ? h->world : 1, 0);
  int rc = comm_allgather_host(h, &n, sizeof(uint64_t), counts.data());
  if (rc != HML_OK) return rc;
  uint64_t nmax = 0, total = 0;
  for (uint64_t c : counts) {
    nmax = c > nmax ? c : nmax;
    total += c;
  }
  std::vector<unsigned char> mine((size_t)nmax * elem + 16, 0), all(((size_t)nmax * elem + 16) * counts.size());
  if (n) memcpy(mine.data(), send, (size_t)n * elem);
  rc = comm_allgather_host(h, mine.data(), mine.size(), all.data());
  if (rc != HML_OK) return rc;
  out.resize((size_t)total * elem);
  size_t o = 0;
  for (size_t r = 0; r < counts.size(); ++r) {
    memcpy(out.data() + o, all.data() + r * mine.size(), (size_t)counts[r] * elem);
    o += (size_t)counts[r] * elem;
  }
  return HML_OK;
}

int hml_get_segments(hml_t* h, uint64_t* nsegments, uint64_t* seg_size, int16_t* seg_state, uint64_t capacity) {
  if (!h || !nsegments) return HML_ERR_ARG;
  if (!h->states_valid) return fail(h, HML_ERR_STATE, "no sweep has been run on the current block structure");
  CK(cudaSetDevice(h->device));
  const bool seg = h->world > 1;
  if (h->nblocks == 0 && !seg) {
    *nsegments = 0;
    return HML_OK;
  }
  const bool want = seg_size && seg_state;
  if (seg) {
    // Collective: the runs of the whole sequence, identical on every rank.  A rank's run at local position 0 is a run
    // of the sequence only if its first block begins there in a state other than the previous rank's last one.
    if (!h->g_valid) {
      int rc = ensure_runs(h, true);
      if (rc != HML_OK) return rc;
      const uint64_t n = h->nsegs;
      std::vector<uint32_t> st(n);
      std::vector<int16_t> ss(n);
      uint32_t border[2] = {0, 0};
      CK(cudaMemcpyAsync(st.data(), h->seg_starts, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(ss.data(), h->seg_states, n * sizeof(int16_t), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(border, h->run_border, sizeof(border), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      struct Run {
        uint64_t start;  // position in the whole sequence
        int64_t state;
      };
      std::vector<Run> mine;
      mine.reserve(n);
      for (uint64_t i = 0; i < n; ++i) {
        if (i == 0 && h->rank > 0 && !border[1]) continue;  // continues the previous rank's run
        mine.push_back({h->seg_start + st[i], ss[i]});
      }
      std::vector<unsigned char> all;
      std::vector<uint64_t> counts;
      rc = comm_allgatherv_host(h, mine.data(), mine.size(), sizeof(Run), all, counts);
      if (rc != HML_OK) return rc;
      const Run* runs = reinterpret_cast<const Run*>(all.data());
      const uint64_t total = all.size() / sizeof(Run);
      h->g_seg_size.resize(total);
      h->g_seg_state.resize(total);
      for (uint64_t i = 0; i < total; ++i) {
        h->g_seg_size[i] = (i + 1 < total ? runs[i + 1].start : h->T_global) - runs[i].start;
        h->g_seg_state[i] = (int16_t)runs[i].state;
      }
      h->g_valid = true;
    }
    const uint64_t n = h->g_seg_size.size();
    *nsegments = n;
    if (!want) return HML_OK;
    if (n > capacity) return fail(h, HML_ERR_CAPACITY, "segment buffer too small");
    memcpy(seg_size, h->g_seg_size.data(), n * sizeof(uint64_t));
    memcpy(seg_state, h->g_seg_state.data(), n * sizeof(int16_t));
    return HML_OK;
  }
  int rc = ensure_runs(h, want);
  if (rc != HML_OK) return rc;
  const uint64_t n = h->nsegs;
  *nsegments = n;
  if (!want) return HML_OK;
  if (n > capacity) return fail(h, HML_ERR_CAPACITY, "segment buffer too small");
  std::vector<uint32_t> st(n);
  CK(cudaMemcpyAsync(st.data(), h->seg_starts, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(seg_state, h->seg_states, n * sizeof(int16_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const uint64_t end = h->T;
  for (uint64_t i = 0; i < n; ++i) seg_size[i] = (uint64_t)(i + 1 < n ? st[i + 1] : end) - st[i];
  return HML_OK;
}

int hml_comm_allgather(hml_t* h, const void* send_host, uint64_t bytes, void* recv_host) {
  if (!h || !send_host || !recv_host || bytes == 0) return HML_ERR_ARG;
  CK(cudaSetDevice(h->device));
  return comm_allgather_host(h, send_host, (size_t)bytes, recv_host);
}

// ---- state marginals accumulated on the device

static int mg_reserve(hml_t* h, uint64_t segments, uint64_t runs) {
  if (segments > h->mg_cap) {
    const uint64_t cap = segments + segments / 2 + 1024;
    for (int k = 0; k < 2; ++k) {
      uint32_t* np = nullptr;
      uint16_t* nc = nullptr;
      CK(cudaMalloc((void**)&np, cap * sizeof(uint32_t)));
      CK(cudaMalloc((void**)&nc, cap * (size_t)h->mg_K * sizeof(uint16_t)));
      if (k == h->mg_cur && h->mg_n) {  // keep what has been accumulated
        CK(cudaMemcpyAsync(np, h->mg_pos[k], h->mg_n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(nc, h->mg_cnt[k], h->mg_n * (size_t)h->mg_K * sizeof(uint16_t), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
      }
      dev_free(h->mg_pos[k]);
      dev_free(h->mg_cnt[k]);
      h->mg_pos[k] = np;
      h->mg_cnt[k] = nc;
    }
    h->mg_cap = cap;
    CK(dev_alloc(h->mg_run_of_old, cap));
  }
  if (runs > h->mg_runs_cap) {
    const uint64_t cap = runs + runs / 2 + 1024;
    CK(dev_alloc(h->mg_olds_below, cap));
    CK(dev_alloc(h->mg_flags, cap + 1));
    h->mg_runs_cap = cap;
  }
  return HML_OK;
}

int hml_marginals_reset(hml_t* h, int K) {
  if (!h) return HML_ERR_ARG;
  if (K < 1 || K > HML_MAX_STATES) return fail(h, HML_ERR_ARG, "number of states must be in [1, 32]");
  if (h->T == 0) return fail(h, HML_ERR_STATE, "no data loaded");
  CK(cudaSetDevice(h->device));
  for (int k = 0; k < 2; ++k) {
    dev_free(h->mg_pos[k]);
    dev_free(h->mg_cnt[k]);
  }
  h->mg_cap = 0;
  h->mg_K = K;
  h->mg_cur = 0;
  h->mg_n = 0;
  h->mg_iterations = 0;
  if (!h->mg_n_dev) CK(dev_alloc(h->mg_n_dev, 2));
  if (!h->mg_n_host) CK(cudaMallocHost((void**)&h->mg_n_host, sizeof(uint32_t)));
  // room for one segment per 512 positions up front (the refinement has about as many segments as a sweep has
  // equal-state runs; growing later means freeing and allocating while tens of GB are resident: ~100 ms)
  uint64_t guess = h->T / 512;
  if (guess > h->capacity) guess = h->capacity;
  if (guess < (1u << 16)) guess = 1u << 16;
  int rc = mg_reserve(h, guess, guess / 2);
  if (rc != HML_OK) return rc;
  // one segment covering the whole sequence, all counts zero (StateMarginals.hpp:24-33)
  CK(cudaMemsetAsync(h->mg_pos[0], 0, sizeof(uint32_t), h->stream));
  CK(cudaMemsetAsync(h->mg_cnt[0], 0, (size_t)K * sizeof(uint16_t), h->stream));
  const uint32_t one[2] = {1u, 1u};
  CK(cudaMemcpyAsync(h->mg_n_dev, one, sizeof(one), cudaMemcpyHostToDevice, h->stream));
  if (h->run_border) CK(cudaMemsetAsync(h->run_border, 0, 2 * sizeof(uint32_t), h->stream));
  h->g_mg_valid = false;
  CK(cudaStreamSynchronize(h->stream));
  *h->mg_n_host = 1;
  h->mg_n = 1;
  h->mg_pending = false;
  return HML_OK;
}

// segments as of now: the count copy queued by the last hml_marginals_add has landed once the stream has been
// synchronised (every sweep does that); otherwise wait for it
static int mg_current_segments(hml_t* h, uint64_t* n) {
  if (h->mg_pending) {
    CK(cudaStreamSynchronize(h->stream));
    h->mg_pending = false;
  }
  if (h->mg_n_host) h->mg_n = *h->mg_n_host;
  *n = h->mg_n;
  return HML_OK;
}

int hml_marginals_add(hml_t* h) {
  if (!h) return HML_ERR_ARG;
  if (h->mg_K == 0) return fail(h, HML_ERR_STATE, "hml_marginals_reset has not been called");
  if (!h->states_valid) return fail(h, HML_ERR_STATE, "no sweep has been run on the current block structure");
  if (h->last_K > h->mg_K) return fail(h, HML_ERR_ARG, "the last sweep had more states than the marginals hold");
  if (h->mg_iterations >= 32767) return fail(h, HML_ERR_CAPACITY, "marginal counts are 16-bit (marginal_t): 32767 iterations");
  const bool seg = h->world > 1;
  if (h->nblocks == 0 && !seg) return HML_OK;
  CK(cudaSetDevice(h->device));
  // Nothing here waits for the device: the runs are formed and merged by kernels that read their counts from device
  // memory; the host only needs upper bounds (segments so far + blocks of the sweep) to size buffers and grids.
  // Segment mode: collective (the ranks trade the state of their last block so that runs continue across borders);
  // each rank keeps the marginals of its own positions, hml_marginals_get merges them.
  int rc = ensure_runs(h, true, false, true);
  if (rc != HML_OK) return rc;
  uint64_t n = 0;
  rc = mg_current_segments(h, &n);
  if (rc != HML_OK) return rc;
  const uint64_t m_upper = h->nblocks + (seg ? 1 : 0);  // a run is at least one block (+ the virtual run of a rank > 0)
  if (n + m_upper >= (1ull << 32)) return fail(h, HML_ERR_CAPACITY, "too many marginal segments for 32-bit indices");
  rc = mg_reserve(h, n + m_upper, m_upper);
  if (rc != HML_OK) return rc;
  const int cur = h->mg_cur, nxt = cur ^ 1;
  uint64_t ntiles = (h->nblocks + 1023) / 1024;
  if (seg && ntiles == 0) ntiles = 1;
  launch_marginals_merge(h->mg_pos[cur], h->mg_n_dev + cur, n, h->mg_cnt[cur], h->seg_starts, h->seg_states,
                         h->seg_counts + ntiles, m_upper, h->mg_run_of_old, h->mg_olds_below, h->mg_flags, h->mg_K,
                         h->mg_pos[nxt], h->mg_cnt[nxt], h->mg_n_dev + nxt, h->sms, h->stream);
  h->launches += 3;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->mg_n_host, h->mg_n_dev + nxt, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  h->mg_pending = true;
  h->mg_cur = nxt;
  h->mg_iterations++;
  h->g_mg_valid = false;
  return HML_OK;
}

// Segment mode: the marginals of the whole sequence from the ranks' own ones (collective).  The lists are
// concatenated in rank order; the segment that begins at a rank's first observation joins the previous rank's last
// segment unless that position started a run of the sequence in some recorded iteration (run_border[0]) — in every
// iteration the same run then covered both pieces, so their counts agree (checked).
static int mg_merge_global(hml_t* h) {
  if (h->g_mg_valid) return HML_OK;
  uint64_t n = 0;
  int rc = mg_current_segments(h, &n);
  if (rc != HML_OK) return rc;
  const int K = h->mg_K;
  const size_t elem = 8 + 2 * (size_t)K;  // start (global, 8 bytes) + K counts
  std::vector<uint32_t> pos(n);
  std::vector<uint16_t> cnt(n * (size_t)K);
  uint32_t border[2] = {0, 0};
  CK(cudaMemcpyAsync(pos.data(), h->mg_pos[h->mg_cur], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(cnt.data(), h->mg_cnt[h->mg_cur], n * (size_t)K * sizeof(uint16_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(border, h->run_border, sizeof(border), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  std::vector<unsigned char> mine(n * elem);
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t start = h->seg_start + pos[i];
    if (i == 0 && h->rank > 0 && !border[0]) start = ~0ull;  // marks "joins the previous segment"
    memcpy(mine.data() + i * elem, &start, 8);
    memcpy(mine.data() + i * elem + 8, cnt.data() + i * (size_t)K, 2 * (size_t)K);
  }
  std::vector<unsigned char> all;
  std::vector<uint64_t> counts;
  rc = comm_allgatherv_host(h, mine.data(), n, elem, all, counts);
  if (rc != HML_OK) return rc;
  const uint64_t total = all.size() / elem;
  std::vector<uint64_t> starts;
  starts.reserve(total);
  h->g_mg_counts.clear();
  h->g_mg_counts.reserve(total * (size_t)K);
  for (uint64_t i = 0; i < total; ++i) {
    uint64_t start;
    memcpy(&start, all.data() + i * elem, 8);
    const uint16_t* c = reinterpret_cast<const uint16_t*>(all.data() + i * elem + 8);
    if (start == ~0ull) {
      if (starts.empty()) return fail(h, HML_ERR_STATE, "marginal merge: a continued segment without a predecessor");
      for (int s2 = 0; s2 < K; ++s2)
        if (h->g_mg_counts[(starts.size() - 1) * (size_t)K + s2] != (int32_t)c[s2])
          return fail(h, HML_ERR_STATE, "marginal merge: a run crossing a rank border has different counts on its two sides");
      continue;
    }
    starts.push_back(start);
    for (int s2 = 0; s2 < K; ++s2) h->g_mg_counts.push_back((int32_t)c[s2]);
  }
  h->g_mg_size.resize(starts.size());
  for (size_t i = 0; i < starts.size(); ++i) h->g_mg_size[i] = (i + 1 < starts.size() ? starts[i + 1] : h->T_global) - starts[i];
  h->g_mg_valid = true;
  return HML_OK;
}

int hml_marginals_info(hml_t* h, uint64_t* nsegments, uint64_t* iterations, int* K) {
  if (!h) return HML_ERR_ARG;
  if (nsegments) {
    CK(cudaSetDevice(h->device));
    if (h->world > 1 && h->mg_K) {
      if (h->g_mg_valid) {
        *nsegments = h->g_mg_size.size();
      } else {
        // the count alone needs 16 bytes per rank: own segments, and whether the one at position 0 is a real border
        uint64_t mine[2] = {0, 0};
        int rc = mg_current_segments(h, &mine[0]);
        if (rc != HML_OK) return rc;
        uint32_t border[2] = {0, 0};
        CK(cudaMemcpyAsync(border, h->run_border, sizeof(border), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        mine[1] = border[0];
        std::vector<uint64_t> all(2 * (size_t)h->world);
        rc = comm_allgather_host(h, mine, sizeof(mine), all.data());
        if (rc != HML_OK) return rc;
        uint64_t n = 0;
        for (int r = 0; r < h->world; ++r) n += all[2 * r] - ((r > 0 && !all[2 * r + 1]) ? 1 : 0);
        *nsegments = n;
      }
    } else {
      int rc = mg_current_segments(h, nsegments);
      if (rc != HML_OK) return rc;
    }
  }
  if (iterations) *iterations = h->mg_iterations;
  if (K) *K = h->mg_K;
  return HML_OK;
}

int hml_marginals_get(hml_t* h, uint64_t* seg_size, int32_t* counts, uint64_t capacity) {
  if (!h || !seg_size || !counts) return HML_ERR_ARG;
  if (h->mg_K == 0) return fail(h, HML_ERR_STATE, "hml_marginals_reset has not been called");
  CK(cudaSetDevice(h->device));
  if (h->world > 1) {
    int rc = mg_merge_global(h);
    if (rc != HML_OK) return rc;
    const uint64_t n = h->g_mg_size.size();
    if (capacity < n) return fail(h, HML_ERR_CAPACITY, "buffer smaller than the number of marginal segments");
    memcpy(seg_size, h->g_mg_size.data(), n * sizeof(uint64_t));
    memcpy(counts, h->g_mg_counts.data(), n * (size_t)h->mg_K * sizeof(int32_t));
    return HML_OK;
  }
  uint64_t n = 0;
  int rc0 = mg_current_segments(h, &n);
  if (rc0 != HML_OK) return rc0;
  if (capacity < n) return fail(h, HML_ERR_CAPACITY, "buffer smaller than the number of marginal segments");
  const int K = h->mg_K;
  std::vector<uint32_t> pos(n);
  std::vector<uint16_t> cnt(n * (size_t)K);
  CK(cudaMemcpyAsync(pos.data(), h->mg_pos[h->mg_cur], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(cnt.data(), h->mg_cnt[h->mg_cur], n * (size_t)K * sizeof(uint16_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (uint64_t i = 0; i < n; ++i) {
    seg_size[i] = (uint64_t)(i + 1 < n ? pos[i + 1] : h->T) - pos[i];
    for (int s = 0; s < K; ++s) counts[i * K + s] = cnt[i * (size_t)K + s];
  }
  return HML_OK;
}

int hml_get_rows(hml_t* h, double* rows, uint64_t capacity_rows) {
  if (!h || !rows) return HML_ERR_ARG;
  if (!h->rows_valid) return fail(h, HML_ERR_STATE, "the last sweep did not keep its forward rows");
  if (capacity_rows < h->nblocks + 1) return fail(h, HML_ERR_CAPACITY, "row buffer too small");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(rows, h->rows, (h->nblocks + 1) * h->last_K * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return HML_OK;
}

int hml_comm_unique_id(uint8_t id[HML_UNIQUE_ID_BYTES]) {
  if (!id) return HML_ERR_ARG;
  static_assert(sizeof(ncclUniqueId) == HML_UNIQUE_ID_BYTES, "NCCL unique id size");
  if (!g_nccl.load()) {
    g_create_error = g_nccl.error;
    return HML_ERR_CUDA;
  }
  ncclUniqueId u;
  const ncclResult_t r = g_nccl.GetUniqueId(&u);
  if (r != ncclSuccess) {
    g_create_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r);
    return HML_ERR_CUDA;
  }
  memcpy(id, &u, sizeof(u));
  return HML_OK;
}

int hml_comm_init(hml_t* h, int rank, int world, const uint8_t id[HML_UNIQUE_ID_BYTES]) {
  if (!h || !id) return HML_ERR_ARG;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail(h, HML_ERR_ARG, "invalid rank / world size");
  if (h->comm) return fail(h, HML_ERR_STATE, "the handle already joined a communicator");
  if (h->T) return fail(h, HML_ERR_STATE, "join the communicator before loading data");
  if (world == 1) return HML_OK;
  if (!g_nccl.load()) return fail(h, HML_ERR_CUDA, g_nccl.error);
  CK(cudaSetDevice(h->device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  CKN(g_nccl.CommInitRank(&h->comm, world, u, rank));
  h->rank = rank;
  h->world = world;
  const size_t seg_words = kHeadWords + 4 + kOpDoubles + (size_t)world * (kHeadWords + 4 + kOpDoubles);
  CK(dev_alloc(h->seg_dev, seg_words));
  CK(cudaMemsetAsync(h->seg_dev, 0, seg_words * sizeof(double), h->stream));
  CK(dev_alloc(h->stats_gather, (size_t)world * kOutWords));
  CK(dev_alloc(h->run_states, 8 + (size_t)world));
  CK(cudaMemsetAsync(h->run_states, 0, (8 + (size_t)world) * sizeof(unsigned long long), h->stream));
  CK(dev_alloc(h->run_border, 2));
  CK(cudaMemsetAsync(h->run_border, 0, 2 * sizeof(uint32_t), h->stream));
  CK(cudaMallocHost((void**)&h->stats_gather_host, (size_t)world * kOutWords * 8));
  CK(cudaStreamSynchronize(h->stream));
  return setup_p2p(h);
}

int hml_segment_plan(uint64_t T, int world, int rank, uint64_t* start, uint64_t* len) {
  if (!start || !len || world < 1 || rank < 0 || rank >= world) return HML_ERR_ARG;
  if (world > 1 && T < (uint64_t)kTile * world) return HML_ERR_ARG;
  if (world == 1) {
    *start = 0;
    *len = T;
    return HML_OK;
  }
  segment_plan(T, world, rank, start, len);
  return HML_OK;
}

int hml_load_segment_f32_device_md(hml_t* h, const float* x_dev, uint64_t len, uint64_t T, uint32_t nr_dims,
                                   float weight_multiplier) {
  if (!h) return HML_ERR_ARG;
  if (!x_dev) return fail(h, HML_ERR_ARG, "x must not be NULL");
  if (nr_dims == 0) return fail(h, HML_ERR_ARG, "Number of dimensions must be positive!");
  if (nr_dims > HML_MAX_DIMS)
    return fail(h, HML_ERR_ARG, "at most " + std::to_string(HML_MAX_DIMS) + " data dimensions are supported");
  if (T >= (1ull << 32)) return fail(h, HML_ERR_ARG, "a sequence holds fewer than 2^32 observations");
  CK(cudaSetDevice(h->device));
  if (h->world <= 1)
    return len == T ? load_common(h, x_dev, T, weight_multiplier, (int)nr_dims)
                    : fail(h, HML_ERR_ARG, "len != T without a communicator");
  return load_segment_common(h, x_dev, len, T, weight_multiplier, (int)nr_dims);
}

int hml_load_segment_f32_device(hml_t* h, const float* x_dev, uint64_t len, uint64_t T, float weight_multiplier) {
  return hml_load_segment_f32_device_md(h, x_dev, len, T, 1, weight_multiplier);
}

int hml_load_segment_f32_md(hml_t* h, const float* x_host, uint64_t len, uint64_t T, uint32_t nr_dims, float weight_multiplier) {
  if (!h) return HML_ERR_ARG;
  if (!x_host) return fail(h, HML_ERR_ARG, "x must not be NULL");
  if (len == 0 || nr_dims == 0) return fail(h, HML_ERR_ARG, "Input vector for breakpoint weights is empty!");
  CK(cudaSetDevice(h->device));
  float* xd = nullptr;
  CK(dev_alloc(xd, len * nr_dims));
  cudaError_t e = cudaMemcpyAsync(xd, x_host, len * nr_dims * sizeof(float), cudaMemcpyHostToDevice, h->stream);
  if (e != cudaSuccess) {
    dev_free(xd);
    return fail(h, HML_ERR_CUDA, cudaGetErrorString(e));
  }
  const int rc = hml_load_segment_f32_device_md(h, xd, len, T, nr_dims, weight_multiplier);
  cudaStreamSynchronize(h->stream);
  dev_free(xd);
  return rc;
}

int hml_load_segment_f32(hml_t* h, const float* x_host, uint64_t len, uint64_t T, float weight_multiplier) {
  if (!h) return HML_ERR_ARG;
  if (!x_host) return fail(h, HML_ERR_ARG, "x must not be NULL");
  if (len == 0) return fail(h, HML_ERR_ARG, "Input vector for breakpoint weights is empty!");
  CK(cudaSetDevice(h->device));
  float* xd = nullptr;
  CK(dev_alloc(xd, len));
  cudaError_t e = cudaMemcpyAsync(xd, x_host, len * sizeof(float), cudaMemcpyHostToDevice, h->stream);
  if (e != cudaSuccess) {
    dev_free(xd);
    return fail(h, HML_ERR_CUDA, cudaGetErrorString(e));
  }
  const int rc = hml_load_segment_f32_device(h, xd, len, T, weight_multiplier);
  cudaStreamSynchronize(h->stream);
  dev_free(xd);
  return rc;
}

int hml_segment_info(const hml_t* h, int* rank, int* world, uint64_t* seg_start, uint64_t* seg_len,
                     uint64_t* first_block, uint64_t* global_blocks) {
  if (!h) return HML_ERR_ARG;
  if (rank) *rank = h->rank;
  if (world) *world = h->world;
  if (seg_start) *seg_start = h->seg_start;
  if (seg_len) *seg_len = h->T;
  if (first_block) *first_block = h->first_block;
  if (global_blocks) *global_blocks = h->world > 1 ? h->global_blocks : h->nblocks;
  return HML_OK;
}

int hml_exchange_transport(const hml_t* h, int* transport) {
  if (!h || !transport) return HML_ERR_ARG;
  *transport = h->world <= 1 ? HML_EXCHANGE_NONE : (h->p2p ? HML_EXCHANGE_PEER : HML_EXCHANGE_NCCL);
  return HML_OK;
}

int hml_set_detect_mode(hml_t* h, int mode) {
  if (!h) return HML_ERR_ARG;
  if (mode != HML_DETECT_STREAM && mode != HML_DETECT_PYRAMID && mode != HML_DETECT_CANDIDATES)
    return fail(h, HML_ERR_ARG, "unknown detection mode");
  h->detect_mode = mode;
  return HML_OK;
}

int hml_set_forward_mode(hml_t* h, int mode) {
  if (!h) return HML_ERR_ARG;
  if (mode != HML_FORWARD_AUTO && mode != HML_FORWARD_OPERATORS && mode != HML_FORWARD_SPECULATIVE)
    return fail(h, HML_ERR_ARG, "unknown forward mode");
  h->forward_mode = mode;
  h->spec_skip = h->spec_streak = h->spec_good = 0;
  h->spec_level = 0;
  h->spec_patience = 64;
  h->spec_lowered = false;
  return HML_OK;
}

int hml_forward_info(hml_t* h, int* mode, uint64_t* speculative_sweeps, uint64_t* failures, int* piece_blocks,
                     int* warmup_blocks) {
  if (!h) return HML_ERR_ARG;
  if (mode) *mode = h->forward_mode;
  if (speculative_sweeps) *speculative_sweeps = h->spec_sweeps;
  if (failures) *failures = h->spec_failures;
  const int KP = h->KP ? h->KP : 2;
  if (piece_blocks) *piece_blocks = spec_sub_of(KP, h->spec_level);
  if (warmup_blocks) *warmup_blocks = spec_warm_of(KP, h->spec_level);
  return HML_OK;
}

int hml_detect_info(hml_t* h, int* mode, uint64_t* hot_subblocks) {
  if (!h) return HML_ERR_ARG;
  if (mode) *mode = h->detect_mode;
  if (hot_subblocks) {
    *hot_subblocks = 0;
    if (h->detect_mode == HML_DETECT_CANDIDATES) {
      *hot_subblocks = h->cand_valid ? h->cand_n : 0;  // entries the last candidate pass looked at
    } else if (h->T && h->detect_scratch) {
      CK(cudaSetDevice(h->device));
      unsigned long long v = 0;
      CK(cudaMemcpyAsync(&v, detect_hot_count_ptr(h->detect_scratch, h->T), sizeof(v), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      *hot_subblocks = v;
    }
  }
  return HML_OK;
}

int hml_set_timing(hml_t* h, int on) {
  if (!h) return HML_ERR_ARG;
  h->timing = on != 0;
  return HML_OK;
}

int hml_get_timing(hml_t* h, int* nstages, const char** names, float* ms, int capacity) {
  if (!h || !nstages) return HML_ERR_ARG;
  *nstages = (int)h->last_names.size();
  for (int i = 0; i < *nstages && i < capacity; ++i) {
    if (names) names[i] = h->last_names[i].c_str();
    if (ms) ms[i] = h->last_ms[i];
  }
  return HML_OK;
}

int hml_launch_count(const hml_t* h, uint64_t* n) {
  if (!h || !n) return HML_ERR_ARG;
  *n = h->launches;
  return HML_OK;
}

int hml_sync(hml_t* h) {
  if (!h) return HML_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  return HML_OK;
}

int hml_get_stream(hml_t* h, void** stream) {
  if (!h || !stream) return HML_ERR_ARG;
  *stream = (void*)h->stream;
  return HML_OK;
}

}  // extern "C"
