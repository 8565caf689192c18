// hammlet_b200 — internal kernel launch interface (host side of hammlet_b200/csrc/*.cu).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace hml {

// ---- load (hml_load.cu)
void upload_level_norms(const float* host64, cudaStream_t s);
void launch_maxlet_level(const float* in, uint64_t n_valid, uint64_t n_pos, uint64_t stride, int level0, float* coeffs,
                         float* tile_sums, cudaStream_t s);
void launch_bp_weights(const float* c, uint64_t T, float mult, float* w, int sms, cudaStream_t s);
void launch_sum_odd(const float* c, uint64_t T, double* partial, int nblocks, cudaStream_t s);
void launch_integral_cells(const float* x, uint64_t T, double2* pq, double2* cell_tot, cudaStream_t s);
// multivariate input: plane[t] = x[t * nr_dims + dim]; dst[i] = max(dst[i], src[i]) (maxlet, wavelet.hpp:155-160)
void launch_deinterleave(const float* x, uint64_t T, int nr_dims, int dim, float* plane, int sms, cudaStream_t s);
void launch_max_combine(float* dst, const float* src, uint64_t n, int sms, cudaStream_t s);
// segment mode (one sequence split over ranks): edge coefficients c[len - 2^k], k < 12, for the next rank
void launch_pack_edge(const float* coeffs_local, uint64_t len, float* edge16, cudaStream_t s);
// breakpoint weights of the local segment from local coefficients, the replicated table of the
// coefficients at multiples of 4096 (ctop[m] = c[4096 m]) and the 12 edge coefficients of the previous rank
void launch_bp_weights_segment(const float* c_local, const float* ctop, const float* halo, uint64_t seg_start,
                               uint64_t len, uint64_t T, float mult, float* w, int sms, cudaStream_t s);

// ---- boundary detection (hml_detect.cu)
typedef void (*stage_cb_t)(void* user, const char* name);
size_t detect_scratch_bytes(uint64_t T);
// flags -> counts -> ordered block starts; returns the number of kernels launched.  smax != nullptr: pyramid
// mode (only sub-blocks of 32 weights whose maximum reaches the threshold are read), else the streaming kernel.
int launch_detect(const float* w, const uint16_t* smax, uint64_t T, float thr, int force_first, void* scratch,
                  uint32_t* starts, uint64_t capacity, unsigned long long* nblocks_out, cudaStream_t s, stage_cb_t cb,
                  void* user);
// candidate mode: the positions (ascending) and weights of everything that can be a boundary while thr >= floor
// cand_pq (may be null: multivariate data) receives the integral pair pq[position] of every candidate
void launch_cand_gather(const float* w, const double2* pq, const uint32_t* starts, uint32_t n, float* cand_w,
                        uint32_t* cand_pos, double2* cand_pq, int sms, cudaStream_t s);
uint32_t cand_ctas(uint32_t nc);
// spq (capacity + 1 entries; ignored without cand_pq) receives the integral pairs of the block starts, in block order,
// and pq[T] behind the last one: SweepBuffers::spq for the block statistics of this structure
int launch_detect_candidates(const float* cand_w, const uint32_t* cand_pos, const double2* cand_pq, uint32_t nc, float thr,
                             uint32_t* cta_scratch, uint32_t scratch_ctas, uint32_t* starts, double2* spq, const double2* pq,
                             uint64_t capacity, uint64_t T, unsigned long long* nblocks_out, cudaStream_t s, stage_cb_t cb,
                             void* user, const struct SweepBuffers* head, unsigned long long head_seq);
// head != null (split sequence with peer mailboxes): the scatter's last CTA also forms the head partial of the rank's
// segment and runs the head exchange number head_seq (what launch_seg_head does as a kernel of its own)
size_t pyramid_entries(uint64_t T);  // bf16 entries, one per 32 weights, padded to whole spans
void launch_build_pyramid(const float* w, uint64_t T, uint16_t* smax, int sms, cudaStream_t s);
const unsigned long long* detect_hot_count_ptr(const void* scratch, uint64_t T);

// ---- peer-memory carry exchange of the segment-split mode (hml_p2p.cu)
constexpr int kP2PMaxWorld = 64;
constexpr int kP2PSlots = 4;            // heads, operators, maps, statistics
constexpr size_t kP2PPayload = 12288;   // largest payload: the result block of a K = 32 sweep on 5-dimensional data
constexpr size_t kP2PEntry = 2 * kP2PPayload;  // on the wire every 4 payload bytes travel with a 4-byte sequence tag
constexpr unsigned long long kP2PTimeoutNs = 30ull * 1000ull * 1000ull * 1000ull;
struct P2PPeers {
  unsigned char* box[kP2PMaxWorld];  // mailbox of every rank as mapped into this process (own one included)
};
// entry written by rank `src` for exchange `slot`, sequence parity `parity`, inside any mailbox
__host__ __device__ inline size_t p2p_entry_offset(int parity, int slot, int src, int world) {
  return ((size_t)(parity * kP2PSlots + slot) * world + src) * kP2PEntry;
}
inline size_t p2p_mailbox_bytes(int world) { return 2 * (size_t)kP2PSlots * world * kP2PEntry; }
// what the device side needs, resident in global memory (SegInfo::p2p points at it)
struct P2PDev {
  P2PPeers peers;
  unsigned int* timeout_flag;  // mapped host word raised by an exchange that gave up waiting
  int rank, world;
};
enum { kSlotHeads = 0, kSlotOps = 1, kSlotMaps = 2, kSlotStats = 3 };
// all-gather of `bytes` (a multiple of 8, <= kP2PPayload) per rank into recv (rank-major), one single-CTA kernel
void launch_p2p_exchange(const P2PDev* d, int slot, uint64_t seq, const void* send, size_t bytes, void* recv,
                         cudaStream_t s);

// ---- block-level sweep kernels (hml_sweep.cu)
constexpr int kMaxDims = 5;  // HML_MAX_DIMS
struct ModelHost {  // what the C ABI receives, validated
  int K;
  int use_self;
  double mean[32], var[32], A[32 * 32], pi[32];  // univariate: per state (= per parameter)
  // multivariate (D > 1): parameters resolved through the mapping, per state and dimension
  int D;
  double mean_sd[kMaxDims][32], var_sd[kMaxDims][32];
};

// Device buffers of one handle that the sweep kernels touch.  All per-block arrays are stored in
// the chunk-interleaved order defined in hml_sweep.cu (Layout::perm).
// Segment mode (world > 1): gathered carries of all ranks and this rank's send slots.
// head record of a rank in segment mode: {blocks of the rank, length of the head = observations in front of the rank's
// first boundary, then (sum x, sum x^2) of the head per data dimension}
constexpr int kHeadWords = 2 + 2 * kMaxDims;
struct SegInfo {
  int rank, world;
  const double* heads;   // world x kHeadWords
  const double* ops;     // world x (KP*KP + KP): segment operator mantissas (row-major) then row exponents
  const uint64_t* maps;  // world x 4 words: segment map f_first o ... o f_last (byte-packed)
  double* send_head;
  double* send_op;
  uint64_t* send_map;
  unsigned long long* overflow;  // set if this rank's block arrays were too small
  // peer-memory exchange embedded in the producing kernels (null: the caller exchanges between launches)
  const P2PDev* p2p;
  const unsigned long long* stats_send;  // the rank's result block
  unsigned long long* stats_recv;        // world x stats_words
};

struct SweepBuffers {
  SegInfo seg;
  // inputs resident since load
  const double2* pq;        // T+1 cell-local running sums of (x, x^2)
  const double4* cell_pref; // per cell: double-double exclusive prefix of cell totals (hi_x, lo_x, hi_q, lo_q)
  // multivariate data: D planes of pq / cell_pref / bS, `*_stride` elements apart (bS: `capacity` apart)
  int D;
  uint64_t pq_stride, cell_stride;
  // block structure
  const uint32_t* starts;   // capacity+1, natural order
  const double2* spq;       // capacity+1, natural order: pq[starts[b]] (and pq[T] behind the last block) when the block
                            // list came from the candidate list (univariate data), else null: gather from pq
  const unsigned long long* nblocks;  // device scalar
  uint64_t capacity;        // blocks the per-block arrays can hold
  uint32_t* bN;             // block sizes
  double2* bS;              // block (sum x, sum x^2)
  // per sweep
  double* e;                // KP per block: exp(E_s - maxE)
  double* maxE;             // per block (only filled for loglik)
  double* alpha;            // KP per block: normalised forward vector alpha_t
  uint8_t* maps;            // KPB bytes per block: backward map j -> state
  uint8_t* states;          // sampled state per block
  double* chunk_ops;        // per chunk KP*KP
  int* chunk_exp;           // per chunk KP
  uint8_t* chunk_maps;      // per chunk KPB: composed map of the LATER chunks of the same tile
  uint8_t* chunk_submaps;   // per chunk 3 x KPB: composed maps of the chunk's blocks from step 8, 16 and 24 on
  uint8_t* tile_maps;       // per tile KPB: composed map of the tile
  uint8_t* tile_qin;        // per tile: state of the block following the tile
  unsigned* tickets;        // kTickets arrival counters (zero between launches): the CTA that arrives last finishes the step
  // pinned host memory (device-addressable): where the sweep's last kernel leaves the result block — the blocks of all
  // ranks when it ran the statistics exchange itself — so that no copy has to follow it; result_words words per rank
  unsigned long long* result_host;
  uint32_t result_words;
  double* tile_ops;         // per tile KP*KP
  int* tile_exp;            // per tile KP
  double* tile_ain;         // per tile KP: normalised forward vector entering the tile
  double* group_ops;        // scratch of the tile scan: operators of groups of tiles (256 entries)
  int* group_exp;
  double* group_ain;        // forward vector entering each group of tiles (K > 8, multi-CTA tile scan)
  double* wide_ops;         // K > 8: per-CTA scratch of k_fwd_chunks_wide (tree nodes of the tile operator), or null
  int* wide_exp;
  double* rows;             // (capacity+1)*K, natural order, only with KEEP_ROWS (may be null)
  const double* replay_u;   // device copy of replay uniforms (may be null)
  double* partials;         // reduce scratch
  unsigned long long* out_u64;  // [0..KP) stat_n, [KP..KP+KP*KP) trans, then [fallbacks]
  double* out_f64;          // [0..KP) sum, [KP..2KP) sumsq, [2KP] loglik; then per extra dimension d >= 1:
                            // [2KP+1 + (d-1)*2KP ..) per-state sum and sumsq of dimension d
};

int padded_states(int K);          // KP for K (0 if unsupported)
int map_bytes(int KP);             // KPB
int chunks_per_tile(int KP);       // C
constexpr int kChunkLen = 32;      // L
constexpr int kTileBlocks = 1024;  // L * C: per-block arrays are sized in multiples of this
size_t reduce_partials_doubles(int KP, int grid);
constexpr int kWideCtasPerSm = 2;  // CTAs of k_fwd_chunks_wide launched per SM (each owns a scratch area)
size_t wide_scratch_doubles(int KP);  // per CTA; 0 for K <= 8
size_t wide_scratch_ints(int KP);

// ---- the whole sweep in one persistent kernel + device-resident Gibbs chain (hml_fused.cuh; K <= 8, univariate,
// single handle, candidate-list detection, at most kFusedMaxTiles tiles)
constexpr int kFusedThreads = 256;
constexpr int kFusedMaxTiles = 64;
constexpr int kChainMaxStates = 8;

enum { kChainOk = 0, kChainThreshold = 1, kChainCapacity = 2, kChainFallback = 3, kChainNumeric = 4 };

// Device-resident Gibbs chain (one per handle): parameters, priors, RNG position, status of the last launch.
struct ChainDev {
  double mean[kChainMaxStates], var[kChainMaxStates], A[kChainMaxStates * kChainMaxStates], pi[kChainMaxStates];
  // derived from them (make_model): 1 / (2 var), log sd + mean^2 / (2 var), log A_ss (0 without self transitions)
  double inv2var[kChainMaxStates], lognorm[kChainMaxStates], loga[kChainMaxStates];
  float prior_theta[kChainMaxStates][4];  // NIG hyper-parameters alpha, beta, mu0, nu per state (real_t = float)
  float prior_trans, prior_self, prior_pi;
  float thr;         // threshold of the current parameters, BreakpointArray.hpp:195-199
  float cand_floor;  // the candidate list serves thresholds >= this
  uint32_t T;        // observations of the whole sequence
  int K, use_self;
  unsigned long long seed, sweep;  // Philox key; sweeps sampled so far (the counter of the parameter streams)
  unsigned int abort_code, sweeps_done;
  unsigned long long nblocks_seen;  // block count that did not fit (kChainCapacity)
  // One word per phase of a sweep: a phase reports through its own word, which is read after the grid barrier that
  // ends the phase and is not written again before the next sweep — so all threads of the grid take the same decision
  // (a single word could be raised by a CTA that is already a phase ahead while a slower one still reads it).
  unsigned int phase_abort[8];
};

struct FusedArgs {
  ChainDev* chain;
  const float* cand_w;
  const uint32_t* cand_pos;
  const double2* cand_pq;
  uint32_t nc;
  uint32_t T_local;
  uint32_t* cta_count;  // gridDim.x words
  unsigned* barriers;   // 1 + kFusedMaxTiles words, zero at launch: grid barrier, then one barrier per tile
  double* qtot;         // 4 * kFusedMaxTiles operators of K*K + K doubles: the quarter-tile totals
  unsigned long long* qmap;  // 4 * kFusedMaxTiles words: the quarter-tile maps
  double* subops;       // per quarter 32 sub-chunk prefix operators (K*K + K doubles): CTAs with several quarters per phase
  unsigned long long* submaps;  // per quarter 32 sub-chunk suffix maps: the same
  unsigned long long* phase_ns;  // 16 words or null: globaltimer stamps of CTA 0 at the phase borders of the last sweep
  int nsweeps;
  int sample_params;    // 0: the model stays as it is (single sweeps with a caller-provided model)
  int philox_sweep_from_chain;  // uniforms keyed by chain->sweep (chain mode) or by `sweep` (single sweep)
  unsigned long long seed, sweep;
};

// fills inv2var / lognorm / loga from mean / var / A (host side: hml_chain_set, single fused sweeps)
inline void chain_derive(ChainDev* c) {
  for (int i = 0; i < c->K; ++i) {
    c->inv2var[i] = 1.0 / (2.0 * c->var[i]);
    c->lognorm[i] = log(sqrt(c->var[i])) + c->mean[i] * c->mean[i] / (2 * c->var[i]);
    c->loga[i] = c->use_self ? log(c->A[i * c->K + i]) : 0.0;
  }
}
// cooperative launch of a.nsweeps sweeps on `grid` CTAs; cudaError_t as int, or -2 for an unsupported K
int launch_sweep_fused(int KP, const SweepBuffers& b, const FusedArgs& a, int grid, cudaStream_t s);
int fused_max_grid(int KP, int sms);
// the parameter phase alone (after a sweep of the multi-kernel path): theta, pi, A of the next sweep from the result block
int launch_chain_params(int KP, ChainDev* ch, const unsigned long long* out_u64, const double* out_f64, cudaStream_t s);

struct SweepLaunch {
  uint32_t flags;      // HML_SWEEP_* bits
  bool gather;         // recompute block sums from the integral arrays
  bool mixture;
  uint64_t seed, sweep;
  int sms;
  uint64_t nblocks_hint;  // upper bound used to size grids (capacity if unknown)
  // segment mode: all-gathers the named carry (send slot -> gathered array) on the stream; 0 on success
  int (*exchange)(void* user, int which);
  // kernels that embed the exchange (seg.p2p != null) ask for the sequence number of their collective
  unsigned long long (*next_seq)(void* user, int which);
  void* exchange_user;
  uint32_t stats_words;  // 8-byte words of the result block travelling in the statistics exchange (0: not fused)
  bool speculate;        // forward filter by guessed chunk starts + repair pass (result word KP + KP*KP + 1 counts failures)
  int spec_warm;         // blocks in front of a piece its guess is pushed through
  int spec_sub;          // K <= 8: blocks per piece of the speculative pass (8, 16 or 32)
  bool* result_on_host;  // set when the last kernel wrote the result block(s) to SweepBuffers::result_host
};
// Levels of the speculative pass: (piece length, warm-up).  Short pieces give a latency-bound recursion more warps;
// data on which the filter forgets slowly needs long warm-ups, which only pay with long pieces.  Past the last level
// the operator scan is cheaper.
constexpr int kSpecLevels = 3;
inline int spec_sub_of(int KP, int level) { return KP <= 8 ? (level == 0 ? 8 : level == 1 ? 16 : 32) : 32; }
inline int spec_warm_of(int KP, int level) { return KP <= 8 ? (level == 0 ? 4 : level == 1 ? 16 : 64) : (level == 0 ? 8 : level == 1 ? 32 : 128); }
enum { kTicketScatter = 0, kTicketFixup = 1, kTicketChunkMaps = 2, kTicketReduce = 3, kTickets = 8 };
enum { kExchangeHeads = 0, kExchangeOps = 1, kExchangeMaps = 2, kExchangeStats = 3 };
// segment mode: head partial of this rank -> seg.send_head (to be all-gathered before the block statistics)
// seq != 0: the kernel also runs the head exchange itself (seg.p2p != null)
void launch_seg_head(const SweepBuffers& b, uint64_t seg_len, unsigned long long seq, cudaStream_t s);

// Enqueues all block-level kernels of one sweep on `s`; returns the number of kernels launched.
// stage_cb(name) is called before each stage so the caller can drop timing events.
int launch_sweep(const ModelHost& m, const SweepBuffers& b, const SweepLaunch& l, cudaStream_t s, stage_cb_t cb,
                 void* user);

// Exact sequential forward pass (one thread) + backward + reductions; used after a uniform fallback.
int launch_sweep_sequential(const ModelHost& m, const SweepBuffers& b, const SweepLaunch& l, cudaStream_t s);

// Copies per-block arrays out of the interleaved order: dst_states (int16), dst_sum/dst_sumsq.
void launch_unpermute(const SweepBuffers& b, int KP, uint64_t nblocks, int16_t* dst_states, double* dst_sum,
                      double* dst_sumsq, cudaStream_t s);
// Run-length view of the sampled states: tile_counts[0..ntiles] <- exclusive offsets of the runs that start in each
// 1024-block tile, [ntiles] = number of runs (2 launches); then (start position, state) per run (1 launch).
// Segment mode: `last_states` holds the all-gathered state of every rank's last block (launch_segments_last_state
// fills this rank's word before the exchange); ranks > 0 always emit a run start at local position 0 (a virtual run in
// the previous rank's state unless their first block begins there) and record in border[1] / accumulate in border[0]
// whether that position started a run of the whole sequence.  ntiles = max(1, ceil(nblocks / 1024)) then.
void launch_segments_last_state(const SweepBuffers& b, unsigned long long* send, cudaStream_t s);
void launch_segments_count(const SweepBuffers& b, uint64_t nblocks, const unsigned long long* last_states,
                           uint32_t* tile_counts, int sms, cudaStream_t s);
void launch_segments_write(const SweepBuffers& b, uint64_t nblocks, const unsigned long long* last_states, uint32_t* border,
                           int accumulate, const uint32_t* tile_offsets, uint32_t* seg_start, int16_t* seg_state, int sms,
                           cudaStream_t s);
// State marginals on the device: merges the run starts R[m] / run states of one recorded iteration into the
// refinement (P[n], cnt[n x K]) -> (P2, cnt2) with new_flags[m] = number of new boundaries (3 launches).
// The segment count n and the run count m are read from device memory (*n_ptr, *m_ptr), the new segment count is
// written to *n_out; n_upper / m_upper only size the grid.
void launch_marginals_merge(const uint32_t* P, const uint32_t* n_ptr, uint64_t n_upper, const uint16_t* cnt, const uint32_t* R,
                            const int16_t* rstate, const uint32_t* m_ptr, uint64_t m_upper, uint32_t* run_of_old,
                            uint32_t* olds_below, uint32_t* new_flags, int K, uint32_t* P2, uint16_t* cnt2, uint32_t* n_out,
                            int sms, cudaStream_t s);
// Only block sums (no model): gather statistics for the current starts.
void launch_block_stats(const SweepBuffers& b, int KP, uint64_t nblocks_hint, int sms, cudaStream_t s);

}  // namespace hml
