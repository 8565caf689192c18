/* hammlet_b200 — C ABI of the B200-native hot path of HaMMLET
 * (forward-backward Gibbs sweep over dynamically wavelet-compressed blocks).
 *
 * The reference (wiedenhoeft/HaMMLET) has no plugin / FFI boundary: `hammlet` is one translation
 * unit of C++ templates.  This header is the boundary the new build introduces;
 * each entry point names the reference code it replaces (paths relative to the reference's src/).
 * The C++ model surface in hammlet_b200/host/ (Blocks, Statistics, Emissions, StateSequence, ...)
 * and the ctypes binding in hammlet_b200/capi.py call exactly these functions.
 *
 * Conventions: plain C types; the caller owns host buffers, the library owns device memory; no
 * exceptions cross the boundary — every function returns HML_OK (0) or a negative error code and
 * hml_last_error() gives the message (host wrappers rethrow it as std::runtime_error, matching
 * the reference's error behaviour, main.cpp:467-474).  A handle is not thread-safe (the
 * reference is single-threaded).  There is NO CPU fallback: without a CUDA device hml_create
 * fails.  One handle = one sequence (or one contiguous shard of a sequence) on one GPU.
 *
 * Numerics: breakpoint weights are fp32 and bit-identical to the reference's; block boundaries
 * and all counts are exact; block statistics, emission terms and the trellis are fp64.
 */
#ifndef HAMMLET_B200_H
#define HAMMLET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hml_ctx hml_t;

enum {
  HML_OK = 0,
  HML_ERR_CUDA = -1,     /* a CUDA runtime call or kernel failed */
  HML_ERR_ARG = -2,      /* invalid argument */
  HML_ERR_STATE = -3,    /* call order: no data loaded / no blocks created / no sweep run */
  HML_ERR_CAPACITY = -4, /* a caller buffer is too small */
  HML_ERR_NUMERIC = -5   /* negative backward variable etc. (ForwardBackward.hpp:147-149) */
};

#define HML_MAX_STATES 32
#define HML_MAX_DIMS 5 /* data dimensions of multivariate input (`-s C p d`: states = p^d <= HML_MAX_STATES) */

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Creates a context on CUDA device `device` (own stream, scratch).  Fails if no device exists. */
int hml_create(hml_t** out, int device);
int hml_destroy(hml_t* h);
/* Message of the last error on this handle (h == NULL: last error of hml_create). */
const char* hml_last_error(const hml_t* h);
const char* hml_version(void);

/* ---- load: replaces MaxletTransform (wavelet.hpp:97-188), HaarBreakpointWeights
 *      (wavelet.hpp:68-93), the weight multiplier (main.cpp:332-334), the Statistics<IntegralArray>
 *      constructor (Statistics/IntegralArray.hpp:136-191) and the Blocks<BreakpointArray>
 *      constructor (Blocks/BreakpointArray.hpp:130-184; its skip pointers are not needed here). */

/* x: T values in host memory (univariate).  T < 2^32. */
int hml_load_f32(hml_t* h, const float* x_host, uint64_t T, float weight_multiplier);
/* Same with x already in device memory of the context's device (not modified, not retained). */
int hml_load_f32_device(hml_t* h, const float* x_dev, uint64_t T, float weight_multiplier);
/* Multivariate data (main.cpp:114-137 `-s C p d`, wavelet.hpp:131-163): x holds T positions x nr_dims values,
 * position-major as they stand in the input stream.  The breakpoint weights come from the maxlet transform (the
 * maximum over the dimensions of the absolute Haar coefficients, wavelet.hpp:155-160); integral arrays are kept per
 * dimension (Statistics/IntegralArray.hpp:176-182).  nr_dims = 1 is hml_load_f32.  A handle that joined a communicator
 * loads its share with hml_load_segment_f32_md. */
int hml_load_f32_md(hml_t* h, const float* x_host, uint64_t T, uint32_t nr_dims, float weight_multiplier);
int hml_load_f32_device_md(hml_t* h, const float* x_dev, uint64_t T, uint32_t nr_dims, float weight_multiplier);
int hml_nr_dims(const hml_t* h, uint32_t* nr_dims);
int hml_size(const hml_t* h, uint64_t* T);
/* Noise estimate of main.cpp:303-311: mean of the level-1 |detail| coefficients / sqrt(2/pi). */
int hml_sigma_hat(hml_t* h, double* sigma_hat);
/* Parity/debug: copy the fp32 breakpoint weights (after the multiplier) or the maxlet
 * coefficients (before HaarBreakpointWeights) to host memory; n = T. */
int hml_get_weights(hml_t* h, float* dst_host, uint64_t n);
int hml_get_coeffs(hml_t* h, float* dst_host, uint64_t n);

/* ---- blocks: replaces Blocks::createBlocks/initForward/next (Blocks/BreakpointArray.hpp:189-235)
 *      and Statistics::setStats/addBlockStats (Statistics/IntegralArray.hpp:104-124,198-212). */

/* Block boundaries {0} U {t : !(w[t] < threshold)} and per-block (N, sum x, sum x^2).
 * The threshold is the caller's fp32 value, sqrt(2 log T * min var) evaluated on the host exactly
 * as BreakpointArray.hpp:195-199 does; it is never recomputed on the device. */
int hml_create_blocks(hml_t* h, float threshold, uint64_t* nblocks);
/* How the boundary positions are found (the result is the same set):
 *   HML_DETECT_STREAM   every weight is read each time (4 bytes/observation, the HBM-roofline formulation);
 *   HML_DETECT_PYRAMID  the device analogue of the reference's skip pointers
 *                       (Blocks/BreakpointArray.hpp:150-182): a max pyramid over sub-blocks of 32 weights is
 *                       built at load, and only sub-blocks whose maximum reaches the threshold are read.
 *   HML_DETECT_CANDIDATES (default) the threshold of a Gibbs chain moves by a hair from sweep to sweep, so the
 *                       positions whose weight is not below a floor (0.75 x the threshold the list was built for) are
 *                       kept, in order, with their weights; while threshold >= floor one coalesced pass over that list
 *                       (8 bytes per candidate, ~1.3 candidates per block) finds the boundaries.  The list is rebuilt
 *                       by a pyramid pass when the threshold drops below the floor or the list has become much longer
 *                       than the block list.  Thresholds <= 0, NaN and inf go through the pyramid pass.
 * hml_detect_info reports the mode and the number of sub-blocks (pyramid) or candidates (candidate mode) the last
 * detection pass had to read. */
enum { HML_DETECT_STREAM = 0, HML_DETECT_PYRAMID = 1, HML_DETECT_CANDIDATES = 2 };
int hml_set_detect_mode(hml_t* h, int mode);
int hml_detect_info(hml_t* h, int* mode, uint64_t* hot_subblocks);
/* How the forward filter (ForwardBackward.hpp:64-125) of a sweep is parallelised.
 *   HML_FORWARD_OPERATORS    scan of K x K operators over chunks, tiles and the sequence: exact for any data, 2 K^3 flop
 *                            per block.
 *   HML_FORWARD_SPECULATIVE  every piece of 8, 16 or 32 blocks runs the reference's vector recursion (2 K^2 flop per
 *                            block) from a guessed start (uniform pushed through the blocks in front of the piece); a
 *                            second pass restarts every piece from the last row of its predecessor and rewrites rows
 *                            until the new row is parallel to the stored one (every component within 1e-13 relative).
 *                            If a piece has not met its guess by its last block the sweep is run again through the
 *                            operator scan, so the result never depends on the assumption that the filter forgets its
 *                            start.
 *   HML_FORWARD_AUTO (default) speculative.  A failed sweep is repeated through the operator scan and the following
 *                            sweeps use longer pieces behind longer warm-ups: (8, 4) -> (16, 16) -> (32, 64) blocks for
 *                            K <= 8, (32, 8) -> (32, 32) -> (32, 128) above; back down after 64 good sweeps (4 x as
 *                            many each time that turns out wrong).  If the last level is not enough either — blocks
 *                            of one or two observations with levels a sigma apart can take hundreds of blocks to
 *                            forget — the operator scan takes the next 1, 3, 7, ... 63 sweeps before the next attempt.
 * A split sequence speculates too (the first chunk of a rank is repaired from the last row of the rank before: one
 * all-gather of K + 1 words instead of the K x K segment operators), except when the log-likelihood is asked for.
 * Mixture sweeps have no forward filter; the fused kernel uses the operator scan.  hml_forward_info: the mode, the
 * number of speculative sweeps so far, how many of them had to be repeated, and the piece length and warm-up the next
 * speculative sweep would use (any pointer may be NULL). */
enum { HML_FORWARD_AUTO = 0, HML_FORWARD_OPERATORS = 1, HML_FORWARD_SPECULATIVE = 2 };
int hml_set_forward_mode(hml_t* h, int mode);
int hml_forward_info(hml_t* h, int* mode, uint64_t* speculative_sweeps, uint64_t* failures, int* piece_blocks,
                     int* warmup_blocks);
int hml_nr_blocks(const hml_t* h, uint64_t* nblocks);
/* Copies the current block structure to host: starts[nblocks] (block b = [starts[b], starts[b+1])
 * with starts[nblocks] = T implied), sum[nblocks], sumsq[nblocks].  Any pointer may be NULL. */
int hml_get_blocks(hml_t* h, uint32_t* starts, double* sum, double* sumsq, uint64_t capacity);
/* Block sums of data dimension `dim` (y.suffStat(dim), Emissions.hpp:60-64); dim 0 is what hml_get_blocks returns. */
int hml_get_block_sums(hml_t* h, uint32_t dim, double* sum, double* sumsq, uint64_t capacity);

/* ---- sweeps: replace StateSequence<ForwardBackward>::sample (StateSequence/ForwardBackward.hpp:
 *      16-213, incl. Trellis.hpp) and StateSequence<Mixture>::sample (StateSequence/Mixture.hpp:
 *      31-144) up to, but not including, the O(K^2) conjugate updates, which stay on the host. */

typedef struct {
  int32_t K;               /* number of states, 2..HML_MAX_STATES (univariate: state == parameter) */
  int32_t use_self_transitions; /* 0 with -S (main.cpp:157) */
  const double* mean;      /* P   theta.value()[p].mean()   (P = K for univariate data) */
  const double* var;       /* P   theta.value()[p].var()   */
  const double* A;         /* K*K row-major A(i,j)         */
  const double* pi;        /* K   pi.valueVector()         */
  /* Multivariate data only (leave zero / NULL otherwise): the handle's nr_dims, the number P of emission
   * parameters and mapping[s * nr_dims + d] = parameter used by state s in dimension d (Mapping.hpp:53-137). */
  int32_t nr_dims;
  int32_t nr_params;
  const int32_t* mapping;
} hml_model;

typedef struct {
  uint64_t nblocks;           /* blocks of the structure the sweep ran on */
  uint64_t uniform_fallbacks; /* ForwardBackward.hpp:106-111 events ("[WARNING] Uniform sampling...") */
  double loglik;              /* sum_t (max_s E_t(s) + log forwardSum_t); only if HML_SWEEP_LOGLIK */
  /* caller-provided arrays, filled on return */
  double* stat_sum;           /* P    per-parameter sum x  (ForwardBackward.hpp:189-191; P = K for univariate data) */
  double* stat_sumsq;         /* P    per-parameter sum x^2 */
  uint64_t* stat_n;           /* P    per-parameter number of observations (Kahan term count) */
  uint64_t* trans;            /* K*K  transition counts incl. N-1 self transitions per block and the
                                      phantom 0 -> q0 transition (:182-184) */
  uint64_t* counts;           /* K    state occupancy (:185) */
} hml_sweep_out;

enum {
  HML_SWEEP_DYNAMIC = 1, /* re-derive the block structure from `threshold` first (HMM.hpp:100-102) */
  HML_SWEEP_LOGLIK = 2,  /* also accumulate the forward log-likelihood */
  HML_SWEEP_KEEP_ROWS = 4, /* keep the forward rows (B+1)xK for hml_get_rows (parity/debug) */
  HML_SWEEP_FUSED = 8 /* hml_fb_sweep only: run the sweep as ONE persistent cooperative kernel (boundaries from the
                         candidate list, emission terms, forward filter, backward sampling, statistics; one CTA per tile
                         of 1024 blocks, grid-wide barriers instead of kernel boundaries) — the building block of the
                         device-resident chain below, 3-4x faster than the 13-kernel sweep for block structures of up to
                         65 536 blocks.  Same arrays, same numerics, same uniforms (Philox counters or replayed): the
                         result is that of the multi-kernel sweep.  K <= 8, univariate data, single handle, candidate
                         detection mode, no log-likelihood / kept rows; sweeps that do not qualify at run time (more
                         blocks than 64 tiles, a threshold the candidate list cannot serve, a vanished forward sum) take
                         the multi-kernel path by themselves. */
};

/* One FBG sweep.  Uniforms: counter-based Philox4x32-10 keyed by (seed, sweep_index, block), or —
 * replay mode — `replay_uniforms` (host, n_replay >= nblocks) consumed in the reference's order,
 * i.e. entry 0 samples the LAST block (ForwardBackward.hpp:140-162).  Replay needs the block count
 * up front, so it cannot be combined with HML_SWEEP_DYNAMIC: call hml_create_blocks first. */
int hml_fb_sweep(hml_t* h, const hml_model* m, uint32_t flags, float threshold, uint64_t seed, uint64_t sweep_index,
                 const double* replay_uniforms, uint64_t n_replay, hml_sweep_out* out);
/* One mixture sweep (per-block independent draw); replay uniforms are consumed in block order. */
int hml_mix_sweep(hml_t* h, const hml_model* m, uint32_t flags, float threshold, uint64_t seed, uint64_t sweep_index,
                  const double* replay_uniforms, uint64_t n_replay, hml_sweep_out* out);

/* ---- device-resident Gibbs chain: sampleHMM (HMM.hpp:99-121) without a host round trip per sweep ----------------
 *
 * theta, pi and A, their conjugate hyper-parameters and the random streams live in device memory.  After every sweep
 * one CTA performs the Normal-Inverse-Gamma and Dirichlet updates from the sweep's statistics (Conjugate.hpp:120-205,
 * real_t = float like the reference; pi from the occupancy counts, ForwardBackward.hpp:211) and draws theta_k
 * (var = 1 / Gamma(alpha, 1 / beta), mean ~ N(mu0, sqrt(var / nu)); Theta.hpp:203-211, Distribution.hpp:76-87), pi and
 * the rows of A (normalised Gamma draws, Distribution.hpp:116-139) for the next one, then the threshold
 * sqrtf(2 logf(T) min var) (BreakpointArray.hpp:195-199); the posteriors fall back to the priors after every draw
 * (Theta.hpp:209).  Draws come from Philox streams keyed by (seed, sweep number, draw): a chain is reproducible and does
 * not depend on how its sweeps were batched, but it is not the reference's mt19937 stream — `hammlet -replay` and
 * hml_fb_sweep with host parameters keep that.  hml_chain_run(n) puts n sweeps on the device and returns when they are
 * done: as ONE launch of the persistent kernel of HML_SWEEP_FUSED where that applies (nothing crosses PCIe between
 * sweeps), sweep by sweep through the multi-kernel path otherwise (more than 64 tiles of blocks, a segment-split
 * sequence, a threshold outside the candidate list: the parameter phase then still runs on the device, and the host only
 * forwards the parameters).  K <= 8, univariate data.  Every rank of a split sequence runs the same chain. */
int hml_chain_init(hml_t* h, int K, const float nig_prior[4] /* alpha, beta, mu0, nu for every state: main.cpp:348-362 */,
                   float trans, float self_trans /* main.cpp:146-155 */, float alpha_pi /* main.cpp:165-166 */, uint64_t seed,
                   int use_self_transitions);
/* Parameters as of the last hml_chain_* call; any pointer may be NULL.  A is K*K row-major. */
int hml_chain_set(hml_t* h, const double* mean, const double* var, const double* A, const double* pi);
int hml_chain_get(hml_t* h, double* mean, double* var, double* A, double* pi, float* threshold, uint64_t* sweeps_so_far);
/* nsweeps dynamic forward-backward sweeps.  *fused_sweeps (may be NULL) = how many of them ran inside the persistent
 * kernel; `last` (may be NULL) receives the statistics of the last sweep.  Afterwards hml_get_states / hml_get_segments /
 * hml_marginals_add see the last sweep's state sequence, as after hml_fb_sweep. */
int hml_chain_run(hml_t* h, uint64_t nsweeps, uint64_t* fused_sweeps, hml_sweep_out* last);

/* ---- records: inputs of Records::record(state, N) (Records.hpp:155-235) -------------------- */

/* Per-block sampled states of the last sweep (marginal_t = int16, includes.hpp:13). */
int hml_get_states(hml_t* h, int16_t* states, uint64_t capacity);
/* The last sweep's state sequence merged into maximal equal-state runs, as Records::record forms
 * them (Records.hpp:166-188): seg_size[i] observations in state seg_state[i].  Call with NULL
 * arrays to get *nsegments only.  Segment mode: a COLLECTIVE call (every rank makes it after the same sweep) that
 * returns the runs of the WHOLE sequence on every rank — a run that crosses a rank border is one entry, as it is one
 * segment in the reference's files. */
int hml_get_segments(hml_t* h, uint64_t* nsegments, uint64_t* seg_size, int16_t* seg_state, uint64_t capacity);
/* State marginals accumulated on the device: StateMarginals::addRecord / save (StateMarginals.hpp:51-137,268-310).
 * The structure is the reference's: the common refinement of all recorded segmentations, one count per state and
 * segment (marginal_t = int16, so at most 32767 iterations).  hml_marginals_add merges the equal-state runs of the
 * last sweep into it without leaving the device and without waiting for it (six small kernels queued on the handle's
 * stream; the counts they need live in device memory); only hml_marginals_get moves data to the
 * host: seg_size[n] and counts[n * K] (row-major), which printed as `size TAB c_0 TAB ... c_{S-1}` with S = highest
 * recorded label + 1 is the reference's marginals file.  Loading new data or calling hml_marginals_reset starts over.
 * Segment mode: every rank accumulates the marginals of its own positions on its own device; hml_marginals_add is
 * collective (the ranks trade the state of their last block — 8 bytes — so that runs continue across rank borders), and
 * hml_marginals_info(nsegments) / hml_marginals_get are collective and return the marginals of the WHOLE sequence on
 * every rank: the ranks' lists concatenated, with the segment at a rank's first observation joined to its left
 * neighbour unless a run of some recorded iteration really started there.  The result is the list a single handle
 * holding the whole sequence accumulates (tests/mgpu_worker.py compares them entry by entry). */
int hml_marginals_reset(hml_t* h, int K);
int hml_marginals_add(hml_t* h);
int hml_marginals_info(hml_t* h, uint64_t* nsegments, uint64_t* iterations, int* K);
int hml_marginals_get(hml_t* h, uint64_t* seg_size, int32_t* counts, uint64_t capacity);
/* Parity/debug: forward rows of the last HML_SWEEP_KEEP_ROWS sweep, (nblocks+1) x K, as the
 * backward pass finds them (row 0 = pi; rows < nblocks carry the self-transition rescale of
 * ForwardBackward.hpp:115-119). */
int hml_get_rows(hml_t* h, double* rows, uint64_t capacity_rows);

/* ---- multi-GPU: one sequence split into contiguous segments -----------------
 *
 * The reference is single-process; this mode has no counterpart there.  One handle (one process,
 * one GPU) per rank; rank r owns the observations [start_r, start_r + len_r) given by
 * hml_segment_plan (segments are aligned to 4096 observations).  A block belongs to the rank where
 * it starts.  Collectives are NCCL all-gathers of a few hundred bytes on the handle's stream:
 * at load the per-tile Haar sums (so every rank derives the top levels of the transform) and the
 * edge coefficients; per sweep (i) the partial block in front of each rank's first boundary,
 * (ii) one KxK forward operator per rank, (iii) one K->K backward map per rank, (iv) the per-rank
 * statistics, which every rank sums in rank order.  Every rank must make the same sequence of
 * calls with the same model, threshold, seed and flags; all ranks then return identical
 * hml_sweep_out contents (so host-side parameter draws stay in lock-step), identical to what a
 * single handle holding the whole sequence returns (counts, states) or within fp64 rounding
 * (sums, log-likelihood).  Independent sequences need none of this: use one handle each. */

#define HML_UNIQUE_ID_BYTES 128
/* rank 0 creates the id, the caller distributes it (torch.distributed, MPI, a file, ...). */
int hml_comm_unique_id(uint8_t id[HML_UNIQUE_ID_BYTES]);
/* Joins the communicator (collective over all ranks).  libnccl.so.2 is loaded at this point. */
int hml_comm_init(hml_t* h, int rank, int world, const uint8_t id[HML_UNIQUE_ID_BYTES]);
/* Observations of rank `rank` for a sequence of T observations; fails if T < 4096 * world. */
int hml_segment_plan(uint64_t T, int world, int rank, uint64_t* start, uint64_t* len);
/* Collective load: x holds this rank's observations [start, start+len) per hml_segment_plan
 * (host or device memory respectively); T is the length of the whole sequence. */
int hml_load_segment_f32(hml_t* h, const float* x_host, uint64_t len, uint64_t T, float weight_multiplier);
int hml_load_segment_f32_device(hml_t* h, const float* x_dev, uint64_t len, uint64_t T, float weight_multiplier);
/* The same for multivariate data: x holds this rank's len positions x nr_dims values, position-major (hml_load_f32_md).
 * Per dimension the ranks exchange their tile sums (the levels of the transform above 4096 observations); the partial
 * block in front of a rank's first boundary travels with one (sum x, sum x^2) pair per dimension. */
int hml_load_segment_f32_md(hml_t* h, const float* x_host, uint64_t len, uint64_t T, uint32_t nr_dims, float weight_multiplier);
int hml_load_segment_f32_device_md(hml_t* h, const float* x_dev, uint64_t len, uint64_t T, uint32_t nr_dims,
                                   float weight_multiplier);
/* After a sweep in segment mode: this rank's first global block index and the global block count
 * (hml_sweep_out.nblocks is the global count; hml_nr_blocks / hml_get_* stay rank-local, with
 * block starts reported as global positions). */
int hml_segment_info(const hml_t* h, int* rank, int* world, uint64_t* seg_start, uint64_t* seg_len,
                     uint64_t* first_block, uint64_t* global_blocks);

/* All-gather of `bytes` bytes of host memory per rank over the handle's communicator (recv_host: world x bytes, rank
 * order); a plain copy without a communicator.  For the host side of a split sequence: the block lists behind the
 * automatic priors, per-block outputs of recorded iterations.  Collective. */
int hml_comm_allgather(hml_t* h, const void* send_host, uint64_t bytes, void* recv_host);

/* How the per-sweep carries travel: peer mailboxes written over NVLink by one exchange kernel per rank
 * (CUDA IPC mappings made in hml_comm_init), or NCCL all-gathers when peer mapping is unavailable or the
 * environment says HML_EXCHANGE=nccl.  The choice is collective: all ranks use the same transport. */
#define HML_EXCHANGE_NONE 0
#define HML_EXCHANGE_PEER 1
#define HML_EXCHANGE_NCCL 2
int hml_exchange_transport(const hml_t* h, int* transport);

/* ---- measurement ---------------------------------------------------------------------------- */

/* With timing on, every kernel stage of a sweep is bracketed by CUDA events on the context's
 * stream.  hml_get_timing returns the stage names and the milliseconds of the LAST sweep. */
int hml_set_timing(hml_t* h, int on);
int hml_get_timing(hml_t* h, int* nstages, const char** names, float* ms, int capacity);
/* With timing on, CTA 0 of the persistent kernel stamps %globaltimer at the 13 phase borders of a sweep (model, candidate
 * count | barrier | scatter | barrier | statistics + emission + chunk operators | barrier | forward rows + maps | barrier
 * | states + statistics | barrier | final sums + parameter draws | barrier): nanoseconds of the last fused sweep. */
int hml_chain_phase_ns(hml_t* h, uint64_t stamps[16]);
/* Number of kernel launches issued by this handle since creation. */
int hml_launch_count(const hml_t* h, uint64_t* n);
/* Blocks until all work queued on the handle's stream has finished. */
int hml_sync(hml_t* h);
/* The context's cudaStream_t (as void*), so callers can bracket calls with their own CUDA events. */
int hml_get_stream(hml_t* h, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* HAMMLET_B200_H */
