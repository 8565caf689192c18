/* hammlet_b200 — C entry points of the C++ host side (hammlet_b200/host/): the Gibbs chain that
 * src/main.cpp of the reference sets up after loading (theta, A, pi, their conjugate hyper-parameters,
 * the shared mt19937) and sampleHMM (HMM.hpp:60-125) driving it.  The per-sweep device work goes
 * through include/hammlet_b200.h; the O(K^2) conjugate updates (Conjugate.hpp:120-205) and parameter
 * draws (Theta.hpp:203-211, Initial.hpp:35-40, Transitions.hpp:75-79, Distribution.hpp:76-139) run
 * here in real_t = float with libstdc++ <random>, as in the reference.  Used by bench.py and by
 * non-C++ callers that want whole runs instead of single sweeps; the hammlet executable uses the
 * same classes directly.  Exported by libhammlet_b200.so. */
#ifndef HAMMLET_HOST_H
#define HAMMLET_HOST_H

#include "hammlet_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hammlet_chain hammlet_chain;

/* Automatic NIG hyper-parameters {alpha, beta, mu0, nu} (AutoPriors.hpp:18-110; main.cpp:348-352)
 * from the loaded sequence (single handle; for a segment-split sequence gather the block lists and
 * use the caller-side formula, see hammlet_b200/gibbs.py). */
int hammlet_auto_prior(hml_t* dev, float s2, float p, float prior_out[4]);

/* K states (univariate, identity mapping), every state with the NIG prior `prior` (main.cpp:354-362),
 * Dirichlet rows with off-diagonal `trans` and diagonal `self_trans` (main.cpp:146-155), Dirichlet
 * initial distribution `alpha_pi` (main.cpp:165-166), RNG seeded like -R (main.cpp:107-108).
 * theta, pi and A are drawn from their priors (main.cpp:393-406).  `dev` stays owned by the caller. */
int hammlet_chain_create(hammlet_chain** out, hml_t* dev, int K, const float prior[4], float trans, float self_trans,
                         float alpha_pi, uint32_t seed);
void hammlet_chain_destroy(hammlet_chain* c);
const char* hammlet_chain_error(const hammlet_chain* c); /* c == NULL: error of hammlet_chain_create */

/* Current theta (mean, var: K each), A (K*K row-major) and pi (K); any pointer may be NULL. */
int hammlet_chain_get(hammlet_chain* c, float* mean, float* var, float* A, float* pi);
int hammlet_chain_set(hammlet_chain* c, const float* mean, const float* var, const float* A, const float* pi);

/* sampleHMM (HMM.hpp:99-121) for `iterations` sweeps without recording: method 'F' (forward-backward)
 * or 'M' (mixture); dynamic != 0 re-derives the blocks from theta each sweep, otherwise the current
 * structure is kept ("S" token, main.cpp:407-414).  Returns the block count of the last sweep. */
int hammlet_chain_run(hammlet_chain* c, char method, uint64_t iterations, int dynamic, int use_self_transitions,
                      uint64_t* nblocks_last);

/* Integer statistics of the chain's most recent sweep as the sampler received them from the device
 * (ForwardBackward.hpp:177-200): block count, occupancy counts[K], transition counts trans[K*K] (incl. the phantom
 * 0 -> q0), observations per emission parameter stat_n[K].  Each of the three sums to the sequence length — the
 * invariant bench.py asserts on the timed run.  Any pointer may be NULL. */
int hammlet_chain_last_sweep(hammlet_chain* c, uint64_t* nblocks, uint64_t* counts, uint64_t* trans, uint64_t* stat_n);

/* The same with recording (HMM.hpp:104-119): every `thinning`-th sweep (thinning > 0) the sampled state sequence joins
 * the chain's state marginals (Records::record -> StateMarginals::addRecord, Records.hpp:155-235,
 * StateMarginals.hpp:51-137), kept in memory.  The device hands over one (size, state) entry per equal-state run
 * (hml_get_segments).  *marginal_segments = segments of the common refinement so far. */
int hammlet_chain_run_recorded(hammlet_chain* c, char method, uint64_t iterations, uint64_t thinning, int dynamic,
                               int use_self_transitions, uint64_t* nblocks_last, uint64_t* marginal_segments);
/* hammlet_chain_run on n independent chains (one per sequence, e.g. one per chromosome: the reference would be
 * started once per sequence), `threads` of them at a time on host threads.  Every chain keeps its own handle, CUDA
 * stream, parameters and RNG stream, so the result of each chain is what hammlet_chain_run alone gives; running
 * several at once lets the latency-bound kernels of one chain overlap the wide kernels of another (3.9x on 24
 * chromosome-length sequences on one B200).  Returns HML_ERR_STATE if any chain failed (hammlet_chain_error tells). */
int hammlet_chains_run(hammlet_chain** chains, int n, int threads, char method, uint64_t iterations, int dynamic,
                       int use_self_transitions);
/* Writes the marginals accumulated so far in the reference's file format (StateMarginals.hpp:268-310). */
int hammlet_chain_save_marginals(hammlet_chain* c, const char* path);

#ifdef __cplusplus
}
#endif
#endif /* HAMMLET_HOST_H */
