// hammlet_b200 host side — C entry points around the C++ model surface (include/hammlet_host.h).
#include "../../include/hammlet_host.h"

#include <atomic>
#include <memory>
#include <thread>

#include "StateSequence.hpp"

namespace {
std::string g_chain_error;
typedef Statistics<IntegralArray, Normal> S;
typedef Blocks<BreakpointArray> B;
}  // namespace

// Everything main.cpp keeps on its stack between load and the sampling scheme (main.cpp:107-166,338-362).
struct hammlet_chain {
  rng_t rng;
  DeviceSequence sequence;
  S ia;
  B blocks;
  Emissions<S, B> y;
  Mapping mapping;
  Transitions<DirichletVector> A;
  TransitionHyperParam<DirichletParamVector> tau_A;
  Initial<Dirichlet> pi;
  InitialHyperParam<DirichletParam> tau_pi;
  ThetaHyperParam<NormalInverseGammaParam> tau_theta;
  Theta<NormalInverseGamma> theta;
  Records records;
  DeviceMarginals deviceMarginals;
  std::string error;
  size_t K;

  hammlet_chain(hml_t* dev, size_t nrStates, const std::vector<std::vector<real_t>>& priors, real_t trans, real_t selfTrans,
                real_t alphaPi, uint32_t seed)
      : rng(seed),
        sequence(dev),
        ia(sequence, 1),
        blocks(sequence),
        y(ia, blocks),
        mapping(1, nrStates, combinations),
        A(nrStates, rng),
        tau_A(nrStates, trans, selfTrans),
        pi(nrStates, rng),
        tau_pi(nrStates, alphaPi),
        tau_theta(priors),
        theta(tau_theta, 1, combinations, rng),
        records(sequence.size(), "hammlet-", ".csv", nrStates),
        deviceMarginals(sequence, nrStates),
        K(nrStates) {
    records.setMarginalsSink(&deviceMarginals);
    // a leading "P" is implied (main.cpp:393-406)
    theta.sample(tau_theta);
    pi.sample(tau_pi);
    A.sample(tau_A);
  }
};

extern "C" {

const char* hammlet_chain_error(const hammlet_chain* c) { return c ? c->error.c_str() : g_chain_error.c_str(); }

int hammlet_auto_prior(hml_t* dev, float s2, float p, float prior_out[4]) {
  try {
    if (!dev || !prior_out) throw std::runtime_error("NULL argument");
    DeviceSequence seq(dev);
    S ia(seq, 1);
    B blocks(seq);
    Emissions<S, B> y(ia, blocks);
    const std::vector<real_t> v = autoPrior((real_t)s2, (real_t)p, y, seq.noiseStdev());
    for (int i = 0; i < 4; ++i) prior_out[i] = (float)v[i];
    return HML_OK;
  } catch (std::exception& e) {
    g_chain_error = e.what();
    return HML_ERR_ARG;
  }
}

int hammlet_chain_create(hammlet_chain** out, hml_t* dev, int K, const float prior[4], float trans, float self_trans,
                         float alpha_pi, uint32_t seed) {
  if (out) *out = nullptr;
  try {
    if (!out || !dev || !prior) throw std::runtime_error("NULL argument");
    if (K < 2 || K > HML_MAX_STATES) throw std::runtime_error("number of states must be in [2, 32]");
    const std::vector<real_t> one(prior, prior + 4);
    *out = new hammlet_chain(dev, (size_t)K, std::vector<std::vector<real_t>>((size_t)K, one), trans, self_trans, alpha_pi, seed);
    return HML_OK;
  } catch (std::exception& e) {
    g_chain_error = e.what();
    return HML_ERR_ARG;
  }
}

void hammlet_chain_destroy(hammlet_chain* c) { delete c; }

int hammlet_chain_get(hammlet_chain* c, float* mean, float* var, float* A, float* pi) {
  if (!c) return HML_ERR_ARG;
  const std::vector<real_t> pv = c->pi.valueVector();
  for (size_t s = 0; s < c->K; ++s) {
    if (mean) mean[s] = (float)c->theta.value()[s].mean();
    if (var) var[s] = (float)c->theta.value()[s].var();
    if (pi) pi[s] = (float)pv[s];
    if (A)
      for (size_t j = 0; j < c->K; ++j) A[s * c->K + j] = (float)c->A(s, j);
  }
  return HML_OK;
}

int hammlet_chain_set(hammlet_chain* c, const float* mean, const float* var, const float* A, const float* pi) {
  if (!c) return HML_ERR_ARG;
  try {
    for (size_t s = 0; s < c->K; ++s) {
      if (mean && var) c->theta.param(s).setValue((real_t)mean[s], (real_t)var[s]);
      if (pi) c->pi.values()[s] = (real_t)pi[s];
      if (A)
        for (size_t j = 0; j < c->K; ++j) c->A(s, j) = (real_t)A[s * c->K + j];
    }
    return HML_OK;
  } catch (std::exception& e) {
    c->error = e.what();
    return HML_ERR_ARG;
  }
}

static void chain_run(hammlet_chain* c, char method, uint64_t iterations, uint64_t thinning, int dynamic,
                      int use_self_transitions, uint64_t* nblocks_last) {
  if (!dynamic && c->blocks.dirty()) c->y.createBlocks(c->theta);  // "S": freeze the structure of the current theta
  if (method == 'F') {
    StateSequence<ForwardBackward> q(c->rng);
    sampleHMM(c->y, q, c->theta, c->tau_theta, c->A, c->tau_A, c->pi, c->tau_pi, c->mapping, (size_t)iterations,
              (size_t)thinning, c->records, dynamic != 0, use_self_transitions != 0);
  } else if (method == 'M') {
    StateSequence<Mixture> q(c->rng);
    sampleHMM(c->y, q, c->theta, c->tau_theta, c->A, c->tau_A, c->pi, c->tau_pi, c->mapping, (size_t)iterations,
              (size_t)thinning, c->records, dynamic != 0, use_self_transitions != 0);
  } else {
    throw std::runtime_error(std::string("Unknown sampling type ") + method + "!");
  }
  if (nblocks_last) {
    uint64_t gb = 0;
    c->sequence.check(hml_segment_info(c->sequence.handle(), nullptr, nullptr, nullptr, nullptr, nullptr, &gb));
    *nblocks_last = gb;
  }
}

int hammlet_chain_run(hammlet_chain* c, char method, uint64_t iterations, int dynamic, int use_self_transitions,
                      uint64_t* nblocks_last) {
  if (!c) return HML_ERR_ARG;
  try {
    chain_run(c, method, iterations, 0, dynamic, use_self_transitions, nblocks_last);
    return HML_OK;
  } catch (std::exception& e) {
    c->error = e.what();
    return HML_ERR_STATE;
  }
}

int hammlet_chain_run_recorded(hammlet_chain* c, char method, uint64_t iterations, uint64_t thinning, int dynamic,
                               int use_self_transitions, uint64_t* nblocks_last, uint64_t* marginal_segments) {
  if (!c) return HML_ERR_ARG;
  try {
    chain_run(c, method, iterations, thinning, dynamic, use_self_transitions, nblocks_last);
    if (marginal_segments) *marginal_segments = c->records.nrMarginalSegments();
    return HML_OK;
  } catch (std::exception& e) {
    c->error = e.what();
    return HML_ERR_STATE;
  }
}

int hammlet_chain_last_sweep(hammlet_chain* c, uint64_t* nblocks, uint64_t* counts, uint64_t* trans, uint64_t* stat_n) {
  if (!c) return HML_ERR_ARG;
  const DeviceSequence::LastSweep& l = c->sequence.lastSweep;
  if (l.counts.size() != c->K) {
    c->error = "no sweep has been run on this chain";
    return HML_ERR_STATE;
  }
  if (nblocks) *nblocks = l.nblocks;
  for (size_t s = 0; s < c->K; ++s) {
    if (counts) counts[s] = l.counts[s];
    if (stat_n) stat_n[s] = l.statN[s];
    if (trans)
      for (size_t j = 0; j < c->K; ++j) trans[s * c->K + j] = l.trans[s * c->K + j];
  }
  return HML_OK;
}

// Independent sequences: every chain owns its handle, its CUDA stream, its parameters and its RNG,
// so chains never interact; `threads` host threads each take the next unfinished chain (longest sequence first).
// A single chain of 1e4-1e5 blocks is a string of latency-bound kernels; several at a time fill the device.
int hammlet_chains_run(hammlet_chain** chains, int n, int threads, char method, uint64_t iterations, int dynamic,
                       int use_self_transitions) {
  if (!chains || n < 0) return HML_ERR_ARG;
  for (int i = 0; i < n; ++i)
    if (!chains[i]) return HML_ERR_ARG;
  if (n == 0) return HML_OK;
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return chains[a]->sequence.size() > chains[b]->sequence.size(); });
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  std::atomic<int> next(0), failed(0);
  auto work = [&]() {
    for (;;) {
      const int k = next.fetch_add(1);
      if (k >= n) return;
      hammlet_chain* c = chains[order[k]];
      try {
        chain_run(c, method, iterations, 0, dynamic, use_self_transitions, nullptr);
      } catch (std::exception& e) {
        c->error = e.what();
        failed.fetch_add(1);
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work);
  work();
  for (std::thread& t : pool) t.join();
  return failed.load() ? HML_ERR_STATE : HML_OK;
}

int hammlet_chain_save_marginals(hammlet_chain* c, const char* path) {
  if (!c || !path) return HML_ERR_ARG;
  try {
    std::ofstream f;  // a split sequence: every rank takes part in the merge, rank 0 writes the file
    if (c->sequence.rank() == 0) {
      f.open(path);
      if (!f.is_open()) throw std::runtime_error(std::string("Cannot write to file ") + path + "!");
    }
    c->records.saveMarginals(f);
    return HML_OK;
  } catch (std::exception& e) {
    c->error = e.what();
    return HML_ERR_STATE;
  }
}

}  // extern "C"
