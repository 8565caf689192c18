// hammlet_b200 host side — state-sequence samplers and the Gibbs driver.
//
// Same surface as the reference's StateSequence<ForwardBackward> / StateSequence<Mixture>
// (src/StateSequence.hpp, src/StateSequence/ForwardBackward.hpp, src/StateSequence/Mixture.hpp),
// Trellis (src/Trellis.hpp) and sampleHMM (src/HMM.hpp:60-125).  sample() has the reference's
// signature and side effects — it updates tau_theta, tau_A, tau_pi and feeds Records — but the
// forward filter, backward sampling and the statistics pass run on the device in one
// hml_fb_sweep / hml_mix_sweep call; only the O(K^2) conjugate updates happen here.
//
// Randomness.  Default: every sweep draws one 64-bit key from the shared mt19937 and the device
// derives its per-block uniforms from it with Philox (counter-based).  Replay mode
// (setReplay(true)): the per-block 53-bit uniforms themselves are drawn from the shared mt19937
// with std::generate_canonical, in the order the reference consumes them, so a run is
// draw-for-draw comparable with the reference built with the same real_t.
#pragma once

#include "Emissions.hpp"
#include "Records.hpp"

// The trellis lives on the device; this class keeps the reference's name for the forward rows and
// gives read access to them for diagnostics (rows as the backward pass sees them).
class Trellis {
  std::vector<double> mRows;
  size_t mNrStates = 2;

 public:
  Trellis(const Trellis&) = delete;
  Trellis() {}
  void setNrStates(size_t K) { mNrStates = K; }
  size_t size() const { return mNrStates ? mRows.size() / mNrStates : 0; }
  double operator()(size_t t, size_t d) const {
    if (d >= mNrStates) throw std::runtime_error("Trellis dimension index out of bounds!");
    return mRows.at(t * mNrStates + d);
  }
  void clear() { mRows.clear(); }
  void fetch(DeviceSequence& seq, size_t nrBlocks) {
    mRows.resize((nrBlocks + 1) * mNrStates);
    seq.check(hml_get_rows(seq.handle(), mRows.data(), nrBlocks + 1));
  }
};

// StateMarginals on the device (include/hammlet_b200.h: hml_marginals_*): recorded iterations never leave the GPU;
// the marginals file is written from one device-to-host copy at the end, in the format of StateMarginals::save
// (StateMarginals.hpp:268-310: `size TAB c_0 ... TAB c_{S-1}` with S = highest recorded label + 1).
class DeviceMarginals : public MarginalsSink {
  DeviceSequence& mSeq;

 public:
  DeviceMarginals(DeviceSequence& seq, size_t nrStates) : mSeq(seq) {
    mSeq.check(hml_marginals_reset(mSeq.handle(), (int)nrStates));
  }
  void addIteration() override { mSeq.check(hml_marginals_add(mSeq.handle())); }
  size_t nrSegments() override {
    uint64_t n = 0;
    mSeq.check(hml_marginals_info(mSeq.handle(), &n, nullptr, nullptr));
    return n;
  }
  void save(std::ofstream& ofs) override {
    uint64_t n = 0, iterations = 0;
    int K = 0;
    mSeq.check(hml_marginals_info(mSeq.handle(), &n, &iterations, &K));
    std::vector<uint64_t> size(n);
    std::vector<int32_t> counts(n * (size_t)K);
    mSeq.check(hml_marginals_get(mSeq.handle(), size.data(), counts.data(), n));
    size_t nrLabels = 0;  // highest recorded label + 1
    for (uint64_t i = 0; i < n; ++i)
      for (int s = K; s > (int)nrLabels; --s)
        if (counts[i * K + s - 1] != 0) {
          nrLabels = s;
          break;
        }
    std::string line;
    for (uint64_t i = 0; i < n; ++i) {
      uint64_t sum = 0;
      line = std::to_string(size[i]);
      for (size_t s = 0; s < nrLabels; ++s) {
        line += '\t';
        line += std::to_string(counts[i * K + s]);
        sum += counts[i * K + s];
      }
      line += '\n';
      ofs << line;
      if (sum != iterations)
        throw std::runtime_error("Sum of marginals (" + std::to_string(sum) + ") does not match the number of iterations (" +
                                 std::to_string(iterations) + ")!");
    }
    ofs.flush();
  }
};

template <typename Type>
class StateSequence {
  std::vector<marginal_t> mStates;
  std::vector<uint64_t> mRunSize;     // equal-state runs of the last recorded iteration
  std::vector<marginal_t> mRunState;
  rng_t& mRNG;
  Trellis mTrellis;
  bool mReplay = false;
  bool mKeepTrellis = false;
  double mLogLikelihood = NAN;

  template <typename ThetaType, typename TransitionsType, typename InitialType>
  static void fillModel(const ThetaType& theta, const TransitionsType& A, const InitialType& pi, const Mapping& mapping,
                        bool useSelf, std::vector<double>& mean, std::vector<double>& var, std::vector<double>& a,
                        std::vector<double>& p, std::vector<int32_t>& map, hml_model& m) {
    const size_t K = A.nrStates();
    const size_t D = mapping.nrDataDims();
    a.resize(K * K);
    p = std::vector<double>(K);
    const std::vector<real_t> pv = pi.valueVector();
    for (size_t s = 0; s < K; ++s) {
      p[s] = pv[s];
      for (size_t j = 0; j < K; ++j) a[s * K + j] = A(s, j);
    }
    m = hml_model{};
    m.K = (int32_t)K;
    m.use_self_transitions = useSelf ? 1 : 0;
    if (D == 1) {
      // univariate data: the device sees one (mean, var) per state, resolved through the mapping here
      mean.resize(K);
      var.resize(K);
      for (size_t s = 0; s < K; ++s) {
        const auto& param = theta.value()[mapping[s][0]];
        mean[s] = param.mean();
        var[s] = param.var();
      }
    } else {
      // multivariate data: the P shared parameters and the state -> parameter mapping go to the device
      // (Mapping.hpp:89-117); it returns per-parameter statistics (ForwardBackward.hpp:189-191)
      const size_t P = theta.nrParams();
      mean.resize(P);
      var.resize(P);
      for (size_t prm = 0; prm < P; ++prm) {
        mean[prm] = theta.value()[prm].mean();
        var[prm] = theta.value()[prm].var();
      }
      map.resize(K * D);
      for (size_t s = 0; s < K; ++s)
        for (size_t d = 0; d < D; ++d) map[s * D + d] = (int32_t)mapping[s][d];
      m.nr_dims = (int32_t)D;
      m.nr_params = (int32_t)P;
      m.mapping = map.data();
    }
    m.mean = mean.data();
    m.var = var.data();
    m.A = a.data();
    m.pi = p.data();
  }

 public:
  StateSequence(const StateSequence&) = delete;
  StateSequence(rng_t& RNG) : mRNG(RNG) {}

  void setReplay(bool on) { mReplay = on; }
  rng_t& rng() { return mRNG; }
  // forward-backward sampling with Philox uniforms may run whole on the device (sampleHMMOnDevice)
  bool deviceChainAllowed() const { return !kIsMixture && !mReplay && !mKeepTrellis; }
  void setKeepTrellis(bool on) { mKeepTrellis = on; }
  const Trellis& trellis() const { return mTrellis; }
  double logLikelihood() const { return mLogLikelihood; }

  template <typename StatsStructure, typename StatsType, typename BlocksType, typename ThetaType, typename TauThetaType,
            typename TransitionsType, typename TauAType, typename InitialType, typename TauPiType>
  void sample(Emissions<Statistics<StatsStructure, StatsType>, Blocks<BlocksType>>& y, const ThetaType& theta,
              TauThetaType& tau_theta, const TransitionsType& A, TauAType& tau_A, const InitialType& pi, TauPiType& tau_pi,
              const Mapping& mapping, Records& records, const bool doRecord, const bool useSelfTransitions);

  size_t size() const { return mStates.size(); }
  const std::vector<marginal_t>& states() const { return mStates; }
  const marginal_t& operator[](const size_t s) const {
    if (s >= mStates.size()) throw std::runtime_error("State sequence index " + std::to_string(s) + " out of bounds!");
    return mStates[s];
  }
  void clear() {
    std::vector<marginal_t>().swap(mStates);
    mTrellis.clear();
  }

 private:
  static constexpr bool kIsMixture = std::is_same<Type, Mixture>::value;
};

template <typename Type>
template <typename StatsStructure, typename StatsType, typename BlocksType, typename ThetaType, typename TauThetaType,
          typename TransitionsType, typename TauAType, typename InitialType, typename TauPiType>
void StateSequence<Type>::sample(Emissions<Statistics<StatsStructure, StatsType>, Blocks<BlocksType>>& y,
                                 const ThetaType& theta, TauThetaType& tau_theta, const TransitionsType& A, TauAType& tau_A,
                                 const InitialType& pi, TauPiType& tau_pi, const Mapping& mapping, Records& records,
                                 const bool doRecord, const bool useSelfTransitions) {
  const size_t nrStates = A.nrStates();
  const size_t nrParams = tau_theta.nrParams();
  auto& blocks = y.blocks();
  DeviceSequence& seq = blocks.sequence();

  std::vector<double> mean, var, a, p;
  std::vector<int32_t> map;
  hml_model model;
  fillModel(theta, A, pi, mapping, useSelfTransitions, mean, var, a, p, map, model);

  // per-parameter statistics: one per state for univariate data, nrParams with a multivariate mapping
  const size_t nrStatSlots = std::max(nrStates, nrParams);
  std::vector<double> statSum(nrStatSlots), statSq(nrStatSlots);
  std::vector<uint64_t> statN(nrStatSlots), trans(nrStates * nrStates), counts(nrStates);
  hml_sweep_out out;
  out.stat_sum = statSum.data();
  out.stat_sumsq = statSq.data();
  out.stat_n = statN.data();
  out.trans = trans.data();
  out.counts = counts.data();

  uint32_t flags = mKeepTrellis && !kIsMixture ? (HML_SWEEP_KEEP_ROWS | HML_SWEEP_LOGLIK) : 0;
  std::vector<double> uniforms;
  uint64_t key = 0;
  if (mReplay) {
    // one 53-bit uniform per block from the shared stream, as Trellis::sample / discrete_distribution draw them
    const size_t nb = blocks.materialize();
    uniforms.resize(nb);
    for (size_t b = 0; b < nb; ++b) uniforms[b] = std::generate_canonical<double, 53>(mRNG);
  } else {
    key = ((uint64_t)mRNG() << 32) | (uint64_t)mRNG();
    if (blocks.dirty()) flags |= HML_SWEEP_DYNAMIC;
  }
  auto fn = kIsMixture ? hml_mix_sweep : hml_fb_sweep;
  seq.check(fn(seq.handle(), &model, flags, (float)blocks.threshold(), key, 0, mReplay ? uniforms.data() : nullptr,
               uniforms.size(), &out));
  blocks.markBuilt();
  seq.lastSweep.nblocks = out.nblocks;
  seq.lastSweep.counts = counts;
  seq.lastSweep.trans = trans;
  seq.lastSweep.statN.assign(statN.begin(), statN.begin() + nrParams);
  for (uint64_t i = 0; i < out.uniform_fallbacks && seq.rank() == 0; ++i) std::cout << "[WARNING] Uniform sampling of forward variables!" << std::endl;
  mLogLikelihood = out.loglik;
  if (mKeepTrellis && !kIsMixture) {
    mTrellis.setNrStates(nrStates);
    mTrellis.fetch(seq, out.nblocks);
  }

  // ---- posterior updates (ForwardBackward.hpp:203-211, Mixture.hpp:131-139)
  for (size_t prm = 0; prm < nrParams; ++prm) {
    if (statN[prm] > 0) {
      const SufficientStatistics<StatsType> s((real_t)statSum[prm], (real_t)statSq[prm]);
      tau_theta.addObservation(s, statN[prm], prm);
    }
  }
  SufficientStatistics<CategoricalVector> transitions(nrStates);
  SufficientStatistics<Categorical> stateCounts(nrStates);
  for (size_t i = 0; i < nrStates; ++i) {
    stateCounts[i] = counts[i];
    for (size_t j = 0; j < nrStates; ++j) transitions[i][j] = trans[i * nrStates + j];
  }
  tau_A.addObservation(transitions);
  tau_pi.addObservation(stateCounts);

  // ---- records (ForwardBackward.hpp:193-195).  Unless the per-block sizes are written out (-O B), the device
  // merges equal-state neighbours into runs (Records.hpp:166-188 does the same block by block) and one entry per
  // run travels to the host: ~T / mean segment length entries instead of one per block.
  if (doRecord && !mKeepTrellis && records.canRecordOnDevice()) {
    records.recordIterationOnDevice(out.nblocks);  // marginals only: the iteration never leaves the device
    return;
  }
  if (doRecord && !records.wantsBlocks() && !mKeepTrellis) {
    uint64_t nruns = 0;
    seq.check(hml_get_segments(seq.handle(), &nruns, nullptr, nullptr, 0));
    mRunSize.resize(nruns);
    mRunState.resize(nruns);
    seq.check(hml_get_segments(seq.handle(), &nruns, mRunSize.data(), mRunState.data(), nruns));
    for (uint64_t i = 0; i < nruns; ++i) records.recordRun((size_t)mRunState[i], (size_t)mRunSize[i], i == 0 ? out.nblocks : 0);
    return;
  }
  // block by block, in order
  if (doRecord || (!kIsMixture && mKeepTrellis)) {
    uint64_t local = 0;  // a split sequence: the rank's own blocks, then everybody's in rank order
    seq.check(hml_nr_blocks(seq.handle(), &local));
    mStates.resize(local);
    if (local) seq.check(hml_get_states(seq.handle(), mStates.data(), mStates.size()));
    if (seq.split()) mStates = seq.allgatherv(mStates);
  }
  if (doRecord) {
    blocks.fetch(false);
    const std::vector<uint32_t>& st = blocks.starts();
    const size_t T = y.size();
    for (size_t b = 0; b < st.size(); ++b) {
      const size_t N = (b + 1 < st.size() ? st[b + 1] : T) - st[b];
      records.record((size_t)mStates[b], N);
    }
  }
}

// sampleHMM on the device-resident chain (include/hammlet_b200.h: hml_chain_*): the conjugate updates and the draws of
// theta, pi and A happen on the device after every sweep (Philox streams keyed by one 64-bit draw from the shared
// mt19937), a run of sweeps between two recorded iterations is ONE launch of the persistent sweep kernel, and nothing
// crosses PCIe in between.  Forward-backward sampling with dynamic blocks on a single univariate sequence of at most 8
// states whose block structure fits the persistent kernel (64 tiles); recorded iterations need nothing on the host but
// the marginals (and the compression / parameters files).  Returns false — nothing done — where that does not apply:
// -replay, mixture sampling, static blocks, multivariate data, a split sequence, per-iteration sequence / block output.
// HAMMLET_HOST_PARAMS=1 in the environment keeps the host-side parameter draws.
template <typename EmissionsType, typename ThetaType, typename ThetaParamType, typename TransitionType,
          typename TransitionParamType, typename InitialType, typename InitialParamType>
bool sampleHMMOnDevice(EmissionsType& y, rng_t& rng, ThetaType& theta, ThetaParamType& tau_theta, TransitionType& A,
                       TransitionParamType& tau_A, InitialType& pi, InitialParamType& tau_pi, const Mapping& mapping,
                       const size_t iterations, const size_t thinning, Records& records, const bool useSelfTransitions) {
  auto& blocks = y.blocks();
  DeviceSequence& seq = blocks.sequence();
  const size_t K = A.nrStates();
  if (std::getenv("HAMMLET_HOST_PARAMS") || iterations == 0 || seq.split() || seq.nrDim() != 1 || mapping.nrDataDims() != 1 ||
      tau_theta.nrParams() != K || K < 2 || K > 8)
    return false;
  const bool recording = thinning > 0 && thinning <= iterations;
  if (recording && !records.canRecordOnDevice()) return false;
  // one prior for all states, one off-diagonal and one diagonal transition prior, one initial prior (what main.cpp builds)
  const auto& p0 = tau_theta.prior(0);
  for (size_t s = 0; s < K; ++s) {
    const auto& ps = tau_theta.prior(s);
    if (ps.alpha() != p0.alpha() || ps.beta() != p0.beta() || ps.mu0() != p0.mu0() || ps.nu() != p0.nu()) return false;
    if (tau_pi.prior()[s] != tau_pi.prior()[0]) return false;
    for (size_t j = 0; j < K; ++j)
      if (tau_A.prior()[s][j] != (s == j ? tau_A.prior()[0][0] : tau_A.prior()[0][1])) return false;
  }
  // does the block structure of the current parameters fit the persistent kernel?  (one detection pass; a chain whose
  // structure outgrows it later still runs: those sweeps take the multi-kernel path, parameter phase on the device)
  y.createBlocks(theta);
  if (blocks.materialize() > (size_t)60000) return false;
  const float prior[4] = {(float)p0.alpha(), (float)p0.beta(), (float)p0.mu0(), (float)p0.nu()};
  const uint64_t seed = ((uint64_t)rng() << 32) | (uint64_t)rng();
  seq.check(hml_chain_init(seq.handle(), (int)K, prior, (float)tau_A.prior()[0][1], (float)tau_A.prior()[0][0],
                           (float)tau_pi.prior()[0], seed, useSelfTransitions ? 1 : 0));
  std::vector<double> mean(K), var(K), a(K * K), p(K);
  const std::vector<real_t> pv = pi.valueVector();
  for (size_t s = 0; s < K; ++s) {
    mean[s] = theta.value()[s].mean();
    var[s] = theta.value()[s].var();
    p[s] = pv[s];
    for (size_t j = 0; j < K; ++j) a[s * K + j] = A(s, j);
  }
  seq.check(hml_chain_set(seq.handle(), mean.data(), var.data(), a.data(), p.data()));
  auto pull = [&]() {  // theta, pi, A as they stand on the device
    seq.check(hml_chain_get(seq.handle(), mean.data(), var.data(), a.data(), p.data(), nullptr, nullptr));
    for (size_t s = 0; s < K; ++s) {
      theta.param(s).setValue((real_t)mean[s], (real_t)var[s]);
      pi.values()[s] = (real_t)p[s];
      for (size_t j = 0; j < K; ++j) A(s, j) = (real_t)a[s * K + j];
    }
  };
  if (thinning > iterations)
    std::cout << "[WARNING] Thinning parameter is larger than number of iterations. No data will be recorded!" << std::endl;
  size_t done = 0;
  std::vector<double> statSum(K), statSq(K);
  std::vector<uint64_t> statN(K), trans(K * K), counts(K);
  while (done < iterations) {
    const size_t batch = recording ? std::min(thinning, iterations - done) : iterations - done;
    hml_sweep_out last{};
    last.stat_sum = statSum.data();
    last.stat_sumsq = statSq.data();
    last.stat_n = statN.data();
    last.trans = trans.data();
    last.counts = counts.data();
    seq.check(hml_chain_run(seq.handle(), batch, nullptr, &last));
    done += batch;
    seq.lastSweep.nblocks = last.nblocks;
    seq.lastSweep.counts = counts;
    seq.lastSweep.trans = trans;
    seq.lastSweep.statN = statN;
    if (recording && done % thinning == 0) {  // HMM.hpp:104-119
      records.recordIterationOnDevice(last.nblocks);
      pull();
      records.record(theta);
    }
  }
  pull();
  blocks.createBlocks(theta);  // the host's view: threshold of the final parameters, structure to be rebuilt on use
  return true;
}

// The Gibbs driver, reference: HMM.hpp:60-125.
template <typename StateSequenceType, typename EmissionsType, typename ThetaType, typename ThetaParamType,
          typename TransitionType, typename TransitionParamType, typename InitialType, typename InitialParamType>
void sampleHMM(EmissionsType& y, StateSequenceType& q, ThetaType& theta, ThetaParamType& tau_theta, TransitionType& A,
               TransitionParamType& tau_A, InitialType& pi, InitialParamType& tau_pi, const Mapping& mapping,
               const size_t iterations, const size_t thinning, Records& records, const bool dynamic = true,
               const bool useSelfTransitions = true) {
  if (dynamic && q.deviceChainAllowed() &&
      sampleHMMOnDevice(y, q.rng(), theta, tau_theta, A, tau_A, pi, tau_pi, mapping, iterations, thinning, records,
                        useSelfTransitions))
    return;
  if (thinning > iterations)
    std::cout << "[WARNING] Thinning parameter is larger than number of iterations. No data will be recorded!" << std::endl;
  for (size_t i = 0; i < iterations; ++i) {
    if (dynamic) y.createBlocks(theta);
    const bool doRecord = thinning > 0 && ((i + 1) % thinning == 0);
    q.sample(y, theta, tau_theta, A, tau_A, pi, tau_pi, mapping, records, doRecord, useSelfTransitions);
    theta.sample(tau_theta);
    pi.sample(tau_pi);
    A.sample(tau_A);
    if (doRecord) records.record(theta);
  }
}
