// hammlet_b200 host side — model parameters and conjugate algebra (O(K^2) per sweep, stays on the host).
//
// Same type surface as the reference: tag classes (src/Tags.hpp), Observation<> value types
// (src/Observation.hpp), SufficientStatistics<> (src/SufficientStatistics.hpp), Conjugate<>
// (src/Conjugate.hpp), Distribution<> (src/Distribution.hpp), Mapping (src/Mapping.hpp), Theta /
// ThetaHyperParam (src/Theta.hpp, src/ThetaHyperParam.hpp), Transitions (src/Transitions.hpp), Initial
// (src/Initial.hpp).  Only the members the FBG / mixture path uses are provided.  Arithmetic follows the
// reference expression by expression (mixed real_t/double promotions included) because it decides the
// sampled parameters; draws use libstdc++ <random> on the shared mt19937, in the reference's order.
#pragma once

#include "Parser.hpp"
#include "base.hpp"

// ---------------------------------------------------------------------------------------- tags
enum MappingType { combinations, independent };
class Normal {};
class NormalParam {};
using NormalInverseGamma = NormalParam;
class NormalInverseGammaParam {};
class Categorical {};
class CategoricalParam {};
using Dirichlet = CategoricalParam;
class DirichletParam {};
class CategoricalVector {};
class CategoricalParamVector {};
using DirichletVector = CategoricalParamVector;
class DirichletParamVector {};
class ForwardBackward {};
class Mixture {};
class IntegralArray {};
class BreakpointArray {};

template <typename T> class Observation;
template <typename T> class SufficientStatistics;
template <typename T> class Distribution;
template <typename T> class Conjugate;

template <>
inline MappingType convertType(const std::string& s) {
  if (s == "combinations" || s == "C") return combinations;
  throw std::runtime_error("Unknown mapping type " + s + "!");
}

// ---------------------------------------------------------------------------------------- observations

template <>
class Observation<NormalParam> {  // (mean, variance) of one emission distribution
  real_t mMean = NAN, mVar = NAN, mStdev = NAN, mPrec = NAN;

 public:
  Observation() {}
  Observation(real_t mean, real_t var) { setValue(mean, var); }
  void setValue(real_t mean, real_t var) {
    if (!std::isfinite(mean)) throw std::runtime_error("Mean (" + std::to_string(mean) + ") must be set to a finite value!");
    mMean = mean;
    if (!std::isfinite(var)) throw std::runtime_error("Variance(" + std::to_string(var) + ") must be set to a finite value!");
    if (var <= 0) throw std::runtime_error("Variance (" + std::to_string(var) + ") must be positive!");
    mVar = var;
    mPrec = 1 / var;
    mStdev = std::sqrt(var);
  }
  real_t mean() const { return mMean; }
  real_t var() const { return mVar; }
  real_t stdev() const { return mStdev; }
  real_t prec() const { return mPrec; }
  size_t nrDim() const { return 1; }
  std::string str(const std::string sep = "\t", const std::string finalSep = "") const {
    return std::to_string(mMean) + sep + std::to_string(mVar) + finalSep;
  }
  friend std::ostream& operator<<(std::ostream& o, const Observation<NormalParam>& d) { return o << d.str(); }
};

template <>
class Observation<NormalInverseGammaParam> {
  real_t mAlpha, mBeta, mMu0, mNu;

 public:
  Observation() : mAlpha(NAN), mBeta(NAN), mMu0(NAN), mNu(NAN) {}
  Observation(const std::vector<real_t>& v) {
    if (v.size() != 4) throw std::runtime_error("Parameter vector for Normal-Inverse Gamma must have 4 elements!");
    setValue(v[0], v[1], v[2], v[3]);
  }
  Observation(real_t alpha, real_t beta, real_t mu0, real_t nu) { setValue(alpha, beta, mu0, nu); }
  void setValue(real_t alpha, real_t beta, real_t mu0, real_t nu) {
    if (alpha <= 0) throw std::runtime_error("Alpha (" + std::to_string(alpha) + ") must be positive!");
    if (beta <= 0) throw std::runtime_error("Beta (" + std::to_string(beta) + ") must be positive!");
    if (nu <= 0) throw std::runtime_error("Nu (" + std::to_string(nu) + ")must be positive!");
    if (!std::isfinite(mu0)) throw std::runtime_error("Mu0 (" + std::to_string(mu0) + ")  must be finite!");
    mAlpha = alpha;
    mBeta = beta;
    mMu0 = mu0;
    mNu = nu;
  }
  real_t alpha() const { return mAlpha; }
  real_t beta() const { return mBeta; }
  real_t mu0() const { return mMu0; }
  real_t nu() const { return mNu; }
  std::string str(const std::string sep = "\t", const std::string finalSep = "") const {
    return std::to_string(mAlpha) + sep + std::to_string(mBeta) + sep + std::to_string(mMu0) + sep + std::to_string(mNu) + finalSep;
  }
  friend std::ostream& operator<<(std::ostream& o, const Observation<NormalInverseGammaParam>& d) { return o << d.str(); }
};

template <>
class Observation<Dirichlet> {  // a probability vector
  std::vector<real_t> mProbs;

 public:
  Observation() {}
  explicit Observation(size_t size) : mProbs(size, NAN) {}
  Observation(const std::vector<real_t>& v) : mProbs(v) {}
  size_t domainSize() const { return mProbs.size(); }
  real_t& operator[](size_t i) { return mProbs.at(i); }
  const real_t& operator[](size_t i) const { return mProbs.at(i); }
  const std::vector<real_t>& probs() const { return mProbs; }
  std::vector<real_t>& probs() { return mProbs; }
  std::string str(const std::string sep = "\t", const std::string finalSep = "") const {
    return hammlet::concat(mProbs, sep, finalSep);
  }
  friend std::ostream& operator<<(std::ostream& o, const Observation<Dirichlet>& d) { return o << d.str(); }
};

template <>
class Observation<DirichletVector> {  // rows of a transition matrix
  std::vector<Observation<Dirichlet>> mRows;

 public:
  Observation() {}
  explicit Observation(size_t size) : mRows(size, Observation<Dirichlet>(size)) {}
  size_t nrDim() const { return mRows.size(); }
  Observation<Dirichlet>& operator[](size_t i) { return mRows.at(i); }
  const Observation<Dirichlet>& operator[](size_t i) const { return mRows.at(i); }
  const real_t& operator()(size_t from, size_t to) const {
    if (from >= mRows.size()) throw std::runtime_error("Dirichlet vector row index out of bounds!");
    if (to >= mRows[from].domainSize()) throw std::runtime_error("Dirichlet vector column index out of bounds!");
    return mRows[from][to];
  }
  real_t& operator()(size_t from, size_t to) {
    return const_cast<real_t&>(static_cast<const Observation<DirichletVector>&>(*this)(from, to));
  }
  std::string str(const std::string sep = "\t", const std::string finalSep = "") const {
    return hammlet::concat(mRows, sep, finalSep);
  }
};

template <>
class Observation<DirichletParam> {  // concentration parameters
  std::vector<real_t> mAlphas;

 public:
  Observation() {}
  Observation(size_t size, real_t value) {
    if (value <= 0) throw std::runtime_error("All values in Dirichlet parameters must be greater than zero!");
    mAlphas.assign(size, value);
  }
  Observation(size_t size, real_t value, size_t specialIndex, real_t specialValue) {
    if (specialIndex >= size) throw std::runtime_error("Special index out of bounds for Dirichlet parameter.");
    if (value <= 0 || specialValue <= 0)
      throw std::runtime_error("All values in Dirichlet parameters must be greater than zero!");
    mAlphas.assign(size, value);
    mAlphas[specialIndex] = specialValue;
  }
  size_t domainSize() const { return mAlphas.size(); }
  const std::vector<real_t>& alphas() const { return mAlphas; }
  real_t& operator[](size_t i) { return mAlphas.at(i); }
  const real_t& operator[](size_t i) const { return mAlphas.at(i); }
  std::string str(const std::string sep = "\t", const std::string finalSep = "") const {
    return hammlet::concat(mAlphas, sep, finalSep);
  }
  friend std::ostream& operator<<(std::ostream& o, const Observation<DirichletParam>& d) { return o << d.str(); }
};

template <>
class Observation<DirichletParamVector> {
  std::vector<Observation<DirichletParam>> mRows;

 public:
  Observation() {}
  // off-diagonal entries `value`, diagonal `diagonalValue` (main.cpp:154-155: `-t trans self`)
  Observation(size_t size, real_t value, real_t diagonalValue) {
    for (size_t r = 0; r < size; ++r) mRows.push_back(Observation<DirichletParam>(size, value, r, diagonalValue));
  }
  size_t nrDim() const { return mRows.size(); }
  Observation<DirichletParam>& operator[](size_t i) { return mRows.at(i); }
  const Observation<DirichletParam>& operator[](size_t i) const { return mRows.at(i); }
  std::string str(const std::string sep = "\t", const std::string finalSep = "") const {
    return hammlet::concat(mRows, sep, finalSep);
  }
};

// ---------------------------------------------------------------------------------------- sufficient statistics

template <>
class SufficientStatistics<Normal> {  // (sum x, sum x^2)
  real_t mSum = 0, mSumSq = 0;

 public:
  SufficientStatistics() {}
  SufficientStatistics(real_t singleValue) : mSum(singleValue), mSumSq(singleValue * singleValue) {}
  SufficientStatistics(real_t sum, real_t sumSq) : mSum(sum), mSumSq(sumSq) {}
  void addObs(real_t x) {
    mSum += x;
    mSumSq += x * x;
  }
  real_t sum() const { return mSum; }
  real_t sumSq() const { return mSumSq; }
  size_t nrDim() const { return 1; }
  void clear() { mSum = mSumSq = 0; }
};

template <>
class SufficientStatistics<Categorical> {  // counts per category
  std::vector<size_t> mCounts;

 public:
  explicit SufficientStatistics(size_t domainsize) : mCounts(domainsize, 0) {}
  size_t domainSize() const { return mCounts.size(); }
  size_t& operator[](size_t i) { return mCounts[i]; }
  const size_t& operator[](size_t i) const { return mCounts[i]; }
  void clear() { mCounts.assign(mCounts.size(), 0); }
};

template <>
class SufficientStatistics<CategoricalVector> {  // K x K transition counts
  std::vector<SufficientStatistics<Categorical>> mCounts;

 public:
  explicit SufficientStatistics(size_t nrdim, size_t domainsize = 0) {
    mCounts.assign(nrdim, SufficientStatistics<Categorical>(domainsize ? domainsize : nrdim));
  }
  size_t nrDim() const { return mCounts.size(); }
  SufficientStatistics<Categorical>& operator[](int i) { return mCounts[i]; }
  const SufficientStatistics<Categorical>& operator[](int i) const { return mCounts[i]; }
  void clear() {
    for (auto& c : mCounts) c.clear();
  }
};

inline real_t sampleMean(const SufficientStatistics<Normal>& s, size_t N) {
  if (N <= 0) throw std::runtime_error("Cannot calculate mean from zero observations!");
  const double n = N;
  return s.sum() / n;
}
inline real_t sampleVariance(const SufficientStatistics<Normal>& s, size_t N) {
  if (N <= 0) throw std::runtime_error("Cannot calculate variance from zero observations!");
  const double n = N;
  const double avg = sampleMean(s, N);
  return s.sumSq() / n - (avg * avg);
}

// EFD.hpp:35-38
inline real_t logNormalizer(const Observation<NormalParam>& p) {
  return std::log(p.stdev()) + p.mean() * p.mean() / (2 * p.var());
}

// ---------------------------------------------------------------------------------------- conjugate pairs

template <typename ParamType>
class Conjugate {
  Observation<ParamType> mPrior, mPosterior;

 public:
  template <typename... Types>
  Conjugate(Types... args) : mPrior(args...), mPosterior(args...) {}
  Conjugate(Observation<ParamType> prior) : mPrior(prior), mPosterior(prior) {}
  const Observation<ParamType>& prior() const { return mPrior; }
  const Observation<ParamType>& posterior() const { return mPosterior; }
  Observation<ParamType>& posterior() { return mPosterior; }
  void reset() { mPosterior = mPrior; }

  template <typename ObsType>
  inline void addObservation(const SufficientStatistics<ObsType>& obs);
  template <typename ObsType>
  inline void addObservation(const SufficientStatistics<ObsType>& obs, const size_t N);
};

// Normal-Inverse-Gamma update from (sum, sum of squares, count), reference: Conjugate.hpp:120-168
template <>
template <>
inline void Conjugate<NormalInverseGammaParam>::addObservation(const SufficientStatistics<Normal>& obs, const size_t counts) {
  const real_t sum = obs.sum();
  const real_t sumSq = obs.sumSq();
  if (counts == 0) {
    if (sumSq > 0) throw std::runtime_error("Sufficient statistics contain values, but no observation count!");
    std::cout << "[WARNING] No observation count for sufficient statistics, there might be an index error!" << std::endl;
    return;
  }
  if (sumSq < 0)
    throw std::runtime_error("Sum of squares is negative (" + std::to_string(sumSq) + ") for " + std::to_string(counts) +
                             " observations!");
  const double N = (double)counts;
  const real_t xbar = sum / N;
  const real_t alpha = mPosterior.alpha(), beta = mPosterior.beta(), mu0 = mPosterior.mu0(), nu = mPosterior.nu();
  real_t ssN = (sum * sum) / N;  // can exceed sumSq by rounding: clamp so the variance term stays >= 0
  if (ssN > sumSq) ssN = sumSq;
  mPosterior.setValue(alpha + N / 2.0,
                      beta + ((sumSq + (N * nu / (N + nu)) * ((xbar - mu0) * (xbar - mu0))) - ssN) / 2.0,
                      (nu * mu0 + sum) / (nu + N), nu + N);
}

// Dirichlet updates, reference: Conjugate.hpp:177-205
template <>
template <>
inline void Conjugate<DirichletParamVector>::addObservation(const SufficientStatistics<CategoricalVector>& countMatrix) {
  if (countMatrix.nrDim() != mPosterior.nrDim())
    throw std::runtime_error("Dimensions of count matrix (" + std::to_string(countMatrix.nrDim()) +
                             ") and posterior observations (" + std::to_string(mPosterior.nrDim()) + ") do not match!");
  for (size_t d = 0; d < countMatrix.nrDim(); ++d)
    for (size_t c = 0; c < countMatrix[d].domainSize(); ++c) mPosterior[d][c] += countMatrix[d][c];
}
template <>
template <>
inline void Conjugate<DirichletParam>::addObservation(const SufficientStatistics<Categorical>& obs) {
  if (obs.domainSize() != mPosterior.domainSize())
    throw std::runtime_error("Domain size of observations (" + std::to_string(obs.domainSize()) +
                             ") does not match that of posterior (" + std::to_string(mPosterior.domainSize()) + ")!");
  for (size_t i = 0; i < mPosterior.domainSize(); ++i) mPosterior[i] += obs[i];
}

template <typename ParamType>
using TransitionHyperParam = Conjugate<ParamType>;
template <typename ParamType>
using InitialHyperParam = Conjugate<ParamType>;

// ---------------------------------------------------------------------------------------- distributions

template <typename DistType>
class Distribution {
  rng_t& mRNG;

 public:
  Distribution(rng_t& RNG) : mRNG(RNG) {}
  template <typename ParamType>
  void resample(Observation<DistType>& obs, const Observation<ParamType>& param);
};

// var = 1 / Gamma(alpha, 1/beta); mean ~ N(mu0, sqrt(var / nu))   (Distribution.hpp:76-87)
template <>
template <>
inline void Distribution<NormalInverseGamma>::resample(Observation<NormalInverseGamma>& obs,
                                                       const Observation<NormalInverseGammaParam>& param) {
  std::gamma_distribution<real_t> gamma(param.alpha(), 1.0 / param.beta());
  real_t var = 1.0 / gamma(mRNG);
  std::normal_distribution<real_t> normal(param.mu0(), std::sqrt(var / param.nu()));
  real_t mean = normal(mRNG);
  obs.setValue(mean, var);
}

// Dirichlet via normalised Gamma(alpha_i, 1) draws (Distribution.hpp:116-139)
inline void dirichlet_sample(std::vector<real_t>& probs, const std::vector<real_t>& alphas, rng_t& RNG) {
  if (probs.size() != alphas.size())
    throw std::runtime_error("Number of parameters must match the domain size of the Dirichlet RV!");
  real_t sum = 0;
  for (size_t d = 0; d < alphas.size(); ++d) {
    std::gamma_distribution<real_t> dist(alphas[d], 1.0);
    const real_t g = dist(RNG);
    probs[d] = g;
    sum += g;
  }
  for (auto& p : probs) p /= sum;
}
template <>
template <>
inline void Distribution<Dirichlet>::resample(Observation<Dirichlet>& obs, const Observation<DirichletParam>& param) {
  if (obs.domainSize() != param.domainSize())
    throw std::runtime_error("Domain sizes of Dirichlet random variable (" + std::to_string(obs.domainSize()) +
                             ") and the parameters requested for sampling (" + std::to_string(param.domainSize()) +
                             ") do not match!");
  dirichlet_sample(obs.probs(), param.alphas(), mRNG);
}
template <>
template <>
inline void Distribution<DirichletVector>::resample(Observation<DirichletVector>& obs,
                                                    const Observation<DirichletParamVector>& param) {
  if (obs.nrDim() != param.nrDim())
    throw std::runtime_error("Dimensions of Dirichlet random variable (" + std::to_string(obs.nrDim()) +
                             ") and the parameters requested for sampling (" + std::to_string(param.nrDim()) + ") do not match!");
  for (size_t d = 0; d < obs.nrDim(); ++d) dirichlet_sample(obs[d].probs(), param[d].alphas(), mRNG);
}

// ---------------------------------------------------------------------------------------- mapping

inline size_t nrOfStates(size_t nrDataDim, size_t nrParam, MappingType mappingType) {
  if (mappingType != combinations) throw std::runtime_error("Mapping type not implemented!");
  const size_t result = (size_t)std::pow(nrParam, nrDataDim);
  if (result <= 1) throw std::runtime_error("Requested parameters would yield an HMM with less than 2 states!");
  return result;
}

// Mapping[s][d] = emission parameter used by state s in data dimension d (Mapping.hpp:53-137)
class Mapping {
  std::vector<std::vector<size_t>> mValue;
  size_t mNrDataDim, mNrParams, mNrStates;

 public:
  Mapping(size_t nrdatadim, size_t nrparams, MappingType mappingType)
      : mNrDataDim(nrdatadim), mNrParams(nrparams), mNrStates(nrOfStates(nrdatadim, nrparams, mappingType)) {
    if (mNrDataDim <= 0) throw std::runtime_error("Number of data dimensions must be positive!");
    if (mNrParams <= 0) throw std::runtime_error("Number of parameters must be positive!");
    for (size_t x = 0; x < mNrStates; ++x) {  // states = nrParams-ary numbers, least significant digit first
      std::vector<size_t> digits;
      for (size_t d = 0, n = x; d < nrdatadim; ++d, n /= mNrParams) digits.push_back(n % mNrParams);
      mValue.push_back(digits);
    }
  }
  const std::vector<size_t>& operator[](size_t state) const { return mValue[state]; }
  size_t nrStates() const { return mNrStates; }
  size_t nrParams() const { return mNrParams; }
  size_t nrDataDims() const { return mNrDataDim; }
};

// ---------------------------------------------------------------------------------------- theta, A, pi

template <typename ParamType>  // NormalInverseGammaParam
class ThetaHyperParam {
  std::vector<Conjugate<ParamType>> mParams;

 public:
  ThetaHyperParam(const std::vector<std::vector<real_t>>& hyperparams) {
    if (hyperparams.empty())
      throw std::runtime_error("Number of emission hyperparameters must be positive! Did you forget to provide them, or to use -a?");
    for (const auto& hp : hyperparams) mParams.push_back(Conjugate<ParamType>(Observation<ParamType>(hp)));
  }
  size_t nrParams() const { return mParams.size(); }
  template <typename EmissionsType>
  void addObservation(const SufficientStatistics<EmissionsType>& suffStat, const size_t N, const size_t dim) {
    mParams[dim].addObservation(suffStat, N);
  }
  const Observation<ParamType>& posterior(size_t d) const { return mParams[d].posterior(); }
  const Observation<ParamType>& prior(size_t d) const { return mParams[d].prior(); }
  void reset() {
    for (auto& p : mParams) p.reset();
  }
};

template <typename ParamType>  // NormalInverseGamma (= NormalParam)
class Theta {
  size_t mNrDataDim;
  std::vector<Observation<ParamType>> mParams;
  Mapping mMapping;
  Distribution<ParamType> mDist;

 public:
  Theta(const Theta&) = delete;
  template <typename HyperParamType>
  Theta(ThetaHyperParam<HyperParamType>& tau_theta, size_t nrdatadim, MappingType mappingType, rng_t& RNG)
      : mNrDataDim(nrdatadim), mParams(tau_theta.nrParams()), mMapping(nrdatadim, tau_theta.nrParams(), mappingType), mDist(RNG) {
    sample(tau_theta);  // initialised by a draw from the prior (Theta.hpp:126-127)
  }
  real_t logNormalizer(size_t state) const {
    real_t result = 0;
    for (size_t p : mMapping[state]) result += ::logNormalizer(mParams[p]);
    return result;
  }
  const std::vector<Observation<ParamType>>& value() const { return mParams; }
  size_t nrDataDim() const { return mNrDataDim; }
  size_t nrStates() const { return mMapping.nrStates(); }
  size_t nrParams() const { return mMapping.nrParams(); }
  const std::vector<size_t>& mapping(size_t state) const { return mMapping[state]; }
  // draw every parameter from its posterior, then posterior <- prior (Theta.hpp:203-211)
  template <typename ThetaParamType>
  void sample(ThetaHyperParam<ThetaParamType>& tau_theta) {
    for (size_t d = 0; d < mParams.size(); ++d) mDist.resample(mParams[d], tau_theta.posterior(d));
    tau_theta.reset();
  }
  // smallest emission variance: drives the wavelet threshold (Theta.hpp:226-234)
  real_t thresholdValue() const {
    real_t result = hammlet::inf;
    for (const auto& p : mParams) result = std::min(result, p.var());
    return result;
  }
  std::string str(const std::string& sep = "\t", const std::string& finalSep = "") const {
    return hammlet::concat(mParams, sep, finalSep);
  }
  // direct access for tests / replay harnesses
  Observation<ParamType>& param(size_t p) { return mParams[p]; }
};
template <typename ParamType>
std::ostream& operator<<(std::ostream& o, const Theta<ParamType>& t) {
  return o << t.str();
}

template <typename DistType>  // DirichletVector
class Transitions {
  size_t mNrStates;
  Distribution<DistType> mDist;
  Observation<DistType> mValue;

 public:
  Transitions(const Transitions&) = delete;
  Transitions(size_t nrStates, rng_t& RNG) : mNrStates(nrStates), mDist(RNG), mValue(nrStates) {}
  const real_t& operator()(size_t from, size_t to) const { return mValue(from, to); }
  real_t& operator()(size_t from, size_t to) { return mValue(from, to); }
  size_t nrStates() const { return mNrStates; }
  std::string str() const { return mValue.str(); }
  template <typename TransitionParamType>
  void sample(TransitionHyperParam<TransitionParamType>& tau_A) {
    mDist.resample(mValue, tau_A.posterior());
    tau_A.reset();
  }
};

template <typename DistType>  // Dirichlet
class Initial {
  Observation<DistType> mValue;
  Distribution<DistType> mDist;

 public:
  Initial(const Initial&) = delete;
  Initial(size_t nrStates, rng_t& RNG) : mValue(nrStates), mDist(RNG) {}
  template <typename InitialHyperParamType>
  void sample(InitialHyperParamType& tau_pi) {
    mDist.resample(mValue, tau_pi.posterior());
    tau_pi.reset();
  }
  std::vector<real_t> valueVector() const { return mValue.probs(); }
  std::vector<real_t>& values() { return mValue.probs(); }
  size_t nrStates() const { return mValue.domainSize(); }
  std::string str() const { return mValue.str(); }
};
