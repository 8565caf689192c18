// hammlet_b200 host side — compressed emissions: the data of one sequence resident on one B200.
//
// Mirrors the reference's layer L2 (src/Emissions.hpp, src/Blocks/BreakpointArray.hpp,
// src/Statistics/IntegralArray.hpp, src/wavelet.hpp, src/AutoPriors.hpp).  The reference builds host
// vectors (maxlet coefficients -> breakpoint weights, integral arrays, skip pointers) and iterates
// blocks on the CPU; here DeviceSequence::load hands the parsed values to hml_load_f32 once, and
// Blocks / Statistics / Emissions are views onto that device-resident sequence:
//   * Blocks::createBlocks(thr) only records the fp32 threshold; the samplers pass it to the device,
//     which derives the block structure inside the sweep (no host round trip);
//   * the iterator protocol (initForward / next / start / end / blockSize / suffStat / nrBlocks) is kept
//     for callers that walk blocks on the host (autoPrior does): it materialises the block list through
//     hml_create_blocks + hml_get_blocks on first use.
// Differences in construction are listed in INTEGRATION.md.
#pragma once

#include <cstdlib>
#include <memory>

#include "../../include/hammlet_b200.h"
#include "FastParse.hpp"
#include "Model.hpp"

// Thin RAII owner of an hml_t; every C-ABI failure becomes std::runtime_error, like every reference error.
class DeviceSequence {
  hml_t* mHandle = nullptr;
  bool mOwns = true;
  uint64_t mSize = 0;
  size_t mNrDim = 1;
  double mSigmaHat = 0;
  int mRank = 0, mWorld = 1;  // > 1: this handle holds one segment of a sequence split over several GPUs

 public:
  DeviceSequence(const DeviceSequence&) = delete;
  explicit DeviceSequence(int device = 0) {
    if (hml_create(&mHandle, device) != HML_OK) throw std::runtime_error(hml_last_error(nullptr));
  }
  // view of a handle that the caller created and loaded (and keeps owning); for a segment of a split
  // sequence size() is the length of the whole sequence, which is what the threshold depends on
  explicit DeviceSequence(hml_t* loaded) : mHandle(loaded), mOwns(false) {
    if (!loaded) throw std::runtime_error("NULL device handle");
    check(hml_size(mHandle, &mSize));
    check(hml_sigma_hat(mHandle, &mSigmaHat));
    uint32_t d = 1;
    check(hml_nr_dims(mHandle, &d));
    mNrDim = d;
    check(hml_segment_info(mHandle, &mRank, &mWorld, nullptr, nullptr, nullptr, nullptr));
  }
  ~DeviceSequence() {
    if (mOwns) hml_destroy(mHandle);
  }
  void check(int rc) const {
    if (rc != HML_OK) throw std::runtime_error(hml_last_error(mHandle));
  }
  hml_t* handle() const { return mHandle; }
  size_t size() const { return mSize; }     // positions
  size_t nrDim() const { return mNrDim; }   // values per position

  // MaxletTransform + HaarBreakpointWeights + weight multiplier + integral arrays, on the device
  // (wavelet.hpp:97-188, :68-93; main.cpp:332-334; IntegralArray.hpp:136-191).  `values` holds the stream as read:
  // nrDim values per position (wavelet.hpp:131-136).
  void load(const std::vector<float>& values, float weightMultiplier, size_t nrDim = 1) {
    if (values.empty()) throw std::runtime_error("Input vector for breakpoint weights is empty!");
    if (values.size() % nrDim != 0)
      throw std::runtime_error("Input stream did not contain enough values to fill all dimensions at last position!");
    if (nrDim == 1)
      check(hml_load_f32(mHandle, values.data(), values.size(), weightMultiplier));
    else
      check(hml_load_f32_md(mHandle, values.data(), values.size() / nrDim, (uint32_t)nrDim, weightMultiplier));
    mSize = values.size() / nrDim;
    mNrDim = nrDim;
    check(hml_sigma_hat(mHandle, &mSigmaHat));
  }
  // noise estimate from the finest detail coefficients (main.cpp:303-311)
  double noiseStdev() const { return mSigmaHat; }

  // ---- one sequence split into contiguous segments, one per GPU and process (the reference has no
  // counterpart).  Every rank runs the same host code on the same parameters; only rank 0 writes files.
  int rank() const { return mRank; }
  int world() const { return mWorld; }
  bool split() const { return mWorld > 1; }
  // joins the communicator (collective) and loads this rank's slice of `values` (the whole sequence as read: nrDim
  // values per position)
  void loadSegment(const std::vector<float>& values, float weightMultiplier, int rank, int world,
                   const uint8_t id[HML_UNIQUE_ID_BYTES], size_t nrDim = 1) {
    if (values.empty()) throw std::runtime_error("Input vector for breakpoint weights is empty!");
    if (values.size() % nrDim != 0)
      throw std::runtime_error("Input stream did not contain enough values to fill all dimensions at last position!");
    const uint64_t T = values.size() / nrDim;
    check(hml_comm_init(mHandle, rank, world, id));
    uint64_t start = 0, len = 0;
    if (hml_segment_plan(T, world, rank, &start, &len) != HML_OK)
      throw std::runtime_error("Sequence too short to split over " + std::to_string(world) + " devices (4096 observations each at least)!");
    check(hml_load_segment_f32_md(mHandle, values.data() + start * nrDim, len, T, (uint32_t)nrDim, weightMultiplier));
    mSize = T;
    mNrDim = nrDim;
    mRank = rank;
    mWorld = world;
    check(hml_sigma_hat(mHandle, &mSigmaHat));
  }
  // concatenation, in rank order, of every rank's `mine` (collective; a copy without a communicator)
  template <typename T>
  std::vector<T> allgatherv(const std::vector<T>& mine) const {
    if (mWorld <= 1) return mine;
    std::vector<uint64_t> counts(mWorld);
    const uint64_t n = mine.size();
    check(hml_comm_allgather(mHandle, &n, sizeof(uint64_t), counts.data()));
    uint64_t nmax = 0, total = 0;
    for (uint64_t c : counts) {
      nmax = std::max(nmax, c);
      total += c;
    }
    std::vector<T> padded(nmax + 1), all((nmax + 1) * (size_t)mWorld);
    std::copy(mine.begin(), mine.end(), padded.begin());
    check(hml_comm_allgather(mHandle, padded.data(), (nmax + 1) * sizeof(T), all.data()));
    std::vector<T> out;
    out.reserve(total);
    for (int r = 0; r < mWorld; ++r) out.insert(out.end(), all.begin() + r * (nmax + 1), all.begin() + r * (nmax + 1) + counts[r]);
    return out;
  }

  // integer statistics of the most recent sweep on this sequence as the sampler received them (ForwardBackward.hpp:
  // 177-200): callers that only drive whole runs (bench.py) check their invariants — transition counts, occupancy and
  // per-parameter counts each sum to the sequence length
  struct LastSweep {
    uint64_t nblocks = 0;
    std::vector<uint64_t> counts, trans, statN;
  };
  LastSweep lastSweep;
};

// Reads whitespace-separated numbers with the result of `input >> v` (wavelet.hpp:131), through the
// multi-threaded parser of FastParse.hpp, and loads them.
inline std::vector<float> readValues(std::istream& input, const size_t reserveT = 0,
                                     fastparse::Format format = fastparse::Format::Auto) {
  if (!input) throw std::runtime_error("Cannot read input file or stream!");
  std::vector<float> values;
  values.reserve(reserveT);
  if (std::getenv("HAMMLET_SLOW_PARSE")) {  // the reference's own extraction loop, kept for comparison
    float v;
    while (input >> v) values.push_back(v);
  } else {
    const std::string bytes = fastparse::slurp(input);
    fastparse::parseAny(bytes.data(), bytes.size(), format, values);
  }
  return values;
}
// a file: text, gzip'd text (recognised by its magic number) or raw little-endian float32 (`-F f32`)
inline std::vector<float> readValuesFile(const std::string& path, fastparse::Format format = fastparse::Format::Auto) {
  if (std::getenv("HAMMLET_SLOW_PARSE")) {
    std::ifstream fin(path);
    if (!fin) throw std::runtime_error("Cannot read from input file " + path + "!");
    return readValues(fin);
  }
  const std::string bytes = fastparse::slurpFile(path);
  std::vector<float> values;
  fastparse::parseAny(bytes.data(), bytes.size(), format, values);
  return values;
}
inline void MaxletTransform(std::istream& input, DeviceSequence& seq, const size_t nrDim, const float weightMultiplier,
                            const size_t reserveT = 0) {
  if (nrDim <= 0) throw std::runtime_error("Number of dimensions must be positive!");
  seq.load(readValues(input, reserveT), weightMultiplier, nrDim);
}

template <typename T> class Blocks;
template <typename T, typename StatsType> class Statistics;
template <typename DataStructure, typename DistType> class Emissions;

template <>
class Blocks<BreakpointArray> {
  DeviceSequence& mSeq;
  real_t mThreshold = 0;
  bool mDirty = true;         // threshold changed since the device last built the structure
  // host copy for the iterator protocol
  std::vector<uint32_t> mStarts;
  std::vector<double> mSum, mSumSq;  // dimension-major: [dim * nrBlocks + block]
  bool mHostValid = false;
  bool mIterating = false;
  size_t mCursor = 0, mBlockStart = 0, mBlockEnd = 0, mBlockCounter = 0;

 public:
  Blocks(const Blocks&) = delete;
  explicit Blocks(DeviceSequence& seq) : mSeq(seq) {
    if (seq.size() <= 0) throw std::runtime_error("Input vector for breakpoint weights is empty!");
  }
  DeviceSequence& sequence() const { return mSeq; }

  void createBlocks(real_t threshold) {
    mThreshold = threshold;
    mDirty = true;
    mHostValid = false;
  }
  // threshold from the smallest emission variance, in real_t like BreakpointArray.hpp:195-199
  template <typename ParamType>
  void createBlocks(const Theta<ParamType>& param) {
    createBlocks(std::sqrt(2 * std::log((real_t)mSeq.size()) * param.thresholdValue()));
  }
  real_t threshold() const { return mThreshold; }
  bool dirty() const { return mDirty; }
  // the samplers call this after a sweep that rebuilt the structure on the device
  void markBuilt() { mDirty = false; }

  // make the device structure match the current threshold (static mode, replay mode, host iteration)
  size_t materialize() {
    if (mDirty) {
      uint64_t n = 0;
      mSeq.check(hml_create_blocks(mSeq.handle(), (float)mThreshold, &n));
      mDirty = false;
      mHostValid = false;
    }
    // blocks of the whole sequence (for a split sequence hml_nr_blocks counts the rank's own)
    uint64_t n = 0;
    mSeq.check(hml_segment_info(mSeq.handle(), nullptr, nullptr, nullptr, nullptr, nullptr, &n));
    return n;
  }
  void fetch(bool stats) {
    materialize();
    if (mHostValid && (!stats || !mSum.empty())) return;
    uint64_t n = 0;  // blocks on this handle
    mSeq.check(hml_nr_blocks(mSeq.handle(), &n));
    mStarts.resize(n);
    if (stats) {
      const size_t D = mSeq.nrDim();
      mSum.resize(n * D);
      mSumSq.resize(n * D);
      mSeq.check(hml_get_blocks(mSeq.handle(), mStarts.data(), mSum.data(), mSumSq.data(), n));
      for (size_t d = 1; d < D; ++d)
        mSeq.check(hml_get_block_sums(mSeq.handle(), (uint32_t)d, mSum.data() + d * n, mSumSq.data() + d * n, n));
    } else {
      mSum.clear();
      mSumSq.clear();
      mSeq.check(hml_get_blocks(mSeq.handle(), mStarts.data(), nullptr, nullptr, n));
    }
    if (mSeq.split()) {
      // every rank sees the block list of the whole sequence (starts are positions in the whole sequence already;
      // a block that crosses a rank border is listed, with its complete sums, by the rank where it starts)
      const size_t D = stats ? mSeq.nrDim() : 0;
      std::vector<std::vector<double>> sums, sqs;
      for (size_t d = 0; d < D; ++d) {
        sums.push_back(mSeq.allgatherv(std::vector<double>(mSum.begin() + d * n, mSum.begin() + (d + 1) * n)));
        sqs.push_back(mSeq.allgatherv(std::vector<double>(mSumSq.begin() + d * n, mSumSq.begin() + (d + 1) * n)));
      }
      mStarts = mSeq.allgatherv(mStarts);
      mSum.clear();
      mSumSq.clear();
      for (size_t d = 0; d < D; ++d) {
        mSum.insert(mSum.end(), sums[d].begin(), sums[d].end());
        mSumSq.insert(mSumSq.end(), sqs[d].begin(), sqs[d].end());
      }
    }
    mHostValid = true;
  }
  const std::vector<uint32_t>& starts() const { return mStarts; }

  void initForward() {
    fetch(true);
    mIterating = true;
    mCursor = mBlockStart = mBlockEnd = mBlockCounter = 0;
  }
  bool next() {
    if (mBlockEnd >= mSeq.size()) {
      mIterating = false;
      return false;
    }
    mCursor = mBlockCounter++;
    mBlockStart = mStarts[mCursor];
    mBlockEnd = mCursor + 1 < mStarts.size() ? mStarts[mCursor + 1] : mSeq.size();
    return true;
  }
  size_t start() const { return mBlockStart; }
  size_t end() const { return mBlockEnd; }
  size_t blockSize() const { return mBlockEnd - mBlockStart; }
  size_t size() const { return mSeq.size(); }
  size_t pos() const {
    if (mBlockCounter == 0) throw std::runtime_error("No blocks created yet, position is undefined!");
    return mBlockCounter - 1;
  }
  size_t nrBlocks() const {
    if (mIterating) throw std::runtime_error("Cannot determine size of block structure before all blocks have been seen!");
    return mBlockCounter;
  }
  double currentSum(size_t dim = 0) const { return mSum[dim * mStarts.size() + mCursor]; }
  double currentSumSq(size_t dim = 0) const { return mSumSq[dim * mStarts.size() + mCursor]; }
};

template <>
class Statistics<IntegralArray, Normal> {
  DeviceSequence& mSeq;
  std::vector<SufficientStatistics<Normal>> mCurrent;  // one per data dimension (IntegralArray.hpp:198-212)

 public:
  Statistics(const Statistics&) = delete;
  Statistics(DeviceSequence& seq, const size_t nrDim) : mSeq(seq), mCurrent(nrDim) {
    if (seq.size() <= 0) throw std::runtime_error("Input vector for breakpoint weights is empty!");
    if (nrDim != seq.nrDim())
      throw std::runtime_error("Cannot infer data dimension, the loaded sequence has " + std::to_string(seq.nrDim()) +
                               " values per position!");
  }
  // block sums come from the device's fp64 integral arrays, rounded once to real_t
  template <typename B>
  void setStats(const Blocks<B>& blocks) {
    for (size_t dim = 0; dim < mCurrent.size(); ++dim)
      mCurrent[dim] = SufficientStatistics<Normal>((real_t)blocks.currentSum(dim), (real_t)blocks.currentSumSq(dim));
  }
  const SufficientStatistics<Normal>& suffStat(size_t dim) const { return mCurrent[dim]; }
  size_t nrDim() const { return mCurrent.size(); }
  size_t size() const { return mSeq.size(); }
};

template <typename S, typename T, typename B>
class Emissions<Statistics<S, T>, Blocks<B>> {
  Statistics<S, T>& mStats;
  Blocks<B>& mBlocks;

 public:
  Emissions(Statistics<S, T>& stats, Blocks<B>& blocks) : mStats(stats), mBlocks(blocks) {
    if (mStats.size() != mBlocks.size())
      throw std::runtime_error("Block structure and statistics have different number of data points!");
  }
  Statistics<S, T>& stats() { return mStats; }
  Blocks<B>& blocks() { return mBlocks; }
  const Blocks<B>& blocks() const { return mBlocks; }
  void createBlocks(real_t thresh) { mBlocks.createBlocks(thresh); }
  template <typename ParamType>
  void createBlocks(const Theta<ParamType>& theta) { mBlocks.createBlocks(theta); }
  size_t nrBlocks() const { return mBlocks.nrBlocks(); }
  size_t nrDim() const { return mStats.nrDim(); }
  size_t start() const { return mBlocks.start(); }
  size_t end() const { return mBlocks.end(); }
  size_t blockSize() const { return mBlocks.blockSize(); }
  size_t size() const { return mBlocks.size(); }
  void initForward() { mBlocks.initForward(); }
  bool next() {
    if (!mBlocks.next()) return false;
    mStats.setStats(mBlocks);
    return true;
  }
  const SufficientStatistics<T>& suffStat(size_t dim) const { return mStats.suffStat(dim); }
};

// ---------------------------------------------------------------------------------------- automatic priors

// closed-form NIG hyper-parameters (reference: AutoPriors.hpp:18-80)
inline std::vector<real_t> NormalInverseGammaAutoPrior(real_t s2, real_t p, real_t dataMean, real_t dataVar) {
  if (p < 0 || p > 1) throw std::runtime_error("Parameter p for automatic priors is a probability and must be in [0,1]!");
  if (s2 <= 0) throw std::runtime_error("Parameter s2  for automatic priors is a variance and must be positive!");
  if (dataVar <= 0) throw std::runtime_error("Data variance provided to autoprior must be positive!");
  const real_t M1 = 0.3361, M2 = -0.0042, M3 = -0.0201;
  const real_t b = -std::log(p);
  const real_t alpha = 2.0;
  const real_t beta = s2 * ((2.0 * std::sqrt(b)) / (M1 * std::sqrt(b) + std::sqrt(2.0) * (M2 * b * std::exp(M3 * std::sqrt(b)) + 1)) + b);
  const real_t mu0 = dataMean;
  const real_t nu = beta / dataVar;
  if (beta <= 0) throw std::runtime_error("Autoprior yields non-positive beta!");
  if (nu <= 0) throw std::runtime_error("Autoprior yields non-positive nu!");
  if (!std::isfinite(beta)) throw std::runtime_error("Autoprior yields non-finite beta!");
  if (!std::isfinite(mu0)) throw std::runtime_error("Autoprior yields non-finite mu0!");
  if (!std::isfinite(nu)) throw std::runtime_error("Autoprior yields non-finite nu!");
  return std::vector<real_t>{alpha, beta, mu0, nu};
}

// one pass over the blocks at sqrt(2 log T) * sigma-hat; mean and variance of the block means
// (reference: AutoPriors.hpp:86-110)
template <typename Stats, typename BlocksT>
std::vector<real_t> autoPrior(real_t s2, real_t p, Emissions<Statistics<Stats, Normal>, BlocksT>& y, const double noiseStdev) {
  // (the reference calls initForward before createBlocks; there createBlocks only stores the threshold,
  // so the order is immaterial — here initForward fetches the block list and must come second)
  y.createBlocks(std::sqrt(2 * std::log((double)y.blocks().size())) * noiseStdev);
  y.initForward();
  SufficientStatistics<Normal> muStats;
  while (y.next())
    for (size_t dim = 0; dim < y.nrDim(); ++dim) muStats.addObs(y.suffStat(dim).sum() / y.blockSize());
  const size_t N = y.nrBlocks() * y.nrDim();
  const double blocksMean = sampleMean(muStats, N);
  const double blocksVariance = sampleVariance(muStats, N);
  return NormalInverseGammaAutoPrior(s2, p, blocksMean, blocksVariance);
}
