// hammlet_b200 host side — output recording with the reference's file formats
// (reference: src/Records.hpp, src/StateMarginals.hpp; formats in SURVEY.md App. C).
//
// StateMarginals keeps, like the reference, the common refinement of all recorded segmentations with
// one count per state and segment.  The reference threads a run-length code through two deques; here
// the refinement is rebuilt by a linear merge of the previous segment list with the incoming runs
// (two flat arrays that swap roles per recorded iteration).  Observable behaviour is the same: the
// saved lines, nrSegments() and internalSize() (the length the reference's code would have).
#pragma once

#include "Model.hpp"

class StateMarginals {
  struct SegList {
    std::vector<size_t> size;
    std::vector<marginal_t> count;  // row-major, stride columns per segment
    void clear() {
      size.clear();
      count.clear();
    }
  };
  const size_t mSize;
  size_t mStride;
  SegList mCur, mNext;
  size_t mFront = 0;           // first unconsumed segment of mCur
  size_t mFrontRemaining = 0;  // observations left in it
  size_t mNrIterations = 0;
  size_t mNrStates = 0;        // highest recorded label + 1
  size_t mNrSegments = 1;

  // length of the reference's code for one segment: a negative count per non-zero state, a positive
  // "next state" marker whenever states are skipped, and the terminating zero (StateMarginals.hpp:51-137)
  size_t codeLength(const marginal_t* row) const {
    size_t len = 1, expected = 0;
    for (size_t s = 0; s < mStride; ++s)
      if (row[s] != 0) {
        len += (s != expected) ? 2 : 1;
        expected = s + 1;
      }
    return len;
  }
  void widen(size_t columns) {
    for (SegList* l : {&mCur, &mNext}) {
      std::vector<marginal_t> wide(l->size.size() * columns, 0);
      for (size_t i = 0; i < l->size.size(); ++i)
        std::copy(l->count.begin() + i * mStride, l->count.begin() + (i + 1) * mStride, wide.begin() + i * columns);
      l->count.swap(wide);
    }
    mStride = columns;
  }

 public:
  StateMarginals(const StateMarginals&) = delete;
  explicit StateMarginals(size_t size, size_t nrStates = 2) : mSize(size), mStride(std::max<size_t>(nrStates, 1)) {
    mCur.size.push_back(size);
    mCur.count.assign(mStride, 0);
    mFrontRemaining = size;
  }
  size_t size() const { return mSize; }

  // `blockSize` further observations of the iteration being recorded are in `state`
  void addRecord(const marginal_t state, size_t blockSize, const marginal_t count = 1) {
    if ((size_t)state >= mNrStates) mNrStates = state + 1;
    if ((size_t)state >= mStride) widen(state + 1);
    while (blockSize > 0) {
      if (mFront >= mCur.size.size()) throw std::runtime_error("Empty count queue, this is a bug!");
      const size_t take = std::min(blockSize, mFrontRemaining);
      mNext.size.push_back(take);
      mNext.count.insert(mNext.count.end(), mCur.count.begin() + mFront * mStride, mCur.count.begin() + (mFront + 1) * mStride);
      mNext.count[mNext.count.size() - mStride + state] += count;
      if (blockSize < mFrontRemaining) {  // the run ends inside the front segment: it is split
        mFrontRemaining -= blockSize;
        mNrSegments++;
        break;
      }
      blockSize -= mFrontRemaining;
      mFront++;
      mFrontRemaining = mFront < mCur.size.size() ? mCur.size[mFront] : 0;
    }
    if (mFront >= mCur.size.size()) {  // the iteration covered all positions
      std::swap(mCur, mNext);
      mNext.clear();
      mFront = 0;
      mFrontRemaining = mCur.size[0];
      mNrIterations++;
    }
  }

  size_t nrSegments() const { return mNrSegments; }
  size_t nrIterations() const { return mNrIterations; }
  size_t internalSize() const {
    size_t n = 0;
    for (size_t i = mFront; i < mCur.size.size(); ++i) n += codeLength(&mCur.count[i * mStride]);
    for (size_t i = 0; i < mNext.size.size(); ++i) n += codeLength(&mNext.count[i * mStride]);
    return n;
  }

  // one line per segment: size TAB count_0 TAB ... count_{S-1}, S = highest recorded label + 1
  void save(std::ofstream& ofs, size_t chunkSize = 1) const {
    if (chunkSize <= 0) throw std::runtime_error("Chunk size must be at least one!");
    if (mFront != 0 || !mNext.size.empty())
      throw std::runtime_error("Cannot output incomplete marginals, currently processing block " + std::to_string(mFront) + "!");
    for (size_t i = 0; i < mCur.size.size(); ++i) {
      size_t iterations = 0;
      ofs << mCur.size[i];
      for (size_t s = 0; s < mNrStates; ++s) {
        const marginal_t c = mCur.count[i * mStride + s];
        ofs << "\t" << c;
        iterations += c;
      }
      ofs << std::endl;
      if (iterations != mNrIterations)
        throw std::runtime_error("Sum of marginals (" + std::to_string(iterations) + ") does not match the number of iterations (" +
                                 std::to_string(mNrIterations) + ")!");
    }
  }
};

// Marginals that are kept elsewhere — on the device (DeviceMarginals in StateSequence.hpp, over hml_marginals_*).
// Records only says when the iteration just sampled is to be added and asks for the file at the end, so this
// header needs nothing from the C ABI (records_tool links without the CUDA library).
struct MarginalsSink {
  virtual ~MarginalsSink() {}
  virtual void addIteration() = 0;
  virtual void save(std::ofstream& ofs) = 0;
  virtual size_t nrSegments() = 0;
};

class Records {
  MarginalsSink* mSink = nullptr;
  bool mSinkUsed = false;
  size_t mNrObservedPos = 0, mNrBlocks = 0, mNrSegments = 0;
  size_t mSegmentState = 0, mSegmentSize = 0;
  const size_t mSize;
  std::string mPrefix, mSuffix;
  StateMarginals mMarginals;
  bool mRecordMarginals = true, mRecordBlocks = false, mRecordCompression = false, mRecordSequences = false,
       mRecordTheta = false, mRecordSegments = false;
  std::ofstream mMarginalsFile, mSequenceFile, mBlocksFile, mThetaFile, mCompressionsFile, mSegmentFile;
  bool mClosed = false;
  // A rank > 0 of a split sequence records like rank 0 (the same calls in the same order, some of them collective)
  // but owns no files: its streams stay closed, so everything written to them is dropped.
  bool mMute = false;

  void setRecordX(std::ofstream& file, const std::string& type, bool& member, const bool flag, const bool overwrite) {
    member = flag;
    if (member && !file.is_open() && !mMute) {
      const std::string filename = mPrefix + type + mSuffix;
      if (hammlet::fileExists(filename) && !overwrite)
        throw std::runtime_error("File " + filename + " already exists! Use -w to allow overwrite!");
      file.open(filename.c_str());
      if (!file.is_open()) throw std::runtime_error("Cannot write to file " + filename + "!");
    }
  }

 public:
  Records(const Records&) = delete;
  Records(size_t T, std::string prefix, std::string suffix, const size_t nrStates)
      : mSize(T), mPrefix(prefix), mSuffix(suffix), mMarginals(T, nrStates) {}
  ~Records() {
    try {
      close();
    } catch (...) {
    }
  }
  void close() {
    if (mClosed) return;
    mClosed = true;
    if (mRecordMarginals) {
      if (mMarginalsFile.is_open() || mMute) saveMarginals(mMarginalsFile);
      mMarginalsFile.close();
    }
    if (mRecordSequences) mSequenceFile.close();
    if (mRecordBlocks) mBlocksFile.close();
    if (mRecordTheta) mThetaFile.close();
    if (mRecordCompression) mCompressionsFile.close();
    if (mRecordSegments) mSegmentFile.close();
  }
  void setRecordMarginals(bool b, bool overwrite = false) { setRecordX(mMarginalsFile, "marginals", mRecordMarginals, b, overwrite); }
  void setRecordBlocks(bool b, bool overwrite = false) { setRecordX(mBlocksFile, "blocks", mRecordBlocks, b, overwrite); }
  void setRecordCompression(bool b, bool overwrite = false) { setRecordX(mCompressionsFile, "compression", mRecordCompression, b, overwrite); }
  void setRecordStateSequence(bool b, bool overwrite = false) { setRecordX(mSequenceFile, "sequences", mRecordSequences, b, overwrite); }
  void setRecordTheta(bool b, bool overwrite = false) { setRecordX(mThetaFile, "parameters", mRecordTheta, b, overwrite); }
  void setRecordSegments(bool b, bool overwrite = false) { setRecordX(mSegmentFile, "segments", mRecordSegments, b, overwrite); }
  bool wantsBlocks() const { return mRecordBlocks; }
  void setMute(bool on) { mMute = on; }  // before any setRecord*()

  // ---- marginals accumulated on the device
  void setMarginalsSink(MarginalsSink* sink) { mSink = sink; }
  // a recorded iteration needs nothing on the host but the marginals (and the block count for the compression file)
  bool canRecordOnDevice() const {
    return mSink && mRecordMarginals && !mRecordBlocks && !mRecordSequences && !mRecordSegments && mNrObservedPos == 0 &&
           mMarginals.nrIterations() == 0;
  }
  // the iteration just sampled joins the device-side marginals (StateMarginals::addRecord for all its runs at once)
  void recordIterationOnDevice(const size_t nrBlocks) {
    mSink->addIteration();
    mSinkUsed = true;
    if (mRecordCompression) mCompressionsFile << ((double)mSize) / ((double)nrBlocks) << std::endl;
  }
  size_t nrMarginalSegments() const { return mSinkUsed ? mSink->nrSegments() : mMarginals.nrSegments(); }
  void saveMarginals(std::ofstream& ofs) const {
    if (mSinkUsed)
      mSink->save(ofs);
    else
      mMarginals.save(ofs);
  }

  template <typename ThetaType>
  void record(const Theta<ThetaType>& theta) {
    if (mRecordTheta) mThetaFile << theta << std::endl;
  }

  // one block of the iteration being recorded: N observations in `state` (Records.hpp:155-235).
  // Equal-state neighbours merge into segments; a finished segment goes to the marginals and, as
  // "size:state", to the sequences file; the line ends when all T positions have been seen.
  // A run of `nrBlocks` consecutive blocks in one state, N observations in total: what nrBlocks calls of
  // record(state, n_i) amount to when the per-block sizes are not written out (the device merges the runs,
  // hml_get_segments).  `nrBlocks` may be 0 for all runs but the first of an iteration as long as the counts add up
  // to the iteration's block count (only the total enters the compression file).
  void recordRun(const size_t state, const size_t N, const size_t nrBlocks) {
    if (mRecordBlocks) throw std::runtime_error("Block sizes are being recorded: blocks must be recorded one by one!");
    if (mSinkUsed) throw std::runtime_error("The marginals of this run accumulate on the device!");
    const bool first = mNrObservedPos == 0;
    if (first) {
      mSegmentState = state;
      mSegmentSize = N;
    } else if (state != mSegmentState) {
      flushSegment(false);
      mSegmentState = state;
      mSegmentSize = N;
      mNrSegments++;
    } else {
      mSegmentSize += N;
    }
    mNrBlocks += nrBlocks;
    mNrObservedPos += N;
    if (mNrObservedPos >= mSize) {
      if (mNrObservedPos > mSize) throw std::runtime_error("Cannot record block, exceeding data size!");
      if (mRecordCompression) mCompressionsFile << ((double)mSize) / ((double)mNrBlocks) << std::endl;
      if (mRecordSegments) mSegmentFile << mMarginals.nrSegments() << "\t" << mMarginals.internalSize() << std::endl;
      flushSegment(true);
      mNrObservedPos = mNrBlocks = mSegmentSize = mNrSegments = 0;
    }
  }
  const StateMarginals& marginals() const { return mMarginals; }

  void record(const size_t state, const size_t N) {
    if (mSinkUsed) throw std::runtime_error("The marginals of this run accumulate on the device!");
    const bool firstBlock = mNrBlocks == 0;
    if (firstBlock) {
      mSegmentState = state;
      mSegmentSize = N;
    } else if (state != mSegmentState) {
      flushSegment(false);
      mSegmentState = state;
      mSegmentSize = N;
      mNrSegments++;
    } else {
      mSegmentSize += N;
    }
    mNrBlocks++;
    mNrObservedPos += N;
    const bool lineEnd = mNrObservedPos >= mSize;
    if (lineEnd) {
      if (mNrObservedPos > mSize) throw std::runtime_error("Cannot record block, exceeding data size!");
      if (mRecordCompression) mCompressionsFile << ((double)mSize) / ((double)mNrBlocks) << std::endl;
      if (mRecordSegments) mSegmentFile << mMarginals.nrSegments() << "\t" << mMarginals.internalSize() << std::endl;
      flushSegment(true);
      mNrObservedPos = mNrBlocks = mSegmentSize = mNrSegments = 0;
    }
    if (mRecordBlocks) mBlocksFile << (firstBlock ? "" : "\t") << N << (lineEnd ? "\n" : "");
  }

 private:
  void flushSegment(bool lineEnd) {
    if (mRecordMarginals) mMarginals.addRecord(mSegmentState, mSegmentSize);
    if (mRecordSequences) mSequenceFile << (mNrSegments > 0 ? "\t" : "") << mSegmentSize << ":" << mSegmentState << (lineEnd ? "\n" : "");
  }
};
