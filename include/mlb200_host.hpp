// mlb200_host.hpp -- host-side glue for the caller of the boundary:
//
//   mlb::DSPBuffer                 single-producer / single-consumer float ring, the role of the
//                                  reference's ml::DSPBuffer (source/DSP/MLDSPBuffer.h:20-384)
//   mlb::BatchedSignalProcessBuffer serves a main loop that asks for arbitrary chunk sizes, like
//                                  ml::SignalProcessBuffer (source/app/MLSignalProcessBuffer.cpp:36-90),
//                                  but computes ALL the 64-frame vectors a callback needs in ONE
//                                  batched call (= one GPU launch) instead of one call per vector.
//
// Plain C++17, no GPU code: this is host logic above the C ABI.  Behavioural contract taken from
// the reference (write clobbers the oldest data when full, vector reads return silence and consume
// nothing when fewer than 64 samples are available, power-of-two storage with a 2x distance mask so
// that "full" and "empty" are distinguishable) and checked against it by tests/cpp/test_ringbuffer.cpp.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstring>
#include <functional>
#include <vector>

#include "mlb200.h"

namespace mlb
{
class DSPBuffer
{
  std::vector<float> store_;
  size_t capacity_{0};      // power of two
  size_t wrap_{0};          // capacity - 1: index -> storage position
  size_t span_{0};          // 2 * capacity - 1: indices live on [0, 2 * capacity)
  std::atomic<size_t> head_{0};  // next write (producer owned)
  std::atomic<size_t> tail_{0};  // next read (consumer owned)

  void copyIn(size_t at, const float* src, size_t n)
  {
    const size_t pos = at & wrap_;
    const size_t first = std::min(n, capacity_ - pos);
    std::memcpy(store_.data() + pos, src, first * sizeof(float));
    if (n > first) std::memcpy(store_.data(), src + first, (n - first) * sizeof(float));
  }
  void copyOut(size_t at, float* dst, size_t n) const
  {
    const size_t pos = at & wrap_;
    const size_t first = std::min(n, capacity_ - pos);
    std::memcpy(dst, store_.data() + pos, first * sizeof(float));
    if (n > first) std::memcpy(dst + first, store_.data(), (n - first) * sizeof(float));
  }

 public:
  DSPBuffer() = default;
  explicit DSPBuffer(int samples) { resize(samples); }
  DSPBuffer(const DSPBuffer& o) : store_(o.store_), capacity_(o.capacity_), wrap_(o.wrap_), span_(o.span_)
  {
    head_.store(o.head_.load());
    tail_.store(o.tail_.load());
  }

  // storage for at least `samples` floats, rounded up to a power of two, never below one vector
  size_t resize(int samples)
  {
    head_ = tail_ = 0;
    size_t c = MLB_BLOCK;
    while (c < (size_t)std::max(samples, 1)) c <<= 1;
    store_.assign(c, 0.f);
    capacity_ = c, wrap_ = c - 1, span_ = 2 * c - 1;
    return c;
  }
  void clear() { tail_.store(head_.load(std::memory_order_acquire), std::memory_order_release); }

  size_t getReadAvailable() const
  {
    return (head_.load(std::memory_order_acquire) - tail_.load(std::memory_order_relaxed)) & span_;
  }
  size_t getWriteAvailable() const { return capacity_ - getReadAvailable(); }

  // append n samples; when they do not fit, the oldest samples are dropped (the buffer stays full)
  void write(const float* src, size_t n)
  {
    if (n > capacity_)
    {
      // more than the ring can hold (the reference writes out of bounds here): keep the newest part
      src += n - capacity_;
      n = capacity_;
    }
    const bool overflow = getWriteAvailable() < n;
    const size_t h = head_.load(std::memory_order_relaxed);
    copyIn(h, src, n);
    const size_t nh = (h + n) & span_;
    head_.store(nh, std::memory_order_release);
    if (overflow) tail_.store((nh - capacity_) & span_, std::memory_order_release);
  }
  // pop up to n samples, returns how many were delivered
  size_t read(float* dst, size_t n)
  {
    n = std::min(n, getReadAvailable());
    const size_t t = tail_.load(std::memory_order_relaxed);
    copyOut(t, dst, n);
    tail_.store((t + n) & span_, std::memory_order_release);
    return n;
  }
  // pop exactly one 64-sample vector; silence (and nothing consumed) when fewer are available
  bool readVector(float* dst64)
  {
    if (getReadAvailable() < MLB_BLOCK)
    {
      std::memset(dst64, 0, MLB_BLOCK * sizeof(float));
      return false;
    }
    read(dst64, MLB_BLOCK);
    return true;
  }
  void discard(size_t n)
  {
    n = std::min(n, getReadAvailable());
    tail_.store((tail_.load(std::memory_order_relaxed) + n) & span_, std::memory_order_release);
  }
  // Overlap-add writer (reference: DSPBuffer::writeWithOverlapAdd, source/DSP/MLDSPBuffer.h:288-320): ADD n samples
  // to what is stored from the write index on, zero the following n - overlap samples for the next window, and
  // advance the write index by n - overlap.  Partial windows are never written (needs 2 n - overlap of space).
  void writeWithOverlapAdd(const float* src, size_t n, size_t overlap)
  {
    if (getWriteAvailable() < n * 2 - overlap) return;
    const size_t h = head_.load(std::memory_order_acquire);
    for (size_t i = 0; i < n; ++i) store_[(h + i) & wrap_] += src[i];
    const size_t after = (h + n) & span_;
    for (size_t i = 0; i < n - overlap; ++i) store_[(after + i) & wrap_] = 0.f;
    head_.store((after - overlap) & span_, std::memory_order_release);
  }
  // Overlapping reader (:323-340): deliver up to n samples (the last `overlap` of them may lie beyond the write
  // index), then advance the read index by delivered - overlap.
  void readWithOverlap(float* dst, size_t n, size_t overlap)
  {
    n = std::min(n, getReadAvailable() + overlap);
    const size_t t = tail_.load(std::memory_order_acquire);
    copyOut(t, dst, n);
    tail_.store((t + n - overlap) & span_, std::memory_order_release);
  }
  // copy the newest n samples without consuming anything; nothing happens when fewer are available
  // (reference: DSPBuffer::peekMostRecent, source/DSP/MLDSPBuffer.h:344-384)
  void peekMostRecent(float* dst, size_t n) const
  {
    const size_t avail = getReadAvailable();
    if (avail < n) return;
    copyOut((tail_.load(std::memory_order_relaxed) + (avail - n)) & span_, dst, n);
  }
};

// PublishedSignal: a signal handed from the DSP side to outside code such as displays, decimated by
// 2^octavesDown and stored frame-major in a DSPBuffer (reference: SignalProcessor::PublishedSignal,
// source/app/MLSignalProcessor.h:28-105, MLSignalProcessor.cpp:11-38).  The rows it takes are host rows,
// e.g. planes read back from mlb_graph_process_host.
class PublishedSignal
{
  std::vector<float> rotate_;
  DSPBuffer buffer_;
  size_t channels_{0};
  int octavesDown_{0};
  int downsampleCtr_{0};

 public:
  PublishedSignal(int maxFrames, int maxVoices, int channels, int octavesDown)
      : rotate_((size_t)maxFrames * channels), channels_((size_t)channels), octavesDown_(octavesDown)
  {
    buffer_.resize(maxFrames * channels * maxVoices);
  }
  size_t getNumChannels() const { return channels_; }
  int getAvailableFrames() const { return (int)(channels_ ? buffer_.getReadAvailable() / channels_ : 0); }
  int getReadAvailable() const { return (int)buffer_.getReadAvailable(); }

  // rows: `channels` rows of 64 samples (one voice's DSPVectorArray<CHANNELS>); every 2^octavesDown-th frame
  // of the first `frames` frames is written, the channels of a frame next to each other (writeQuick, .h:60-84)
  void writeQuick(const float* rows, size_t frames, size_t /*voice*/ = 0)
  {
    size_t framesWritten = 0;
    for (size_t f = 0; f < frames; ++f)
    {
      if (++downsampleCtr_ >= (1 << octavesDown_))
      {
        for (size_t j = 0; j < channels_; ++j) rotate_[framesWritten * channels_ + j] = rows[j * MLB_BLOCK + f];
        ++framesWritten;
        downsampleCtr_ = 0;
      }
    }
    if (framesWritten) buffer_.write(rotate_.data(), framesWritten * channels_);
  }
  // one frame of `channels` contiguous values (writeQuickVert, .h:87-97)
  void writeQuickVert(const float* frame, size_t channels, size_t /*voice*/ = 0)
  {
    if (++downsampleCtr_ >= (1 << octavesDown_))
    {
      buffer_.write(frame, channels);
      downsampleCtr_ = 0;
    }
  }
  size_t readLatest(float* dst, size_t framesRequested)  // .cpp:19-28
  {
    const size_t avail = buffer_.getReadAvailable();
    if (avail > framesRequested * channels_) buffer_.discard(avail - framesRequested * channels_);
    return buffer_.read(dst, framesRequested * channels_);
  }
  void peekLatest(float* dst, size_t framesRequested) const { buffer_.peekMostRecent(dst, framesRequested * channels_); }
  size_t read(float* dst, size_t framesRequested) { return buffer_.read(dst, framesRequested * channels_); }
};

// in [n_vectors][n_inputs][64] -> out [n_vectors][n_outputs][64]: every vector the host callback needs,
// in one call (for a DeviceBank: one kernel launch with n_blocks = n_vectors)
using BatchProcessFn = std::function<void(const float* in, float* out, int n_vectors)>;

class BatchedSignalProcessBuffer
{
  std::vector<DSPBuffer> in_, out_;
  std::vector<float> inBlocks_, outBlocks_;
  size_t maxFrames_;

 public:
  BatchedSignalProcessBuffer(size_t inputs, size_t outputs, size_t maxFrames)
      : in_(inputs), out_(outputs), maxFrames_(maxFrames)
  {
    for (auto& b : in_) b.resize((int)maxFrames);
    for (auto& b : out_) b.resize((int)maxFrames);
    const size_t maxVectors = maxFrames / MLB_BLOCK + 2;
    inBlocks_.assign(maxVectors * std::max<size_t>(1, inputs) * MLB_BLOCK, 0.f);
    outBlocks_.assign(maxVectors * std::max<size_t>(1, outputs) * MLB_BLOCK, 0.f);
  }

  // Same contract as ml::SignalProcessBuffer::process: buffer `frames` of every external input,
  // compute whole vectors until `frames` of output are available, deliver them.  Returns the number
  // of vectors computed by this call (0 when the output ring already held enough).
  int process(const float* const* externalInputs, float* const* externalOutputs, int frames,
              const BatchProcessFn& fn)
  {
    const size_t nIn = in_.size(), nOut = out_.size();
    if (nOut < 1 || !externalOutputs || frames > (int)maxFrames_ || frames < 0) return 0;
    for (size_t c = 0; c < nIn; ++c)
      if (externalInputs && externalInputs[c]) in_[c].write(externalInputs[c], (size_t)frames);

    const size_t have = out_[0].getReadAvailable();
    int nVec = 0;
    if (have < (size_t)frames) nVec = (int)(((size_t)frames - have + MLB_BLOCK - 1) / MLB_BLOCK);
    if (nVec > 0)
    {
      // the reference pops one input vector per computed vector; with too little input it gets silence
      for (int i = 0; i < nVec; ++i)
        for (size_t c = 0; c < nIn; ++c) in_[c].readVector(&inBlocks_[((size_t)i * nIn + c) * MLB_BLOCK]);
      fn(nIn ? inBlocks_.data() : nullptr, outBlocks_.data(), nVec);
      for (int i = 0; i < nVec; ++i)
        for (size_t c = 0; c < nOut; ++c) out_[c].write(&outBlocks_[((size_t)i * nOut + c) * MLB_BLOCK], MLB_BLOCK);
    }
    for (size_t c = 0; c < nOut; ++c)
      if (externalOutputs[c]) out_[c].read(externalOutputs[c], (size_t)frames);
    return nVec;
  }
};
}  // namespace mlb
