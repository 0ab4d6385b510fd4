// tests/cpp/test_trace.cpp -- the reference's own example process functions, compiled against the tracing
// layer (include/mlb200_trace.hpp) with ONLY the namespace changed:
//   examples/audio-and-midi/sine.cpp:16-36    (SineExampleState, sineProcess)
//   examples/audio-and-midi/reverb.cpp:12-124 (AaltoverbState, initializeReverb, processVector)
// plus the SURVEY 8c plumbing chain 0.5 * Lopass{0.1, 1.0}(SineGen.clear()(440/48000)) and a swept shelf.
//
//   test_trace dump <case>                      trace only (no GPU needed), print the graph as JSON
//   test_trace run <case> <instances> <blocks> <in.bin> <out.bin>
//                                               compile for <instances> copies, process <blocks> vectors:
//                                               in [T][n_in][V][64] f32 -> out [T][n_out][V][64] f32
//   test_trace stream <case> <frames> <in.bin> <out.bin>
//                                               ONE instance through TracedProcessor::process with host
//                                               buffers of 37 / 100 / 512 / ... frames (SignalProcessBuffer)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mlb200_trace.hpp"

using namespace mlb::tr;  // the reference says: using namespace ml;

// ---------------------------------------------------------------- sine.cpp:9-36
constexpr int kSampleRate = 48000;
constexpr float kOutputGain = 0.1f;

struct SineExampleState
{
  SineGen s1, s2;
};

void sineProcess(AudioContext* ctx, void* state)
{
  auto procState = static_cast<SineExampleState*>(state);

  // Running the sine generators makes DSPVectors as output.
  // The input parameter is omega: the frequency in Hz divided by the sample rate.
  // The output sines are multiplied by the gain.
  ctx->outputs[0] = procState->s1(220.f / kSampleRate) * kOutputGain;
  ctx->outputs[1] = procState->s2(275.f / kSampleRate) * kOutputGain;
}

// ---------------------------------------------------------------- reverb.cpp:17-124
constexpr float kDecayLo = 0.8, kDecayHi = 20;
Projection unityToDecay(projections::unityToLogParam({kDecayLo, kDecayHi}));

struct AaltoverbState
{
  // parameter smoothers
  LinearGlide mSmoothFeedback;
  LinearGlide mSmoothDelay;

  // reverb machinery
  Allpass<PitchbendableDelay> mAp1, mAp2, mAp3, mAp4;
  Allpass<PitchbendableDelay> mAp5, mAp6, mAp7, mAp8, mAp9, mAp10;
  PitchbendableDelay mDelayL, mDelayR;

  // feedback storage
  DSPVector mvFeedbackL, mvFeedbackR;
};

void initializeReverb(AaltoverbState& r)
{
  // set fixed parameters for reverb
  r.mSmoothFeedback.setGlideTimeInSamples(0.1f * kSampleRate);
  r.mSmoothDelay.setGlideTimeInSamples(0.1f * kSampleRate);

  // set allpass filter coefficients
  r.mAp1.mGain = 0.75f;
  r.mAp2.mGain = 0.70f;
  r.mAp3.mGain = 0.625f;
  r.mAp4.mGain = 0.625f;
  r.mAp5.mGain = r.mAp6.mGain = 0.7f;
  r.mAp7.mGain = r.mAp8.mGain = 0.6f;
  r.mAp9.mGain = r.mAp10.mGain = 0.5f;

  // allocate delay memory
  r.mAp1.setMaxDelayInSamples(500.f);
  r.mAp2.setMaxDelayInSamples(500.f);
  r.mAp3.setMaxDelayInSamples(1000.f);
  r.mAp4.setMaxDelayInSamples(1000.f);
  r.mAp5.setMaxDelayInSamples(2600.f);
  r.mAp6.setMaxDelayInSamples(2600.f);
  r.mAp7.setMaxDelayInSamples(8000.f);
  r.mAp8.setMaxDelayInSamples(8000.f);
  r.mAp9.setMaxDelayInSamples(10000.f);
  r.mAp10.setMaxDelayInSamples(10000.f);
  r.mDelayL.setMaxDelayInSamples(3500.f);
  r.mDelayR.setMaxDelayInSamples(3500.f);
}

void processVector(AudioContext* ctx, void* stateData)
{
  AaltoverbState* r = static_cast<AaltoverbState*>(stateData);

  const float sr = kSampleRate;
  const float RT60const = 0.001f;

  // size and decay parameters from 0-1. It will be more interesting to change these over time in some way.
  float sizeU = 0.5f;
  float decayU = 0.5f;

  // generate delay and feedback scalars
  float decayTime = unityToDecay(decayU);
  float decayIterations = decayTime / (sizeU * 0.5);
  float feedback = (decayU < 1.0f) ? powf(RT60const, 1.0f / decayIterations) : 1.0f;

  // generate smoothed delay time and feedback gain vectors
  DSPVector vSmoothDelay = r->mSmoothDelay(sizeU * 2.0f);
  DSPVector vSmoothFeedback = r->mSmoothFeedback(feedback);

  // get the minimum possible delay in samples, which is the length of a DSPVector.
  DSPVector vMin(kFloatsPerDSPVector);

  // get smoothed allpass times in samples
  DSPVector delayParamInSamples = sr * vSmoothDelay;
  DSPVector vt1 = max(0.00476 * delayParamInSamples, vMin);
  DSPVector vt2 = max(0.00358 * delayParamInSamples, vMin);
  DSPVector vt3 = max(0.00973 * delayParamInSamples, vMin);
  DSPVector vt4 = max(0.00830 * delayParamInSamples, vMin);
  DSPVector vt5 = max(0.029 * delayParamInSamples, vMin);
  DSPVector vt6 = max(0.021 * delayParamInSamples, vMin);
  DSPVector vt7 = max(0.078 * delayParamInSamples, vMin);
  DSPVector vt8 = max(0.090 * delayParamInSamples, vMin);
  DSPVector vt9 = max(0.111 * delayParamInSamples, vMin);
  DSPVector vt10 = max(0.096 * delayParamInSamples, vMin);

  // sum stereo inputs and diffuse with four allpass filters in series
  DSPVector monoInput = (ctx->inputs[0] + ctx->inputs[1]);
  DSPVector diffusedInput = r->mAp4(r->mAp3(r->mAp2(r->mAp1(monoInput, vt1), vt2), vt3), vt4);

  // get delay times in samples, subtracting the constant delay of one DSPVector and clamping to zero
  DSPVector vDelayTimeL = max(0.0313 * delayParamInSamples - vMin, DSPVector(0.f));
  DSPVector vDelayTimeR = max(0.0371 * delayParamInSamples - vMin, DSPVector(0.f));

  // sum diffused input with feedback, and apply late diffusion of two more allpass filters to each channel
  DSPVector vTapL = r->mAp7(r->mAp5(diffusedInput + r->mDelayL(r->mvFeedbackL, vDelayTimeL), vt5), vt7);
  DSPVector vTapR = r->mAp8(r->mAp6(diffusedInput + r->mDelayR(r->mvFeedbackR, vDelayTimeR), vt6), vt8);

  // apply final allpass filter and gain, and store the feedback
  r->mvFeedbackR = r->mAp9(vTapL, vt9) * vSmoothFeedback;
  r->mvFeedbackL = r->mAp10(vTapR, vt10) * vSmoothFeedback;

  // write the stereo outputs
  ctx->outputs[0] = vTapL;
  ctx->outputs[1] = vTapR;
}

// ---------------------------------------------------------------- SURVEY 8c plumbing chain (config 1)
struct ChainState
{
  SineGen osc;
  Lopass lp;
};
void chainProcess(AudioContext* ctx, void* state)
{
  auto s = static_cast<ChainState*>(state);
  ctx->outputs[0] = s->lp(s->osc(440.f / 48000.f)) * 0.5f;
}

// ---------------------------------------------------------------- a swept shelf, the way vcoeffs is meant to be used
struct ShelfState
{
  NoiseGen noise;
  LoShelf shelf;
  Bank<SineGen, 2> pair;
};
void shelfProcess(AudioContext* ctx, void* state)
{
  auto s = static_cast<ShelfState*>(state);
  auto vc = LoShelf::vcoeffs({0.05f, 0.7f, dBToGain(-6.f)}, {0.21f, 1.3f, dBToGain(9.f)});
  DSPVectorArray<2> f;
  f.row(0) = DSPVector(0.01f), f.row(1) = ctx->inputs[0];
  auto two = s->pair(f);
  ctx->outputs[0] = s->shelf(s->noise() * 0.25f + two.constRow(0) * two.constRow(1), vc);
}

// ---------------------------------------------------------------- the shared body (also compiled against the reference)
#include "kitchen_body.h"
void kitchenProcessFn(AudioContext* ctx, void* state) { kitchenProcess(ctx, state); }

struct Case
{
  const char* name;
  size_t nIn, nOut;
  SignalProcessFn fn;
};

int main(int argc, char** argv)
{
  if (argc < 3)
  {
    std::fprintf(stderr, "usage: test_trace dump|run|stream <sine|reverb|chain|shelf|kitchen> ...\n");
    return 2;
  }
  const std::string mode = argv[1], which = argv[2];
  SineExampleState sine;
  AaltoverbState verb;
  initializeReverb(verb);
  ChainState chain;
  chain.osc.clear();
  chain.lp.coeffs = Lopass::makeCoeffs(0.1f, 1.0f);
  ShelfState shelf;
  shelf.noise.setSeed(7);
  KitchenState kitchen;
  kitchenInit(kitchen);

  size_t nIn = 0, nOut = 0;
  SignalProcessFn fn = nullptr;
  void* state = nullptr;
  if (which == "sine") nIn = 0, nOut = 2, fn = sineProcess, state = &sine;
  if (which == "reverb") nIn = 2, nOut = 2, fn = processVector, state = &verb;  // the example sums inputs[0] + inputs[1]
  if (which == "chain") nIn = 0, nOut = 1, fn = chainProcess, state = &chain;
  if (which == "shelf") nIn = 1, nOut = 1, fn = shelfProcess, state = &shelf;
  if (which == "kitchen") nIn = 2, nOut = 2, fn = kitchenProcessFn, state = &kitchen;
  if (!fn) return 2;
  try
  {
    AudioContext ctx(nIn, nOut, kSampleRate);
    TracedProcessor proc;
    proc.trace(&ctx, fn, state);
    if (mode == "dump")
    {
      proc.dump(stdout);
      return 0;
    }
    auto readAll = [](const char* path, std::vector<float>& v)
    {
      FILE* f = std::fopen(path, "rb");
      if (!f) return false;
      const size_t n = std::fread(v.data(), 4, v.size(), f);
      std::fclose(f);
      return n == v.size();
    };
    if (mode == "run" && argc == 7)
    {
      const int V = std::atoi(argv[3]), T = std::atoi(argv[4]);
      std::vector<float> in((size_t)T * nIn * V * 64), out((size_t)T * nOut * V * 64);
      if (nIn && !readAll(argv[5], in)) return 3;
      proc.compile(V);
      // two launches: state, delay lines and feedback rows carry over
      const int T1 = T / 2;
      if (T1 > 0) proc.processBlocks(nIn ? in.data() : nullptr, out.data(), nullptr, T1);
      proc.processBlocks(nIn ? in.data() + (size_t)T1 * nIn * V * 64 : nullptr, out.data() + (size_t)T1 * nOut * V * 64,
                         nullptr, T - T1);
      FILE* f = std::fopen(argv[6], "wb");
      std::fwrite(out.data(), 4, out.size(), f);
      std::fclose(f);
      std::printf("ran %s: %d instances x %d vectors on %s\n", which.c_str(), V, T, proc.kernelName());
      return 0;
    }
    if (mode == "stream" && argc == 6)
    {
      const int frames = std::atoi(argv[3]);
      std::vector<float> in((size_t)nIn * frames), out((size_t)nOut * frames);
      if (nIn && !readAll(argv[4], in)) return 3;
      proc.compile(1);
      static const int sizes[] = {37, 100, 512, 64, 1, 333};
      int done = 0, k = 0;
      while (done < frames)
      {
        const int n = std::min(sizes[k++ % 6], frames - done);
        std::vector<const float*> ip(nIn);
        std::vector<float*> op(nOut);
        for (size_t c = 0; c < nIn; ++c) ip[c] = in.data() + c * frames + done;
        for (size_t c = 0; c < nOut; ++c) op[c] = out.data() + c * frames + done;
        proc.process(nIn ? ip.data() : nullptr, op.data(), n);
        done += n;
      }
      FILE* f = std::fopen(argv[5], "wb");
      std::fwrite(out.data(), 4, out.size(), f);
      std::fclose(f);
      std::printf("streamed %s: %d frames\n", which.c_str(), frames);
      return 0;
    }
  }
  catch (const mlb::Error& e)
  {
    std::printf("mlb error %d: %s\n", e.code, e.what());
    return e.code == MLB_ERR_NO_DEVICE ? 77 : 1;
  }
  return 2;
}
