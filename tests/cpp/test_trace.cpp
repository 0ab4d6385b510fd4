// tests/cpp/test_trace.cpp -- the reference's own example process functions, compiled against the tracing
// layer (include/mlb200_trace.hpp) with ONLY the namespace changed:
//   examples/audio-and-midi/sine.cpp:9-36     (constants, SineExampleState, sineProcess)
//   examples/audio-and-midi/reverb.cpp:12-124 (constants, unityToDecay, AaltoverbState, initializeReverb, processVector)
// (their text is included from tests/cpp/_ref/*.inc, generated from the reference at build time, not committed),
// plus the SURVEY 8c plumbing chain 0.5 * Lopass{0.1, 1.0}(SineGen.clear()(440/48000)), a swept shelf, and
// tests/cpp/kitchen_body.h and tests/cpp/upsample_body.h (a stateful process function inside Upsample2xFunction<1>),
// which are also compiled against the reference itself.
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

// ---------------------------------------------------------------- the reference's own example bodies
// examples/audio-and-midi/sine.cpp and reverb.cpp, everything between `using namespace ml;` and `int main()`,
// byte for byte: extracted at build time by tests/cpp/make_example_bodies.py into tests/cpp/_ref/ (git-ignored;
// the reference's text is not part of this repository).  The only change is the namespace they see.
#if __has_include("_ref/sine_body.inc") && __has_include("_ref/reverb_body.inc")
#define HAVE_REFERENCE_EXAMPLES 1
namespace sine_example
{
using namespace mlb::tr;  // the reference says: using namespace ml;
#include "_ref/sine_body.inc"
}  // namespace sine_example
namespace reverb_example
{
using namespace mlb::tr;  // the reference says: using namespace ml;
#include "_ref/reverb_body.inc"
}  // namespace reverb_example
#else
#define HAVE_REFERENCE_EXAMPLES 0
#endif

using namespace mlb::tr;

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
// a process function with state inside Upsample2xFunction<1> (also compiled against the reference)
#include "upsample_body.h"
void upsampleProcessFn(AudioContext* ctx, void* state) { upsampleProcess(ctx, state); }
// oversampled loops between an Upsampler and a Downsampler (also compiled against the reference)
#include "oversample_body.h"
void oversampleProcessFn(AudioContext* ctx, void* state) { oversampleProcess(ctx, state); }
// the functor spellings the other bodies leave out (also compiled against the reference)
#define FDN_SIZE_DELAYS(fdn, times) (void)0
#include "rest_body.h"
#undef FDN_SIZE_DELAYS
void restProcessFn(AudioContext* ctx, void* state) { restProcess(ctx, state); }
// DSPVectorArray<ROWS> as a value and the row operations (also compiled against the reference)
#include "rows_body.h"
void rowsProcessFn(AudioContext* ctx, void* state) { rowsProcess(ctx, state); }
// FDN<4> and FDN<6>, recorded as the nodes they are made of (also compiled against the reference)
#define FDN_SIZE_DELAYS(fdn, times) (void)0  // mlb::tr::FDN sizes its delay lines in setDelaysInSamples
#include "fdn_body.h"
#undef FDN_SIZE_DELAYS
void fdnProcessFn(AudioContext* ctx, void* state) { fdnProcess(ctx, state); }
// a functor with a delay ring cannot be called twice per vector (its ring is written once per vector)
struct TwiceState
{
  IntegerDelay delay{100};
};
void twiceProcess(AudioContext* ctx, void* state)
{
  auto s = static_cast<TwiceState*>(state);
  ctx->outputs[0] = s->delay(DSPVector(0.01f)) + s->delay(DSPVector(0.02f));
}
// ... and so does a functor inside the process function of a Downsample2xFunction
struct HalfRateState
{
  Downsample2xFunction<1> downer;
  OnePole smooth;
};
void halfRateProcess(AudioContext* ctx, void* state)
{
  auto s = static_cast<HalfRateState*>(state);
  ctx->outputs[0] = s->downer([&](const DSPVector v) { return s->smooth(v); }, DSPVector(0.5f));
}

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
    std::fprintf(stderr, "usage: test_trace dump|run|stream <sine|reverb|chain|shelf|kitchen|upsample|fdn|rows|rest|oversample|twice|halfrate> ...\n");
    return 2;
  }
  const std::string mode = argv[1], which = argv[2];
#if HAVE_REFERENCE_EXAMPLES
  sine_example::SineExampleState sine;
  reverb_example::AaltoverbState verb;
  reverb_example::initializeReverb(verb);
#endif
  ChainState chain;
  chain.osc.clear();
  chain.lp.coeffs = Lopass::makeCoeffs(0.1f, 1.0f);
  ShelfState shelf;
  shelf.noise.setSeed(7);
  KitchenState kitchen;
  kitchenInit(kitchen);
  UpsampleState upsample;
  upsampleInit(upsample);
  FdnState fdn;
  fdnInit(fdn);
  RowsState rows;
  rowsInit(rows);
  RestState rest;
  restInit(rest);
  OversampleState oversample;
  oversampleInit(oversample);
  TwiceState twice;
  HalfRateState halfRate;

  size_t nIn = 0, nOut = 0;
  SignalProcessFn fn = nullptr;
  void* state = nullptr;
#if HAVE_REFERENCE_EXAMPLES
  if (which == "sine") nIn = 0, nOut = 2, fn = sine_example::sineProcess, state = &sine;
  if (which == "reverb") nIn = 2, nOut = 2, fn = reverb_example::processVector, state = &verb;  // the example sums inputs[0] + inputs[1]
#else
  if (which == "sine" || which == "reverb")
  {
    std::printf("reference example bodies not generated (tests/cpp/make_example_bodies.py)\n");
    return 4;
  }
#endif
  if (which == "chain") nIn = 0, nOut = 1, fn = chainProcess, state = &chain;
  if (which == "shelf") nIn = 1, nOut = 1, fn = shelfProcess, state = &shelf;
  if (which == "kitchen") nIn = 2, nOut = 2, fn = kitchenProcessFn, state = &kitchen;
  if (which == "upsample") nIn = 2, nOut = 2, fn = upsampleProcessFn, state = &upsample;
  if (which == "fdn") nIn = 2, nOut = 2, fn = fdnProcessFn, state = &fdn;
  if (which == "rows") nIn = 1, nOut = 2, fn = rowsProcessFn, state = &rows;
  if (which == "rest") nIn = 2, nOut = 2, fn = restProcessFn, state = &rest;
  if (which == "oversample") nIn = 1, nOut = 2, fn = oversampleProcessFn, state = &oversample;
  if (which == "twice") nIn = 0, nOut = 1, fn = twiceProcess, state = &twice;
  if (which == "halfrate") nIn = 0, nOut = 1, fn = halfRateProcess, state = &halfRate;
  if (!fn) return 2;
  try
  {
    AudioContext ctx(nIn, nOut, 48000);
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
