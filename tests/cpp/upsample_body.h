// tests/cpp/upsample_body.h -- ONE process function, compiled twice from this very file (like kitchen_body.h):
//   * against the reference itself (oracle/ref/mlref.cpp: `using namespace ml;`), and
//   * against the tracing layer (tests/cpp/test_trace.cpp: `using namespace mlb::tr;`).
// A process function with STATE run at twice the rate by Upsample2xFunction<1> (MLDSPFunctional.h:114-160), the way
// the reference's tutorial wraps a sine generator (examples/tutorial/dspOpsExample.cpp:100-102): the oscillator, the
// filters, the envelope and the glide inside `fn` are called twice per vector.  One input row (frequency in cycles
// per sample at the ORIGINAL rate), one gate row; output 0 = that, output 1 = a stateless process function run at HALF
// the rate by Downsample2xFunction<1> (MLDSPFunctional.h:166-223), then a HalfBandFilter pair used directly
// (MLDSPFilters.h:1245-1310) plus a TempoLock (MLDSPFilters.h:1478-1579).
#pragma once

static volatile float kUpsampleParams[] = {0.11f, 0.8f, 0.004f, 0.05f, 1.3f, 4.f, 0.002f, 0.01f, 0.7f, 0.02f};
inline float up(int i) { return kUpsampleParams[i]; }

struct UpsampleState
{
  Upsample2xFunction<1> upper;
  Downsample2xFunction<1> downer;
  SineGen osc;
  SawGen saw;
  Lopass lp;
  Bell bell;
  OnePole smooth;
  ADSR env;
  LinearGlide glide;
  NoiseGen noise;  // called once per vector, outside fn
  // the half-band filter used directly, and a clock follower
  HalfBandFilter hbUp, hbDown;
  PhasorGen clock;
  TempoLock lock;
};

inline void upsampleInit(UpsampleState& s)
{
  s.osc.clear();
  s.noise.setSeed(77);
  s.lp.coeffs = Lopass::makeCoeffs(up(0), up(1));
  s.bell.coeffs = Bell::makeCoeffs(up(3), up(4), dBToGain(up(5)));
  s.smooth.coeffs = OnePole::makeCoeffs(up(2));
  s.env.coeffs = ADSR::calcCoeffs(up(6), up(7), up(8), up(9), 96000.f);
  s.glide.setGlideTimeInSamples(300.f);
}

template <class UPSAMPLE_CTX>
inline void upsampleProcess(UPSAMPLE_CTX* ctx, void* state)
{
  UpsampleState* s = static_cast<UpsampleState*>(state);
  DSPVector freq = ctx->inputs[0];
  DSPVector gate = ctx->inputs[1];
  DSPVector dither = s->noise() * 0.001f;
  // everything inside fn runs at 2x: half the per-sample frequency, and every functor ticks twice per vector
  auto fn = [&](const DSPVector f2)
  {
    DSPVector e = s->env(abs(f2) * 0.f + 1.f) * s->glide(0.75f);
    DSPVector o = s->osc(f2 * 0.5f) + s->saw(f2 * 0.25f) * 0.3f;
    return s->smooth(s->bell(s->lp(o * e)));
  };
  ctx->outputs[0] = s->upper(fn, freq) * gate + dither;
  // a waveshaper at half the rate: stateless, as the half-rate wrapper requires here
  auto shaper = [&](const DSPVector v) { return clamp(v * 3.f, DSPVector(-1.f), DSPVector(1.f)) * 0.5f; };
  DSPVector half = s->downer(shaper, ctx->outputs[0] + gate * 0.25f);
  // HalfBandFilter by hand: both 2x halves, scaled differently, back down; TempoLock follows a clock at 3/2
  DSPVector a = s->hbUp.upsampleFirstHalf(half), b = s->hbUp.upsampleSecondHalf(half);
  DSPVector locked = s->lock(s->clock(freq * 0.01f), 1.5f, 1.f / 48000.f);
  ctx->outputs[1] = s->hbDown.downsample(a * 0.75f, b * 1.25f) + locked * 0.1f;
}
