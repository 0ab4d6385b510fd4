// tests/cpp/kitchen_body.h -- ONE process function, compiled twice from this very file:
//   * against the reference itself (oracle/ref/mlref.cpp: `using namespace ml;`), and
//   * against the tracing layer (tests/cpp/test_trace.cpp: `using namespace mlb::tr;`).
// It exercises the functor / operator spellings that the two example programs do not: every generator, the
// whole SVF family with coeffs = makeCoeffs(...), the one-pole family, envelope followers, ADSR, both glides,
// integer / fractional delays, Allpass<>, swept shelves through vcoeffs, float <-> DSPVector mixing, compound
// assignment, min / max / clamp / lerp / transcendental ops, and a DSPVector member carried between calls.
// KITCHEN_CTX is the context type (something with `inputs` and `outputs` indexable by int).
#pragma once

// Every numeric parameter goes through a volatile read, so that neither build folds a makeCoeffs / dBToGain call at
// compile time (GCC folds libm calls on literals with exact rounding, which can differ from glibc's run-time result
// in the last bit -- seen on HiShelf::vcoeffs with literal arguments).
static volatile float kKitchenParams[] = {0.09f, 0.7f,  0.002f, 1.2f, 0.05f, 0.3f,  0.01f, 0.9f,  4.f,   0.07f, 0.8f,
                                          -5.f,  0.003f, 0.045f, 0.002f, 0.01f, 0.6f,  0.02f, 0.001f, 0.002f, 0.1f,
                                          0.8f,  -3.f,  0.2f,   1.1f, 6.f};
inline float kp(int i) { return kKitchenParams[i]; }

struct KitchenState
{
  SineGen osc;
  SawGen saw;
  PulseGen pulse;
  NoiseGen noise;
  TickGen tick;
  Lopass lp;
  Hipass hp;
  Bandpass bp;
  LoShelf lshelf;
  HiShelf hshelf;
  Bell bell;
  OnePole smooth;
  DCBlocker dc;
  ADSR env;
  Peak peak;
  RMS rms;
  Integrator integ;
  LinearGlide glide;
  IntegerDelay idelay;
  FractionalDelay fdelay;
  Allpass<IntegerDelay> ap;
  DSPVector carried;  // read before it is written: state between calls
};

inline void kitchenInit(KitchenState& s)
{
  const float sr = 48000.f;
  s.osc.clear();
  s.noise.setSeed(1234);
  s.lp.coeffs = Lopass::makeCoeffs(kp(0), kp(1));
  s.hp.coeffs = Hipass::makeCoeffs(kp(2), kp(3));
  s.bp.coeffs = Bandpass::makeCoeffs(kp(4), kp(5));
  s.lshelf.coeffs = LoShelf::makeCoeffs({kp(6), kp(7), dBToGain(kp(8))});
  s.bell.coeffs = Bell::makeCoeffs(kp(9), kp(10), dBToGain(kp(11)));
  s.smooth.coeffs = OnePole::makeCoeffs(kp(12));
  s.dc.coeffs = DCBlocker::makeCoeffs(kp(13));
  s.env.coeffs = ADSR::calcCoeffs(kp(14), kp(15), kp(16), kp(17), sr);
  s.peak.coeffs = Peak::makeCoeffs(kp(18));
  s.peak.peakHoldSamples = 200;
  s.rms.coeffs = RMS::makeCoeffs(kp(19));
  s.integ.mLeak = 0.01f;
  s.glide.setGlideTimeInSamples(256.f);
  s.idelay.setMaxDelayInSamples(400.f);
  s.idelay.setDelayInSamples(173);
  s.fdelay.setMaxDelayInSamples(300.f);
  s.fdelay.setDelayInSamples(117.37f);
  s.ap.setMaxDelayInSamples(500.f);
  s.ap.setDelayInSamples(211.f);
  s.ap.mGain = 0.6f;
}

template <class KITCHEN_CTX>
inline void kitchenProcess(KITCHEN_CTX* ctx, void* state)
{
  KitchenState* s = static_cast<KitchenState*>(state);
  const float sr = 48000.f;

  // control: a gate row and a frequency row come in; everything else is made here
  DSPVector gate = ctx->inputs[0];
  DSPVector freq = ctx->inputs[1];
  DSPVector envelope = s->env(gate);
  DSPVector detune = s->glide(0.5f) * 0.01f + 1.f;

  // oscillators
  DSPVector voice = s->osc(freq * detune) + s->saw(freq * 0.5f) * 0.5f + s->pulse(freq * 2.f, 0.3f) * 0.25f;
  voice += s->noise() * 0.05f;
  DSPVector clicks = s->tick(110.f / sr);

  // filters with fixed coefficients
  DSPVector y = s->bell(s->lshelf(s->hp(s->lp(voice * envelope))));
  y = y + s->bp(clicks) * 0.5f;

  // a shelf swept over the block through vcoeffs (two host designs, interpolated per sample)
  auto vc = HiShelf::vcoeffs({kp(20), kp(21), dBToGain(kp(22))}, {kp(23), kp(24), dBToGain(kp(25))});
  y = s->hshelf(y, vc);

  // delays and an allpass, with the row carried over from the previous call mixed in
  DSPVector wet = s->ap(s->fdelay(s->idelay(y))) * 0.4f + s->carried * 0.3f;
  s->carried = s->dc(wet);

  // an integrator and some elementwise maths: output 0 stays free of the envelope followers
  DSPVector shaped = clamp(sin(wet * 2.f) * 0.7f + lerp(y, wet, 0.25f), DSPVector(-1.f), DSPVector(1.f));
  DSPVector slow = s->smooth(abs(shaped)) - s->integ(shaped * 0.001f);
  shaped *= min(abs(wet) + 0.1f, DSPVector(1.f));
  shaped /= sqrt(wet * wet + 1.f);

  // followers (Peak and RMS use the CPU's 12-bit rsqrt approximation: tolerance-only on any other hardware)
  DSPVector level = max(s->peak(y), s->rms(y) * 1.5f);

  ctx->outputs[0] = shaped + slow * 0.5f;
  ctx->outputs[1] = level + exp(level * -2.f) * 0.1f;
}
