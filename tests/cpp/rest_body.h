// tests/cpp/rest_body.h -- ONE process function, compiled twice from this very file (like kitchen_body.h):
//   * against the reference itself (oracle/ref/mlref.cpp: `using namespace ml;`), and
//   * against the tracing layer (tests/cpp/test_trace.cpp: `using namespace mlb::tr;`).
// The functor spellings the other shared bodies and the two examples leave out: ImpulseGen, OneShotGen, Interpolator1,
// Allpass1, Differentiator, FractionalDelay (fixed and per-sample delay), IntegerDelay (per-sample delay),
// Allpass<FractionalDelay>, PitchbendableDelay used directly, and the fused FDN<8>.
// Inputs: an audio row and a delay-time row (samples); two output rows.
// FDN_SIZE_DELAYS(fdn, times): see fdn_body.h.
#pragma once

static volatile float kRestParams[] = {0.7f, 77.3f, 300.f, 500.f, 400.f, 0.45f, 131.5f, 200.f, 600.f, 0.004f, 0.0007f, 0.3f};
inline float rp(int i) { return kRestParams[i]; }

struct RestState
{
  ImpulseGen imp;
  OneShotGen shot;
  Interpolator1 interp;
  Allpass1 ap1{0.f};
  Differentiator diff;
  FractionalDelay fixedDelay, sweptDelay;
  IntegerDelay steppedDelay;
  Allpass<FractionalDelay> apf;
  PitchbendableDelay bend;
  FDN<8> fdn;
};

inline void restInit(RestState& s)
{
  s.shot.trigger();
  s.ap1.coeffs = Allpass1::makeCoeffs(rp(0));
  s.fixedDelay.setMaxDelayInSamples(rp(2));
  s.fixedDelay.setDelayInSamples(rp(1));
  s.sweptDelay.setMaxDelayInSamples(rp(3));
  s.steppedDelay.setMaxDelayInSamples(rp(4));
  s.apf.mGain = rp(5);
  s.apf.setMaxDelayInSamples(rp(7));
  s.apf.setDelayInSamples(rp(6));
  s.bend.setMaxDelayInSamples(rp(8));
  const std::array<float, 8> times{{67.f, 73.f, 91.f, 103.f, 131.f, 157.f, 179.f, 199.f}};
  FDN_SIZE_DELAYS(s.fdn, times);
  s.fdn.setDelaysInSamples(times);
  s.fdn.setFilterCutoffs({{0.1f, 0.2f, 0.3f, 0.4f, 0.1f, 0.2f, 0.3f, 0.4f}});
  s.fdn.mFeedbackGains = {{0.5f, 0.5f, 0.5f, 0.5f, 0.5f, 0.5f, 0.5f, 0.5f}};
}

template <class REST_CTX>
inline void restProcess(REST_CTX* ctx, void* state)
{
  RestState* s = static_cast<RestState*>(state);
  DSPVector x = ctx->inputs[0];
  DSPVector d = ctx->inputs[1];
  DSPVector clicks = s->imp(rp(9)) + s->shot(rp(10)) * 0.5f;
  DSPVector level = s->interp(rp(11));
  DSPVector y = s->ap1(s->diff(x) * 0.5f + clicks);
  y = s->fixedDelay(y) + s->sweptDelay(y, d) * 0.5f;
  y = y + s->steppedDelay(x, d * 0.7f) * 0.25f;
  y = s->apf(y);
  DSPVector bent = s->bend(y, d * 1.1f);
  auto wet = s->fdn(bent * 0.3f);
  ctx->outputs[0] = wet.constRow(0) + y * level;
  ctx->outputs[1] = wet.constRow(1) + bent * 0.1f;
}
