// tests/cpp/oversample_body.h -- ONE process function, compiled twice from this very file (like kitchen_body.h):
//   * against the reference itself (oracle/ref/mlref.cpp: `using namespace ml;`), and
//   * against the tracing layer (tests/cpp/test_trace.cpp: `using namespace mlb::tr;`).
// Oversampled loops inside one process call, the use Upsampler / Downsampler (MLDSPFilters.h:1316-1473) are made for:
// write a vector, read 2^octaves vectors at the higher rate, run a nonlinearity AND stateful functors on each, write
// them to the Downsampler, read one vector back.  4x (two octaves) on output 0, 2x on output 1; plus an oscillator
// simply called twice in the vector (it ticks twice, as any reference functor would).
#pragma once

static volatile float kOversampleParams[] = {0.21f, 0.9f, 0.02f, 3.f, 0.35f, 1.1f};
inline float op_(int i) { return kOversampleParams[i]; }

struct OversampleState
{
  Upsampler up4{2};
  Downsampler down4{2};
  Upsampler up2{1};
  Downsampler down2{1};
  Lopass lp;       // called four times per vector
  OnePole smooth;  // called four times per vector
  DCBlocker dc;    // called twice per vector
  SineGen osc;     // called twice per vector, outside any resampler
};

inline void oversampleInit(OversampleState& s)
{
  s.osc.clear();
  s.lp.coeffs = Lopass::makeCoeffs(op_(0), op_(1));
  s.smooth.coeffs = OnePole::makeCoeffs(op_(2));
  s.dc.coeffs = DCBlocker::makeCoeffs(0.045f * op_(5));
}

template <class OVERSAMPLE_CTX>
inline void oversampleProcess(OVERSAMPLE_CTX* ctx, void* state)
{
  OversampleState* s = static_cast<OversampleState*>(state);
  DSPVector x = ctx->inputs[0];
  // 4x: waveshaper, then a filter and a smoother that run at the high rate
  s->up4.write(x);
  for (int i = 0; i < 4; ++i)
  {
    DSPVector hi = s->up4.read();
    DSPVector shaped = clamp(hi * op_(3), DSPVector(-1.f), DSPVector(1.f));
    s->down4.write(s->smooth(s->lp(shaped)));
  }
  ctx->outputs[0] = s->down4.read();
  // 2x on a signal made of an oscillator that is called twice in this vector
  DSPVector pair = s->osc(0.003f) * 0.5f + s->osc(0.003f) * 0.25f;
  s->up2.write(pair + x * op_(4));
  for (int i = 0; i < 2; ++i)
  {
    DSPVector hi = s->up2.read();
    s->down2.write(s->dc(hi * hi * hi));
  }
  ctx->outputs[1] = s->down2.read();
}
