// tests/cpp/fdn_body.h -- ONE process function, compiled twice from this very file (like kitchen_body.h):
//   * against the reference itself (oracle/ref/mlref.cpp: `using namespace ml;`), and
//   * against the tracing layer (tests/cpp/test_trace.cpp: `using namespace mlb::tr;`).
// FDN<SIZE> for sizes other than the fused 8: a four-line and a six-line network side by side (the tutorial's
// `FDN<4> f`, examples/tutorial/dspOpsExample.cpp:61-72), fed by the two input rows; outputs: left and right sums of both.
// FDN_SIZE_DELAYS(fdn, times) is supplied by the includer: the reference never sizes an FDN's delay lines,
// so its build reaches in and sizes them for exactly their length; the tracing layer does that in setDelaysInSamples.
#pragma once

static volatile float kFdnParams[] = {67.f, 73.f, 91.f, 103.f, 0.1f, 0.2f, 0.3f, 0.4f, 0.5f,
                                      131.f, 157.f, 179.f, 199.f, 223.f, 251.f, 0.15f, 0.45f};
inline float fp(int i) { return kFdnParams[i]; }

struct FdnState
{
  FDN<4> four;
  FDN<6> six;
};

inline void fdnInit(FdnState& s)
{
  const std::array<float, 4> t4{{fp(0), fp(1), fp(2), fp(3)}};
  FDN_SIZE_DELAYS(s.four, t4);
  s.four.setDelaysInSamples(t4);
  s.four.setFilterCutoffs({{fp(4), fp(5), fp(6), fp(7)}});
  s.four.mFeedbackGains = {{fp(8), fp(8), fp(8), fp(8)}};
  const std::array<float, 6> t6{{fp(9), fp(10), fp(11), fp(12), fp(13), fp(14)}};
  FDN_SIZE_DELAYS(s.six, t6);
  s.six.setDelaysInSamples(t6);
  s.six.setFilterCutoffs({{fp(15), fp(15), fp(15), fp(15), fp(15), fp(15)}});
  s.six.mFeedbackGains = {{fp(16), fp(16), fp(16), fp(16), fp(16), fp(16)}};
}

template <class FDN_CTX>
inline void fdnProcess(FDN_CTX* ctx, void* state)
{
  FdnState* s = static_cast<FdnState*>(state);
  DSPVector a = ctx->inputs[0], b = ctx->inputs[1];
  auto y4 = s->four(a);
  auto y6 = s->six(b * 0.5f + y4.constRow(0) * 0.1f);
  ctx->outputs[0] = y4.constRow(0) + y6.constRow(0);
  ctx->outputs[1] = y4.constRow(1) + y6.constRow(1);
}
