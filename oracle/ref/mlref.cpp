// oracle/ref/mlref.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Runs voice graphs through the UNMODIFIED reference implementation
// (madronalib's header-only SSE DSP layer), compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/libmlref.so.  No reference
// source is copied into this repository: this file only #includes the headers
// where they lie and calls the reference's own functors.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load the resulting library.  The product path
// (madronalib_b200/) never does.
//
// Build flags that matter (see oracle/Makefile and SURVEY.md section 8c):
//   -fno-strict-aliasing   the reference reads DSPVectorArrayInt through float*
//   -ffp-contract=off      scalar recurrences must stay mul-then-add (SSE2)
//   -include cstdint -include cstddef   libstdc++ 13 needs them (MLDSPMathSSE.h:68-72)
//   -fno-access-control    read/write private functor state (mOmega32, ic1eq, ...)
//                          without touching the reference sources
#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "MLDSPOps.h"
#include "MLDSPFilters.h"
#include "MLDSPGens.h"
#include "MLDSPFunctional.h"

#include "mlb200.h"

using namespace ml;

namespace
{
struct OpInfo
{
  int nin, nst, nco;
};

bool opInfo(int op, OpInfo& oi)
{
  switch (op)
  {
#define MLB_X_CASE(NAME, id, nin, nst, nco) \
  case id:                                  \
    oi = {nin, nst, nco};                   \
    return true;
    MLB_OP_TABLE(MLB_X_CASE)
#undef MLB_X_CASE
  }
  return false;
}

inline uint32_t f2u(float f)
{
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}
inline float u2f(uint32_t u)
{
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

// reinterpret a float row as an int row and back (the reference's int vectors
// share storage with float vectors, MLDSPOps.h:392-421)
inline DSPVectorInt asInt(const DSPVector& x)
{
  DSPVectorInt r;
  std::memcpy(r.getBufferInt(), x.getConstBuffer(), sizeof(float) * kFloatsPerDSPVector);
  return r;
}
inline DSPVector asFloat(const DSPVectorInt& x)
{
  DSPVector r;
  std::memcpy(r.getBuffer(), x.getConstBufferInt(), sizeof(float) * kFloatsPerDSPVector);
  return r;
}

// One voice's processor for one node.  Holds the reference functor by value.
struct Proc
{
  virtual ~Proc() {}
  virtual void setCoefs(const float*) {}
  virtual void loadState(const uint32_t*) {}
  virtual void storeState(uint32_t*) const {}
  // returns output row; out2 only for FDN8
  virtual DSPVector run(const DSPVector* const* in, DSPVector* out2) = 0;
};

struct PParam : Proc
{
  float k{0};
  void setCoefs(const float* c) override { k = c[0]; }
  DSPVector run(const DSPVector* const*, DSPVector*) override { return DSPVector(k); }
};
struct PNoise : Proc
{
  NoiseGen g;
  void loadState(const uint32_t* s) override { g.setSeed(s[0]); }
  void storeState(uint32_t* s) const override { s[0] = g.mSeed; }
  DSPVector run(const DSPVector* const*, DSPVector*) override { return g(); }
};
struct PPhasor : Proc
{
  PhasorGen g;
  void loadState(const uint32_t* s) override { g.clear(s[0]); }
  void storeState(uint32_t* s) const override { s[0] = g.mOmega32; }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return g(*in[0]); }
};
struct PSine : Proc
{
  SineGen g;
  void loadState(const uint32_t* s) override { g._phasor.clear(s[0]); }
  void storeState(uint32_t* s) const override { s[0] = g._phasor.mOmega32; }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return g(*in[0]); }
};
struct PSaw : Proc
{
  SawGen g;
  void loadState(const uint32_t* s) override { g._phasor.clear(s[0]); }
  void storeState(uint32_t* s) const override { s[0] = g._phasor.mOmega32; }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return g(*in[0]); }
};
struct PPulse : Proc
{
  PulseGen g;
  void loadState(const uint32_t* s) override { g._phasor.clear(s[0]); }
  void storeState(uint32_t* s) const override { s[0] = g._phasor.mOmega32; }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return g(*in[0], *in[1]); }
};
struct PTick : Proc
{
  TickGen g;
  void loadState(const uint32_t* s) override { g.mOmega = u2f(s[0]); }
  void storeState(uint32_t* s) const override { s[0] = f2u(g.mOmega); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return g(*in[0]); }
};

#define SVF_STATE(f)                                                                 \
  void loadState(const uint32_t* s) override { f.ic1eq = u2f(s[0]), f.ic2eq = u2f(s[1]); } \
  void storeState(uint32_t* s) const override { s[0] = f2u(f.ic1eq), s[1] = f2u(f.ic2eq); }

struct PLopass : Proc
{
  Lopass f;
  void setCoefs(const float* c) override { f.coeffs = {c[0], c[1], c[2]}; }
  SVF_STATE(f)
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PHipass : Proc
{
  Hipass f;
  void setCoefs(const float* c) override { f.coeffs = {c[0], c[1], c[2], c[3]}; }
  SVF_STATE(f)
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PBandpass : Proc
{
  Bandpass f;
  void setCoefs(const float* c) override { f.coeffs = {c[0], c[1], c[2]}; }
  SVF_STATE(f)
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PLoShelf : Proc
{
  LoShelf f;
  void setCoefs(const float* c) override { f.coeffs = {c[0], c[1], c[2], c[3], c[4]}; }
  SVF_STATE(f)
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PHiShelf : Proc
{
  HiShelf f;
  void setCoefs(const float* c) override { f.coeffs = {c[0], c[1], c[2], c[3], c[4], c[5]}; }
  SVF_STATE(f)
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PBell : Proc
{
  Bell f;
  void setCoefs(const float* c) override { f.coeffs = {c[0], c[1], c[2], c[3]}; }
  SVF_STATE(f)
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
// ---- coefficient-row (modulated) forms ----
// Lopass has no operator taking coefficient ROWS (only (vx, omega, k), which designs them inside, F:136-152):
// this node restates that operator's loop on rows handed in -- the same eight lines, the reference's types.
// tests pin it to the reference's own operator()(vx, omega, k) through mlref_lopass_mod below.
struct PLopassV : Proc
{
  Lopass f;
  SVF_STATE(f)
  DSPVector run(const DSPVector* const* in, DSPVector*) override
  {
    const DSPVector &vx = *in[0], &g0 = *in[1], &g1 = *in[2], &g2 = *in[3];
    DSPVector vy;
    for (int n = 0; n < kFloatsPerDSPVector; ++n)
    {
      float v0 = vx[n];
      float t0 = v0 - f.ic2eq;
      float t1 = g0[n] * t0 + g1[n] * f.ic1eq;
      float t2 = g2[n] * t0 + g0[n] * f.ic1eq;
      float v2 = t2 + f.ic2eq;
      f.ic1eq += 2.0f * t1;
      f.ic2eq += 2.0f * t2;
      vy[n] = v2;
    }
    return vy;
  }
};
// Lopass::operator()(vx, omega, k) itself (makeCoeffsVec with glibc sinf inside)
struct PLopassMod : Proc
{
  Lopass f;
  SVF_STATE(f)
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0], *in[1], *in[2]); }
};
// LoShelf / HiShelf::operator()(vx, vc) called directly with the rows as the DSPVectorArray
struct PLoShelfV : Proc
{
  LoShelf f;
  SVF_STATE(f)
  DSPVector run(const DSPVector* const* in, DSPVector*) override
  {
    DSPVectorArray<5> vc;
    for (int i = 0; i < 5; ++i) vc.row(i) = *in[1 + i];
    return f(*in[0], vc);
  }
};
struct PHiShelfV : Proc
{
  HiShelf f;
  SVF_STATE(f)
  DSPVector run(const DSPVector* const* in, DSPVector*) override
  {
    DSPVectorArray<6> vc;
    for (int i = 0; i < 6; ++i) vc.row(i) = *in[1 + i];
    return f(*in[0], vc);
  }
};
struct PRamp : Proc
{
  DSPVector run(const DSPVector* const* in, DSPVector*) override
  {
    return interpolateDSPVectorLinear((*in[0])[0], (*in[1])[0]);
  }
};
struct POnePole : Proc
{
  OnePole f;
  void setCoefs(const float* c) override { f.coeffs = {c[0], c[1]}; }
  void loadState(const uint32_t* s) override { f.y1 = u2f(s[0]); }
  void storeState(uint32_t* s) const override { s[0] = f2u(f.y1); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PDCBlocker : Proc
{
  DCBlocker f;
  void setCoefs(const float* c) override { f.coeffs = c[0]; }
  void loadState(const uint32_t* s) override { f.x1 = u2f(s[0]), f.y1 = u2f(s[1]); }
  void storeState(uint32_t* s) const override { s[0] = f2u(f.x1), s[1] = f2u(f.y1); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PDifferentiator : Proc
{
  Differentiator f;
  void loadState(const uint32_t* s) override { f._x1 = u2f(s[0]); }
  void storeState(uint32_t* s) const override { s[0] = f2u(f._x1); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PIntegrator : Proc
{
  Integrator f;
  void setCoefs(const float* c) override { f.mLeak = c[0]; }
  void loadState(const uint32_t* s) override { f.y1 = u2f(s[0]); }
  void storeState(uint32_t* s) const override { s[0] = f2u(f.y1); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};

// FDN<8> as shipped never allocates its IntegerDelay buffers; the
// oracle allocates them through the delays' own public setMaxDelayInSamples,
// reaching the private array with -fno-access-control.  Everything else is the
// reference's FDN<8>::operator() untouched.
struct PFDN8 : Proc
{
  FDN<8> f;
  void setCoefs(const float* c) override
  {
    for (int n = 0; n < 8; ++n)
    {
      f.mFilters[n].coeffs = {c[n], c[8 + n]};
      f.mFeedbackGains[n] = c[16 + n];
      int len = static_cast<int>(c[24 + n]);
      f.mDelays[n].setMaxDelayInSamples(static_cast<float>(len));
      f.mDelays[n].setDelayInSamples(len);
    }
  }
  void loadState(const uint32_t* s) override
  {
    for (int n = 0; n < 8; ++n) f.mFilters[n].y1 = u2f(s[n]);
  }
  void storeState(uint32_t* s) const override
  {
    for (int n = 0; n < 8; ++n) s[n] = f2u(f.mFilters[n].y1);
  }
  DSPVector run(const DSPVector* const* in, DSPVector* out2) override
  {
    DSPVectorArray<2> y = f(*in[0]);
    *out2 = y.constRow(1);  // sumR
    return y.constRow(0);   // sumL
  }
};


// ---- section 8(f) row 2: the rest of the L2 functor set ----
struct POneShot : Proc
{
  OneShotGen g;
  void loadState(const uint32_t* s) override { g.mOmega32 = s[0], g.mGate = s[1], g.mOmegaPrev = s[2]; }
  void storeState(uint32_t* s) const override { s[0] = g.mOmega32, s[1] = g.mGate, s[2] = g.mOmegaPrev; }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return g(*in[0]); }
};
struct PImpulse : Proc
{
  ImpulseGen g;
  void loadState(const uint32_t* s) override { g._omega = u2f(s[0]), g._outputCounter = (int)s[1]; }
  void storeState(uint32_t* s) const override { s[0] = f2u(g._omega), s[1] = (uint32_t)g._outputCounter; }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return g(*in[0]); }
};
struct PPeak : Proc
{
  Peak f;
  void setCoefs(const float* c) override
  {
    f.coeffs = {c[0], c[1]};
    f.peakHoldSamples = static_cast<int>(c[2]);
  }
  void loadState(const uint32_t* s) override { f.y1 = u2f(s[0]), f.peakHoldCounter = (int)s[1]; }
  void storeState(uint32_t* s) const override { s[0] = f2u(f.y1), s[1] = (uint32_t)f.peakHoldCounter; }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PRMS : Proc
{
  RMS f;
  void setCoefs(const float* c) override { f.coeffs = {c[0], c[1]}; }
  void loadState(const uint32_t* s) override { f.y1 = u2f(s[0]); }
  void storeState(uint32_t* s) const override { s[0] = f2u(f.y1); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PADSR : Proc
{
  ADSR f;
  void setCoefs(const float* c) override { f.coeffs = {c[0], c[1], c[2], c[3]}; }
  void loadState(const uint32_t* s) override
  {
    f.y = u2f(s[0]), f.y1 = u2f(s[1]), f.x1 = u2f(s[2]), f.threshold = u2f(s[3]);
    f.target = u2f(s[4]), f.k = u2f(s[5]), f.amp = u2f(s[6]), f.segment = (int)s[7];
  }
  void storeState(uint32_t* s) const override
  {
    s[0] = f2u(f.y), s[1] = f2u(f.y1), s[2] = f2u(f.x1), s[3] = f2u(f.threshold);
    s[4] = f2u(f.target), s[5] = f2u(f.k), s[6] = f2u(f.amp), s[7] = (uint32_t)f.segment;
  }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PAllpass1 : Proc
{
  Allpass1 f{0.f};
  void setCoefs(const float* c) override { f.coeffs = c[0]; }
  void loadState(const uint32_t* s) override { f.x1 = u2f(s[0]), f.y1 = u2f(s[1]); }
  void storeState(uint32_t* s) const override { s[0] = f2u(f.x1), s[1] = f2u(f.y1); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PGlide : Proc
{
  LinearGlide g;
  void setCoefs(const float* c) override
  {
    g.mVectorsPerGlide = static_cast<int>(c[0]);
    g.mDyPerVector = c[1];
  }
  void loadState(const uint32_t* s) override
  {
    g.mStepVec = DSPVector(u2f(s[0]));
    g.mTargetValue = u2f(s[1]);
    g.mVectorsRemaining = (int)s[2];
  }
  void storeState(uint32_t* s) const override
  {
    s[0] = f2u(g.mStepVec[0]), s[1] = f2u(g.mTargetValue), s[2] = (uint32_t)g.mVectorsRemaining;
  }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return g((*in[0])[0]); }
};
struct PInterp1 : Proc
{
  Interpolator1 g;
  void loadState(const uint32_t* s) override { g.currentValue = u2f(s[0]); }
  void storeState(uint32_t* s) const override { s[0] = f2u(g.currentValue); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return g((*in[0])[0]); }
};
struct PSampleGlide : Proc
{
  SampleAccurateLinearGlide g;
  void setCoefs(const float* c) override
  {
    g.mSamplesPerGlide = static_cast<int>(c[0]);
    g.mDyPerSample = c[1];
  }
  void loadState(const uint32_t* s) override
  {
    g.mCurrValue = u2f(s[0]), g.mStepValue = u2f(s[1]), g.mTargetValue = u2f(s[2]);
    g.mSamplesRemaining = (int)s[3];
  }
  void storeState(uint32_t* s) const override
  {
    s[0] = f2u(g.mCurrValue), s[1] = f2u(g.mStepValue), s[2] = f2u(g.mTargetValue);
    s[3] = (uint32_t)g.mSamplesRemaining;
  }
  DSPVector run(const DSPVector* const* in, DSPVector*) override
  {
    DSPVector y;
    for (int n = 0; n < kFloatsPerDSPVector; ++n) y[n] = g.nextSample((*in[0])[n]);
    return y;
  }
};
struct PIntDelay : Proc
{
  IntegerDelay d;
  void setCoefs(const float* c) override
  {
    d.setMaxDelayInSamples(c[1]);
    d.setDelayInSamples(static_cast<int>(c[0]));
  }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return d(*in[0]); }
};
struct PIntDelayVar : Proc
{
  IntegerDelay d;
  void setCoefs(const float* c) override { d.setMaxDelayInSamples(c[0]); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return d(*in[0], *in[1]); }
};
#define AP1_STATE(ap)                                                                   \
  void loadState(const uint32_t* s) override { (ap).x1 = u2f(s[0]), (ap).y1 = u2f(s[1]); } \
  void storeState(uint32_t* s) const override { s[0] = f2u((ap).x1), s[1] = f2u((ap).y1); }
struct PFracDelay : Proc
{
  FractionalDelay d;
  void setCoefs(const float* c) override
  {
    d.setMaxDelayInSamples(c[1]);
    d.setDelayInSamples(c[0]);
  }
  AP1_STATE(d.mAllpassSection)
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return d(*in[0]); }
};
struct PFracDelayVar : Proc
{
  FractionalDelay d;
  void setCoefs(const float* c) override { d.setMaxDelayInSamples(c[0]); }
  AP1_STATE(d.mAllpassSection)
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return d(*in[0], *in[1]); }
};
inline void pbLoad(PitchbendableDelay& d, const uint32_t* s)
{
  FractionalDelay* fd[2] = {&d.mDelay1, &d.mDelay2};
  for (int i = 0; i < 2; ++i)
  {
    fd[i]->mAllpassSection.x1 = u2f(s[4 * i]);
    fd[i]->mAllpassSection.y1 = u2f(s[4 * i + 1]);
    fd[i]->mIntegerDelay.mIntDelayInSamples = (int)s[4 * i + 2];
    fd[i]->mAllpassSection.coeffs = u2f(s[4 * i + 3]);
  }
}
inline void pbStore(const PitchbendableDelay& d, uint32_t* s)
{
  const FractionalDelay* fd[2] = {&d.mDelay1, &d.mDelay2};
  for (int i = 0; i < 2; ++i)
  {
    s[4 * i] = f2u(fd[i]->mAllpassSection.x1);
    s[4 * i + 1] = f2u(fd[i]->mAllpassSection.y1);
    s[4 * i + 2] = (uint32_t)fd[i]->mIntegerDelay.mIntDelayInSamples;
    s[4 * i + 3] = f2u(fd[i]->mAllpassSection.coeffs);
  }
}
struct PPitchbend : Proc
{
  PitchbendableDelay d;
  void setCoefs(const float* c) override { d.setMaxDelayInSamples(c[0]); }
  void loadState(const uint32_t* s) override { pbLoad(d, s); }
  void storeState(uint32_t* s) const override { pbStore(d, s); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return d(*in[0], *in[1]); }
};
struct PAllpassInt : Proc
{
  Allpass<IntegerDelay> f;
  void setCoefs(const float* c) override
  {
    f.mGain = c[0];
    f.setMaxDelayInSamples(c[2]);
    f.setDelayInSamples(c[1]);
  }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PAllpassFrac : Proc
{
  Allpass<FractionalDelay> f;
  void setCoefs(const float* c) override
  {
    f.mGain = c[0];
    f.setMaxDelayInSamples(c[2]);
    f.setDelayInSamples(c[1]);
  }
  AP1_STATE(f.mDelay.mAllpassSection)
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0]); }
};
struct PAllpassPB : Proc
{
  Allpass<PitchbendableDelay> f;
  void setCoefs(const float* c) override
  {
    f.mGain = c[0];
    f.setMaxDelayInSamples(c[1]);
  }
  void loadState(const uint32_t* s) override { pbLoad(f.mDelay, s); }
  void storeState(uint32_t* s) const override { pbStore(f.mDelay, s); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0], *in[1]); }
};
// HalfBandFilter (F:1245-1310): the four Allpass1 sections and b1
inline void hbLoad(HalfBandFilter& f, const uint32_t* s)
{
  Allpass1* ap[4] = {&f.apa0, &f.apa1, &f.apb0, &f.apb1};
  for (int i = 0; i < 4; ++i) ap[i]->x1 = u2f(s[2 * i]), ap[i]->y1 = u2f(s[2 * i + 1]);
  f.b1 = u2f(s[8]);
}
inline void hbStore(const HalfBandFilter& f, uint32_t* s)
{
  const Allpass1* ap[4] = {&f.apa0, &f.apa1, &f.apb0, &f.apb1};
  for (int i = 0; i < 4; ++i) s[2 * i] = f2u(ap[i]->x1), s[2 * i + 1] = f2u(ap[i]->y1);
  s[8] = f2u(f.b1);
}
struct PHalfBandUp : Proc
{
  HalfBandFilter f;
  void loadState(const uint32_t* s) override { hbLoad(f, s); }
  void storeState(uint32_t* s) const override { hbStore(f, s); }
  DSPVector run(const DSPVector* const* in, DSPVector* out2) override
  {
    DSPVector a = f.upsampleFirstHalf(*in[0]);
    *out2 = f.upsampleSecondHalf(*in[0]);
    return a;
  }
};
struct PHalfBandDown : Proc
{
  HalfBandFilter f;
  void loadState(const uint32_t* s) override { hbLoad(f, s); }
  void storeState(uint32_t* s) const override { hbStore(f, s); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f.downsample(*in[0], *in[1]); }
};
// Downsample2xFunction<1> split at its process function (MLDSPFunctional.h:166-223): the IN half performs
// the statements before fn(...), the OUT half those after it, on the reference's own member objects.
struct PDown2xIn : Proc
{
  HalfBandFilter downer;
  DSPVector inputBuffer;
  bool phase{false};
  void loadState(const uint32_t* s) override { hbLoad(downer, s), phase = s[9] != 0; }
  void storeState(uint32_t* s) const override { hbStore(downer, s), s[9] = phase ? 1u : 0u; }
  DSPVector run(const DSPVector* const* in, DSPVector*) override
  {
    DSPVector y;  // zero when there is nothing to hand to fn
    if (phase)
      y = downer.downsample(inputBuffer, *in[0]);
    else
      inputBuffer = *in[0];
    phase = !phase;
    return y;
  }
};
struct PDown2xOut : Proc
{
  HalfBandFilter upper;
  DSPVector outputBuffer;
  bool phase{false};
  void loadState(const uint32_t* s) override { hbLoad(upper, s), phase = s[9] != 0; }
  void storeState(uint32_t* s) const override { hbStore(upper, s), s[9] = phase ? 1u : 0u; }
  DSPVector run(const DSPVector* const* in, DSPVector*) override
  {
    DSPVector y;
    if (phase)
    {
      y = upper.upsampleFirstHalf(*in[0]);
      outputBuffer = upper.upsampleSecondHalf(*in[0]);
    }
    else
      y = outputBuffer;
    phase = !phase;
    return y;
  }
};
struct PTempoLock : Proc
{
  TempoLock f;
  float isr{0};
  void setCoefs(const float* c) override { isr = c[0]; }
  void loadState(const uint32_t* s) override { f._omega = u2f(s[0]), f._x1v = u2f(s[1]); }
  void storeState(uint32_t* s) const override { s[0] = f2u(f._omega), s[1] = f2u(f._x1v); }
  DSPVector run(const DSPVector* const* in, DSPVector*) override { return f(*in[0], (*in[1])[0], isr); }
};
// a DSPVector member kept between processVector calls (reverb.cpp:34,115-116)
struct PFeedback : Proc
{
  DSPVector held{0.f};
  DSPVector run(const DSPVector* const*, DSPVector*) override { return held; }
};

struct PStateless : Proc
{
  int op;
  explicit PStateless(int o) : op(o) {}
  DSPVector run(const DSPVector* const* in, DSPVector*) override
  {
    const DSPVector& a = *in[0];
    switch (op)
    {
      case MLB_OP_SQRT: return sqrt(a);
      case MLB_OP_SQRT_APPROX: return sqrtApprox(a);
      case MLB_OP_ABS: return abs(a);
      case MLB_OP_SIGN: return sign(a);
      case MLB_OP_SIGNBIT: return signBit(a);
      case MLB_OP_SIN: return sin(a);
      case MLB_OP_COS: return cos(a);
      case MLB_OP_LOG: return log(a);
      case MLB_OP_EXP: return exp(a);
      case MLB_OP_LOG2: return log2(a);
      case MLB_OP_EXP2: return exp2(a);
      case MLB_OP_SIN_APPROX: return sinApprox(a);
      case MLB_OP_COS_APPROX: return cosApprox(a);
      case MLB_OP_EXP_APPROX: return expApprox(a);
      case MLB_OP_LOG_APPROX: return logApprox(a);
      case MLB_OP_LOG2_APPROX: return log2Approx(a);
      case MLB_OP_EXP2_APPROX: return exp2Approx(a);
      case MLB_OP_FRACTIONAL_PART: return fractionalPart(a);
      case MLB_OP_ROUND_F2I: return asFloat(roundFloatToInt(a));
      case MLB_OP_TRUNC_F2I: return asFloat(truncateFloatToInt(a));
      case MLB_OP_INT_TO_FLOAT: return intToFloat(asInt(a));
      case MLB_OP_UNSIGNED_TO_FLOAT: return unsignedIntToFloat(asInt(a));
      default: break;
    }
    const DSPVector& b = *in[1];
    switch (op)
    {
      case MLB_OP_ADD: return add(a, b);
      case MLB_OP_SUBTRACT: return subtract(a, b);
      case MLB_OP_MULTIPLY: return multiply(a, b);
      case MLB_OP_DIVIDE: return divide(a, b);
      case MLB_OP_DIVIDE_APPROX: return divideApprox(a, b);
      case MLB_OP_POW: return pow(a, b);
      case MLB_OP_POW_APPROX: return powApprox(a, b);
      case MLB_OP_MIN: return min(a, b);
      case MLB_OP_MAX: return max(a, b);
      case MLB_OP_EQUAL: return asFloat(equal(a, b));
      case MLB_OP_NOT_EQUAL: return asFloat(notEqual(a, b));
      case MLB_OP_GREATER_THAN: return asFloat(greaterThan(a, b));
      case MLB_OP_GREATER_EQUAL: return asFloat(greaterThanOrEqual(a, b));
      case MLB_OP_LESS_THAN: return asFloat(lessThan(a, b));
      case MLB_OP_LESS_EQUAL: return asFloat(lessThanOrEqual(a, b));
      case MLB_OP_ADD_INT32: return asFloat(addInt32(asInt(a), asInt(b)));
      case MLB_OP_SUBTRACT_INT32: return asFloat(subtractInt32(asInt(a), asInt(b)));
      default: break;
    }
    const DSPVector& c = *in[2];
    switch (op)
    {
      case MLB_OP_LERP: return lerp(a, b, c);
      case MLB_OP_INVERSE_LERP: return inverseLerp(a, b, c);
      case MLB_OP_CLAMP: return clamp(a, b, c);
      case MLB_OP_WITHIN: return within(a, b, c);
      case MLB_OP_SELECT: return select(a, b, asInt(c));
      default: break;
    }
    return DSPVector(0.f);
  }
};

Proc* makeProc(int op)
{
  switch (op)
  {
    case MLB_OP_PARAM: return new PParam;
    case MLB_OP_NOISE: return new PNoise;
    case MLB_OP_PHASOR: return new PPhasor;
    case MLB_OP_SINE: return new PSine;
    case MLB_OP_SAW: return new PSaw;
    case MLB_OP_PULSE: return new PPulse;
    case MLB_OP_TICK: return new PTick;
    case MLB_OP_LOPASS: return new PLopass;
    case MLB_OP_HIPASS: return new PHipass;
    case MLB_OP_BANDPASS: return new PBandpass;
    case MLB_OP_LOSHELF: return new PLoShelf;
    case MLB_OP_HISHELF: return new PHiShelf;
    case MLB_OP_BELL: return new PBell;
    case MLB_OP_LOPASS_V: return new PLopassV;
    case MLB_OP_LOPASS_MOD: return new PLopassMod;
    case MLB_OP_LOSHELF_V: return new PLoShelfV;
    case MLB_OP_HISHELF_V: return new PHiShelfV;
    case MLB_OP_RAMP: return new PRamp;
    case MLB_OP_ONEPOLE: return new POnePole;
    case MLB_OP_DCBLOCKER: return new PDCBlocker;
    case MLB_OP_DIFFERENTIATOR: return new PDifferentiator;
    case MLB_OP_INTEGRATOR: return new PIntegrator;
    case MLB_OP_FDN8: return new PFDN8;
    case MLB_OP_ONESHOT: return new POneShot;
    case MLB_OP_IMPULSE: return new PImpulse;
    case MLB_OP_PEAK: return new PPeak;
    case MLB_OP_RMS: return new PRMS;
    case MLB_OP_ADSR: return new PADSR;
    case MLB_OP_ALLPASS1: return new PAllpass1;
    case MLB_OP_GLIDE: return new PGlide;
    case MLB_OP_INTERPOLATOR1: return new PInterp1;
    case MLB_OP_SAMPLE_GLIDE: return new PSampleGlide;
    case MLB_OP_INTEGER_DELAY: return new PIntDelay;
    case MLB_OP_INTEGER_DELAY_VAR: return new PIntDelayVar;
    case MLB_OP_FRACTIONAL_DELAY: return new PFracDelay;
    case MLB_OP_FRACTIONAL_DELAY_VAR: return new PFracDelayVar;
    case MLB_OP_PITCHBEND_DELAY: return new PPitchbend;
    case MLB_OP_ALLPASS_INT: return new PAllpassInt;
    case MLB_OP_ALLPASS_FRAC: return new PAllpassFrac;
    case MLB_OP_ALLPASS_PB: return new PAllpassPB;
    case MLB_OP_FEEDBACK_READ: return new PFeedback;
    case MLB_OP_HALFBAND_UP: return new PHalfBandUp;
    case MLB_OP_HALFBAND_DOWN: return new PHalfBandDown;
    case MLB_OP_TEMPO_LOCK: return new PTempoLock;
    case MLB_OP_DOWN2X_IN: return new PDown2xIn;
    case MLB_OP_DOWN2X_OUT: return new PDown2xOut;
    case MLB_OP_INPUT:
    case MLB_OP_FEEDBACK_WRITE:
    case MLB_OP_HALFBAND_UP_2:
    case MLB_OP_FDN8_R: return nullptr;
    default: return new PStateless(op);
  }
}

struct Graph
{
  std::vector<mlb_node> nodes;
  std::vector<int> outs;
  std::vector<int> stOff, coOff;
  std::vector<int> again;  // MLB_AGAIN target per node (the node whose functor OBJECT this node calls once more), or -1
  int nState{0}, nCoef{0}, nIn{0}, V{0};
  // procs[v][node]
  std::vector<std::vector<std::unique_ptr<Proc>>> procs;
};
}  // namespace

extern "C"
{
struct mlref_graph
{
  Graph g;
};

int mlref_abi_version() { return MLB_ABI_VERSION; }

// Create per-voice reference functors for a graph.  coef: [n_coef_words][V].
mlref_graph* mlref_graph_create(const mlb_node* nodes, int n_nodes, const int32_t* outs, int n_out,
                                int V, const float* coef)
{
  auto* h = new mlref_graph;
  Graph& g = h->g;
  g.nodes.assign(nodes, nodes + n_nodes);
  g.outs.assign(outs, outs + n_out);
  g.V = V;
  g.stOff.resize(n_nodes);
  g.coOff.resize(n_nodes);
  g.again.assign(n_nodes, -1);
  for (int i = 0; i < n_nodes; ++i)
  {
    OpInfo oi;
    if (!opInfo(nodes[i].op, oi))
    {
      delete h;
      return nullptr;
    }
    // MLB_AGAIN (mlb200.h): the same reference functor object is simply called again -- no second object
    if (nodes[i].iarg < 0 && nodes[i].op != MLB_OP_INPUT && nodes[i].op != MLB_OP_PARAM &&
        nodes[i].op != MLB_OP_FEEDBACK_WRITE)
    {
      const int t = MLB_AGAIN_TARGET(nodes[i].iarg);
      if (t >= i || nodes[t].op != nodes[i].op || g.again[t] >= 0 || (oi.nst == 0 && oi.nco == 0))
      {
        delete h;
        return nullptr;
      }
      g.again[i] = t;
      g.stOff[i] = g.stOff[t];
      g.coOff[i] = g.coOff[t];
      continue;
    }
    g.stOff[i] = g.nState;
    g.coOff[i] = g.nCoef;
    g.nState += oi.nst;
    g.nCoef += oi.nco;
    if (nodes[i].op == MLB_OP_INPUT) g.nIn = std::max(g.nIn, nodes[i].iarg + 1);
    if (nodes[i].op == MLB_OP_FEEDBACK_WRITE &&
        (nodes[i].iarg < 0 || nodes[i].iarg >= i || nodes[nodes[i].iarg].op != MLB_OP_FEEDBACK_READ))
    {
      delete h;
      return nullptr;
    }
  }
  g.procs.resize(V);
  std::vector<float> c(64);
  for (int v = 0; v < V; ++v)
  {
    g.procs[v].resize(n_nodes);
    for (int i = 0; i < n_nodes; ++i)
    {
      OpInfo oi;
      opInfo(nodes[i].op, oi);
      if (g.again[i] >= 0) continue;  // no object of its own
      g.procs[v][i].reset(makeProc(nodes[i].op));
      if (g.procs[v][i] && oi.nco > 0)
      {
        for (int k = 0; k < oi.nco; ++k) c[k] = coef[(size_t)(g.coOff[i] + k) * V + v];
        g.procs[v][i]->setCoefs(c.data());
      }
    }
  }
  return h;
}

void mlref_graph_destroy(mlref_graph* h) { delete h; }

void mlref_graph_set_state(mlref_graph* h, const uint32_t* state)
{
  Graph& g = h->g;
  uint32_t s[16];
  for (int v = 0; v < g.V; ++v)
    for (size_t i = 0; i < g.nodes.size(); ++i)
    {
      OpInfo oi;
      opInfo(g.nodes[i].op, oi);
      if (!g.procs[v][i] || oi.nst == 0) continue;
      for (int k = 0; k < oi.nst; ++k) s[k] = state[(size_t)(g.stOff[i] + k) * g.V + v];
      g.procs[v][i]->loadState(s);
    }
}

void mlref_graph_get_state(mlref_graph* h, uint32_t* state)
{
  Graph& g = h->g;
  uint32_t s[16];
  for (int v = 0; v < g.V; ++v)
    for (size_t i = 0; i < g.nodes.size(); ++i)
    {
      OpInfo oi;
      opInfo(g.nodes[i].op, oi);
      if (!g.procs[v][i] || oi.nst == 0) continue;
      g.procs[v][i]->storeState(s);
      for (int k = 0; k < oi.nst; ++k) state[(size_t)(g.stOff[i] + k) * g.V + v] = s[k];
    }
}

// Process T blocks.  in [T][n_in][V][64], out [T][n_out][V][64] (may be null),
// mix [T][n_out][64] (may be null): voices summed left to right, v = 0..V-1,
// exactly as addRows does (MLDSPOps.h:1349-1359).
void mlref_graph_process(mlref_graph* h, const float* in, float* out, float* mix, int T,
                         int nthreads)
{
  Graph& g = h->g;
  const int V = g.V;
  const int N = (int)g.nodes.size();
  const int nOut = (int)g.outs.size();
  if (nthreads < 1) nthreads = 1;
  nthreads = std::min(nthreads, std::max(1, V));

  // per-voice outputs are needed for the ordered mix; if the caller wants mix
  // but not out, use a scratch out.
  std::vector<float> scratch;
  float* o = out;
  if (!o && mix)
  {
    scratch.resize((size_t)T * nOut * V * 64);
    o = scratch.data();
  }

  auto worker = [&](int v0, int v1)
  {
    std::vector<DSPVector> rows(N);
    std::vector<DSPVector> rows2(N);  // second output (FDN8 sumR)
    for (int v = v0; v < v1; ++v)
    {
      for (int t = 0; t < T; ++t)
      {
        for (int i = 0; i < N; ++i)
        {
          const mlb_node& nd = g.nodes[i];
          if (nd.op == MLB_OP_INPUT)
          {
            rows[i] = DSPVector(in + (((size_t)t * g.nIn + nd.iarg) * V + v) * 64);
            continue;
          }
          if (nd.op == MLB_OP_FDN8_R || nd.op == MLB_OP_HALFBAND_UP_2)
          {
            rows[i] = rows2[nd.in[0]];
            continue;
          }
          if (nd.op == MLB_OP_FEEDBACK_WRITE)
          {
            static_cast<PFeedback*>(g.procs[v][nd.iarg].get())->held = rows[nd.in[0]];
            rows[i] = rows[nd.in[0]];
            continue;
          }
          const DSPVector* ins[MLB_MAX_INS] = {};
          for (int k = 0; k < MLB_MAX_INS; ++k)
            if (nd.in[k] >= 0) ins[k] = &rows[nd.in[k]];
          rows[i] = g.procs[v][g.again[i] >= 0 ? g.again[i] : i]->run(ins, &rows2[i]);
        }
        if (o)
          for (int c = 0; c < nOut; ++c)
            store(rows[g.outs[c]], o + (((size_t)t * nOut + c) * V + v) * 64);
      }
    }
  };

  if (nthreads == 1)
    worker(0, V);
  else
  {
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; ++i)
    {
      int v0 = (int)((long long)V * i / nthreads), v1 = (int)((long long)V * (i + 1) / nthreads);
      th.emplace_back(worker, v0, v1);
    }
    for (auto& t : th) t.join();
  }

  if (mix)
  {
    for (int t = 0; t < T; ++t)
      for (int c = 0; c < nOut; ++c)
      {
        DSPVector acc{0.f};
        for (int v = 0; v < V; ++v)
          acc = add(acc, DSPVector(o + (((size_t)t * nOut + c) * V + v) * 64));
        store(acc, mix + ((size_t)t * nOut + c) * 64);
      }
  }
}

// ---- coefficient design straight from the reference's makeCoeffs ----
void mlref_coeffs_lopass(float omega, float k, float* o)
{
  auto c = Lopass::makeCoeffs(omega, k);
  o[0] = c[0], o[1] = c[1], o[2] = c[2];
}
// Lopass::makeCoeffsVec, F:97-115; out = rows g0, g1, g2
void mlref_coeffs_lopass_vec(const float* omega, const float* k, float* out)
{
  auto vc = Lopass::makeCoeffsVec(DSPVector(omega), DSPVector(k));
  for (int i = 0; i < 3; ++i) store(vc.constRow(i), out + 64 * i);
}
// LoShelf / HiShelf::vcoeffs(p0, p1) = interpolateCoeffsLinear(makeCoeffs(p0), makeCoeffs(p1)), F:283-286,364-367
void mlref_loshelf_vcoeffs(const float* p0, const float* p1, float* out)
{
  auto vc = LoShelf::vcoeffs({p0[0], p0[1], p0[2]}, {p1[0], p1[1], p1[2]});
  for (int i = 0; i < 5; ++i) store(vc.constRow(i), out + 64 * i);
}
void mlref_hishelf_vcoeffs(const float* p0, const float* p1, float* out)
{
  auto vc = HiShelf::vcoeffs({p0[0], p0[1], p0[2]}, {p1[0], p1[1], p1[2]});
  for (int i = 0; i < 6; ++i) store(vc.constRow(i), out + 64 * i);
}
// Lopass::operator()(vx, omega, k) for ONE voice over T blocks, from cleared state: x, omega, k, out [T][64]
void mlref_lopass_mod(int T, const float* x, const float* omega, const float* k, float* out)
{
  Lopass lp;
  for (int t = 0; t < T; ++t)
    store(lp(DSPVector(x + 64 * t), DSPVector(omega + 64 * t), DSPVector(k + 64 * t)), out + 64 * t);
}
void mlref_coeffs_hipass(float omega, float k, float* o)
{
  auto c = Hipass::makeCoeffs(omega, k);
  o[0] = c.g0, o[1] = c.g1, o[2] = c.g2, o[3] = c.k;
}
void mlref_coeffs_bandpass(float omega, float k, float* o)
{
  auto c = Bandpass::makeCoeffs(omega, k);
  o[0] = c.g0, o[1] = c.g1, o[2] = c.g2;
}
void mlref_coeffs_loshelf(float omega, float k, float A, float* o)
{
  auto c = LoShelf::makeCoeffs({omega, k, A});
  for (int i = 0; i < 5; ++i) o[i] = c[i];
}
void mlref_coeffs_hishelf(float omega, float k, float A, float* o)
{
  auto c = HiShelf::makeCoeffs({omega, k, A});
  for (int i = 0; i < 6; ++i) o[i] = c[i];
}
void mlref_coeffs_bell(float omega, float k, float A, float* o)
{
  auto c = Bell::makeCoeffs(omega, k, A);
  o[0] = c.a1, o[1] = c.a2, o[2] = c.a3, o[3] = c.m1;
}
void mlref_coeffs_onepole(float omega, float* o)
{
  auto c = OnePole::makeCoeffs(omega);
  o[0] = c.a0, o[1] = c.b1;
}
float mlref_coeffs_dcblocker(float omega) { return DCBlocker::makeCoeffs(omega); }
float mlref_db_to_gain(float dB) { return dBToGain(dB); }
void mlref_coeffs_peak(float omega, float* o)
{
  auto c = Peak::makeCoeffs(omega);
  o[0] = c.a0, o[1] = c.b1;
}
void mlref_coeffs_rms(float omega, float* o)
{
  auto c = RMS::makeCoeffs(omega);
  o[0] = c.a0, o[1] = c.b1;
}
void mlref_coeffs_adsr(float a, float d, float s, float r, float sr, float* o)
{
  auto c = ADSR::calcCoeffs(a, d, s, r, sr);
  o[0] = c.ka, o[1] = c.kd, o[2] = c.s, o[3] = c.kr;
}
float mlref_coeffs_allpass1(float d) { return Allpass1::makeCoeffs(d); }
void mlref_impulse_table(float* o)
{
  ImpulseGen g;
  for (int i = 0; i < 17; ++i) o[i] = g._table[i];
}
void mlref_coeffs_glide(float timeInSamples, float* o)
{
  LinearGlide g;
  g.setGlideTimeInSamples(timeInSamples);
  o[0] = static_cast<float>(g.mVectorsPerGlide), o[1] = g.mDyPerVector;
}
void mlref_coeffs_sample_glide(float timeInSamples, float* o)
{
  SampleAccurateLinearGlide g;
  g.setGlideTimeInSamples(timeInSamples);
  o[0] = static_cast<float>(g.mSamplesPerGlide), o[1] = g.mDyPerSample;
}

// sizes the survey pins (Appendix A) -- lets a test check this really is the reference
int mlref_sizeof(int which)
{
  switch (which)
  {
    case 0: return (int)sizeof(DSPVector);
    case 1: return (int)sizeof(Lopass);
    case 2: return (int)sizeof(OnePole);
    case 3: return (int)sizeof(SineGen);
    case 4: return (int)sizeof(IntegerDelay);
    case 5: return (int)sizeof(FDN<8>);
  }
  return -1;
}

// ---- the reference's own chain loop, as a user would write it (CPU baseline) ----
// struct Voice{SineGen s; Lopass lp;};  out_v = lp(s(freq_v)) * gain
// (SURVEY 8d "CPU baseline beside it").  Contract R:
// in/out [T][V][64].  Returns seconds of wall time for the processing loop.
double mlref_chain_sine_lopass_gain(int V, int T, const float* in, float* out,
                                    const float* coef3 /*[3][V]*/, const float* gain /*[V]*/,
                                    uint32_t* phase /*[V] io*/, float* ic /*[2][V] io*/,
                                    int nthreads, int repeats)
{
  struct Voice
  {
    SineGen s;
    Lopass lp;
  };
  std::vector<Voice> voices(V);
  for (int v = 0; v < V; ++v)
  {
    voices[v].s._phasor.clear(phase[v]);
    voices[v].lp.coeffs = {coef3[v], coef3[V + v], coef3[2 * V + v]};
    voices[v].lp.ic1eq = ic[v];
    voices[v].lp.ic2eq = ic[V + v];
  }
  if (nthreads < 1) nthreads = 1;
  nthreads = std::min(nthreads, std::max(1, V));
  if (repeats < 1) repeats = 1;
  // `repeats` passes over the same T input blocks inside one thread launch (state carries on),
  // so that thread start-up is amortised when this loop is used as the CPU baseline.
  auto worker = [&](int v0, int v1)
  {
    for (int r = 0; r < repeats; ++r)
      for (int t = 0; t < T; ++t)
        for (int v = v0; v < v1; ++v)
        {
          const size_t off = ((size_t)t * V + v) * 64;
          DSPVector y = voices[v].lp(voices[v].s(DSPVector(in + off))) * DSPVector(gain[v]);
          store(y, out + off);
        }
  };
  auto t0 = std::chrono::steady_clock::now();
  if (nthreads == 1)
    worker(0, V);
  else
  {
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; ++i)
      th.emplace_back(worker, (int)((long long)V * i / nthreads),
                      (int)((long long)V * (i + 1) / nthreads));
    for (auto& t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  for (int v = 0; v < V; ++v)
  {
    phase[v] = voices[v].s._phasor.mOmega32;
    ic[v] = voices[v].lp.ic1eq;
    ic[V + v] = voices[v].lp.ic2eq;
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---- the Aaltoverb example's processVector, called as its author wrote it ----
// (examples/audio-and-midi/reverb.cpp:21-123; the example itself needs RtAudio, so the
// per-vector body is driven from here with the same functor members, the same
// expressions and the same evaluation order.)  One reverb; in/out [T][2][64].
// Used to check that graph_aaltoverb() really is that example, and as its CPU baseline.
double mlref_aaltoverb(int T, const float* in, float* out, float sizeU2, float feedback,
                       float glideSamples, int repeats)
{
  struct Verb
  {
    LinearGlide smoothFeedback, smoothDelay;
    Allpass<PitchbendableDelay> ap[10];
    PitchbendableDelay delayL, delayR;
    DSPVector fbL, fbR;
  } r;
  static const float gains[10] = {0.75f, 0.70f, 0.625f, 0.625f, 0.7f, 0.7f, 0.6f, 0.6f, 0.5f, 0.5f};
  static const float maxd[10] = {500.f, 500.f, 1000.f, 1000.f, 2600.f, 2600.f, 8000.f, 8000.f, 10000.f, 10000.f};
  r.smoothFeedback.setGlideTimeInSamples(glideSamples);
  r.smoothDelay.setGlideTimeInSamples(glideSamples);
  for (int i = 0; i < 10; ++i)
  {
    r.ap[i].mGain = gains[i];
    r.ap[i].setMaxDelayInSamples(maxd[i]);
  }
  r.delayL.setMaxDelayInSamples(3500.f);
  r.delayR.setMaxDelayInSamples(3500.f);
  const float sr = 48000;
  if (repeats < 1) repeats = 1;
  auto t0 = std::chrono::steady_clock::now();
  for (int rep = 0; rep < repeats; ++rep)
    for (int t = 0; t < T; ++t)
    {
      DSPVector in0(in + (size_t)t * 128), in1(in + (size_t)t * 128 + 64);
      DSPVector vSmoothDelay = r.smoothDelay(sizeU2);
      DSPVector vSmoothFeedback = r.smoothFeedback(feedback);
      DSPVector vMin(kFloatsPerDSPVector);
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
      DSPVector monoInput = (in0 + in1);
      DSPVector diffusedInput = r.ap[3](r.ap[2](r.ap[1](r.ap[0](monoInput, vt1), vt2), vt3), vt4);
      DSPVector vDelayTimeL = max(0.0313 * delayParamInSamples - vMin, DSPVector(0.f));
      DSPVector vDelayTimeR = max(0.0371 * delayParamInSamples - vMin, DSPVector(0.f));
      DSPVector vTapL = r.ap[6](r.ap[4](diffusedInput + r.delayL(r.fbL, vDelayTimeL), vt5), vt7);
      DSPVector vTapR = r.ap[7](r.ap[5](diffusedInput + r.delayR(r.fbR, vDelayTimeR), vt6), vt8);
      r.fbR = r.ap[8](vTapL, vt9) * vSmoothFeedback;
      r.fbL = r.ap[9](vTapR, vt10) * vSmoothFeedback;
      store(vTapL, out + (size_t)t * 128);
      store(vTapR, out + (size_t)t * 128 + 64);
    }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- tests/cpp/kitchen_body.h compiled against the reference itself: the SAME source the tracing layer compiles
// (tests/cpp/test_trace.cpp), so that "same spelling, same bits" is checked on a body that uses far more functors
// than the two examples.  One instance; in [T][2][64] (gate, freq rows), out [T][2][64].
}  // extern "C"
namespace kitchen_ref
{
using namespace ml;
#include "../../tests/cpp/kitchen_body.h"
struct Ctx
{
  DSPVectorDynamic inputs{2}, outputs{2};
};
}  // namespace kitchen_ref
extern "C"
{
void mlref_kitchen(int T, const float* in, float* out)
{
  kitchen_ref::KitchenState st;
  kitchen_ref::kitchenInit(st);
  kitchen_ref::Ctx ctx;
  for (int t = 0; t < T; ++t)
  {
    ctx.inputs[0] = DSPVector(in + (size_t)t * 128);
    ctx.inputs[1] = DSPVector(in + (size_t)t * 128 + 64);
    kitchen_ref::kitchenProcess(&ctx, &st);
    store(ctx.outputs[0], out + (size_t)t * 128);
    store(ctx.outputs[1], out + (size_t)t * 128 + 64);
  }
}

// ---- FDN<SIZE> itself for SIZE = 4, 6, 16 (MLDSPFilters.h:1162-1239), one voice; in [T][64], out [T][2][64] (sumL, sumR).
// The delays are sized for exactly their length, as for FDN<8> above (FDN never sizes them: SURVEY D7).  What the
// written-out graph of graph.graph_fdn(size) must equal.
}  // extern "C"
namespace
{
template <int SIZE>
void runFdn(int T, const float* in, float* out, const float* times, const float* cutoffs, const float* gains)
{
  FDN<SIZE> f;
  std::array<float, SIZE> t, c;
  for (int n = 0; n < SIZE; ++n) t[n] = times[n], c[n] = cutoffs[n], f.mFeedbackGains[n] = gains[n];
  f.setFilterCutoffs(c);
  for (int n = 0; n < SIZE; ++n)
  {
    int len = static_cast<int>(times[n] - kFloatsPerDSPVector);  // FDN::setDelaysInSamples, :1180-1184
    len = std::max(1, len);
    f.mDelays[n].setMaxDelayInSamples(static_cast<float>(len));
  }
  f.setDelaysInSamples(t);
  for (int b = 0; b < T; ++b)
  {
    DSPVectorArray<2> y = f(DSPVector(in + (size_t)b * 64));
    store(y.constRow(0), out + (size_t)b * 128);
    store(y.constRow(1), out + (size_t)b * 128 + 64);
  }
}
}  // namespace
extern "C"
{
int mlref_fdn_run(int size, int T, const float* in, float* out, const float* times, const float* cutoffs,
                  const float* gains)
{
  switch (size)
  {
    case 4: runFdn<4>(T, in, out, times, cutoffs, gains); return 0;
    case 6: runFdn<6>(T, in, out, times, cutoffs, gains); return 0;
    case 8: runFdn<8>(T, in, out, times, cutoffs, gains); return 0;
    case 16: runFdn<16>(T, in, out, times, cutoffs, gains); return 0;
    default: return 1;
  }
}

// ---- tests/cpp/fdn_body.h compiled against the reference itself (the tracing layer compiles the same file): FDN<4> and
// FDN<6>.  One instance; in [T][2][64], out [T][2][64].
}  // extern "C"
namespace fdn_ref
{
using namespace ml;
// the reference never sizes an FDN's delay lines: size them for exactly their length, as PFDN8 does
#define FDN_SIZE_DELAYS(fdn, times)                                                      \
  for (size_t n_ = 0; n_ < (times).size(); ++n_)                                         \
  (fdn).mDelays[n_].setMaxDelayInSamples(static_cast<float>(std::max(1, static_cast<int>((times)[n_] - kFloatsPerDSPVector))))
#include "../../tests/cpp/fdn_body.h"
#undef FDN_SIZE_DELAYS
struct Ctx
{
  DSPVectorDynamic inputs{2}, outputs{2};
};
}  // namespace fdn_ref
extern "C"
{
void mlref_fdn_body(int T, const float* in, float* out)
{
  fdn_ref::FdnState st;
  fdn_ref::fdnInit(st);
  fdn_ref::Ctx ctx;
  for (int t = 0; t < T; ++t)
  {
    ctx.inputs[0] = DSPVector(in + (size_t)t * 128);
    ctx.inputs[1] = DSPVector(in + (size_t)t * 128 + 64);
    fdn_ref::fdnProcess(&ctx, &st);
    store(ctx.outputs[0], out + (size_t)t * 128);
    store(ctx.outputs[1], out + (size_t)t * 128 + 64);
  }
}

// ---- tests/cpp/oversample_body.h compiled against the reference itself (the tracing layer compiles the same file):
// oversampled loops between an Upsampler and a Downsampler inside one process call.  in [T][64], out [T][2][64].
}  // extern "C"
namespace oversample_ref
{
using namespace ml;
#include "../../tests/cpp/oversample_body.h"
struct Ctx
{
  DSPVectorDynamic inputs{1}, outputs{2};
};
}  // namespace oversample_ref
extern "C"
{
void mlref_oversample_body(int T, const float* in, float* out)
{
  oversample_ref::OversampleState st;
  oversample_ref::oversampleInit(st);
  oversample_ref::Ctx ctx;
  for (int t = 0; t < T; ++t)
  {
    ctx.inputs[0] = DSPVector(in + (size_t)t * 64);
    oversample_ref::oversampleProcess(&ctx, &st);
    store(ctx.outputs[0], out + (size_t)t * 128);
    store(ctx.outputs[1], out + (size_t)t * 128 + 64);
  }
}

// ---- tests/cpp/rest_body.h compiled against the reference itself (the tracing layer compiles the same file): the functor
// spellings the other shared bodies leave out, FDN<8> included.  One instance; in [T][2][64], out [T][2][64].
}  // extern "C"
namespace rest_ref
{
using namespace ml;
#define FDN_SIZE_DELAYS(fdn, times)                                                      \
  for (size_t n_ = 0; n_ < (times).size(); ++n_)                                         \
  (fdn).mDelays[n_].setMaxDelayInSamples(static_cast<float>(std::max(1, static_cast<int>((times)[n_] - kFloatsPerDSPVector))))
#include "../../tests/cpp/rest_body.h"
#undef FDN_SIZE_DELAYS
struct Ctx
{
  DSPVectorDynamic inputs{2}, outputs{2};
};
}  // namespace rest_ref
extern "C"
{
void mlref_rest_body(int T, const float* in, float* out)
{
  rest_ref::RestState st;
  rest_ref::restInit(st);
  rest_ref::Ctx ctx;
  for (int t = 0; t < T; ++t)
  {
    ctx.inputs[0] = DSPVector(in + (size_t)t * 128);
    ctx.inputs[1] = DSPVector(in + (size_t)t * 128 + 64);
    rest_ref::restProcess(&ctx, &st);
    store(ctx.outputs[0], out + (size_t)t * 128);
    store(ctx.outputs[1], out + (size_t)t * 128 + 64);
  }
}

// ---- tests/cpp/rows_body.h compiled against the reference itself (the tracing layer compiles the same file):
// DSPVectorArray<ROWS> as a value, the row operations, Bank with array arguments.  in [T][64], out [T][2][64].
}  // extern "C"
namespace rows_ref
{
using namespace ml;
#include "../../tests/cpp/rows_body.h"
struct Ctx
{
  DSPVectorDynamic inputs{1}, outputs{2};
};
}  // namespace rows_ref
extern "C"
{
void mlref_rows_body(int T, const float* in, float* out)
{
  rows_ref::RowsState st;
  rows_ref::rowsInit(st);
  rows_ref::Ctx ctx;
  for (int t = 0; t < T; ++t)
  {
    ctx.inputs[0] = DSPVector(in + (size_t)t * 64);
    rows_ref::rowsProcess(&ctx, &st);
    store(ctx.outputs[0], out + (size_t)t * 128);
    store(ctx.outputs[1], out + (size_t)t * 128 + 64);
  }
}

// ---- tests/cpp/upsample_body.h compiled against the reference itself (the tracing layer compiles the same file):
// a process function with state run at twice the rate by Upsample2xFunction<1>, a stateless one at half the rate by
// Downsample2xFunction<1>.  One instance; in [T][2][64] (frequency, gate rows), out [T][2][64].
}  // extern "C"
namespace upsample_ref
{
using namespace ml;
#include "../../tests/cpp/upsample_body.h"
struct Ctx
{
  DSPVectorDynamic inputs{2}, outputs{2};
};
}  // namespace upsample_ref
extern "C"
{
void mlref_upsample_body(int T, const float* in, float* out)
{
  upsample_ref::UpsampleState st;
  upsample_ref::upsampleInit(st);
  upsample_ref::Ctx ctx;
  for (int t = 0; t < T; ++t)
  {
    ctx.inputs[0] = DSPVector(in + (size_t)t * 128);
    ctx.inputs[1] = DSPVector(in + (size_t)t * 128 + 64);
    upsample_ref::upsampleProcess(&ctx, &st);
    store(ctx.outputs[0], out + (size_t)t * 128);
    store(ctx.outputs[1], out + (size_t)t * 128 + 64);
  }
}

// ---- Upsample2xFunction<1> (MLDSPFunctional.h:114-160) called as a user would, with the stateless
// process function fn(v) = clamp(v * drive, -1, 1).  One voice; in/out [T][64].  Checks that the
// HALFBAND_UP / HALFBAND_UP_2 / HALFBAND_DOWN graph of workloads.functor_case("upsample2x_clip") is
// that higher-order function.
void mlref_upsample2x_clip(int T, const float* in, float* out, float drive)
{
  Upsample2xFunction<1> upper;
  for (int t = 0; t < T; ++t)
  {
    DSPVector x(in + (size_t)t * 64);
    DSPVector y = upper([&](const DSPVector v) { return clamp(v * DSPVector(drive), DSPVector(-1.f), DSPVector(1.f)); }, x);
    store(y, out + (size_t)t * 64);
  }
}

// ---- Upsample2xFunction<1> with a STATEFUL process function, fn(v) = lp(osc(v * 0.5)): the reference's own
// SineGen and Lopass objects, called twice per vector by the wrapper (MLDSPFunctional.h:138-140), as the tutorial
// does with a sine generator (examples/tutorial/dspOpsExample.cpp:100-102).  One voice; in/out [T][64]; phase0 =
// the SineGen's phasor word, g[3] = the Lopass coefficients.  What an MLB_AGAIN graph
// (workloads.functor_case("upsample2x_osc")) must equal.
void mlref_upsample2x_osc(int T, const float* in, float* out, uint32_t phase0, const float* g3)
{
  Upsample2xFunction<1> upper;
  SineGen osc;
  osc._phasor.clear(phase0);
  Lopass lp;
  lp.coeffs = {g3[0], g3[1], g3[2]};
  for (int t = 0; t < T; ++t)
  {
    DSPVector x(in + (size_t)t * 64);
    DSPVector y = upper([&](const DSPVector v) { return lp(osc(v * DSPVector(0.5f))); }, x);
    store(y, out + (size_t)t * 64);
  }
}

// ---- Upsampler / Downsampler (MLDSPFilters.h:1316-1473), one object per voice ----
struct mlref_resampler
{
  int dir, oct, V;
  std::vector<Upsampler> up;
  std::vector<Downsampler> down;
};
mlref_resampler* mlref_resampler_create(int dir, int oct, int V)
{
  auto* r = new mlref_resampler;
  r->dir = dir, r->oct = oct, r->V = V;
  for (int v = 0; v < V; ++v)
    if (dir == MLB_RESAMPLE_UP)
      r->up.emplace_back(oct);
    else
      r->down.emplace_back(oct);
  return r;
}
void mlref_resampler_destroy(mlref_resampler* r) { delete r; }
// in [T][V][64]; out [T_out][V][64]; returns T_out
int mlref_resampler_process(mlref_resampler* r, const float* in, float* out, int T)
{
  const int V = r->V, N = 1 << r->oct;
  int produced = 0;
  if (r->dir == MLB_RESAMPLE_UP)
  {
    for (int t = 0; t < T; ++t)
      for (int v = 0; v < V; ++v)
      {
        Upsampler& u = r->up[v];
        if (r->oct == 0)
        {
          std::memcpy(out + ((size_t)t * V + v) * 64, in + ((size_t)t * V + v) * 64, 256);  // no buffers at 0 octaves
          continue;
        }
        u.write(DSPVector(in + ((size_t)t * V + v) * 64));
        for (int k = 0; k < N; ++k) store(u.read(), out + (((size_t)t * N + k) * V + v) * 64);
      }
    produced = T * N;
  }
  else
  {
    for (int v = 0; v < V; ++v)
    {
      int n = 0;
      for (int t = 0; t < T; ++t)
        if (r->down[v].write(DSPVector(in + ((size_t)t * V + v) * 64)))
          store(r->down[v].read(), out + ((size_t)(n++) * V + v) * 64);
      produced = n;
    }
  }
  return produced;
}

// Downsample2xFunction<1> (MLDSPFunctional.h:166-223) called as a user would, fn(v) = clamp(v * drive, -1, 1)
void mlref_downsample2x_clip(int T, const float* in, float* out, float drive)
{
  Downsample2xFunction<1> downer;
  for (int t = 0; t < T; ++t)
  {
    DSPVector x(in + (size_t)t * 64);
    DSPVector y = downer([&](const DSPVector v) { return clamp(v * DSPVector(drive), DSPVector(-1.f), DSPVector(1.f)); }, x);
    store(y, out + (size_t)t * 64);
  }
}
}  // extern "C"
