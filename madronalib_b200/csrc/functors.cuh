// functors.cuh -- device restatement of the rest of madronalib's L2 functor set:
// OneShotGen, Peak, RMS, ADSR, Allpass1, LinearGlide, Interpolator1, SampleAccurateLinearGlide,
// IntegerDelay, FractionalDelay, PitchbendableDelay, Allpass<>.  One lane = one voice; scalar state
// in registers across the 64-sample loop; delay rings and 64-float member rows in the graph's
// delay memory (HBM), staged through shared-memory rows.
// Citations: G = source/DSP/MLDSPGens.h, F = source/DSP/MLDSPFilters.h of the reference.
#pragma once
#include "ops.cuh"
#include "tma.cuh"

namespace mlb
{
// ---- register-state functors: one call per sample, st/co = the node's state / coef words ----

// OneShotGen::operator(), G:235-252.  st: mOmega32, mGate, mOmegaPrev
template <bool EX>
MLB_DEV float oneshot_tick(float freq, uint32_t* st)
{
  const int32_t isteps = cvt_round(A<EX>::mul(freq, 4294967296.0f));
  uint32_t om = st[0] + (uint32_t)isteps * st[1];
  if (om < st[2])
  {
    st[1] = 0u;
    om = 0u;
  }
  st[0] = om, st[2] = om;
  return phase_to_phasor(om);  // unsignedIntToFloat(om) * 2^-32, both scalings exact
}

// ImpulseGen::operator(), G:82-102.  The 17-tap table (constructor, G:64-78) is built on the host with
// the reference's libm calls (mlb_impulse_table) and uploaded once.  st: _omega, _outputCounter
__constant__ float c_impulse_table[17];
template <bool EX>
MLB_DEV float impulse_tick(float freq, uint32_t* st)
{
  float om = A<EX>::add(u2f(st[0]), freq);
  int32_t counter = (int32_t)st[1];
  if (om > 1.0f)
  {
    om = A<EX>::sub(om, 1.0f);
    counter = 0;
  }
  float y = 0.f;
  if (counter < 17)
  {
    y = c_impulse_table[counter];
    counter++;
  }
  st[0] = f2u(om), st[1] = (uint32_t)counter;
  return y;
}

// tail of Peak / RMS, F:613,651: select(sqrtApprox(vy), 0, vy > 1e-20); sqrtApprox = x * rsqrt(x)
// (a CPU-defined 12-bit approximation on the reference side -- compared with a tolerance)
template <bool EX>
MLB_DEV float follower_out(float v)
{
  return (v > 1e-20f) ? A<EX>::mul(v, rsqrtf(v)) : 0.f;
}
// Peak::operator() sample loop, F:588-605 (the per-block counter step of F:607-610 is applied by
// the caller).  st: y1, peakHoldCounter; co: a0, b1, peakHoldSamples
template <bool EX>
MLB_DEV float peak_tick(float x, uint32_t* st, const float* co)
{
  const float xs = A<EX>::mul(x, x);
  float y1 = u2f(st[0]);
  if (xs > y1)
  {
    y1 = xs;
    st[1] = (uint32_t)cvt_trunc(co[2]);
  }
  else if ((int32_t)st[1] <= 0)
    y1 = A<EX>::mul_add_mul(co[0], xs, co[1], y1);
  st[0] = f2u(y1);
  return follower_out<EX>(y1);
}
// RMS::operator(), F:638-652
template <bool EX>
MLB_DEV float rms_tick(float x, uint32_t* st, const float* co)
{
  const float y1 = A<EX>::mul_add_mul(co[0], A<EX>::mul(x, x), co[1], u2f(st[0]));
  st[0] = f2u(y1);
  return follower_out<EX>(y1);
}

// ADSR::processSample, F:692-784.  st: y y1 x1 threshold target k amp segment; co: ka kd s kr
template <bool EX>
MLB_DEV float adsr_tick(float x, uint32_t* st, const float* co)
{
  using a = A<EX>;
  enum { SEG_A = 0, SEG_D = 1, SEG_S = 2, SEG_R = 3, SEG_OFF = 4 };
  int32_t segment = (int32_t)st[7];
  if (segment == SEG_OFF && x == 0.f) return 0.f;
  float y = u2f(st[0]), y1 = u2f(st[1]);
  const float x1 = u2f(st[2]);
  float threshold = u2f(st[3]), target = u2f(st[4]), k = u2f(st[5]), amp = u2f(st[6]);
  const bool crossed = ((y1 > threshold) != (y > threshold));
  bool recalc = false;
  if (crossed && segment < SEG_OFF)
  {
    segment++;
    recalc = true;
  }
  const bool trigOn = (x1 == 0.f) && (x > 0.f);
  const bool trigOff = (x1 > 0.f) && (x == 0.f);
  if (trigOn)
  {
    segment = SEG_A;
    amp = x;
    recalc = true;
  }
  else if (trigOff)
  {
    segment = SEG_R;
    recalc = true;
  }
  if (recalc)
  {
    const float sus = co[2];
    float startEnv = 0.f, endEnv = 0.f;
    switch (segment)
    {
      case SEG_A: startEnv = 0.f, endEnv = 1.f, k = co[0]; break;
      case SEG_D: startEnv = 1.f, endEnv = sus, k = co[1]; break;
      case SEG_S: startEnv = sus, endEnv = sus, k = 0.f, y1 = sus, y = sus; break;
      case SEG_R: startEnv = sus, endEnv = 0.f, k = co[3]; break;
      default: startEnv = 0.f, endEnv = 0.f, k = 0.f, y1 = 0.f, y = 0.f; break;
    }
    const float segmentBias = a::mul(a::sub(endEnv, startEnv), 0.1f);
    threshold = endEnv;
    target = a::add(endEnv, segmentBias);
  }
  y1 = y;
  y = a::add(y, a::mul(k, a::sub(target, y)));
  st[0] = f2u(y), st[1] = f2u(y1), st[2] = f2u(x), st[3] = f2u(threshold);
  st[4] = f2u(target), st[5] = f2u(k), st[6] = f2u(amp), st[7] = (uint32_t)segment;
  return a::mul(y, amp);
}

// Allpass1::processSample, F:944-952: y = x1 + (x - y1) * a
template <bool EX>
MLB_DEV float allpass1_tick(float x, float& x1, float& y1, float coeff)
{
  const float y = A<EX>::add(x1, A<EX>::mul(A<EX>::sub(x, y1), coeff));
  x1 = x;
  y1 = y;
  return y;
}

// SampleAccurateLinearGlide::nextSample, G:541-582.  st: curr step target remaining; co: per, dy
template <bool EX>
MLB_DEV float sample_glide_tick(float x, uint32_t* st, const float* co)
{
  float curr = u2f(st[0]), step = u2f(st[1]), target = u2f(st[2]);
  int32_t remaining = (int32_t)st[3];
  const int32_t per = cvt_trunc(co[0]);
  if (x != target)
  {
    target = x;
    remaining = per;
  }
  if (remaining < 0)
  {
  }
  else if (remaining == 0)
  {
    curr = target;
    step = 0.f;
    remaining--;
  }
  else if (remaining == per)
  {
    step = A<EX>::mul(A<EX>::sub(target, curr), co[1]);
    remaining--;
  }
  else
  {
    curr = A<EX>::add(curr, step);
    remaining--;
  }
  st[0] = f2u(curr), st[1] = f2u(step), st[2] = f2u(target), st[3] = (uint32_t)remaining;
  return curr;
}

// FractionalDelay::setDelayInSamples, F:993-1008, with Allpass1::makeCoeffs, F:936-941
template <bool EX>
MLB_DEV void frac_split(float d, int32_t& delayInt, float& apCoeff)
{
  using a = A<EX>;
  const float fDelayInt = floorf(d);
  int32_t di = cvt_trunc(fDelayInt);
  float frac = a::sub(d, fDelayInt);
  if ((frac < 0.618f) && (di > 0))
  {
    frac = a::add(frac, 1.f);
    di -= 1;
  }
  delayInt = di;
  const float xm1 = a::sub(frac, 1.f);
  apCoeff = a::add(a::mul(-0.53f, xm1), a::mul(a::mul(0.24f, xm1), xm1));
}

// kUnityRampVec, G:409-410: (i + 1) / 64
MLB_DEV float unity_ramp(int n) { return __int2float_rn(n + 1) * 0.015625f; }

// ---- IntegerDelay rings in delay memory ----
// IntegerDelay::setMaxDelayInSamples, F:822-830: mask of the ring of 1 << bitsToContain(dMax + 64)
// samples (bitsToContain, MLDSPScalarMath.h:31-36)
MLB_DEV uint32_t ring_mask_of(float max_delay)
{
  const int x = cvt_trunc(floorf(max_delay)) + MLB_BLOCK;
  const int e = (x <= 1) ? 0 : 32 - __clz(x - 1);
  return (1u << e) - 1u;
}

struct RingRef
{
  float* p;       // this lane's ring
  uint32_t mask;  // mLengthMask
  uint32_t w;     // mWriteIndex at the start of the block (a multiple of 64)
};

// a delay of d samples can make the per-sample loop read a slot of THIS block before the loop
// writes it (d mod R > R - 64): such lanes keep the ring's oldest block in a shared row
MLB_DEV bool delay_reads_ahead(int32_t d, uint32_t mask) { return ((uint32_t)d & mask) > mask - 63u; }

// ---- HalfBandFilter, F:1245-1310: two 2-section allpass branches (coefficients F:1305-1306) ----
struct HalfBand
{
  float s[9];  // apa0{x1,y1} apa1{x1,y1} apb0{x1,y1} apb1{x1,y1} b1
  template <bool EX>
  MLB_DEV float a(float x)
  {
    return allpass1_tick<EX>(allpass1_tick<EX>(x, s[0], s[1], 0.07986642623635751f), s[2], s[3], 0.5453536510711322f);
  }
  template <bool EX>
  MLB_DEV float b(float x)
  {
    return allpass1_tick<EX>(allpass1_tick<EX>(x, s[4], s[5], 0.28382934487410993f), s[6], s[7], 0.8344118914807379f);
  }
};

// TempoLock::operator(), F:1494-1578 -- everything before the per-sample loop.  Returns false when
// the input clock is stopped (output row = 0); otherwise omega / dydt drive the loop.
template <bool EX>
MLB_DEV bool tempo_lock_prepare(float x0, float x1s, float dydx, float isr, float& omega, float& x1v, float& dydt)
{
  using a = A<EX>;
  if (x0 == -1.0f)
  {
    omega = -1.0f;
    return false;
  }
  float dxdt;
  if (omega > -1.f)
  {
    float dx = a::sub(x0, x1v);
    if (dx < 0.f) dx = a::add(dx, 1.f);
    dxdt = dx * 0.015625f;  // dx / 64, exact
    dydt = a::mul(dxdt, dydx);
    x1v = x0;
  }
  else
  {
    dxdt = a::sub(x1s, x0);
    dydt = a::mul(dxdt, dydx);
    x1v = a::sub(x0, dxdt * 64.f);
    omega = fmodf(a::mul(x0, dydx), 1.0f);
  }
  bool lock = fabsf(a::sub(dydx, roundf(dydx))) < 0.001f;
  const float rdydx = 1.0f / dydx;
  if (fabsf(a::sub(rdydx, roundf(rdydx))) < 0.001f) lock = true;
  if (lock)
  {
    float error;
    if (dydx >= 1.f)
    {
      const float ref = a::mul(x0, dydx);
      error = a::sub(omega, a::sub(ref, floorf(ref)));
    }
    else
    {
      const float ref = omega / dydx;
      error = a::sub(a::sub(ref, floorf(ref)), x0);
    }
    const float errorDiff = a::sub(roundf(error), error);
    float correction = a::mul(a::mul(errorDiff, isr), 4.0f);
    const float lo = -dydt * 0.5f, hi = dydt;
    correction = (correction < lo) ? lo : (correction > hi ? hi : correction);  // ml::clamp
    dydt = a::add(dydt, correction);
  }
  return true;
}

}  // namespace mlb
