// ops.cuh -- per-lane device restatement of madronalib's DSPVector arithmetic for sm_100a.
//
// One f32 lane of every SSE op of the reference == one scalar op per CUDA thread.
// EXACT mode (template<bool EX = true>) reproduces the reference's SSE2 rounding sequence
// bit for bit: every _mm_mul_ps/_mm_add_ps/_mm_sub_ps is a separately rounded RN operation
// (__fmul_rn/__fadd_rn/__fsub_rn are never contracted by nvcc), IEEE denormals are kept
// (the TU is built with -ftz=false -prec-div=true -prec-sqrt=true).  An FMA is used only
// where it is provably the same rounding (a power-of-two product feeding an add).
// FAST mode lets the compiler contract to FMA (stated tolerance).
//
// Citations: G = reference source/DSP/MLDSPGens.h, F = MLDSPFilters.h, O = MLDSPOps.h,
//            M = MLDSPMathSSE.h, S = MLDSPScalarMath.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mlb200.h"

namespace mlb
{
#define MLB_DEV __device__ __forceinline__

template <bool EX>
struct A
{
  static MLB_DEV float mul(float a, float b) { return EX ? __fmul_rn(a, b) : a * b; }
  static MLB_DEV float add(float a, float b) { return EX ? __fadd_rn(a, b) : a + b; }
  static MLB_DEV float sub(float a, float b) { return EX ? __fsub_rn(a, b) : a - b; }
  // a*b + c*d with two roundings of the products in exact mode
  static MLB_DEV float mul_add_mul(float a, float b, float c, float d)
  {
    if (EX) return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d));
    return fmaf(a, b, c * d);
  }
  // a*b + c
  static MLB_DEV float mad(float a, float b, float c)
  {
    if (EX) return __fadd_rn(__fmul_rn(a, b), c);
    return fmaf(a, b, c);
  }
};

MLB_DEV uint32_t f2u(float f) { return __float_as_uint(f); }
MLB_DEV float u2f(uint32_t u) { return __uint_as_float(u); }

// ---- L0 primitives (M:75-135, 221-241) ----

// _mm_cvtps_epi32 (M:124): RN-even; NaN / |x| >= 2^31 -> 0x80000000 (CUDA would saturate)
MLB_DEV int32_t cvt_round(float x)
{
  int32_t r = __float2int_rn(x);
  return (fabsf(x) < 2147483648.0f) ? r : (int32_t)0x80000000;
}
// _mm_cvttps_epi32 (M:125)
MLB_DEV int32_t cvt_trunc(float x)
{
  int32_t r = __float2int_rz(x);
  return (fabsf(x) < 2147483648.0f) ? r : (int32_t)0x80000000;
}
// vecUnsignedIntToFloat (M:130-135): t = (float)(int)(v >> 1); t + t   (t + t is exact)
MLB_DEV float unsigned_to_float(uint32_t v)
{
  float t = __int2float_rn((int32_t)(v >> 1));
  return __fadd_rn(t, t);
}
// _mm_min_ps/_mm_max_ps (M:80-81): second operand when unordered or equal
MLB_DEV float sse_min(float a, float b) { return a < b ? a : b; }
MLB_DEV float sse_max(float a, float b) { return a > b ? a : b; }
MLB_DEV float mask_f(bool c) { return u2f(c ? 0xFFFFFFFFu : 0u); }

// ---- precise transcendental functions: cephes via sse_mathfun ----

// vecLog, M:308-373
template <bool EX>
MLB_DEV float ml_log(float x)
{
  using a = A<EX>;
  const bool invalid = (x <= 0.0f);
  x = sse_max(x, u2f(0x00800000u));
  int32_t emm0 = (int32_t)(f2u(x) >> 23);
  x = u2f((f2u(x) & ~0x7f800000u) | 0x3f000000u);
  emm0 -= 0x7f;
  float e = __int2float_rn(emm0);
  e = a::add(e, 1.0f);
  const bool m = (x < 0.707106781186547524f);
  float tmp = m ? x : 0.0f;
  x = a::sub(x, 1.0f);
  e = a::sub(e, m ? 1.0f : 0.0f);
  x = a::add(x, tmp);
  float z = a::mul(x, x);
  float y = 7.0376836292E-2f;
  y = a::mad(y, x, -1.1514610310E-1f);
  y = a::mad(y, x, 1.1676998740E-1f);
  y = a::mad(y, x, -1.2420140846E-1f);
  y = a::mad(y, x, 1.4249322787E-1f);
  y = a::mad(y, x, -1.6668057665E-1f);
  y = a::mad(y, x, 2.0000714765E-1f);
  y = a::mad(y, x, -2.4999993993E-1f);
  y = a::mad(y, x, 3.3333331174E-1f);
  y = a::mul(y, x);
  y = a::mul(y, z);
  tmp = a::mul(e, -2.12194440e-4f);
  y = a::add(y, tmp);
  tmp = a::mul(z, 0.5f);
  y = a::sub(y, tmp);
  tmp = a::mul(e, 0.693359375f);
  x = a::add(x, y);
  x = a::add(x, tmp);
  return invalid ? u2f(0xFFFFFFFFu) : x;  // or with all-ones mask, M:371
}

// vecExp, M:389-440
template <bool EX>
MLB_DEV float ml_exp(float x)
{
  using a = A<EX>;
  x = sse_min(x, 88.3762626647949f);
  x = sse_max(x, -88.3762626647949f);
  float fx = a::mul(x, 1.44269504088896341f);
  fx = a::add(fx, 0.5f);
  int32_t emm0 = cvt_trunc(fx);
  float tmp = __int2float_rn(emm0);
  float mask = (tmp > fx) ? 1.0f : 0.0f;
  fx = a::sub(tmp, mask);
  tmp = a::mul(fx, 0.693359375f);
  float z = a::mul(fx, -2.12194440e-4f);
  x = a::sub(x, tmp);
  x = a::sub(x, z);
  z = a::mul(x, x);
  float y = 1.9875691500E-4f;
  y = a::mad(y, x, 1.3981999507E-3f);
  y = a::mad(y, x, 8.3334519073E-3f);
  y = a::mad(y, x, 4.1665795894E-2f);
  y = a::mad(y, x, 1.6666665459E-1f);
  y = a::mad(y, x, 5.0000001201E-1f);
  y = a::mad(y, z, x);
  y = a::add(y, 1.0f);
  emm0 = cvt_trunc(fx);
  uint32_t p2 = ((uint32_t)(emm0 + 0x7f)) << 23;
  return a::mul(y, u2f(p2));
}

// common tail of vecSin / vecCos, M:515-558
template <bool EX>
MLB_DEV float sincos_poly(float x, float y, bool use_sin_poly, uint32_t sign_bit)
{
  using a = A<EX>;
  float xmm1 = a::mul(y, -0.78515625f);
  float xmm2 = a::mul(y, -2.4187564849853515625e-4f);
  float xmm3 = a::mul(y, -3.77489497744594108e-8f);
  x = a::add(x, xmm1);
  x = a::add(x, xmm2);
  x = a::add(x, xmm3);
  float z = a::mul(x, x);
  float yc = 2.443315711809948E-005f;
  yc = a::mad(yc, z, -1.388731625493765E-003f);
  yc = a::mad(yc, z, 4.166664568298827E-002f);
  yc = a::mul(yc, z);
  yc = a::mul(yc, z);
  float tmp = a::mul(z, 0.5f);
  yc = a::sub(yc, tmp);
  yc = a::add(yc, 1.0f);
  float ys = -1.9515295891E-4f;
  ys = a::mad(ys, z, 8.3321608736E-3f);
  ys = a::mad(ys, z, -1.6666654611E-1f);
  ys = a::mul(ys, z);
  ys = a::mad(ys, x, x);
  // (mask & y2) + (~mask & y): one term is +0, so the add is exact except it turns -0 into +0
  float r = a::add(use_sin_poly ? 0.0f : yc, use_sin_poly ? ys : 0.0f);
  return u2f(f2u(r) ^ sign_bit);
}

// vecSin, M:479-559
template <bool EX>
MLB_DEV float ml_sin(float x)
{
  uint32_t sign_bit = f2u(x) & 0x80000000u;
  x = fabsf(x);
  float y = A<EX>::mul(x, 1.27323954473516f);
  uint32_t j = ((uint32_t)cvt_trunc(y) + 1u) & ~1u;
  y = __int2float_rn((int32_t)j);
  sign_bit ^= (j & 4u) << 29;
  return sincos_poly<EX>(x, y, (j & 2u) == 0u, sign_bit);
}

// vecCos, M:562-636
template <bool EX>
MLB_DEV float ml_cos(float x)
{
  x = fabsf(x);
  float y = A<EX>::mul(x, 1.27323954473516f);
  uint32_t j = ((uint32_t)cvt_trunc(y) + 1u) & ~1u;
  y = __int2float_rn((int32_t)j);
  j -= 2u;
  uint32_t sign_bit = (~j & 4u) << 29;
  return sincos_poly<EX>(x, y, (j & 2u) == 0u, sign_bit);
}

// ---- polynomial approximations, M:752-864 ----
template <bool EX>
MLB_DEV float ml_sin_approx(float x)
{
  using a = A<EX>;
  float x2 = a::mul(x, x);
  float p = a::mad(x2, 2.147840177713078446686267852783203125e-6f,
                   -1.92649182281456887722015380859375e-4f);
  p = a::mad(x2, p, 8.30897875130176544189453125e-3f);
  p = a::mad(x2, p, -0.166624367237091064453125f);
  p = a::mad(x2, p, 0.99997937679290771484375f);
  return a::mul(x, p);
}
template <bool EX>
MLB_DEV float ml_cos_approx(float x)
{
  using a = A<EX>;
  float x2 = a::mul(x, x);
  float p = a::mad(x2, 1.8791708498611114919185638427734375e-5f,
                   -1.33926304988563060760498046875e-3f);
  p = a::mad(x2, p, 4.1496001183986663818359375e-2f);
  p = a::mad(x2, p, -0.4997930824756622314453125f);
  p = a::mad(x2, p, 0.999959766864776611328125f);
  return p;
}
template <bool EX>
MLB_DEV float ml_exp_approx(float x)
{
  using a = A<EX>;
  float val2 = a::mad(x, 12102203.1615614f, 1065353216.f);
  float val3 = sse_min(val2, 2139095040.f);
  float val4 = sse_max(val3, 0.0f);
  uint32_t val4i = (uint32_t)cvt_trunc(val4);
  float xu = u2f(val4i & 0x7F800000u);
  float b = u2f((val4i & 0x7FFFFFu) | 0x3F800000u);
  float p = a::mad(b, 1.3671023382430374383648148e-2f, -2.88093587581985443087955e-3f);
  p = a::mad(b, p, 0.168143436463395944830000f);
  p = a::mad(b, p, 0.310670891004095530771135f);
  p = a::mad(b, p, 0.510397365625862338668154f);
  return a::mul(xu, p);
}
template <bool EX>
MLB_DEV float ml_log_approx(float val)
{
  using a = A<EX>;
  uint32_t vi = f2u(val);
  int32_t expi = (int32_t)(vi >> 23);
  float addcst = (val > 0.0f) ? -89.970756366f : u2f(0x00800000u);
  float x = u2f((vi & 0x7FFFFFu) | 0x3F800000u);
  float p = a::mad(x, 3.110401639e-2f, -0.288739945f);
  p = a::mad(x, p, 1.130626167f);
  p = a::mad(x, p, -2.461222105f);
  p = a::mad(x, p, 3.529304993f);
  float poly = a::mul(x, p);
  float acr = a::add(addcst, a::mul(0.69314718055995f, __int2float_rn(expi)));
  return a::add(poly, acr);
}

#define MLB_K_LOG_TWO 0.69314718055994529f    /* O:601 */
#define MLB_K_LOG_TWO_R 1.4426950408889634f   /* O:602 */

// ---- stateless ops on one lane (O:584-614, 640-649, 744-748, 796-856) ----
template <bool EX>
MLB_DEV float op_apply(int op, float x, float b, float c)
{
  using a = A<EX>;
  switch (op)
  {
    case MLB_OP_SQRT: return __fsqrt_rn(x);
    // _mm_rsqrt_ps is a CPU-defined 12-bit approximation: tolerance-only op
    case MLB_OP_SQRT_APPROX: return a::mul(x, rsqrtf(x));
    case MLB_OP_ABS: return u2f(f2u(x) & 0x7FFFFFFFu);
    case MLB_OP_SIGN:
      return u2f(((f2u(x) & 0x80000000u) | 0x3F800000u) & ((x != -0.0f) ? 0xFFFFFFFFu : 0u));
    case MLB_OP_SIGNBIT: return u2f((f2u(x) & 0x80000000u) | 0x3F800000u);
    case MLB_OP_SIN: return ml_sin<EX>(x);
    case MLB_OP_COS: return ml_cos<EX>(x);
    case MLB_OP_LOG: return ml_log<EX>(x);
    case MLB_OP_EXP: return ml_exp<EX>(x);
    case MLB_OP_LOG2: return a::mul(ml_log<EX>(x), MLB_K_LOG_TWO_R);
    case MLB_OP_EXP2: return ml_exp<EX>(a::mul(MLB_K_LOG_TWO, x));
    case MLB_OP_SIN_APPROX: return ml_sin_approx<EX>(x);
    case MLB_OP_COS_APPROX: return ml_cos_approx<EX>(x);
    case MLB_OP_EXP_APPROX: return ml_exp_approx<EX>(x);
    case MLB_OP_LOG_APPROX: return ml_log_approx<EX>(x);
    case MLB_OP_LOG2_APPROX: return a::mul(ml_log_approx<EX>(x), MLB_K_LOG_TWO_R);
    case MLB_OP_EXP2_APPROX: return ml_exp_approx<EX>(a::mul(MLB_K_LOG_TWO, x));
    case MLB_OP_FRACTIONAL_PART: return a::sub(x, __int2float_rn(cvt_trunc(x)));
    case MLB_OP_ROUND_F2I: return u2f((uint32_t)cvt_round(x));
    case MLB_OP_TRUNC_F2I: return u2f((uint32_t)cvt_trunc(x));
    case MLB_OP_INT_TO_FLOAT: return __int2float_rn((int32_t)f2u(x));
    case MLB_OP_UNSIGNED_TO_FLOAT: return unsigned_to_float(f2u(x));
    case MLB_OP_ADD: return a::add(x, b);
    case MLB_OP_SUBTRACT: return a::sub(x, b);
    case MLB_OP_MULTIPLY: return a::mul(x, b);
    case MLB_OP_DIVIDE: return __fdiv_rn(x, b);
    // _mm_rcp_ps is a CPU-defined 12-bit approximation: tolerance-only op
    case MLB_OP_DIVIDE_APPROX: return a::mul(x, __frcp_rn(b));
    case MLB_OP_POW: return ml_exp<EX>(a::mul(ml_log<EX>(x), b));
    case MLB_OP_POW_APPROX: return ml_exp_approx<EX>(a::mul(ml_log_approx<EX>(x), b));
    case MLB_OP_MIN: return sse_min(x, b);
    case MLB_OP_MAX: return sse_max(x, b);
    case MLB_OP_EQUAL: return mask_f(x == b);
    case MLB_OP_NOT_EQUAL: return mask_f(!(x == b));
    case MLB_OP_GREATER_THAN: return mask_f(x > b);
    case MLB_OP_GREATER_EQUAL: return mask_f(x >= b);
    case MLB_OP_LESS_THAN: return mask_f(x < b);
    case MLB_OP_LESS_EQUAL: return mask_f(x <= b);
    case MLB_OP_ADD_INT32: return u2f(f2u(x) + f2u(b));
    case MLB_OP_SUBTRACT_INT32: return u2f(f2u(x) - f2u(b));
    case MLB_OP_LERP: return a::add(x, a::mul(c, a::sub(b, x)));
    case MLB_OP_INVERSE_LERP: return __fdiv_rn(a::sub(c, x), a::sub(b, x));
    case MLB_OP_CLAMP: return sse_min(sse_max(x, b), c);
    case MLB_OP_WITHIN: return mask_f((x >= b) && (x < c));
    case MLB_OP_SELECT: return u2f((f2u(c) & f2u(x)) | (~f2u(c) & f2u(b)));
  }
  return 0.0f;
}

// ---- generators ----

// NoiseGen, G:115,132-142.  u*2 - 3 with u in [1,2): the product is exact, so one FMA is
// the same rounding as mul-then-sub.
MLB_DEV float noise_tick(uint32_t& seed)
{
  seed = seed * 0x0019660Du + 0x3C6EF35Fu;
  uint32_t temp = ((seed >> 9) & 0x007FFFFFu) | 0x3F800000u;
  return __fmaf_rn(u2f(temp), 2.f, -3.f);
}

// PhasorGen step, G:190-199: phase += RN(freq * 2^32) with the cvtps2dq overflow rule
template <bool EX>
MLB_DEV void phase_step(uint32_t& phase, float freq)
{
  float steps = A<EX>::mul(freq, 4294967296.0f);
  phase += (uint32_t)cvt_round(steps);
}
// PhasorGen output, G:202 + M:130-135: 2*(float)(int)(v>>1) * 2^-32.  Both scalings are
// exact powers of two, so (float)(int)(v>>1) * 2^-31 is the same value.
MLB_DEV float phase_to_phasor(uint32_t phase)
{
  return __fmul_rn(__int2float_rn((int32_t)(phase >> 1)), 4.656612873077392578125e-10f);
}

// phasorToSine, G:316-338, with the reference's constexpr-Newton sqrt2 = 0x1.6a0a0ap+0
// (S:224,230-235).  The first multiply is merged with the exact 2^-31 phasor
// scaling: RN(t * (domain * 2^-31)) == RN((t * 2^-31) * domain), no underflow possible.
#define MLB_K_SQRT2 0x1.6a0a0ap+0f        /* 1.41421568393707275390625 (bits 0x3fb50505), not sqrt(2) */
#define MLB_K_FLIP 0x1.6a0a0ap+1f         /* sqrt2 * 2 */
#define MLB_K_DOMAIN_2M31 0x1.6a0a0ap-29f /* (sqrt2 * 4) * 2^-31 */
#define MLB_K_INV_RANGE 0x1.0f876cp+0f    /* 1 / (sqrt2 - sqrt2^3 / 6) = 1.0606601238250732421875 */
#define MLB_K_ONE_SIXTH 0x1.555556p-3f    /* 1.0f / 6.f */
template <bool EX>
MLB_DEV float phase_to_sine(uint32_t phase)
{
  using a = A<EX>;
  float t = __int2float_rn((int32_t)(phase >> 1));
  float omega = a::add(a::mul(t, MLB_K_DOMAIN_2M31), -MLB_K_SQRT2);
  float tri = (omega > MLB_K_SQRT2) ? a::sub(MLB_K_FLIP, omega) : omega;
  float s = a::mul(MLB_K_INV_RANGE, tri);
  float q = a::sub(1.0f, a::mul(a::mul(tri, tri), MLB_K_ONE_SIXTH));
  return a::mul(s, q);
}

// polyBLEP, G:285-311
template <bool EX>
MLB_DEV float poly_blep(float t, float dt)
{
  using a = A<EX>;
  float c = 0.f;
  if (t < dt)
  {
    t = __fdiv_rn(t, dt);
    c = a::sub(a::sub(a::add(t, t), a::mul(t, t)), 1.0f);
  }
  else if (t > a::sub(1.0f, dt))
  {
    t = __fdiv_rn(a::sub(t, 1.0f), dt);
    c = a::add(a::add(a::add(a::mul(t, t), t), t), 1.0f);
  }
  return c;
}

// ---- filters: one sample ----

// Lopass / Hipass / Bandpass core, F:121-131.  ic += 2*t is one FMA (2*t is exact).
template <bool EX>
MLB_DEV void svf_g_core(float v0, float g0, float g1, float g2, float& ic1, float& ic2, float& v1,
                        float& v2)
{
  using a = A<EX>;
  float t0 = a::sub(v0, ic2);
  float t1 = a::mul_add_mul(g0, t0, g1, ic1);
  float t2 = a::mul_add_mul(g2, t0, g0, ic1);
  v1 = a::add(t1, ic1);
  v2 = a::add(t2, ic2);
  ic1 = __fmaf_rn(2.0f, t1, ic1);
  ic2 = __fmaf_rn(2.0f, t2, ic2);
}
// shelves / bell core, F:293-298.  ic = 2*v - ic is one FMA.
template <bool EX>
MLB_DEV void svf_a_core(float v0, float a1, float a2, float a3, float& ic1, float& ic2, float& v1,
                        float& v2)
{
  using a = A<EX>;
  float v3 = a::sub(v0, ic2);
  v1 = a::mul_add_mul(a1, ic1, a2, v3);
  v2 = a::add(a::add(ic2, a::mul(a2, ic1)), a::mul(a3, v3));
  ic1 = __fmaf_rn(2.0f, v1, -ic1);
  ic2 = __fmaf_rn(2.0f, v2, -ic2);
}

// per-sample tick of a stateful 1-in/1-out filter node; st/co point at the node's words
template <bool EX>
MLB_DEV float filter_tick(int op, float x, uint32_t* st, const float* co)
{
  using a = A<EX>;
  switch (op)
  {
    case MLB_OP_LOPASS:
    {
      float ic1 = u2f(st[0]), ic2 = u2f(st[1]), v1, v2;
      svf_g_core<EX>(x, co[0], co[1], co[2], ic1, ic2, v1, v2);
      st[0] = f2u(ic1), st[1] = f2u(ic2);
      return v2;
    }
    case MLB_OP_HIPASS:
    {
      float ic1 = u2f(st[0]), ic2 = u2f(st[1]), v1, v2;
      svf_g_core<EX>(x, co[0], co[1], co[2], ic1, ic2, v1, v2);
      st[0] = f2u(ic1), st[1] = f2u(ic2);
      return a::sub(a::sub(x, a::mul(co[3], v1)), v2);  // F:193
    }
    case MLB_OP_BANDPASS:
    {
      float ic1 = u2f(st[0]), ic2 = u2f(st[1]), v1, v2;
      svf_g_core<EX>(x, co[0], co[1], co[2], ic1, ic2, v1, v2);
      st[0] = f2u(ic1), st[1] = f2u(ic2);
      return v1;
    }
    case MLB_OP_LOSHELF:
    {
      float ic1 = u2f(st[0]), ic2 = u2f(st[1]), v1, v2;
      svf_a_core<EX>(x, co[0], co[1], co[2], ic1, ic2, v1, v2);
      st[0] = f2u(ic1), st[1] = f2u(ic2);
      return a::add(a::add(x, a::mul(co[3], v1)), a::mul(co[4], v2));  // F:299
    }
    case MLB_OP_HISHELF:
    {
      float ic1 = u2f(st[0]), ic2 = u2f(st[1]), v1, v2;
      svf_a_core<EX>(x, co[0], co[1], co[2], ic1, ic2, v1, v2);
      st[0] = f2u(ic1), st[1] = f2u(ic2);
      return a::add(a::mul_add_mul(co[3], x, co[4], v1), a::mul(co[5], v2));  // F:380
    }
    case MLB_OP_BELL:
    {
      float ic1 = u2f(st[0]), ic2 = u2f(st[1]), v1, v2;
      svf_a_core<EX>(x, co[0], co[1], co[2], ic1, ic2, v1, v2);
      st[0] = f2u(ic1), st[1] = f2u(ic2);
      return a::add(x, a::mul(co[3], v1));  // F:438
    }
    case MLB_OP_ONEPOLE:
    {
      float y1 = a::mul_add_mul(co[0], x, co[1], u2f(st[0]));  // F:471
      st[0] = f2u(y1);
      return y1;
    }
    case MLB_OP_DCBLOCKER:
    {
      float x1 = u2f(st[0]), y1 = u2f(st[1]);
      float y0 = a::add(a::sub(x, x1), a::mul(co[0], y1));  // F:506
      st[0] = f2u(x), st[1] = f2u(y0);
      return y0;
    }
    case MLB_OP_DIFFERENTIATOR:
    {
      float y = a::sub(x, u2f(st[0]));  // F:525,530
      st[0] = f2u(x);
      return y;
    }
    case MLB_OP_INTEGRATOR:
    {
      float y1 = u2f(st[0]);
      y1 = a::sub(y1, a::mul(y1, co[0]));  // F:552
      y1 = a::add(y1, x);                  // F:553
      st[0] = f2u(y1);
      return y1;
    }
  }
  return x;
}

// ---- filters whose coefficients are per-sample rows (the reference's modulated forms) ----

// Lopass::makeCoeffsVec for one sample, F:102-113, evaluated on the device.  Same operation sequence as
// the reference; sinf is CUDA's (<= 1 ulp), not glibc's: tolerance-only (LOPASS_MOD).
template <bool EX>
MLB_DEV void lopass_coeffs_device(float omega, float k, float& g0, float& g1, float& g2)
{
  using a = A<EX>;
  omega = sse_min(omega, 0.5f);  // F:101
  k = sse_max(k, 0.01f);         // F:102
  const float piOmega = a::mul(3.1415926535897932384626433f, omega);
  const float s1 = sinf(piOmega);
  const float s2 = sinf(a::mul(2.0f, piOmega));
  const float ks2 = a::mul(k, s2);
  const float nrm = __fdiv_rn(1.0f, a::add(2.f, ks2));
  const float s1s1x2 = a::mul(a::mul(2.0f, s1), s1);  // (2*s1)*s1: the doubling is exact, sign applied below
  g0 = a::mul(s2, nrm);
  g1 = a::mul(a::sub(-s1s1x2, ks2), nrm);
  g2 = a::mul(s1s1x2, nrm);
}

// cv = this sample's coefficient values in the order of the reference's coeffNames enums
template <bool EX>
MLB_DEV float vfilter_tick(int op, float x, uint32_t* st, const float* cv)
{
  using a = A<EX>;
  float ic1 = u2f(st[0]), ic2 = u2f(st[1]), v1, v2, y = x;
  switch (op)
  {
    case MLB_OP_LOPASS_V:  // F:141-149
      svf_g_core<EX>(x, cv[0], cv[1], cv[2], ic1, ic2, v1, v2);
      y = v2;
      break;
    case MLB_OP_LOPASS_MOD:
    {
      float g0, g1, g2;
      lopass_coeffs_device<EX>(cv[0], cv[1], g0, g1, g2);
      svf_g_core<EX>(x, g0, g1, g2, ic1, ic2, v1, v2);
      y = v2;
      break;
    }
    case MLB_OP_LOSHELF_V:  // F:309-316
      svf_a_core<EX>(x, cv[0], cv[1], cv[2], ic1, ic2, v1, v2);
      y = a::add(a::add(x, a::mul(cv[3], v1)), a::mul(cv[4], v2));
      break;
    case MLB_OP_HISHELF_V:  // F:390-397
      svf_a_core<EX>(x, cv[0], cv[1], cv[2], ic1, ic2, v1, v2);
      y = a::add(a::mul_add_mul(cv[3], x, cv[4], v1), a::mul(cv[5], v2));
      break;
  }
  st[0] = f2u(ic1), st[1] = f2u(ic2);
  return y;
}

// interpolateDSPVectorLinear(start, end)[n], O:986-990: columnIndex()*interval + (start + interval)
template <bool EX>
MLB_DEV float ramp_sample(float start, float end, int n)
{
  using a = A<EX>;
  const float interval = __fdiv_rn(a::sub(end, start), 64.0f);
  return a::add(a::mul((float)n, interval), a::add(start, interval));
}

// per-sample tick of a generator node.  in0 = freq, in1 = width (PULSE)
template <bool EX>
MLB_DEV float gen_tick(int op, float in0, float in1, uint32_t* st)
{
  using a = A<EX>;
  switch (op)
  {
    case MLB_OP_NOISE: return noise_tick(st[0]);
    case MLB_OP_PHASOR:
      phase_step<EX>(st[0], in0);
      return phase_to_phasor(st[0]);
    case MLB_OP_SINE:
      phase_step<EX>(st[0], in0);
      return phase_to_sine<EX>(st[0]);
    case MLB_OP_SAW:
    {  // G:362-369: saw = om*2 - 1 (exact product -> FMA); saw - blep
      phase_step<EX>(st[0], in0);
      float om = phase_to_phasor(st[0]);
      float saw = __fmaf_rn(om, 2.f, -1.f);
      return a::sub(saw, poly_blep<EX>(om, in0));
    }
    case MLB_OP_PULSE:
    {  // G:342-358
      phase_step<EX>(st[0], in0);
      float om = phase_to_phasor(st[0]);
      float p = (om >= in1) ? -1.f : 1.f;
      p = a::add(p, poly_blep<EX>(om, in0));
      float t = a::add(a::sub(om, in1), 1.0f);
      float down = a::sub(t, __int2float_rn(cvt_trunc(t)));
      return a::sub(p, poly_blep<EX>(down, in0));
    }
    case MLB_OP_TICK:
    {  // G:36-44
      float om = a::add(u2f(st[0]), in0);
      float y = 0.f;
      if (om > 1.0f)
      {
        om = a::sub(om, 1.0f);
        y = 1.0f;
      }
      st[0] = f2u(om);
      return y;
    }
  }
  return 0.f;
}

}  // namespace mlb
