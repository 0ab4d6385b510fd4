/*
 * oracle/port/mlport.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, CPU-only restatement of madronalib's DSPVector hot path: every
 * function follows the cited reference file:line, one f32 lane at a time
 * (each SSE op on 4 lanes == the same scalar op on each lane).
 * It exists so that the CUDA kernels can be checked on a machine where
 * /root/reference does not exist (the GPU box).  It is pinned against the
 * compiled reference itself (oracle/_ref/libmlref.so) and against the
 * committed golden vectors by tests/test_oracle_*.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  madronalib_b200/ never does.
 *
 * Build: gcc -O2 -ffp-contract=off -msse2 -mfpmath=sse (oracle/Makefile) so
 * that every mul/add is a separately rounded IEEE binary32 operation, like the
 * SSE2 instructions the reference compiles to.
 *
 * Abbreviations: G = source/DSP/MLDSPGens.h, F = source/DSP/MLDSPFilters.h,
 * O = source/DSP/MLDSPOps.h, M = source/DSP/MLDSPMathSSE.h,
 * S = source/DSP/MLDSPScalarMath.h
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h> /* only for _mm_rcp_ss/_mm_rsqrt_ss: the 12-bit hardware approximations */

#include "mlb200.h"

#define NB MLB_BLOCK
#define K_TWO_PI_F 6.2831853071795864769252867f /* ml::kTwoPi, S:23 */

static inline uint32_t f2u(float f)
{
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float u2f(uint32_t u)
{
  float f;
  memcpy(&f, &u, 4);
  return f;
}

/* ------------------------------------------------------------------ */
/* L0 primitives (M:75-135, 221-241)                                   */

/* _mm_cvtps_epi32 (M:124): round to nearest even; NaN / out of range ->
 * "integer indefinite" 0x80000000. */
static inline int32_t cvt_round(float x)
{
  if (!(fabsf(x) < 2147483648.0f)) return INT32_MIN;
  return (int32_t)rintf(x); /* default rounding mode = RN-even */
}
/* _mm_cvttps_epi32 (M:125): truncate, same overflow rule */
static inline int32_t cvt_trunc(float x)
{
  if (!(fabsf(x) < 2147483648.0f)) return INT32_MIN;
  return (int32_t)x;
}
/* vecUnsignedIntToFloat (M:130-135): t = (float)(int)(v >> 1); return t + t */
static inline float unsigned_to_float(uint32_t v)
{
  float t = (float)(int32_t)(v >> 1);
  return t + t;
}
/* _mm_min_ps / _mm_max_ps (M:80-81): second operand when unordered or equal */
static inline float sse_min(float a, float b) { return a < b ? a : b; }
static inline float sse_max(float a, float b) { return a > b ? a : b; }
static inline float mask_f(int c) { return u2f(c ? 0xFFFFFFFFu : 0u); }
static inline float sel_bits(float a, float b, uint32_t m) /* vecSelect M:221-241 */
{
  return u2f((m & f2u(a)) | (~m & f2u(b)));
}
static inline float sse_rcp(float x) { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }
static inline float sse_rsqrt(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }

/* ------------------------------------------------------------------ */
/* precise transcendental functions (cephes via sse_mathfun)            */

/* vecLog, M:308-373 */
static float ml_log(float x)
{
  const int invalid = (x <= 0.0f);
  x = sse_max(x, u2f(0x00800000u)); /* cut off denormals (M:314) */
  int32_t emm0 = (int32_t)(f2u(x) >> 23);
  x = u2f((f2u(x) & ~0x7f800000u) | f2u(0.5f));
  emm0 -= 0x7f;
  float e = (float)emm0;
  e = e + 1.0f;
  const int m = (x < 0.707106781186547524f);
  float tmp = m ? x : 0.0f;
  x = x - 1.0f;
  e = e - (m ? 1.0f : 0.0f);
  x = x + tmp;
  float z = x * x;
  float y = 7.0376836292E-2f;
  y = y * x;
  y = y + -1.1514610310E-1f;
  y = y * x;
  y = y + 1.1676998740E-1f;
  y = y * x;
  y = y + -1.2420140846E-1f;
  y = y * x;
  y = y + 1.4249322787E-1f;
  y = y * x;
  y = y + -1.6668057665E-1f;
  y = y * x;
  y = y + 2.0000714765E-1f;
  y = y * x;
  y = y + -2.4999993993E-1f;
  y = y * x;
  y = y + 3.3333331174E-1f;
  y = y * x;
  y = y * z;
  tmp = e * -2.12194440e-4f;
  y = y + tmp;
  tmp = z * 0.5f;
  y = y - tmp;
  tmp = e * 0.693359375f;
  x = x + y;
  x = x + tmp;
  if (invalid) x = u2f(f2u(x) | 0xFFFFFFFFu); /* M:371 */
  return x;
}

/* vecExp, M:389-440 */
static float ml_exp(float x)
{
  x = sse_min(x, 88.3762626647949f);
  x = sse_max(x, -88.3762626647949f);
  float fx = x * 1.44269504088896341f;
  fx = fx + 0.5f;
  int32_t emm0 = cvt_trunc(fx);
  float tmp = (float)emm0;
  float mask = (tmp > fx) ? 1.0f : 0.0f;
  fx = tmp - mask;
  tmp = fx * 0.693359375f;
  float z = fx * -2.12194440e-4f;
  x = x - tmp;
  x = x - z;
  z = x * x;
  float y = 1.9875691500E-4f;
  y = y * x;
  y = y + 1.3981999507E-3f;
  y = y * x;
  y = y + 8.3334519073E-3f;
  y = y * x;
  y = y + 4.1665795894E-2f;
  y = y * x;
  y = y + 1.6666665459E-1f;
  y = y * x;
  y = y + 5.0000001201E-1f;
  y = y * z;
  y = y + x;
  y = y + 1.0f;
  emm0 = cvt_trunc(fx);
  emm0 = (int32_t)((uint32_t)(emm0 + 0x7f) << 23);
  return y * u2f((uint32_t)emm0);
}

/* shared tail of vecSin / vecCos: M:515-558 / M:591-635 */
static float sincos_poly(float x, float y, uint32_t poly_mask, uint32_t sign_bit)
{
  float xmm1 = y * -0.78515625f;
  float xmm2 = y * -2.4187564849853515625e-4f;
  float xmm3 = y * -3.77489497744594108e-8f;
  x = x + xmm1;
  x = x + xmm2;
  x = x + xmm3;
  y = 2.443315711809948E-005f;
  float z = x * x;
  y = y * z;
  y = y + -1.388731625493765E-003f;
  y = y * z;
  y = y + 4.166664568298827E-002f;
  y = y * z;
  y = y * z;
  float tmp = z * 0.5f;
  y = y - tmp;
  y = y + 1.0f;
  float y2 = -1.9515295891E-4f;
  y2 = y2 * z;
  y2 = y2 + 8.3321608736E-3f;
  y2 = y2 * z;
  y2 = y2 + -1.6666654611E-1f;
  y2 = y2 * z;
  y2 = y2 * x;
  y2 = y2 + x;
  y2 = u2f(poly_mask & f2u(y2));
  y = u2f(~poly_mask & f2u(y));
  y = y + y2;
  return u2f(f2u(y) ^ sign_bit);
}

/* vecSin, M:479-559 */
static float ml_sin(float x)
{
  uint32_t sign_bit = f2u(x) & 0x80000000u;
  x = u2f(f2u(x) & 0x7FFFFFFFu);
  float y = x * 1.27323954473516f;
  int32_t emm2 = cvt_trunc(y);
  emm2 = (int32_t)(((uint32_t)emm2 + 1u) & ~1u);
  y = (float)emm2;
  uint32_t emm0 = ((uint32_t)emm2 & 4u) << 29;
  uint32_t poly_mask = (((uint32_t)emm2 & 2u) == 0u) ? 0xFFFFFFFFu : 0u;
  sign_bit ^= emm0;
  return sincos_poly(x, y, poly_mask, sign_bit);
}

/* vecCos, M:562-636 */
static float ml_cos(float x)
{
  x = u2f(f2u(x) & 0x7FFFFFFFu);
  float y = x * 1.27323954473516f;
  int32_t emm2 = cvt_trunc(y);
  emm2 = (int32_t)(((uint32_t)emm2 + 1u) & ~1u);
  y = (float)emm2;
  emm2 = (int32_t)((uint32_t)emm2 - 2u);
  uint32_t emm0 = (~(uint32_t)emm2 & 4u) << 29;
  uint32_t poly_mask = (((uint32_t)emm2 & 2u) == 0u) ? 0xFFFFFFFFu : 0u;
  return sincos_poly(x, y, poly_mask, emm0);
}

/* ------------------------------------------------------------------ */
/* polynomial approximations, M:752-864                                 */

static float ml_sin_approx(float x) /* M:758-772 */
{
  float x2 = x * x;
  return x * (0.99997937679290771484375f +
              x2 * (-0.166624367237091064453125f +
                    x2 * (8.30897875130176544189453125e-3f +
                          x2 * (-1.92649182281456887722015380859375e-4f +
                                x2 * 2.147840177713078446686267852783203125e-6f))));
}
static float ml_cos_approx(float x) /* M:780-792 */
{
  float x2 = x * x;
  return 0.999959766864776611328125f +
         x2 * (-0.4997930824756622314453125f +
               x2 * (4.1496001183986663818359375e-2f +
                     x2 * (-1.33926304988563060760498046875e-3f +
                           x2 * 1.8791708498611114919185638427734375e-5f)));
}
static float ml_exp_approx(float x) /* M:802-829 */
{
  float val2 = x * 12102203.1615614f + 1065353216.f;
  float val3 = sse_min(val2, 2139095040.f);
  float val4 = sse_max(val3, 0.0f);
  uint32_t val4i = (uint32_t)cvt_trunc(val4);
  float xu = u2f(val4i & 0x7F800000u);
  float b = u2f((val4i & 0x7FFFFFu) | 0x3F800000u);
  return xu * (0.510397365625862338668154f +
               b * (0.310670891004095530771135f +
                    b * (0.168143436463395944830000f +
                         b * (-2.88093587581985443087955e-3f +
                              b * 1.3671023382430374383648148e-2f))));
}
static float ml_log_approx(float val) /* M:839-864 */
{
  uint32_t vi = f2u(val);
  int32_t expi = (int32_t)(vi >> 23);
  /* vecSelect(kLogC1Vec, FLT_MIN, val > 0) */
  float addcst = (val > 0.0f) ? -89.970756366f : u2f(0x00800000u);
  float x = u2f((vi & 0x7FFFFFu) | 0x3F800000u);
  float poly = x * (3.529304993f +
                    x * (-2.461222105f +
                         x * (1.130626167f + x * (-0.288739945f + x * 3.110401639e-2f))));
  float addCstResult = addcst + 0.69314718055995f * (float)expi;
  return poly + addCstResult;
}

/* O:601-604,613-614 */
#define K_LOG_TWO 0.69314718055994529f
#define K_LOG_TWO_R 1.4426950408889634f

/* ------------------------------------------------------------------ */
/* stateless ops on one lane (O:584-614, 640-649, 744-748, 796-856)     */

static float op1(int op, float x)
{
  switch (op)
  {
    case MLB_OP_SQRT: return sqrtf(x);                  /* vecSqrt M:83 */
    case MLB_OP_SQRT_APPROX: return x * sse_rsqrt(x);   /* M:84-85 */
    case MLB_OP_ABS: return u2f(f2u(x) & 0x7FFFFFFFu);  /* M:86 */
    case MLB_OP_SIGN:                                   /* M:88-90 */
      return u2f(((f2u(x) & 0x80000000u) | 0x3F800000u) & (x != -0.0f ? 0xFFFFFFFFu : 0u));
    case MLB_OP_SIGNBIT: return u2f((f2u(x) & 0x80000000u) | 0x3F800000u); /* M:92 */
    case MLB_OP_SIN: return ml_sin(x);
    case MLB_OP_COS: return ml_cos(x);
    case MLB_OP_LOG: return ml_log(x);
    case MLB_OP_EXP: return ml_exp(x);
    case MLB_OP_LOG2: return ml_log(x) * K_LOG_TWO_R;
    case MLB_OP_EXP2: return ml_exp(K_LOG_TWO * x);
    case MLB_OP_SIN_APPROX: return ml_sin_approx(x);
    case MLB_OP_COS_APPROX: return ml_cos_approx(x);
    case MLB_OP_EXP_APPROX: return ml_exp_approx(x);
    case MLB_OP_LOG_APPROX: return ml_log_approx(x);
    case MLB_OP_LOG2_APPROX: return ml_log_approx(x) * K_LOG_TWO_R;
    case MLB_OP_EXP2_APPROX: return ml_exp_approx(K_LOG_TWO * x);
    case MLB_OP_FRACTIONAL_PART: return x - (float)cvt_trunc(x); /* O:825 */
    case MLB_OP_ROUND_F2I: return u2f((uint32_t)cvt_round(x));
    case MLB_OP_TRUNC_F2I: return u2f((uint32_t)cvt_trunc(x));
    case MLB_OP_INT_TO_FLOAT: return (float)(int32_t)f2u(x);
    case MLB_OP_UNSIGNED_TO_FLOAT: return unsigned_to_float(f2u(x));
  }
  return 0.0f;
}

static float op2(int op, float a, float b)
{
  switch (op)
  {
    case MLB_OP_ADD: return a + b;
    case MLB_OP_SUBTRACT: return a - b;
    case MLB_OP_MULTIPLY: return a * b;
    case MLB_OP_DIVIDE: return a / b;
    case MLB_OP_DIVIDE_APPROX: return a * sse_rcp(b);             /* M:79 */
    case MLB_OP_POW: return ml_exp(ml_log(a) * b);                /* O:646 */
    case MLB_OP_POW_APPROX: return ml_exp_approx(ml_log_approx(a) * b);
    case MLB_OP_MIN: return sse_min(a, b);
    case MLB_OP_MAX: return sse_max(a, b);
    case MLB_OP_EQUAL: return mask_f(a == b);
    case MLB_OP_NOT_EQUAL: return mask_f(!(a == b)); /* cmpneq: true when unordered */
    case MLB_OP_GREATER_THAN: return mask_f(a > b);
    case MLB_OP_GREATER_EQUAL: return mask_f(a >= b);
    case MLB_OP_LESS_THAN: return mask_f(a < b);
    case MLB_OP_LESS_EQUAL: return mask_f(a <= b);
    case MLB_OP_ADD_INT32: return u2f(f2u(a) + f2u(b));
    case MLB_OP_SUBTRACT_INT32: return u2f(f2u(a) - f2u(b));
  }
  return 0.0f;
}

static float op3(int op, float a, float b, float c)
{
  switch (op)
  {
    case MLB_OP_LERP: return a + (c * (b - a));           /* O:744 */
    case MLB_OP_INVERSE_LERP: return (c - a) / (b - a);   /* O:745 */
    case MLB_OP_CLAMP: return sse_min(sse_max(a, b), c);  /* M:93 */
    case MLB_OP_WITHIN: return u2f(f2u(mask_f(a >= b)) & f2u(mask_f(a < c))); /* M:94 */
    case MLB_OP_SELECT: return sel_bits(a, b, f2u(c));    /* O:886 */
  }
  return 0.0f;
}

/* ------------------------------------------------------------------ */
/* generators                                                           */

/* NoiseGen::operator(), G:115,132-142 */
static void gen_noise(uint32_t* seed, float* y)
{
  uint32_t s = *seed;
  for (int i = 0; i < NB; ++i)
  {
    s = s * 0x0019660Du + 0x3C6EF35Fu;
    uint32_t temp = ((s >> 9) & 0x007FFFFFu) | 0x3F800000u;
    y[i] = u2f(temp) * 2.f - 3.f;
  }
  *seed = s;
}

/* PhasorGen::operator(), G:187-203.  stepsPerCycle = 2^32, cyclesPerStep = 2^-32 */
static void gen_phasor(uint32_t* omega32, const float* freq, float* y)
{
  uint32_t om = *omega32;
  for (int n = 0; n < NB; ++n)
  {
    float steps = freq[n] * 4294967296.0f;
    int32_t isteps = cvt_round(steps);
    om += (uint32_t)isteps;
    y[n] = unsigned_to_float(om) * (1.0f / 4294967296.0f);
  }
  *omega32 = om;
}

/* phasorToSine, G:316-338.  The constants come from the reference's constexpr
 * Newton sqrt with tolerance 1e-3 (S:224,230-235): sqrt2 = 0x1.6a0a0ap+0, NOT
 * sqrt(2).  Bit patterns verified against the compiled reference. */
#define K_SQRT2 0x1.6a0a0ap+0f
#define K_DOMAIN 0x1.6a0a0ap+2f  /* sqrt2 * 4 */
#define K_FLIP 0x1.6a0a0ap+1f    /* sqrt2 * 2 */
#define K_INV_RANGE 0x1.0f876cp+0f /* 1 / (sqrt2 - sqrt2^3/6) */
#define K_ONE_SIXTH 0x1.555556p-3f
static inline float phasor_to_sine(float ph)
{
  float omega = ph * K_DOMAIN + (-K_SQRT2);
  float tri = (omega > K_SQRT2) ? (K_FLIP - omega) : omega;
  return (K_INV_RANGE * tri) * (1.0f - (tri * tri) * K_ONE_SIXTH);
}

/* polyBLEP, G:285-311 */
static inline float poly_blep(float t, float dt)
{
  float c = 0.f;
  if (t < dt)
  {
    t = t / dt;
    c = t + t - t * t - 1.0f;
  }
  else if (t > 1.0f - dt)
  {
    t = (t - 1.0f) / dt;
    c = t * t + t + t + 1.0f;
  }
  return c;
}

/* TickGen::operator(), G:29-46 */
static void gen_tick(uint32_t* st, const float* freq, float* y)
{
  float om = u2f(*st);
  for (int n = 0; n < NB; ++n)
  {
    y[n] = 0.f;
    om += freq[n];
    if (om > 1.0f)
    {
      om -= 1.0f;
      y[n] = 1.0f;
    }
  }
  *st = f2u(om);
}

/* ------------------------------------------------------------------ */
/* filters                                                              */

/* shared SVF core for Lopass/Hipass/Bandpass (F:121-131,183-194,227-237) */
#define SVF_G_CORE                         \
  float v0 = x[n];                         \
  float t0 = v0 - ic2;                     \
  float t1 = g0 * t0 + g1 * ic1;           \
  float t2 = g2 * t0 + g0 * ic1;

static void flt_lopass(uint32_t* st, const float* c, const float* x, float* y)
{
  float ic1 = u2f(st[0]), ic2 = u2f(st[1]);
  const float g0 = c[0], g1 = c[1], g2 = c[2];
  for (int n = 0; n < NB; ++n)
  {
    SVF_G_CORE
    float v2 = t2 + ic2;
    ic1 += 2.0f * t1;
    ic2 += 2.0f * t2;
    y[n] = v2;
  }
  st[0] = f2u(ic1), st[1] = f2u(ic2);
}
static void flt_hipass(uint32_t* st, const float* c, const float* x, float* y)
{
  float ic1 = u2f(st[0]), ic2 = u2f(st[1]);
  const float g0 = c[0], g1 = c[1], g2 = c[2], k = c[3];
  for (int n = 0; n < NB; ++n)
  {
    SVF_G_CORE
    float v1 = t1 + ic1;
    float v2 = t2 + ic2;
    ic1 += 2.0f * t1;
    ic2 += 2.0f * t2;
    y[n] = v0 - k * v1 - v2;
  }
  st[0] = f2u(ic1), st[1] = f2u(ic2);
}
static void flt_bandpass(uint32_t* st, const float* c, const float* x, float* y)
{
  float ic1 = u2f(st[0]), ic2 = u2f(st[1]);
  const float g0 = c[0], g1 = c[1], g2 = c[2];
  for (int n = 0; n < NB; ++n)
  {
    SVF_G_CORE
    float v1 = t1 + ic1;
    ic1 += 2.0f * t1;
    ic2 += 2.0f * t2;
    y[n] = v1;
  }
  st[0] = f2u(ic1), st[1] = f2u(ic2);
}

/* shared SVF core for the shelves and bell (F:293-298, 374-379, 432-437) */
#define SVF_A_CORE                          \
  float v0 = x[n];                          \
  float v3 = v0 - ic2;                      \
  float v1 = a1 * ic1 + a2 * v3;            \
  float v2 = ic2 + a2 * ic1 + a3 * v3;      \
  ic1 = 2 * v1 - ic1;                       \
  ic2 = 2 * v2 - ic2;

static void flt_loshelf(uint32_t* st, const float* c, const float* x, float* y)
{
  float ic1 = u2f(st[0]), ic2 = u2f(st[1]);
  const float a1 = c[0], a2 = c[1], a3 = c[2], m1 = c[3], m2 = c[4];
  for (int n = 0; n < NB; ++n)
  {
    SVF_A_CORE
    y[n] = v0 + m1 * v1 + m2 * v2;
  }
  st[0] = f2u(ic1), st[1] = f2u(ic2);
}
static void flt_hishelf(uint32_t* st, const float* c, const float* x, float* y)
{
  float ic1 = u2f(st[0]), ic2 = u2f(st[1]);
  const float a1 = c[0], a2 = c[1], a3 = c[2], m0 = c[3], m1 = c[4], m2 = c[5];
  for (int n = 0; n < NB; ++n)
  {
    SVF_A_CORE
    y[n] = m0 * v0 + m1 * v1 + m2 * v2;
  }
  st[0] = f2u(ic1), st[1] = f2u(ic2);
}
static void flt_bell(uint32_t* st, const float* c, const float* x, float* y)
{
  float ic1 = u2f(st[0]), ic2 = u2f(st[1]);
  const float a1 = c[0], a2 = c[1], a3 = c[2], m1 = c[3];
  for (int n = 0; n < NB; ++n)
  {
    SVF_A_CORE
    (void)v2;
    y[n] = v0 + m1 * v1;
  }
  st[0] = f2u(ic1), st[1] = f2u(ic2);
}

/* ---- coefficient-ROW forms: Lopass::operator()(vx, omega, k) after makeCoeffsVec (F:136-152),
 * LoShelf / HiShelf::operator()(vx, vc) (F:304-319, 385-400).  r[i] = coefficient row i. ---- */
static void flt_lopass_v(uint32_t* st, const float* const* r, const float* x, float* y)
{
  float ic1 = u2f(st[0]), ic2 = u2f(st[1]);
  for (int n = 0; n < NB; ++n)
  {
    const float g0 = r[0][n], g1 = r[1][n], g2 = r[2][n];
    SVF_G_CORE
    float v2 = t2 + ic2;
    ic1 += 2.0f * t1;
    ic2 += 2.0f * t2;
    y[n] = v2;
  }
  st[0] = f2u(ic1), st[1] = f2u(ic2);
}
static void flt_loshelf_v(uint32_t* st, const float* const* r, const float* x, float* y)
{
  float ic1 = u2f(st[0]), ic2 = u2f(st[1]);
  for (int n = 0; n < NB; ++n)
  {
    const float a1 = r[0][n], a2 = r[1][n], a3 = r[2][n], m1 = r[3][n], m2 = r[4][n];
    SVF_A_CORE
    y[n] = v0 + m1 * v1 + m2 * v2;
  }
  st[0] = f2u(ic1), st[1] = f2u(ic2);
}
static void flt_hishelf_v(uint32_t* st, const float* const* r, const float* x, float* y)
{
  float ic1 = u2f(st[0]), ic2 = u2f(st[1]);
  for (int n = 0; n < NB; ++n)
  {
    const float a1 = r[0][n], a2 = r[1][n], a3 = r[2][n], m0 = r[3][n], m1 = r[4][n], m2 = r[5][n];
    SVF_A_CORE
    y[n] = m0 * v0 + m1 * v1 + m2 * v2;
  }
  st[0] = f2u(ic1), st[1] = f2u(ic2);
}
static void svf_g(float omega, float k, float* g0, float* g1, float* g2);
/* Lopass::makeCoeffsVec, F:97-115 */
void mlport_coeffs_lopass_vec(const float* omega, const float* k, float* out)
{
  for (int n = 0; n < NB; ++n)
  {
    const float om = omega[n] < 0.5f ? omega[n] : 0.5f; /* _mm_min_ps(omega, 0.5) */
    const float kk = k[n] > 0.01f ? k[n] : 0.01f;       /* _mm_max_ps(k, 0.01) */
    svf_g(om, kk, &out[n], &out[NB + n], &out[2 * NB + n]);
  }
}
/* interpolateDSPVectorLinear, O:986-990 */
static void ramp_row(float start, float end, float* y)
{
  const float interval = (end - start) / (float)NB;
  const float base = start + interval;
  for (int n = 0; n < NB; ++n) y[n] = (float)n * interval + base;
}
void mlport_interpolate_coeffs_linear(const float* c0, const float* c1, int n_coeffs, float* out)
{
  for (int i = 0; i < n_coeffs; ++i) ramp_row(c0[i], c1[i], out + (size_t)i * NB);
}

/* OnePole::operator(), F:466-475 */
static void flt_onepole(uint32_t* st, const float* c, const float* x, float* y)
{
  float y1 = u2f(st[0]);
  const float a0 = c[0], b1 = c[1];
  for (int n = 0; n < NB; ++n)
  {
    y1 = a0 * x[n] + b1 * y1;
    y[n] = y1;
  }
  st[0] = f2u(y1);
}
/* DCBlocker::operator(), F:500-512 */
static void flt_dcblocker(uint32_t* st, const float* c, const float* x, float* y)
{
  float x1 = u2f(st[0]), y1 = u2f(st[1]);
  const float k = c[0];
  for (int n = 0; n < NB; ++n)
  {
    const float x0 = x[n];
    const float y0 = x0 - x1 + k * y1;
    y1 = y0;
    x1 = x0;
    y[n] = y0;
  }
  st[0] = f2u(x1), st[1] = f2u(y1);
}
/* Differentiator::operator(), F:522-534 */
static void flt_differentiator(uint32_t* st, const float* x, float* y)
{
  float x1 = u2f(st[0]);
  y[0] = x[0] - x1;
  for (int n = 1; n < NB; ++n) y[n] = x[n] - x[n - 1];
  st[0] = f2u(x[NB - 1]);
}
/* Integrator::operator(), F:547-557 */
static void flt_integrator(uint32_t* st, const float* c, const float* x, float* y)
{
  float y1 = u2f(st[0]);
  const float leak = c[0];
  for (int n = 0; n < NB; ++n)
  {
    y1 -= y1 * leak;
    y1 += x[n];
    y[n] = y1;
  }
  st[0] = f2u(y1);
}

/* ------------------------------------------------------------------ */
/* IntegerDelay + FDN<8>                                                */

/* bitsToContain, S:31-36 */
static int bits_to_contain(int x)
{
  int e;
  for (e = 0; (1 << e) < x; e++)
    ;
  return e;
}

typedef struct fdn8_mem
{
  float* ring[8];      /* IntegerDelay::mBuffer, F:803 */
  uint32_t mask[8];    /* mLengthMask */
  uint32_t widx[8];    /* mWriteIndex */
  int len[8];          /* mIntDelayInSamples */
  float carry[8][NB];  /* FDN::mDelayInputVectors, F:1167 */
} fdn8_mem;

static void fdn8_mem_init(fdn8_mem* m, const float* coef32)
{
  memset(m, 0, sizeof(*m));
  for (int n = 0; n < 8; ++n)
  {
    int len = (int)coef32[24 + n];
    /* setMaxDelayInSamples(len): newSize = 1 << bitsToContain(dMax + 64), F:822-830 */
    int size = 1 << bits_to_contain(len + NB);
    m->ring[n] = (float*)calloc((size_t)size, sizeof(float));
    m->mask[n] = (uint32_t)size - 1u;
    m->len[n] = len;
  }
}
static void fdn8_mem_free(fdn8_mem* m)
{
  for (int n = 0; n < 8; ++n) free(m->ring[n]);
}

/* IntegerDelay::operator()(vx), F:834-875: block write at w, block read at (w-d)&mask.
 * Sample-by-sample indexing with the mask reproduces both wrap splits. */
static void integer_delay(fdn8_mem* m, int n, const float* x, float* y)
{
  float* buf = m->ring[n];
  const uint32_t mask = m->mask[n];
  const uint32_t w = m->widx[n];
  for (int i = 0; i < NB; ++i) buf[(w + (uint32_t)i) & mask] = x[i];
  const uint32_t r = (w - (uint32_t)m->len[n]) & mask;
  for (int i = 0; i < NB; ++i) y[i] = buf[(r + (uint32_t)i) & mask];
  m->widx[n] = (w + NB) & mask;
}

/* FDN<8>::operator(), F:1195-1238.  st: 8 OnePole y1; coef32: a0[8] b1[8] gain[8] len[8] */
static void fdn8_process(fdn8_mem* m, uint32_t* st, const float* coef32, const float* x,
                         float* outL, float* outR)
{
  float d[8][NB];
  for (int n = 0; n < 8; ++n) integer_delay(m, n, m->carry[n], d[n]); /* F:1198-1201 */

  for (int i = 0; i < NB; ++i)
  {
    /* DSPVector sumR, sumL default-construct to zero (O:153), F:1204-1215 */
    float sumR = 0.f, sumL = 0.f;
    for (int n = 0; n < 8; ++n)
    {
      if (n & 1)
        sumL = sumL + d[n][i];
      else
        sumR = sumR + d[n][i];
    }
    outL[i] = sumL;
    outR[i] = sumR;
    float sum = 0.f; /* F:1223-1228 */
    for (int n = 0; n < 8; ++n) sum = sum + d[n][i];
    sum = sum * (2.0f / 8);
    for (int n = 0; n < 8; ++n) d[n][i] = d[n][i] - sum; /* F:1232 */
  }
  for (int n = 0; n < 8; ++n)
  {
    /* F:1233-1234: filters[n](v) * gain[n] + x */
    float y1 = u2f(st[n]);
    const float a0 = coef32[n], b1 = coef32[8 + n], g = coef32[16 + n];
    for (int i = 0; i < NB; ++i)
    {
      y1 = a0 * d[n][i] + b1 * y1;
      m->carry[n][i] = y1 * g + x[i];
    }
    st[n] = f2u(y1);
  }
}


/* ------------------------------------------------------------------ */
/* the rest of the L2 functor set                    */

/* OneShotGen::operator(), G:235-252 */
static void gen_oneshot(uint32_t* st, const float* freq, float* y)
{
  uint32_t om = st[0], gate = st[1], prev = st[2];
  for (int n = 0; n < NB; ++n)
  {
    int32_t isteps = cvt_round(freq[n] * 4294967296.0f);
    om += (uint32_t)isteps * gate;
    if (om < prev)
    {
      gate = 0;
      om = 0;
    }
    prev = om;
    y[n] = unsigned_to_float(om) * (1.0f / 4294967296.0f);
  }
  st[0] = om, st[1] = gate, st[2] = prev;
}

/* ImpulseGen, G:53-103.  Table (constructor, G:64-78): blackman window over 17 points
 * (makeWindow U:22-26 with projections::linear({0,16},{0,1}), dspwindows::blackman U:34-35), times
 * sinc(2 pi 0.25 (i - 8)), normalised by the row sum (normalize O:1040-1049, sum O:995-1005 with
 * vecSumH M:246-251: (v0 + v2) + (v1 + v3) per SIMD vector, vectors accumulated in order). */
void mlport_impulse_table(float* t17)
{
  float row[NB];
  memset(row, 0, sizeof(row));
  for (int i = 0; i < 17; ++i)
  {
    const float m = (1.f - 0.f) / (16.f - 0.f);
    const float x = m * ((float)i - 0.f) + 0.f;
    const float w = 0.42f - 0.5f * cosf(K_TWO_PI_F * x) + 0.08f * cosf(2.f * K_TWO_PI_F * x);
    const int idx = i - 8;
    const float pi_x = K_TWO_PI_F * 0.25f * idx;
    const float sinc = (idx == 0) ? 1.f : sinf(pi_x) / pi_x;
    row[i] = sinc * w;
  }
  float sum = 0;
  for (int n = 0; n < NB; n += 4) sum += (row[n] + row[n + 2]) + (row[n + 1] + row[n + 3]);
  for (int i = 0; i < 17; ++i) t17[i] = row[i] / sum;
}
/* ImpulseGen::operator(), G:82-102.  st: _omega, _outputCounter */
static void gen_impulse(uint32_t* st, const float* freq, float* y)
{
  static float table[17];
  static int have = 0;
  if (!have)
  {
    mlport_impulse_table(table);
    have = 1;
  }
  float om = u2f(st[0]);
  int32_t counter = (int32_t)st[1];
  for (int n = 0; n < NB; ++n)
  {
    y[n] = 0.f;
    om += freq[n];
    if (om > 1.0f)
    {
      om -= 1.0f;
      counter = 0;
    }
    if (counter < 17)
    {
      y[n] = table[counter];
      counter++;
    }
  }
  st[0] = f2u(om), st[1] = (uint32_t)counter;
}

/* tail shared by Peak and RMS, F:613,651: select(sqrtApprox(vy), 0, vy > 1e-20) */
static inline float follower_out(float v)
{
  return sel_bits(v * sse_rsqrt(v), 0.f, f2u(mask_f(v > (float)(1e-20))));
}
/* Peak::operator(), F:584-614.  st: y1, peakHoldCounter; c: a0, b1, peakHoldSamples */
static void flt_peak(uint32_t* st, const float* c, const float* x, float* y)
{
  float y1 = u2f(st[0]);
  int32_t counter = (int32_t)st[1];
  const float a0 = c[0], b1 = c[1];
  const int32_t hold = (int32_t)c[2];
  for (int n = 0; n < NB; ++n)
  {
    const float xs = x[n] * x[n];
    if (xs > y1)
    {
      y1 = xs;
      counter = hold;
    }
    else if (counter <= 0)
      y1 = a0 * xs + b1 * y1;
    y[n] = follower_out(y1);
  }
  if (counter > 0) counter -= NB;
  st[0] = f2u(y1), st[1] = (uint32_t)counter;
}
/* RMS::operator(), F:638-652 */
static void flt_rms(uint32_t* st, const float* c, const float* x, float* y)
{
  float y1 = u2f(st[0]);
  for (int n = 0; n < NB; ++n)
  {
    y1 = c[0] * (x[n] * x[n]) + c[1] * y1;
    y[n] = follower_out(y1);
  }
  st[0] = f2u(y1);
}

/* ADSR::processSample, F:692-784.  st: y y1 x1 threshold target k amp segment; c: ka kd s kr */
static void flt_adsr(uint32_t* st, const float* c, const float* x, float* out)
{
  float y = u2f(st[0]), y1 = u2f(st[1]), x1 = u2f(st[2]), threshold = u2f(st[3]);
  float target = u2f(st[4]), k = u2f(st[5]), amp = u2f(st[6]);
  int32_t segment = (int32_t)st[7];
  const float ka = c[0], kd = c[1], sus = c[2], kr = c[3];
  enum { A = 0, D = 1, S = 2, R = 3, off = 4 };
  for (int n = 0; n < NB; ++n)
  {
    const float xn = x[n];
    if (segment == off && xn == 0.f)
    {
      out[n] = 0.f;
      continue;
    }
    const int crossed = ((y1 > threshold) != (y > threshold));
    int recalc = 0;
    if (crossed && segment < off)
    {
      segment++;
      recalc = 1;
    }
    const int trigOn = (x1 == 0.f) && (xn > 0.f);
    const int trigOff = (x1 > 0.f) && (xn == 0.f);
    if (trigOn)
    {
      segment = A;
      amp = xn;
      recalc = 1;
    }
    else if (trigOff)
    {
      segment = R;
      recalc = 1;
    }
    if (recalc)
    {
      float startEnv = 0.f, endEnv = 0.f;
      switch (segment)
      {
        case A: startEnv = 0.f, endEnv = 1.f, k = ka; break;
        case D: startEnv = 1.f, endEnv = sus, k = kd; break;
        case S: startEnv = sus, endEnv = sus, k = 0.f, y1 = sus, y = sus; break;
        case R: startEnv = sus, endEnv = 0.f, k = kr; break;
        default: startEnv = 0.f, endEnv = 0.f, k = 0.f, y1 = 0.f, y = 0.f; break;
      }
      const float segmentBias = (endEnv - startEnv) * 0.1f;
      threshold = endEnv;
      target = endEnv + segmentBias;
    }
    x1 = xn;
    y1 = y;
    y = y + k * (target - y);
    out[n] = y * amp;
  }
  st[0] = f2u(y), st[1] = f2u(y1), st[2] = f2u(x1), st[3] = f2u(threshold);
  st[4] = f2u(target), st[5] = f2u(k), st[6] = f2u(amp), st[7] = (uint32_t)segment;
}

/* Allpass1::processSample, F:944-952 */
static inline float allpass1_tick(float* x1, float* y1, float a, float x)
{
  float y = *x1 + (x - *y1) * a;
  *x1 = x;
  *y1 = y;
  return y;
}
static void flt_allpass1(uint32_t* st, const float* c, const float* x, float* y)
{
  float x1 = u2f(st[0]), y1 = u2f(st[1]);
  for (int n = 0; n < NB; ++n) y[n] = allpass1_tick(&x1, &y1, c[0], x[n]);
  st[0] = f2u(x1), st[1] = f2u(y1);
}

/* LinearGlide::operator()(float), G:459-505.  curr = mCurrVec (delay memory);
 * st: step, target, vectorsRemaining; c: vectorsPerGlide, dyPerVector */
static void gen_glide(uint32_t* st, const float* c, float f, float* curr, float* y)
{
  float step = u2f(st[0]), target = u2f(st[1]);
  int32_t remaining = (int32_t)st[2];
  const int32_t per = (int32_t)c[0];
  const float dyPerVector = c[1];
  if (f != target)
  {
    target = f;
    remaining = per;
  }
  if (remaining < 0)
  {
  }
  else if (remaining == 0)
  {
    for (int n = 0; n < NB; ++n) curr[n] = target;
    step = 0.f;
    remaining--;
  }
  else if (remaining == per)
  {
    const float cv = curr[NB - 1];
    const float dydv = (target - cv) * dyPerVector;
    step = dydv;
    for (int n = 0; n < NB; ++n) curr[n] = cv + ((float)(n + 1) / (float)NB) * step; /* kUnityRampVec G:409-410 */
    remaining--;
  }
  else
  {
    for (int n = 0; n < NB; ++n) curr[n] = curr[n] + step;
    remaining--;
  }
  memcpy(y, curr, sizeof(float) * NB);
  st[0] = f2u(step), st[1] = f2u(target), st[2] = (uint32_t)remaining;
}
/* Interpolator1::operator()(float), G:416-422 */
static void gen_interp1(uint32_t* st, float f, float* y)
{
  const float cur = u2f(st[0]);
  const float dydt = f - cur;
  for (int n = 0; n < NB; ++n) y[n] = cur + ((float)(n + 1) / (float)NB) * dydt;
  st[0] = f2u(f);
}
/* SampleAccurateLinearGlide::nextSample, G:541-582 */
static void gen_sample_glide(uint32_t* st, const float* c, const float* x, float* y)
{
  float curr = u2f(st[0]), step = u2f(st[1]), target = u2f(st[2]);
  int32_t remaining = (int32_t)st[3];
  const int32_t per = (int32_t)c[0];
  for (int n = 0; n < NB; ++n)
  {
    if (x[n] != target)
    {
      target = x[n];
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
      step = (target - curr) * c[1];
      remaining--;
    }
    else
    {
      curr += step;
      remaining--;
    }
    y[n] = curr;
  }
  st[0] = f2u(curr), st[1] = f2u(step), st[2] = f2u(target), st[3] = (uint32_t)remaining;
}

/* ---- delay memory of one voice of one node: up to 2 IntegerDelay rings + one 64-float row ---- */
typedef struct delay_mem
{
  float* ring[2];   /* IntegerDelay::mBuffer, F:803 */
  uint32_t mask[2]; /* mLengthMask */
  uint32_t widx[2]; /* mWriteIndex */
  float row[NB];    /* Allpass::vy1 (F:1115), LinearGlide::mCurrVec (G:435), feedback DSPVector */
} delay_mem;

/* IntegerDelay::setMaxDelayInSamples(float d), F:822-830 */
static void ring_alloc(delay_mem* m, int r, float d)
{
  int dMax = (int)floorf(d);
  int size = 1 << bits_to_contain(dMax + NB);
  free(m->ring[r]);
  m->ring[r] = (float*)calloc((size_t)size, sizeof(float));
  m->mask[r] = (uint32_t)size - 1u;
  m->widx[r] = 0;
}
/* IntegerDelay::operator()(vx) with delay d, F:834-875 */
static void ring_block(delay_mem* m, int r, int d, const float* x, float* y)
{
  float* buf = m->ring[r];
  const uint32_t mask = m->mask[r], w = m->widx[r];
  for (int i = 0; i < NB; ++i) buf[(w + (uint32_t)i) & mask] = x[i];
  const uint32_t rd = (w - (uint32_t)d) & mask;
  for (int i = 0; i < NB; ++i) y[i] = buf[(rd + (uint32_t)i) & mask];
  m->widx[r] = (w + NB) & mask;
}
/* IntegerDelay::processSample, F:898-912 */
static inline float ring_tick(delay_mem* m, int r, int d, float x)
{
  float* buf = m->ring[r];
  const uint32_t mask = m->mask[r], w = m->widx[r];
  buf[w] = x;
  const float y = buf[(w - (uint32_t)d) & mask];
  m->widx[r] = (w + 1u) & mask;
  return y;
}
/* FractionalDelay::setDelayInSamples, F:993-1008 + Allpass1::makeCoeffs F:936-941 */
static void frac_split(float d, int32_t* delayInt, float* apCoeff)
{
  float fDelayInt = floorf(d);
  int32_t di = cvt_trunc(fDelayInt);
  float frac = d - fDelayInt;
  if ((frac < 0.618f) && (di > 0))
  {
    frac += 1.f;
    di -= 1;
  }
  *delayInt = di;
  float xm1 = (frac - 1.f);
  *apCoeff = -0.53f * xm1 + 0.24f * xm1 * xm1;
}
/* IntegerDelay::operator()(x, delay), F:877-896 */
static void delay_int_var(delay_mem* m, const float* x, const float* dl, float* y)
{
  for (int n = 0; n < NB; ++n) y[n] = ring_tick(m, 0, cvt_trunc(dl[n]), x[n]);
}
/* FractionalDelay::operator()(vx), F:1014: allpass(integerDelay(vx)) */
static void delay_frac(delay_mem* m, uint32_t* st, float d, const float* x, float* y)
{
  int32_t di;
  float a, t[NB];
  frac_split(d, &di, &a);
  ring_block(m, 0, di, x, t);
  float x1 = u2f(st[0]), y1 = u2f(st[1]);
  for (int n = 0; n < NB; ++n) y[n] = allpass1_tick(&x1, &y1, a, t[n]);
  st[0] = f2u(x1), st[1] = f2u(y1);
}
/* FractionalDelay::operator()(vx, vDelay), F:1033-1042 */
static void delay_frac_var(delay_mem* m, uint32_t* st, const float* x, const float* dl, float* y)
{
  float x1 = u2f(st[0]), y1 = u2f(st[1]);
  for (int n = 0; n < NB; ++n)
  {
    int32_t di;
    float a;
    frac_split(dl[n], &di, &a);
    y[n] = allpass1_tick(&x1, &y1, a, ring_tick(m, 0, di, x[n]));
  }
  st[0] = f2u(x1), st[1] = f2u(y1);
}
/* PitchbendableDelay::operator(), F:1097-1104 with the tick/fade tables of F:1053-1076.
 * st: 2 x {x1, y1, intDelay, apCoeff} */
static void delay_pitchbend(delay_mem* m, uint32_t* st, const float* x, const float* dl, float* y)
{
  float o[2][NB];
  for (int k = 0; k < 2; ++k)
  {
    float x1 = u2f(st[4 * k]), y1 = u2f(st[4 * k + 1]);
    int32_t di = (int32_t)st[4 * k + 2];
    float a = u2f(st[4 * k + 3]);
    for (int n = 0; n < NB; ++n)
    {
      /* delay 1 may change when n % 32 == 16, delay 2 when n % 32 == 0 */
      if ((n & 31) == (k == 0 ? 16 : 0)) frac_split(dl[n], &di, &a);
      o[k][n] = allpass1_tick(&x1, &y1, a, ring_tick(m, k, di, x[n]));
    }
    st[4 * k] = f2u(x1), st[4 * k + 1] = f2u(y1), st[4 * k + 2] = (uint32_t)di, st[4 * k + 3] = f2u(a);
  }
  for (int n = 0; n < NB; ++n)
  {
    const int r = n & 31;
    const float fade = 2.f * (r > 16 ? 1.0f - (float)r / 32.f : (float)r / 32.f);
    y[n] = o[0][n] + fade * (o[1][n] - o[0][n]); /* lerp, O:744 */
  }
}
/* Allpass<DELAY>::operator(), F:1135-1153: the delay call is made by the caller */
static void allpass_pre(const delay_mem* m, float gain, const float* x, float* delayInput, float* y)
{
  const float g = -gain;
  for (int n = 0; n < NB; ++n)
  {
    delayInput[n] = x[n] - m->row[n] * g;
    y[n] = delayInput[n] * g + m->row[n];
  }
}

/* ---- HalfBandFilter, F:1245-1310 ---- */
/* st: apa0{x1,y1} apa1{x1,y1} apb0{x1,y1} apb1{x1,y1} b1.  Coefficients F:1305-1306 (float-converted). */
static const float kHbA0 = 0.07986642623635751f, kHbA1 = 0.5453536510711322f;
static const float kHbB0 = 0.28382934487410993f, kHbB1 = 0.8344118914807379f;
typedef struct hb_state
{
  float s[9];
} hb_state;
static void hb_load(hb_state* h, const uint32_t* st)
{
  for (int i = 0; i < 9; ++i) h->s[i] = u2f(st[i]);
}
static void hb_store(const hb_state* h, uint32_t* st)
{
  for (int i = 0; i < 9; ++i) st[i] = f2u(h->s[i]);
}
static inline float hb_a(hb_state* h, float x) /* apa1(apa0(x)) */
{
  return allpass1_tick(&h->s[2], &h->s[3], kHbA1, allpass1_tick(&h->s[0], &h->s[1], kHbA0, x));
}
static inline float hb_b(hb_state* h, float x) /* apb1(apb0(x)) */
{
  return allpass1_tick(&h->s[6], &h->s[7], kHbB1, allpass1_tick(&h->s[4], &h->s[5], kHbB0, x));
}
/* upsampleFirstHalf then upsampleSecondHalf, F:1248-1270 */
static void hb_upsample(uint32_t* st, const float* x, float* y1, float* y2)
{
  hb_state h;
  hb_load(&h, st);
  for (int i = 0; i < NB / 2; ++i)
  {
    y1[2 * i] = hb_a(&h, x[i]);
    y1[2 * i + 1] = hb_b(&h, x[i]);
  }
  for (int i = NB / 2; i < NB; ++i)
  {
    y2[2 * (i - NB / 2)] = hb_a(&h, x[i]);
    y2[2 * (i - NB / 2) + 1] = hb_b(&h, x[i]);
  }
  hb_store(&h, st);
}
/* downsample(vx1, vx2), F:1272-1294 */
static void hb_downsample(uint32_t* st, const float* x1, const float* x2, float* y)
{
  hb_state h;
  hb_load(&h, st);
  for (int half = 0; half < 2; ++half)
  {
    const float* x = half ? x2 : x1;
    for (int i = 0; i < NB / 2; ++i)
    {
      float a0 = hb_a(&h, x[2 * i]);
      float b0 = hb_b(&h, x[2 * i + 1]);
      y[half * (NB / 2) + i] = (a0 + h.s[8]) * 0.5f;
      h.s[8] = b0;
    }
  }
  hb_store(&h, st);
}

/* TempoLock::operator(), F:1494-1578.  st: _omega, _x1v; dydx = ratio; isr = coef */
static void gen_tempo_lock(uint32_t* st, const float* x, float dydx, float isr, float* y)
{
  float omega = u2f(st[0]), x1v = u2f(st[1]);
  const float x0 = x[0];
  float dxdt = 0.f, dydt = 0.f;
  if (x0 == -1.0f)
  {
    omega = -1.0f;
    for (int i = 0; i < NB; ++i) y[i] = 0.f;
  }
  else
  {
    if (omega > -1.f)
    {
      float dx = x0 - x1v;
      if (dx < 0.f) dx += 1.f;
      dxdt = dx / (float)NB;
      dydt = dxdt * dydx;
      x1v = x0;
    }
    else
    {
      dxdt = x[1] - x0;
      dydt = dxdt * dydx;
      x1v = x0 - dxdt * (float)NB;
      omega = fmodf(x0 * dydx, 1.0f);
    }
    int lock = 0;
    const float lockDist = 0.001f;
    if (fabsf(dydx - roundf(dydx)) < lockDist) lock = 1;
    float rdydx = 1.0f / dydx;
    if (fabsf(rdydx - roundf(rdydx)) < lockDist) lock = 1;
    if (lock)
    {
      float ref, refWrap, error;
      if (dydx >= 1.f)
      {
        ref = x0 * dydx;
        refWrap = ref - floorf(ref);
        error = omega - refWrap;
      }
      else
      {
        ref = omega / dydx;
        refWrap = ref - floorf(ref);
        error = refWrap - x0;
      }
      float errorDiff = roundf(error) - error;
      float correction = errorDiff * isr * 4.0f;
      const float lo = -dydt * 0.5f, hi = dydt * 1.0f;
      correction = (correction < lo) ? lo : (correction > hi ? hi : correction); /* ml::clamp, S:68-72 */
      dydt += correction;
    }
    for (int i = 0; i < NB; ++i)
    {
      y[i] = omega;
      omega += dydt;
      if (omega > 1.0f) omega -= 1.0f;
    }
  }
  st[0] = f2u(omega), st[1] = f2u(x1v);
}

/* ------------------------------------------------------------------ */
/* graph runner                                                         */

static int op_info(int op, int* nin, int* nst, int* nco)
{
  switch (op)
  {
#define MLB_X_CASE(NAME, id, a, b, c) \
  case id:                            \
    *nin = a, *nst = b, *nco = c;     \
    return 1;
    MLB_OP_TABLE(MLB_X_CASE)
#undef MLB_X_CASE
  }
  return 0;
}

typedef struct mlport_graph
{
  mlb_node* nodes;
  int n_nodes;
  int* outs;
  int n_out;
  int* st_off;
  int* co_off;
  int n_state, n_coef, n_in, V;
  uint32_t* state; /* [n_state][V] */
  float* coef;     /* [n_coef][V] */
  fdn8_mem** fdn;  /* [n_nodes] -> array of V, or NULL */
  delay_mem** dmem; /* [n_nodes] -> array of V, or NULL */
  int* again;      /* [n_nodes]: MLB_AGAIN target (the node whose words and member rows this node uses), or -1 */
} mlport_graph;

int mlport_abi_version(void) { return MLB_ABI_VERSION; }

void mlport_graph_destroy(mlport_graph* g)
{
  if (!g) return;
  if (g->fdn)
  {
    for (int i = 0; i < g->n_nodes; ++i)
      if (g->fdn[i])
      {
        for (int v = 0; v < g->V; ++v) fdn8_mem_free(&g->fdn[i][v]);
        free(g->fdn[i]);
      }
    free(g->fdn);
  }
  if (g->dmem)
  {
    for (int i = 0; i < g->n_nodes; ++i)
      if (g->dmem[i] && !(g->again && g->again[i] >= 0))
      {
        for (int v = 0; v < g->V; ++v) free(g->dmem[i][v].ring[0]), free(g->dmem[i][v].ring[1]);
        free(g->dmem[i]);
      }
    free(g->dmem);
  }
  free(g->nodes);
  free(g->outs);
  free(g->st_off);
  free(g->co_off);
  free(g->again);
  free(g->state);
  free(g->coef);
  free(g);
}

mlport_graph* mlport_graph_create(const mlb_node* nodes, int n_nodes, const int32_t* outs,
                                  int n_out, int V, const float* coef)
{
  mlport_graph* g = (mlport_graph*)calloc(1, sizeof(*g));
  g->nodes = (mlb_node*)malloc(sizeof(mlb_node) * (size_t)n_nodes);
  memcpy(g->nodes, nodes, sizeof(mlb_node) * (size_t)n_nodes);
  g->n_nodes = n_nodes;
  g->outs = (int*)malloc(sizeof(int) * (size_t)(n_out > 0 ? n_out : 1));
  for (int i = 0; i < n_out; ++i) g->outs[i] = outs[i];
  g->n_out = n_out;
  g->V = V;
  g->st_off = (int*)calloc((size_t)n_nodes, sizeof(int));
  g->co_off = (int*)calloc((size_t)n_nodes, sizeof(int));
  g->fdn = (fdn8_mem**)calloc((size_t)n_nodes, sizeof(fdn8_mem*));
  g->dmem = (delay_mem**)calloc((size_t)n_nodes, sizeof(delay_mem*));
  g->again = (int*)calloc((size_t)n_nodes, sizeof(int));
  for (int i = 0; i < n_nodes; ++i)
  {
    int a, b, c;
    if (!op_info(nodes[i].op, &a, &b, &c))
    {
      mlport_graph_destroy(g);
      return NULL;
    }
    /* MLB_AGAIN (mlb200.h): a further call of an earlier functor in the same vector uses that node's words */
    g->again[i] = -1;
    if (nodes[i].iarg < 0 && nodes[i].op != MLB_OP_INPUT && nodes[i].op != MLB_OP_PARAM &&
        nodes[i].op != MLB_OP_FEEDBACK_WRITE)
    {
      const int t = MLB_AGAIN_TARGET(nodes[i].iarg);
      const int op = nodes[i].op;
      if (t >= i || nodes[t].op != op || g->again[t] >= 0 || (b == 0 && c == 0) || op == MLB_OP_FEEDBACK_READ ||
          op == MLB_OP_FDN8 || op == MLB_OP_FDN8_R || op == MLB_OP_HALFBAND_UP_2 || op == MLB_OP_DOWN2X_IN ||
          op == MLB_OP_DOWN2X_OUT)
      {
        mlport_graph_destroy(g);
        return NULL;
      }
      g->again[i] = t;
      g->st_off[i] = g->st_off[t];
      g->co_off[i] = g->co_off[t];
      continue;
    }
    g->st_off[i] = g->n_state;
    g->co_off[i] = g->n_coef;
    g->n_state += b;
    g->n_coef += c;
    if (nodes[i].op == MLB_OP_INPUT && nodes[i].iarg + 1 > g->n_in) g->n_in = nodes[i].iarg + 1;
    if ((nodes[i].op == MLB_OP_HALFBAND_UP_2 && nodes[nodes[i].in[0]].op != MLB_OP_HALFBAND_UP) ||
        (nodes[i].op == MLB_OP_FDN8_R && nodes[nodes[i].in[0]].op != MLB_OP_FDN8))
    {
      mlport_graph_destroy(g);
      return NULL;
    }
    if (nodes[i].op == MLB_OP_FEEDBACK_WRITE &&
        (nodes[i].iarg < 0 || nodes[i].iarg >= i || nodes[nodes[i].iarg].op != MLB_OP_FEEDBACK_READ))
    {
      mlport_graph_destroy(g);
      return NULL;
    }
  }
  g->state = (uint32_t*)calloc((size_t)(g->n_state > 0 ? g->n_state : 1) * (size_t)V, 4);
  g->coef = (float*)calloc((size_t)(g->n_coef > 0 ? g->n_coef : 1) * (size_t)V, 4);
  if (coef && g->n_coef) memcpy(g->coef, coef, (size_t)g->n_coef * (size_t)V * 4);
  for (int i = 0; i < n_nodes; ++i)
    if (nodes[i].op == MLB_OP_FDN8)
    {
      g->fdn[i] = (fdn8_mem*)calloc((size_t)V, sizeof(fdn8_mem));
      float c32[32];
      for (int v = 0; v < V; ++v)
      {
        for (int k = 0; k < 32; ++k) c32[k] = g->coef[(size_t)(g->co_off[i] + k) * V + v];
        fdn8_mem_init(&g->fdn[i][v], c32);
      }
    }
  for (int i = 0; i < n_nodes; ++i)
  {
    int n_rows = 0, n_rings = 0;
    switch (nodes[i].op)
    {
#define MLB_X_MEM(NAME, rows, rings) \
  case MLB_OP_##NAME: n_rows = rows, n_rings = rings; break;
      MLB_OP_MEM_TABLE(MLB_X_MEM)
#undef MLB_X_MEM
      default: break;
    }
    if (n_rows == 0 && n_rings == 0) continue;
    if (g->again[i] >= 0)
    {
      if (n_rings)  /* a ring's write index is the vector count: such a functor cannot tick twice per vector */
      {
        mlport_graph_destroy(g);
        return NULL;
      }
      g->dmem[i] = g->dmem[g->again[i]];  /* the functor's member row */
      continue;
    }
    g->dmem[i] = (delay_mem*)calloc((size_t)V, sizeof(delay_mem));
    int a, b, nco;
    op_info(nodes[i].op, &a, &b, &nco);
    for (int v = 0; v < V; ++v)
    {
      if (n_rings == 0) continue;
      /* the last coefficient word is the functor's setMaxDelayInSamples argument */
      float md = g->coef[(size_t)(g->co_off[i] + nco - 1) * V + v];
      switch (nodes[i].op)
      {
        case MLB_OP_ALLPASS_INT:  /* Allpass::setMaxDelayInSamples: d - 64, F:1125-1128 */
          ring_alloc(&g->dmem[i][v], 0, md - (float)NB);
          break;
        case MLB_OP_ALLPASS_FRAC: /* ... then FractionalDelay: floorf(d), F:1010 */
          ring_alloc(&g->dmem[i][v], 0, floorf(md - (float)NB));
          break;
        case MLB_OP_ALLPASS_PB:
          ring_alloc(&g->dmem[i][v], 0, floorf(md - (float)NB));
          ring_alloc(&g->dmem[i][v], 1, floorf(md - (float)NB));
          break;
        case MLB_OP_PITCHBEND_DELAY:
          ring_alloc(&g->dmem[i][v], 0, floorf(md));
          ring_alloc(&g->dmem[i][v], 1, floorf(md));
          break;
        case MLB_OP_FRACTIONAL_DELAY:
        case MLB_OP_FRACTIONAL_DELAY_VAR: ring_alloc(&g->dmem[i][v], 0, floorf(md)); break;
        default: ring_alloc(&g->dmem[i][v], 0, md); break;
      }
    }
  }
  return g;
}

void mlport_graph_set_state(mlport_graph* g, const uint32_t* state)
{
  if (g->n_state) memcpy(g->state, state, (size_t)g->n_state * (size_t)g->V * 4);
}
void mlport_graph_get_state(mlport_graph* g, uint32_t* state)
{
  if (g->n_state) memcpy(state, g->state, (size_t)g->n_state * (size_t)g->V * 4);
}

typedef struct job
{
  mlport_graph* g;
  const float* in;
  float* out;
  int T, v0, v1;
} job;

static void run_voices(mlport_graph* g, const float* in, float* out, int T, int v0, int v1)
{
  const int N = g->n_nodes, V = g->V;
  float(*rows)[NB] = (float(*)[NB])malloc(sizeof(float) * NB * (size_t)N);
  float(*rows2)[NB] = (float(*)[NB])malloc(sizeof(float) * NB * (size_t)N);
  uint32_t st[16];
  float co[64];
  for (int v = v0; v < v1; ++v)
    for (int t = 0; t < T; ++t)
    {
      for (int i = 0; i < N; ++i)
      {
        const mlb_node* nd = &g->nodes[i];
        int nin, nst, nco;
        op_info(nd->op, &nin, &nst, &nco);
        for (int k = 0; k < nst; ++k) st[k] = g->state[(size_t)(g->st_off[i] + k) * V + v];
        for (int k = 0; k < nco; ++k) co[k] = g->coef[(size_t)(g->co_off[i] + k) * V + v];
        const float* a = nd->in[0] >= 0 ? rows[nd->in[0]] : NULL;
        const float* b = nd->in[1] >= 0 ? rows[nd->in[1]] : NULL;
        const float* c = nd->in[2] >= 0 ? rows[nd->in[2]] : NULL;
        float* y = rows[i];
        switch (nd->op)
        {
          case MLB_OP_INPUT:
            memcpy(y, in + (((size_t)t * g->n_in + nd->iarg) * V + v) * NB, sizeof(float) * NB);
            break;
          case MLB_OP_PARAM: /* DSPVector(float k): broadcast, O:157,171-182 */
            for (int n = 0; n < NB; ++n) y[n] = co[0];
            break;
          case MLB_OP_NOISE: gen_noise(&st[0], y); break;
          case MLB_OP_PHASOR: gen_phasor(&st[0], a, y); break;
          case MLB_OP_SINE: /* SineGen, G:380 */
            gen_phasor(&st[0], a, y);
            for (int n = 0; n < NB; ++n) y[n] = phasor_to_sine(y[n]);
            break;
          case MLB_OP_SAW: /* phasorToSaw, G:362-369 */
            gen_phasor(&st[0], a, y);
            for (int n = 0; n < NB; ++n)
            {
              float saw = y[n] * 2.f - 1.f;
              y[n] = saw - poly_blep(y[n], a[n]);
            }
            break;
          case MLB_OP_PULSE: /* phasorToPulse, G:342-358 */
            gen_phasor(&st[0], a, y);
            for (int n = 0; n < NB; ++n)
            {
              float om = y[n], w = b[n];
              float p = (om >= w) ? -1.f : 1.f;
              p = p + poly_blep(om, a[n]);
              float t = om - w + 1.0f;
              float down = t - (float)cvt_trunc(t); /* fractionalPart, O:825 */
              p = p - poly_blep(down, a[n]);
              y[n] = p;
            }
            break;
          case MLB_OP_TICK: gen_tick(&st[0], a, y); break;
          case MLB_OP_LOPASS: flt_lopass(st, co, a, y); break;
          case MLB_OP_HIPASS: flt_hipass(st, co, a, y); break;
          case MLB_OP_BANDPASS: flt_bandpass(st, co, a, y); break;
          case MLB_OP_LOSHELF: flt_loshelf(st, co, a, y); break;
          case MLB_OP_HISHELF: flt_hishelf(st, co, a, y); break;
          case MLB_OP_BELL: flt_bell(st, co, a, y); break;
          case MLB_OP_LOPASS_V:
          case MLB_OP_LOSHELF_V:
          case MLB_OP_HISHELF_V:
          {
            const float* r[MLB_MAX_INS - 1];
            for (int k = 1; k < nin; ++k) r[k - 1] = rows[nd->in[k]];
            if (nd->op == MLB_OP_LOPASS_V) flt_lopass_v(st, r, a, y);
            if (nd->op == MLB_OP_LOSHELF_V) flt_loshelf_v(st, r, a, y);
            if (nd->op == MLB_OP_HISHELF_V) flt_hishelf_v(st, r, a, y);
            break;
          }
          case MLB_OP_LOPASS_MOD: /* makeCoeffsVec inside the operator, F:139 */
          {
            float vc[3 * NB];
            const float* r[3] = {vc, vc + NB, vc + 2 * NB};
            mlport_coeffs_lopass_vec(b, c, vc);
            flt_lopass_v(st, r, a, y);
            break;
          }
          case MLB_OP_RAMP: ramp_row(a[0], b[0], y); break;
          case MLB_OP_ONEPOLE: flt_onepole(st, co, a, y); break;
          case MLB_OP_DCBLOCKER: flt_dcblocker(st, co, a, y); break;
          case MLB_OP_DIFFERENTIATOR: flt_differentiator(st, a, y); break;
          case MLB_OP_INTEGRATOR: flt_integrator(st, co, a, y); break;
          case MLB_OP_FDN8: fdn8_process(&g->fdn[i][v], st, co, a, y, rows2[i]); break;
          case MLB_OP_FDN8_R: memcpy(y, rows2[nd->in[0]], sizeof(float) * NB); break;
          case MLB_OP_ONESHOT: gen_oneshot(st, a, y); break;
          case MLB_OP_IMPULSE: gen_impulse(st, a, y); break;
          case MLB_OP_PEAK: flt_peak(st, co, a, y); break;
          case MLB_OP_RMS: flt_rms(st, co, a, y); break;
          case MLB_OP_ADSR: flt_adsr(st, co, a, y); break;
          case MLB_OP_ALLPASS1: flt_allpass1(st, co, a, y); break;
          case MLB_OP_GLIDE: gen_glide(st, co, a[0], g->dmem[i][v].row, y); break;
          case MLB_OP_INTERPOLATOR1: gen_interp1(st, a[0], y); break;
          case MLB_OP_SAMPLE_GLIDE: gen_sample_glide(st, co, a, y); break;
          case MLB_OP_INTEGER_DELAY: ring_block(&g->dmem[i][v], 0, (int)co[0], a, y); break;
          case MLB_OP_INTEGER_DELAY_VAR: delay_int_var(&g->dmem[i][v], a, b, y); break;
          case MLB_OP_FRACTIONAL_DELAY: delay_frac(&g->dmem[i][v], st, co[0], a, y); break;
          case MLB_OP_FRACTIONAL_DELAY_VAR: delay_frac_var(&g->dmem[i][v], st, a, b, y); break;
          case MLB_OP_PITCHBEND_DELAY: delay_pitchbend(&g->dmem[i][v], st, a, b, y); break;
          case MLB_OP_ALLPASS_INT:
          {
            /* Allpass::setDelayInSamples(d): IntegerDelay::setDelayInSamples(int(d - 64)), F:1123 */
            delay_mem* m = &g->dmem[i][v];
            float din[NB];
            allpass_pre(m, co[0], a, din, y);
            ring_block(m, 0, (int)(co[1] - (float)NB), din, m->row);
            break;
          }
          case MLB_OP_ALLPASS_FRAC:
          {
            delay_mem* m = &g->dmem[i][v];
            float din[NB];
            allpass_pre(m, co[0], a, din, y);
            delay_frac(m, st, co[1] - (float)NB, din, m->row);
            break;
          }
          case MLB_OP_ALLPASS_PB:
          {
            delay_mem* m = &g->dmem[i][v];
            float din[NB], dl[NB];
            allpass_pre(m, co[0], a, din, y);
            for (int n = 0; n < NB; ++n) dl[n] = b[n] - (float)NB; /* F:1151 */
            delay_pitchbend(m, st, din, dl, m->row);
            break;
          }
          case MLB_OP_HALFBAND_UP: hb_upsample(st, a, y, rows2[i]); break;
          case MLB_OP_HALFBAND_UP_2: memcpy(y, rows2[nd->in[0]], sizeof(float) * NB); break;
          case MLB_OP_HALFBAND_DOWN: hb_downsample(st, a, b, y); break;
          case MLB_OP_TEMPO_LOCK: gen_tempo_lock(st, a, b[0], co[0], y); break;
          case MLB_OP_DOWN2X_IN: /* Downsample2xFunction, statements before fn, MLDSPFunctional.h:184-193,213-214 */
          {
            float* buffer = g->dmem[i][v].row;
            memset(y, 0, sizeof(float) * NB);
            if (st[9])
              hb_downsample(st, buffer, a, y);
            else
              memcpy(buffer, a, sizeof(float) * NB);
            st[9] = !st[9];
            break;
          }
          case MLB_OP_DOWN2X_OUT: /* ... statements after fn, :196-205,215-218 */
          {
            float* buffer = g->dmem[i][v].row;
            if (st[9])
            {
              float second[NB];
              hb_upsample(st, a, y, second);
              memcpy(buffer, second, sizeof(second));
            }
            else
              memcpy(y, buffer, sizeof(float) * NB);
            st[9] = !st[9];
            break;
          }
          case MLB_OP_FEEDBACK_READ: memcpy(y, g->dmem[i][v].row, sizeof(float) * NB); break;
          case MLB_OP_FEEDBACK_WRITE:
            memcpy(g->dmem[nd->iarg][v].row, a, sizeof(float) * NB);
            memcpy(y, a, sizeof(float) * NB);
            break;
          default:
            if (nin == 1)
              for (int n = 0; n < NB; ++n) y[n] = op1(nd->op, a[n]);
            else if (nin == 2)
              for (int n = 0; n < NB; ++n) y[n] = op2(nd->op, a[n], b[n]);
            else
              for (int n = 0; n < NB; ++n) y[n] = op3(nd->op, a[n], b[n], c[n]);
        }
        for (int k = 0; k < nst; ++k) g->state[(size_t)(g->st_off[i] + k) * V + v] = st[k];
      }
      if (out)
        for (int c = 0; c < g->n_out; ++c)
          memcpy(out + (((size_t)t * g->n_out + c) * V + v) * NB, rows[g->outs[c]],
                 sizeof(float) * NB);
    }
  free(rows);
  free(rows2);
}

static void* job_main(void* p)
{
  job* j = (job*)p;
  run_voices(j->g, j->in, j->out, j->T, j->v0, j->v1);
  return NULL;
}

/*
 * mix_mode 0: reference order -- voices summed left to right v = 0..V-1 starting
 *             from +0 (addRows, O:1349-1359).
 * mix_mode 1: the device order (DESIGN.md "mix bus") -- every level is a left-to-right
 *             sum starting from +0: 32 consecutive voices -> group; 64 consecutive
 *             groups -> chunk; chunks -> shard (one contiguous shard per GPU);
 *             shards -> total.
 */
void mlport_graph_process(mlport_graph* g, const float* in, float* out, float* mix, int T,
                          int nthreads, int mix_mode, int n_shards)
{
  const int V = g->V;
  float* o = out;
  float* scratch = NULL;
  if (!o && mix)
  {
    scratch = (float*)malloc(sizeof(float) * (size_t)T * (size_t)g->n_out * (size_t)V * NB);
    o = scratch;
  }
  if (nthreads < 1) nthreads = 1;
  if (nthreads > V) nthreads = V > 0 ? V : 1;
  if (nthreads == 1)
    run_voices(g, in, o, T, 0, V);
  else
  {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    job* jobs = (job*)malloc(sizeof(job) * (size_t)nthreads);
    for (int i = 0; i < nthreads; ++i)
    {
      jobs[i].g = g, jobs[i].in = in, jobs[i].out = o, jobs[i].T = T;
      jobs[i].v0 = (int)((long long)V * i / nthreads);
      jobs[i].v1 = (int)((long long)V * (i + 1) / nthreads);
      pthread_create(&th[i], NULL, job_main, &jobs[i]);
    }
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th);
    free(jobs);
  }
  if (mix)
  {
    if (n_shards < 1) n_shards = 1;
    for (int t = 0; t < T; ++t)
      for (int c = 0; c < g->n_out; ++c)
      {
        const float* plane = o + ((size_t)t * g->n_out + c) * (size_t)V * NB;
        float* m = mix + ((size_t)t * g->n_out + c) * NB;
        for (int n = 0; n < NB; ++n)
        {
          if (mix_mode == 0)
          {
            float acc = 0.f;
            for (int v = 0; v < V; ++v) acc = acc + plane[(size_t)v * NB + n];
            m[n] = acc;
          }
          else
          {
            float total = 0.f;
            for (int s = 0; s < n_shards; ++s)
            {
              /* shard s owns voices [s*V/S, (s+1)*V/S) exactly like bench.py's sharding */
              int s0 = (int)((long long)V * s / n_shards), s1 = (int)((long long)V * (s + 1) / n_shards);
              float shard = 0.f;
              for (int c0 = s0; c0 < s1; c0 += 64 * 32) /* chunks of 64 groups */
              {
                int c1 = c0 + 64 * 32 < s1 ? c0 + 64 * 32 : s1;
                float chunk = 0.f;
                for (int g0 = c0; g0 < c1; g0 += 32)
                {
                  int g1 = g0 + 32 < c1 ? g0 + 32 : c1;
                  float grp = 0.f;
                  for (int v = g0; v < g1; ++v) grp = grp + plane[(size_t)v * NB + n];
                  chunk = chunk + grp;
                }
                shard = shard + chunk;
              }
              total = total + shard;
            }
            m[n] = total;
          }
        }
      }
  }
  free(scratch);
}

/* ---- coefficient design (host libm, same calls as the reference) ---- */
#define K_PI 3.1415926535897932384626433f    /* S:24 */
#define K_TWO_PI 6.2831853071795864769252867f /* S:23 */

static void svf_g(float omega, float k, float* g0, float* g1, float* g2)
{
  /* F:85-95 (identical bodies at F:168-178, 212-222) */
  float piOmega = K_PI * omega;
  float s1 = sinf(piOmega);
  float s2 = sinf(2.0f * piOmega);
  float nrm = 1.0f / (2.f + k * s2);
  *g0 = s2 * nrm;
  *g1 = (-2.f * s1 * s1 - k * s2) * nrm;
  *g2 = (2.0f * s1 * s1) * nrm;
}
void mlport_coeffs_lopass(float omega, float k, float* o) { svf_g(omega, k, &o[0], &o[1], &o[2]); }
void mlport_coeffs_hipass(float omega, float k, float* o)
{
  svf_g(omega, k, &o[0], &o[1], &o[2]);
  o[3] = k;
}
void mlport_coeffs_bandpass(float omega, float k, float* o) { svf_g(omega, k, &o[0], &o[1], &o[2]); }
void mlport_coeffs_loshelf(float omega, float k, float A, float* r) /* F:270-281 */
{
  float piOmega = K_PI * omega;
  float g = tanf(piOmega) / sqrtf(A);
  r[0] = 1.f / (1.f + g * (g + k));
  r[1] = g * r[0];
  r[2] = g * r[1];
  r[3] = k * (A - 1.f);
  r[4] = (A * A - 1.f);
}
void mlport_coeffs_hishelf(float omega, float k, float A, float* r) /* F:350-362 */
{
  float piOmega = K_PI * omega;
  float g = tanf(piOmega) * sqrtf(A);
  r[0] = 1.f / (1.f + g * (g + k));
  r[1] = g * r[0];
  r[2] = g * r[1];
  r[3] = A * A;
  r[4] = k * (1.f - A) * A;
  r[5] = (1.f - A * A);
}
void mlport_coeffs_bell(float omega, float k, float A, float* r) /* F:415-425 */
{
  float kc = k / A;
  float piOmega = K_PI * omega;
  float g = tanf(piOmega);
  float a1 = 1.f / (1.f + g * (g + kc));
  float a2 = g * a1;
  float a3 = g * a2;
  float m1 = kc * (A * A - 1.f);
  r[0] = a1, r[1] = a2, r[2] = a3, r[3] = m1;
}
void mlport_coeffs_onepole(float omega, float* r) /* F:458-462 */
{
  float x = expf(-omega * K_TWO_PI);
  r[0] = 1.f - x;
  r[1] = x;
}
float mlport_coeffs_dcblocker(float omega) { return cosf(omega); } /* F:498 */
float mlport_db_to_gain(float dB) { return powf(10.f, dB / 40.f); } /* F:30 */

/* ---- SURVEY 8(f) row 2 coefficient design ---- */
void mlport_coeffs_peak(float omega, float* r) { mlport_coeffs_onepole(omega, r); } /* F:578-582 */
void mlport_coeffs_rms(float omega, float* r) { mlport_coeffs_onepole(omega, r); }  /* F:632-636 */
void mlport_coeffs_adsr(float a, float d, float s, float rel, float sr, float* o) /* F:676-683 */
{
  const float minSegmentTime = 0.0002f;
  const float invSr = 1.0f / sr;
  o[0] = K_TWO_PI * invSr / sse_max(a, minSegmentTime);
  o[1] = K_TWO_PI * invSr / sse_max(d, minSegmentTime);
  o[2] = s;
  o[3] = K_TWO_PI * invSr / sse_max(rel, minSegmentTime);
}
float mlport_coeffs_allpass1(float d) /* F:936-941 */
{
  float xm1 = (d - 1.f);
  return -0.53f * xm1 + 0.24f * xm1 * xm1;
}
void mlport_coeffs_glide(float t, float* o) /* G:443-448 */
{
  int n = (int)(t / NB);
  if (n < 1) n = 1;
  o[0] = (float)n;
  o[1] = 1.0f / ((float)n + 0.f);
}
void mlport_coeffs_sample_glide(float t, float* o) /* G:527-532 */
{
  int n = (int)t;
  if (n < 1) n = 1;
  o[0] = (float)n;
  o[1] = 1.0f / (float)n;
}

/* ------------------------------------------------------------------ */
/* EventsToSignals::Voice (source/app/MLEventsToSignals.{h,cpp}; E = the .cpp) */

typedef struct lin_glide /* LinearGlide, G:433-515 */
{
  float curr[NB];
  uint32_t st[3]; /* step, target, vectorsRemaining */
  float c[2];     /* vectorsPerGlide, dyPerVector */
} lin_glide;

static void lin_glide_init(lin_glide* g) /* member defaults G:435-440 */
{
  memset(g, 0, sizeof(*g));
  g->st[2] = (uint32_t)-1;
  g->c[0] = 32.f;
  g->c[1] = 1.f / 32;
}
static void lin_glide_set_value(lin_glide* g, float f) /* G:451-455 */
{
  g->st[1] = f2u(f);
  g->st[2] = 0;
}

enum { GL_BEND = 0, GL_MOD, GL_X, GL_Y, GL_Z, GL_DRIFT, GL_PRESSURE, GL_COUNT };
enum { ROW_PITCH = 0, ROW_GATE, ROW_VOICE, ROW_Z, ROW_X, ROW_Y, ROW_MOD, ROW_TIME }; /* .h:16-27 */

typedef struct port_voice
{
  /* pitchGlide: SampleAccurateLinearGlide, G:517-590 */
  uint32_t pg[4]; /* curr, step, target, samplesRemaining */
  float pgc[2];   /* samplesPerGlide, dyPerSample */
  lin_glide g[GL_COUNT];
  float vel, pitch, bend, mod, x, y, z;
  uint32_t age, ageStep;
  int nextFrame;
  uint32_t seed;
  int driftCounter, nextDriftTime;
  float curDrift, driftAmount;
  double sr;
  float glideSeconds;
  int pitchGlideTimeInSamples;
  int inhibit, recalc, voiceIndex;
  float pitchBendRange;
  float pressure; /* SmoothedController::inputValue of controllers[128] */
  int midi;
  float out[8][NB];
} port_voice;

static void sa_glide_set_time(port_voice* v, float t) /* G:527-532 */
{
  int n = (int)t;
  if (n < 1) n = 1;
  v->pgc[0] = (float)n;
  v->pgc[1] = 1.0f / n;
}
static float drift_float(uint32_t* seed) /* RandomScalarSource::getFloat, S:189-202 */
{
  *seed = *seed * 0x0019660Du + 0x3C6EF35Fu;
  uint32_t temp = ((*seed >> 9) & 0x007FFFFFu) | 0x3F800000u;
  float f = u2f(temp);
  f *= 2.f;
  f -= 3.f;
  return f;
}
/* SampleAccurateLinearGlide::nextSample for one sample */
static float sa_glide_next(port_voice* v, float x)
{
  float y;

  {
    float curr = u2f(v->pg[0]), step = u2f(v->pg[1]), target = u2f(v->pg[2]);
    int32_t remaining = (int32_t)v->pg[3];
    const int32_t per = (int32_t)v->pgc[0];
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
      step = (target - curr) * v->pgc[1];
      remaining--;
    }
    else
    {
      curr += step;
      remaining--;
    }
    v->pg[0] = f2u(curr), v->pg[1] = f2u(step), v->pg[2] = f2u(target), v->pg[3] = (uint32_t)remaining;
    y = curr;
  }
  return y;
}
/* one output frame: E:134-140 / 185-189 / 225-236 */
static void voice_frame(port_voice* v, int t, float gate)
{
  v->out[ROW_GATE][t] = gate;
  v->out[ROW_PITCH][t] = sa_glide_next(v, v->pitch);
  v->age += v->ageStep;
  v->out[ROW_TIME][t] = (float)((double)v->age / (double)(float)v->sr); /* samplesToSeconds, E:12-18 */
}
static void voice_write_frames(port_voice* v, int endFrame) /* E:129-142 */
{
  for (int t = v->nextFrame; t < endFrame; ++t) voice_frame(v, t, v->vel);
  v->nextFrame = endFrame;
}

static void voice_init(port_voice* v, int voiceIndex, float sr, float glideSeconds, float driftAmount, float bendRange)
{
  memset(v, 0, sizeof(*v));
  v->pg[3] = (uint32_t)-1; /* G:520-525 */
  v->pgc[0] = 32.f, v->pgc[1] = 1.f / 32;
  for (int i = 0; i < GL_COUNT; ++i) lin_glide_init(&v->g[i]);
  v->voiceIndex = voiceIndex;
  /* reset(), E:60-85 */
  v->seed = (uint32_t)(voiceIndex * 232);
  for (int i = GL_BEND; i <= GL_Z; ++i) lin_glide_set_value(&v->g[i], 0.f);
  for (int n = 0; n < NB; ++n) v->out[ROW_VOICE][n] = (float)voiceIndex - 1; /* E:292 */
  v->sr = sr, v->recalc = 1; /* setSampleRate */
  v->glideSeconds = glideSeconds;
  v->driftAmount = driftAmount;
  v->pitchBendRange = bendRange;
}

static void lin_glide_set_time(lin_glide* g, float t) { mlport_coeffs_glide(t, g->c); }

static void voice_begin(port_voice* v) /* E:90-124 */
{
  if (v->recalc)
  {
    v->pitchGlideTimeInSamples = (int)(v->sr * v->glideSeconds);
    if (!v->inhibit) sa_glide_set_time(v, (float)v->pitchGlideTimeInSamples);
    for (int i = GL_BEND; i <= GL_Z; ++i) lin_glide_set_time(&v->g[i], (float)(v->sr * 0.02f));
    lin_glide_set_time(&v->g[GL_DRIFT], (float)(v->sr * 8.0f));
    /* SmoothedController::process, E:274-285: int glideTimeInSamples = sr * kControllerGlideTimeSeconds */
    lin_glide_set_time(&v->g[GL_PRESSURE], (float)(int)(v->sr * 0.02f));
    v->recalc = 0;
  }
  v->nextFrame = 0;
  v->driftCounter += NB;
  if (v->driftCounter >= v->nextDriftTime)
  {
    float d = drift_float(&v->seed);
    float nextTimeMul = 1.0f + fabsf(drift_float(&v->seed));
    v->curDrift = d;
    v->driftCounter = 0;
    v->nextDriftTime = (int)(v->sr * nextTimeMul * 8.0f);
  }
}
static void voice_note(port_voice* v, int type, int time, float v1, float v2, int doGlide, int doReset) /* E:126-220 */
{
  int destTime = time < 0 ? 0 : (time > NB ? NB : time);
  switch (type)
  {
    case MLB_EV_NOTE_ON:
      if (doReset) v->age = 0;
      v->ageStep = 1;
      v->inhibit = !doGlide;
      sa_glide_set_time(v, doGlide ? (float)v->pitchGlideTimeInSamples : 0.f);
      voice_write_frames(v, destTime);
      v->pitch = v1;
      v->vel = v2;
      break;
    case MLB_EV_NOTE_RETRIG:
      if (doReset) v->age = 0;
      v->ageStep = 1;
      if (destTime == 0) destTime++;
      voice_write_frames(v, destTime - 1);
      voice_frame(v, destTime - 1, 0.f);
      v->pitch = v1;
      v->vel = v2;
      v->nextFrame = destTime;
      break;
    case MLB_EV_NOTE_OFF:
      voice_write_frames(v, destTime);
      v->vel = 0.f;
      break;
    default: break;
  }
}
static void voice_end(port_voice* v) /* E:222-262 */
{
  for (int t = v->nextFrame; t < NB; ++t) voice_frame(v, t, v->vel);
  float bend[NB], drift[NB];
  gen_glide(v->g[GL_BEND].st, v->g[GL_BEND].c, v->bend, v->g[GL_BEND].curr, bend);
  gen_glide(v->g[GL_DRIFT].st, v->g[GL_DRIFT].c, v->curDrift, v->g[GL_DRIFT].curr, drift);
  gen_glide(v->g[GL_MOD].st, v->g[GL_MOD].c, v->mod, v->g[GL_MOD].curr, v->out[ROW_MOD]);
  gen_glide(v->g[GL_X].st, v->g[GL_X].c, v->x, v->g[GL_X].curr, v->out[ROW_X]);
  gen_glide(v->g[GL_Y].st, v->g[GL_Y].c, v->y, v->g[GL_Y].curr, v->out[ROW_Y]);
  if (v->vel == 0.f) v->z = 0.f;
  gen_glide(v->g[GL_Z].st, v->g[GL_Z].c, v->z, v->g[GL_Z].curr, v->out[ROW_Z]);
  for (int n = 0; n < NB; ++n)
  {
    float p = v->out[ROW_PITCH][n];
    p = p + (bend[n] * v->pitchBendRange) * (1.f / 12);
    p = p + (drift[n] * v->driftAmount) * 0.02f;
    v->out[ROW_PITCH][n] = p;
  }
  if (v->midi) /* processVector, MIDI: z += smoothed channel pressure, E:432-447 */
  {
    float pr[NB];
    gen_glide(v->g[GL_PRESSURE].st, v->g[GL_PRESSURE].c, v->pressure, v->g[GL_PRESSURE].curr, pr);
    for (int n = 0; n < NB; ++n) v->out[ROW_Z][n] = v->out[ROW_Z][n] + pr[n];
  }
}

typedef struct mlport_voice_bank
{
  int V;
  port_voice* v;
  int32_t* main_voice; /* MPE: index of each voice's main voice or -1; NULL = none */
} mlport_voice_bank;

void mlport_bank_set_main_voices(mlport_voice_bank* b, const int32_t* main_voice)
{
  free(b->main_voice);
  b->main_voice = NULL;
  if (!main_voice) return;
  b->main_voice = (int32_t*)malloc(sizeof(int32_t) * (size_t)b->V);
  memcpy(b->main_voice, main_voice, sizeof(int32_t) * (size_t)b->V);
}

mlport_voice_bank* mlport_bank_create(int V, float sr, const int32_t* voiceIndex, const float* glideSeconds,
                                      const float* driftAmount, const float* pitchBend, unsigned flags)
{
  mlport_voice_bank* b = (mlport_voice_bank*)calloc(1, sizeof(*b));
  b->V = V;
  b->v = (port_voice*)calloc((size_t)V, sizeof(port_voice));
  for (int i = 0; i < V; ++i)
  {
    voice_init(&b->v[i], voiceIndex[i], sr, glideSeconds[i], driftAmount[i], pitchBend[i]);
    b->v[i].midi = (flags & MLB_VOICES_MIDI) != 0;
  }
  return b;
}
void mlport_bank_destroy(mlport_voice_bank* b)
{
  if (!b) return;
  free(b->v);
  free(b->main_voice);
  free(b);
}
double mlport_bank_process(mlport_voice_bank* b, int T, const mlb_voice_events* ev, float* out, int nthreads)
{
  (void)nthreads;
  const int V = b->V;
  for (int i = 0; i < V; ++i)
  {
    port_voice* v = &b->v[i];
    for (int t = 0; t < T; ++t)
    {
      const mlb_voice_events* r = &ev[(size_t)t * V + i];
      voice_begin(v);
      for (int k = 0; k < r->n_events; ++k)
        voice_note(v, r->type[k], r->time[k], r->value1[k], r->value2[k], (r->flags[k] & MLB_EVF_GLIDE) != 0,
                   (r->flags[k] & MLB_EVF_RESET) != 0);
      if (r->set_mask & MLB_SET_BEND) v->bend = r->bend;
      if (r->set_mask & MLB_SET_MOD) v->mod = r->mod;
      if (r->set_mask & MLB_SET_X) v->x = r->x;
      if (r->set_mask & MLB_SET_Y) v->y = r->y;
      if (r->set_mask & MLB_SET_Z) v->z = r->z;
      if (r->set_mask & MLB_SET_PRESSURE) v->pressure = r->pressure;
      voice_end(v);
      for (int row = 0; row < 8; ++row)
        memcpy(out + (((size_t)t * 8 + row) * V + i) * NB, v->out[row], sizeof(float) * NB);
    }
  }
  if (b->main_voice) /* processVector, MPE: channel voices += the main voice's rows, E:448-460 */
  {
    static const int rows[5] = {ROW_PITCH, ROW_X, ROW_Y, ROW_Z, ROW_MOD};
    for (int t = 0; t < T; ++t)
      for (int i = 0; i < V; ++i)
      {
        const int m = b->main_voice[i];
        if (m < 0) continue;
        for (int k = 0; k < 5; ++k)
        {
          float* dst = out + (((size_t)t * 8 + rows[k]) * V + i) * NB;
          const float* src = out + (((size_t)t * 8 + rows[k]) * V + m) * NB;
          for (int n = 0; n < NB; ++n) dst[n] = dst[n] + src[n];
        }
      }
  }
  return 0.0;
}

/* ------------------------------------------------------------------ */
/* Upsampler / Downsampler, F:1316-1473 (one per voice) */
typedef struct mlport_resampler
{
  int dir, oct, V;
  uint32_t* st;      /* [V][oct][9] HalfBandFilter states */
  float* buf;        /* down: [V][2*oct+1][64]; up: [V][1<<oct][64] */
  uint32_t* counter; /* down: [V] */
} mlport_resampler;

mlport_resampler* mlport_resampler_create(int dir, int oct, int V)
{
  mlport_resampler* r = (mlport_resampler*)calloc(1, sizeof(*r));
  r->dir = dir, r->oct = oct, r->V = V;
  const int nbuf = dir == MLB_RESAMPLE_UP ? (1 << oct) : 2 * oct + 1;
  r->st = (uint32_t*)calloc((size_t)V * (size_t)oct * 9 + 1, 4);
  r->buf = (float*)calloc((size_t)V * (size_t)nbuf * NB, 4);
  r->counter = (uint32_t*)calloc((size_t)V, 4);
  return r;
}
void mlport_resampler_destroy(mlport_resampler* r)
{
  if (!r) return;
  free(r->st), free(r->buf), free(r->counter);
  free(r);
}
int mlport_resampler_process(mlport_resampler* r, const float* in, float* out, int T)
{
  const int V = r->V, oct = r->oct, N = 1 << oct;
  int produced = 0;
  for (int v = 0; v < V; ++v)
  {
    uint32_t* st = r->st + (size_t)v * oct * 9;
    if (r->dir == MLB_RESAMPLE_UP)
    {
      float* buf = r->buf + (size_t)v * N * NB; /* Upsampler::write, F:1428-1453 */
      for (int t = 0; t < T; ++t)
      {
        memcpy(buf + (size_t)(N - 1) * NB, in + ((size_t)t * V + v) * NB, sizeof(float) * NB);
        for (int j = 0; j < oct; ++j)
        {
          const int sourceBufs = 1 << j, destBufs = sourceBufs << 1;
          const int srcStart = N - sourceBufs, destStart = N - destBufs;
          for (int i = 0; i < sourceBufs; ++i)
          {
            float src[NB], d1[NB], d2[NB];
            memcpy(src, buf + (size_t)(srcStart + i) * NB, sizeof(src));
            hb_upsample(st + j * 9, src, d1, d2);
            memcpy(buf + (size_t)(destStart + 2 * i) * NB, d1, sizeof(d1));
            memcpy(buf + (size_t)(destStart + 2 * i + 1) * NB, d2, sizeof(d2));
          }
        }
        for (int k = 0; k < N; ++k) /* the 2^octaves reads */
          memcpy(out + (((size_t)t * N + k) * V + v) * NB, buf + (size_t)k * NB, sizeof(float) * NB);
      }
      produced = T * N;
    }
    else
    {
      float* buf = r->buf + (size_t)v * (2 * oct + 1) * NB; /* Downsampler::write, F:1347-1386 */
      uint32_t counter = r->counter[v];
      int n = 0;
      for (int t = 0; t < T; ++t)
      {
        memcpy(buf + (size_t)(counter & 1u) * NB, in + ((size_t)t * V + v) * NB, sizeof(float) * NB);
        uint32_t mask = 1;
        for (int h = 0; h < oct; ++h)
        {
          if (!(counter & mask)) break;
          mask <<= 1;
          const int b1 = (counter & mask) != 0;
          float d[NB];
          hb_downsample(st + h * 9, buf + (size_t)(h * 2) * NB, buf + (size_t)(h * 2 + 1) * NB, d);
          memcpy(buf + (size_t)(h * 2 + 2 + b1) * NB, d, sizeof(d));
        }
        counter = (counter + 1) & (uint32_t)(N - 1);
        if (counter == 0) memcpy(out + ((size_t)(n++) * V + v) * NB, buf + (size_t)(2 * oct) * NB, sizeof(float) * NB);
      }
      r->counter[v] = counter;
      produced = n;
    }
  }
  return produced;
}
