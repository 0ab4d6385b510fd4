// voice_kernel.cuh -- K7: EventsToSignals::Voice for V voices.
// Reference: source/app/MLEventsToSignals.cpp:47-263 (E below), MLEventsToSignals.h:99-168,
// glides source/DSP/MLDSPGens.h:433-590 (G).  One lane = one Voice; a launch runs n_blocks vectors:
// beginProcess (drift), the vector's note events (sample-accurate gate / pitch glide / age), the
// instantaneous controller values, endProcess (vector-accurate LinearGlides, bend and drift added to
// pitch).  Input: one 68-byte mlb_voice_events record per voice and vector; output: up to 8 rows.
// Scalar state is SoA [word][V]; the six LinearGlide::mCurrVec rows are only touched while a glide
// is moving (an idle LinearGlide's row equals its target in every lane of the row).
#pragma once
#include "functors.cuh"

namespace mlb
{
enum VoiceGlide { VG_BEND = 0, VG_MOD, VG_X, VG_Y, VG_Z, VG_DRIFT, VG_COUNT };
enum VoiceState
{
  VS_PG_CURR = 0, VS_PG_STEP, VS_PG_TARGET, VS_PG_REM, VS_PG_PER, VS_PG_DY,  // pitchGlide (G:517-590)
  VS_GL = 6,                       // 6 x {step, target, vectorsRemaining}
  VS_VEL = VS_GL + 3 * VG_COUNT,   // currentVelocity, currentPitch, bend, mod, x, y, z
  VS_PITCH, VS_BEND, VS_MOD, VS_X, VS_Y, VS_Z,
  VS_AGE, VS_AGE_STEP, VS_SEED, VS_DRIFT_COUNTER, VS_NEXT_DRIFT, VS_CUR_DRIFT,
  VS_COUNT
};
enum VoiceCoef { VC_GLIDE_SAMPLES = 0, VC_DRIFT_AMOUNT, VC_BEND_RANGE, VC_VOICE_ROW, VC_COUNT };

struct VoiceArgs
{
  const mlb_voice_events* ev;  // [T][V]
  float* out;                  // [T][8][V][64]
  uint32_t* state;             // [VS_COUNT][V]
  const float* coef;           // [VC_COUNT][V]
  float* grows;                // [VG_COUNT][V][64]  LinearGlide::mCurrVec
  int V, T;
  unsigned row_mask;
  float sr;                    // Voice::sr (a double in the reference, set from this value)
  float gl_per, gl_dy;         // LinearGlide coefficients of bend/mod/x/y/z (sr * kGlideTimeSeconds)
  float dr_per, dr_dy;         // ... of pitchDriftGlide (sr * kDriftTimeSeconds)
};

struct VoiceRegs
{
  uint32_t pg[4];
  float pg_per_f, pg_dy;
  float vel, pitch;
  uint32_t age, age_step;
  int next_frame;
};

// SampleAccurateLinearGlide::setGlideTimeInSamples, G:527-532
MLB_DEV void voice_set_glide_time(VoiceRegs& r, float t)
{
  int n = cvt_trunc(t);
  if (n < 1) n = 1;
  r.pg_per_f = __int2float_rn(n);
  r.pg_dy = __fdiv_rn(1.0f, __int2float_rn(n));
}
// one output frame, E:134-140: gate, glided pitch, age -> seconds (samplesToSeconds, E:12-18)
MLB_DEV void voice_frame(VoiceRegs& r, float sr, int t, float gate_v, float* gate, float* pitch, float* tm)
{
  gate[t] = gate_v;
  const float co[2] = {r.pg_per_f, r.pg_dy};
  pitch[t] = sample_glide_tick<true>(r.pitch, r.pg, co);
  r.age += r.age_step;
  tm[t] = __double2float_rn(__ddiv_rn((double)r.age, (double)sr));
}
MLB_DEV void voice_write_frames(VoiceRegs& r, float sr, int end_frame, float* gate, float* pitch, float* tm)
{
  for (int t = r.next_frame; t < end_frame; ++t) voice_frame(r, sr, t, r.vel, gate, pitch, tm);
  r.next_frame = end_frame;
}

// LinearGlide::operator()(float) for 4 consecutive samples (G:459-505); mode as in run_glide_node
struct GlidePlan
{
  int mode;  // -1 idle, 0 land on target, 1 start, 2 continue
  float step, target, cv;
  int remaining;
};
MLB_DEV GlidePlan glide_plan(uint32_t* st, float f, float per_f, float dy, const float* row)
{
  GlidePlan g;
  g.step = u2f(st[0]), g.target = u2f(st[1]), g.remaining = (int32_t)st[2];
  g.cv = 0.f;
  const int per = cvt_trunc(per_f);
  if (f != g.target)
  {
    g.target = f;
    g.remaining = per;
  }
  if (g.remaining < 0)
    g.mode = -1;
  else if (g.remaining == 0)
  {
    g.mode = 0;
    g.step = 0.f;
    g.remaining--;
  }
  else if (g.remaining == per)
  {
    g.mode = 1;
    g.cv = row[MLB_BLOCK - 1];
    g.step = __fmul_rn(__fsub_rn(g.target, g.cv), dy);
    g.remaining--;
  }
  else
  {
    g.mode = 2;
    g.remaining--;
  }
  st[0] = f2u(g.step), st[1] = f2u(g.target), st[2] = (uint32_t)g.remaining;
  return g;
}
// samples 4q .. 4q+3 of the glide's output row; updates mCurrVec in delay memory when it moves
MLB_DEV float4 glide_quad(const GlidePlan& g, float* row, int q)
{
  float4 y;
  if (g.mode <= 0)
    y = make_float4(g.target, g.target, g.target, g.target);  // idle rows equal their target
  else if (g.mode == 1)
  {
    y.x = __fadd_rn(g.cv, __fmul_rn(unity_ramp(4 * q), g.step));
    y.y = __fadd_rn(g.cv, __fmul_rn(unity_ramp(4 * q + 1), g.step));
    y.z = __fadd_rn(g.cv, __fmul_rn(unity_ramp(4 * q + 2), g.step));
    y.w = __fadd_rn(g.cv, __fmul_rn(unity_ramp(4 * q + 3), g.step));
  }
  else
  {
    y = reinterpret_cast<const float4*>(row)[q];
    y.x = __fadd_rn(y.x, g.step), y.y = __fadd_rn(y.y, g.step), y.z = __fadd_rn(y.z, g.step), y.w = __fadd_rn(y.w, g.step);
  }
  if (g.mode >= 0) reinterpret_cast<float4*>(row)[q] = y;
  return y;
}

__global__ void __launch_bounds__(128) voice_bank_kernel(const VoiceArgs a)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= a.V) return;
  const size_t V = (size_t)a.V;
  uint32_t st[VS_COUNT];
#pragma unroll
  for (int i = 0; i < VS_COUNT; ++i) st[i] = a.state[(size_t)i * V + v];
  const float glide_samples = a.coef[(size_t)VC_GLIDE_SAMPLES * V + v];  // (float)pitchGlideTimeInSamples
  const float drift_amount = a.coef[(size_t)VC_DRIFT_AMOUNT * V + v];
  const float bend_range = a.coef[(size_t)VC_BEND_RANGE * V + v];
  const float voice_row = a.coef[(size_t)VC_VOICE_ROW * V + v];
  float* grow[VG_COUNT];
#pragma unroll
  for (int i = 0; i < VG_COUNT; ++i) grow[i] = a.grows + ((size_t)i * V + v) * MLB_BLOCK;

  VoiceRegs r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.pg[i] = st[VS_PG_CURR + i];
  r.pg_per_f = u2f(st[VS_PG_PER]), r.pg_dy = u2f(st[VS_PG_DY]);
  r.vel = u2f(st[VS_VEL]), r.pitch = u2f(st[VS_PITCH]);
  r.age = st[VS_AGE], r.age_step = st[VS_AGE_STEP];
  float cur[5] = {u2f(st[VS_BEND]), u2f(st[VS_MOD]), u2f(st[VS_X]), u2f(st[VS_Y]), u2f(st[VS_Z])};
  uint32_t seed = st[VS_SEED];
  int drift_counter = (int)st[VS_DRIFT_COUNTER], next_drift = (int)st[VS_NEXT_DRIFT];
  float cur_drift = u2f(st[VS_CUR_DRIFT]);

  float gate[MLB_BLOCK], pitch[MLB_BLOCK], tm[MLB_BLOCK];  // frames can be revisited (retrigger): local rows

  for (int t = 0; t < a.T; ++t)
  {
    // ---- beginProcess, E:90-124 ----
    r.next_frame = 0;
    drift_counter += MLB_BLOCK;
    if (drift_counter >= next_drift)
    {
      const float d = noise_tick(seed);  // RandomScalarSource::getFloat == NoiseGen's LCG step
      const float next_mul = __fadd_rn(1.0f, fabsf(noise_tick(seed)));
      cur_drift = d;
      drift_counter = 0;
      next_drift = __double2int_rz(__dmul_rn(__dmul_rn((double)a.sr, (double)next_mul), 8.0));
    }
    // ---- the vector's note events, E:126-220 ----
    const uint32_t* rec = reinterpret_cast<const uint32_t*>(a.ev + ((size_t)t * V + v));
    const uint32_t head = rec[0], times = rec[1], types = rec[2], flags = rec[3];
    const int n_events = (int)(head & 0xFFu);
    const unsigned set_mask = (head >> 8) & 0xFFu;
    for (int k = 0; k < n_events && k < MLB_VOICE_MAX_EVENTS; ++k)
    {
      int dest = (int)((times >> (8 * k)) & 0xFFu);
      dest = dest > MLB_BLOCK ? MLB_BLOCK : dest;
      const int type = (int)((types >> (8 * k)) & 0xFFu);
      const unsigned fl = (flags >> (8 * k)) & 0xFFu;
      const float v1 = u2f(rec[4 + k]), v2 = u2f(rec[8 + k]);
      if (type == MLB_EV_NOTE_ON)
      {
        if (fl & MLB_EVF_RESET) r.age = 0;
        r.age_step = 1;
        voice_set_glide_time(r, (fl & MLB_EVF_GLIDE) ? glide_samples : 0.f);
        voice_write_frames(r, a.sr, dest, gate, pitch, tm);
        r.pitch = v1;
        r.vel = v2;
      }
      else if (type == MLB_EV_NOTE_RETRIG)
      {
        if (fl & MLB_EVF_RESET) r.age = 0;
        r.age_step = 1;
        if (dest == 0) dest++;
        voice_write_frames(r, a.sr, dest - 1, gate, pitch, tm);
        voice_frame(r, a.sr, dest - 1, 0.f, gate, pitch, tm);  // the retrigger frame: gate 0
        r.pitch = v1;
        r.vel = v2;
        r.next_frame = dest;
      }
      else if (type == MLB_EV_NOTE_OFF)
      {
        voice_write_frames(r, a.sr, dest, gate, pitch, tm);
        r.vel = 0.f;
      }
    }
    if (set_mask & MLB_SET_BEND) cur[0] = u2f(rec[12]);
    if (set_mask & MLB_SET_MOD) cur[1] = u2f(rec[13]);
    if (set_mask & MLB_SET_X) cur[2] = u2f(rec[14]);
    if (set_mask & MLB_SET_Y) cur[3] = u2f(rec[15]);
    if (set_mask & MLB_SET_Z) cur[4] = u2f(rec[16]);
    // ---- endProcess, E:222-262 ----
    for (int f = r.next_frame; f < MLB_BLOCK; ++f) voice_frame(r, a.sr, f, r.vel, gate, pitch, tm);
    if (r.vel == 0.f) cur[4] = 0.f;
    GlidePlan gp[VG_COUNT];
#pragma unroll
    for (int i = 0; i < VG_COUNT; ++i)
      gp[i] = glide_plan(&st[VS_GL + 3 * i], i == VG_DRIFT ? cur_drift : cur[i], i == VG_DRIFT ? a.dr_per : a.gl_per,
                         i == VG_DRIFT ? a.dr_dy : a.gl_dy, grow[i]);
    float* o = a.out + ((size_t)t * MLB_VOICE_ROWS * V + v) * MLB_BLOCK;
    const size_t row_stride = V * MLB_BLOCK;
#pragma unroll 1
    for (int q = 0; q < 16; ++q)
    {
      float4 g[VG_COUNT];
#pragma unroll
      for (int i = 0; i < VG_COUNT; ++i) g[i] = glide_quad(gp[i], grow[i], q);
      float4 p = make_float4(pitch[4 * q], pitch[4 * q + 1], pitch[4 * q + 2], pitch[4 * q + 3]);
      const float k12 = 1.f / 12;
      // pitch += bendGlide * pitchBend * (1/12); pitch += driftSig * driftAmount * kDriftScale  (E:255-261)
      p.x = __fadd_rn(p.x, __fmul_rn(__fmul_rn(g[VG_BEND].x, bend_range), k12));
      p.y = __fadd_rn(p.y, __fmul_rn(__fmul_rn(g[VG_BEND].y, bend_range), k12));
      p.z = __fadd_rn(p.z, __fmul_rn(__fmul_rn(g[VG_BEND].z, bend_range), k12));
      p.w = __fadd_rn(p.w, __fmul_rn(__fmul_rn(g[VG_BEND].w, bend_range), k12));
      p.x = __fadd_rn(p.x, __fmul_rn(__fmul_rn(g[VG_DRIFT].x, drift_amount), 0.02f));
      p.y = __fadd_rn(p.y, __fmul_rn(__fmul_rn(g[VG_DRIFT].y, drift_amount), 0.02f));
      p.z = __fadd_rn(p.z, __fmul_rn(__fmul_rn(g[VG_DRIFT].z, drift_amount), 0.02f));
      p.w = __fadd_rn(p.w, __fmul_rn(__fmul_rn(g[VG_DRIFT].w, drift_amount), 0.02f));
      float4* o4 = reinterpret_cast<float4*>(o) + q;
      const size_t rs4 = row_stride / 4;
      if (a.row_mask & 1u) __stcs(o4 + 0 * rs4, p);                                                                      // kPitch
      if (a.row_mask & 2u) __stcs(o4 + 1 * rs4, make_float4(gate[4 * q], gate[4 * q + 1], gate[4 * q + 2], gate[4 * q + 3]));  // kGate
      if (a.row_mask & 4u) __stcs(o4 + 2 * rs4, make_float4(voice_row, voice_row, voice_row, voice_row));                 // kVoice
      if (a.row_mask & 8u) __stcs(o4 + 3 * rs4, g[VG_Z]);                                                                // kZ
      if (a.row_mask & 16u) __stcs(o4 + 4 * rs4, g[VG_X]);                                                               // kX
      if (a.row_mask & 32u) __stcs(o4 + 5 * rs4, g[VG_Y]);                                                               // kY
      if (a.row_mask & 64u) __stcs(o4 + 6 * rs4, g[VG_MOD]);                                                             // kMod
      if (a.row_mask & 128u) __stcs(o4 + 7 * rs4, make_float4(tm[4 * q], tm[4 * q + 1], tm[4 * q + 2], tm[4 * q + 3]));   // kElapsedTime
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) st[VS_PG_CURR + i] = r.pg[i];
  st[VS_PG_PER] = f2u(r.pg_per_f), st[VS_PG_DY] = f2u(r.pg_dy);
  st[VS_VEL] = f2u(r.vel), st[VS_PITCH] = f2u(r.pitch);
  st[VS_AGE] = r.age, st[VS_AGE_STEP] = r.age_step;
  st[VS_BEND] = f2u(cur[0]), st[VS_MOD] = f2u(cur[1]), st[VS_X] = f2u(cur[2]), st[VS_Y] = f2u(cur[3]), st[VS_Z] = f2u(cur[4]);
  st[VS_SEED] = seed, st[VS_DRIFT_COUNTER] = (uint32_t)drift_counter, st[VS_NEXT_DRIFT] = (uint32_t)next_drift;
  st[VS_CUR_DRIFT] = f2u(cur_drift);
#pragma unroll
  for (int i = 0; i < VS_COUNT; ++i) a.state[(size_t)i * V + v] = st[i];
}

}  // namespace mlb
